"""ncu target: a few predict() calls on a one-frame (307k surfels) and a four-frame model."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from slam_b200.predict import ModelPredictor  # noqa: E402
from tests.support import MODEL_CUTOFF, make_scene  # noqa: E402
from slam_b200.synth import surfels_from_frame  # noqa: E402

scene, intr = make_scene(640, 480)
poses = scene.trajectory(1000)
mp = ModelPredictor(intr["width"], intr["height"], intr["cx"], intr["cy"], intr["fx"], intr["fy"])
depth, rgba = scene.render_frame(poses[5])
d_depth = torch.from_numpy(depth.view(np.int16).copy()).cuda()
d_rgba = torch.from_numpy(rgba).cuda()
nframes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
model = np.concatenate([surfels_from_frame(scene, poses[40 * k], seed=k) for k in range(nframes)])
d_model = torch.from_numpy(model).cuda()
for it in range(4):
    mp.predict(poses[5 + (it % 3)], d_model, len(model), MODEL_CUTOFF, 10.0, 1, 200, d_depth, d_rgba)
mp.lastMs()
