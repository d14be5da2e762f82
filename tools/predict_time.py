"""Timing aid: splat / resolve launch times of the model-prediction producer for models of 1, 4 and 8 frames' worth of surfels."""
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
for nframes in (1, 4, 8):
    model = np.concatenate([surfels_from_frame(scene, poses[40 * k], seed=k) for k in range(nframes)])
    d_model = torch.from_numpy(model).cuda()
    for both in (False, True):
        ts = []
        for it in range(30):
            mp.predict(poses[5 + (it % 3)], d_model, len(model), MODEL_CUTOFF, 10.0, 1, 200, d_depth, d_rgba, write_index_textures=both)
            ts.append(mp.lastMs())
        a = np.array(ts[5:]) * 1e3
        frags = mp.lastFragments()
        cov = (mp.winners()[1] >= 0).mean()
        print(f"surfels {len(model):8d} index_textures {int(both)}: splat {a[:, 0].mean():7.1f} us  resolve {a[:, 1].mean():6.1f} us  fragments {frags:9d} "
              f"({frags / max(a[:, 0].mean(), 1e-9) / 1e3:6.2f} G frag/s)  surfel read {len(model) * 48 / a[:, 0].mean() / 1e3:7.1f} GB/s  covered {cov:.3f}", flush=True)
