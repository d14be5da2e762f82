"""TEST INFRASTRUCTURE -- CPU restatement (numpy / plain Python) of the reference's fern relocaliser, src/lc/Ferns.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path is the CUDA
library.  **Parity unpinned**: the reference holds no golden vectors for this path and its resize step is a GL pass
(gl/Resize.cpp + gl/shaders/resize.frag, no GL context here); the restatement follows the source line by line:

  resize            gl/Resize.cpp:70-154 + resize.frag: texture2D at the centre of every destination pixel of a
                    NEAREST-filtered texture (gl/types.cpp:18, draw == false) = source texel (8x+4, 8y+4)
  encode            Ferns.cpp:108-131 (addFrame), 197-219 (findFrame)
  inverted lists    Ferns.cpp:121-124, 155-161: conservatory[i].ids[code] lists the key frames that hold `code` at fern i
  dissimilarity     Ferns.cpp:136-147 (addFrame, all frames), 227-239 (findFrame, time - srcTime > 300)
  blockHDAware      Ferns.cpp:374-389
  photometricCheck  Ferns.cpp:309-357 (fp32; Eigen's Matrix4f inverse restated with numpy float32)
"""
from __future__ import annotations

import numpy as np

FACTOR = 8
BAD = 255
F32 = np.float32


def resize(img: np.ndarray) -> np.ndarray:
    """[H][W][C] -> [H/8][W/8][C], texel (8x+4, 8y+4)."""
    return np.ascontiguousarray(img[FACTOR // 2::FACTOR, FACTOR // 2::FACTOR])


class FernsOracle:
    def __init__(self, table: np.ndarray, maxDepth: int, photoThresh: float, cx, cy, fx, fy, w, h):
        self.table = np.asarray(table, np.int64)          # [num][x, y, r, g, b, d]
        self.num = len(self.table)
        self.width, self.height = w // FACTOR, h // FACTOR
        self.maxDepth, self.photoThresh = int(maxDepth), F32(photoThresh)
        self.cx, self.cy = F32(cx) / F32(FACTOR), F32(cy) / F32(FACTOR)
        self.invfx = F32(1.0) / (F32(fx) / F32(FACTOR))
        self.invfy = F32(1.0) / (F32(fy) / F32(FACTOR))
        self.ids = [[[] for _ in range(16)] for _ in range(self.num)]   # conservatory[i].ids[code]
        self.frames = []                                                  # dicts: codes, goodCodes, pose, srcTime, rgb, vert, norm

    # Ferns.cpp:108-131
    def encode(self, rgb_small, vert_small):
        codes = np.full(self.num, BAD, np.uint8)
        good = 0
        for i, (x, y, r, g, b, d) in enumerate(self.table):
            z = F32(vert_small[y, x, 2])
            if z > 0:
                pix = rgb_small[y, x]
                codes[i] = (int(pix[0] > r) << 3) | (int(pix[1] > g) << 2) | (int(pix[2] > b) << 1) | int(int(F32(z * F32(1000.0))) > d)
                good += 1
        return codes, good

    def co_occurrences(self, codes):
        co = np.zeros(len(self.frames), np.int64)
        for i in range(self.num):
            if codes[i] != BAD:
                for j in self.ids[i][codes[i]]:
                    co[j] += 1
        return co

    def dissimilarities(self, codes, good):
        co = self.co_occurrences(codes)
        out = np.zeros(len(self.frames), np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            for i, fr in enumerate(self.frames):
                maxCo = F32(min(good, fr["goodCodes"]))
                out[i] = F32(maxCo - F32(co[i])) / maxCo
        return out, co

    def addFrame(self, rgba, vert4, norm4, pose, srcTime, threshold) -> bool:
        rgb, vert, norm = resize(rgba)[..., :3], resize(vert4), resize(norm4)
        codes, good = self.encode(rgb, vert)
        minimum = np.finfo(np.float32).max
        if good > 0:
            dis, _ = self.dissimilarities(codes, good)
            for d in dis:
                if d < minimum:
                    minimum = d
        if (minimum > F32(threshold) or len(self.frames) == 0) and good > 0:
            fid = len(self.frames)
            for i in range(self.num):
                if codes[i] != BAD:
                    self.ids[i][codes[i]].append(fid)
            self.frames.append(dict(codes=codes, goodCodes=good, pose=np.asarray(pose, np.float32).reshape(4, 4), srcTime=int(srcTime), rgb=rgb.copy(),
                                    vert=vert.copy(), norm=norm.copy()))
            return True
        return False

    # Ferns.cpp:374-389
    def blockHDAware(self, c1, c2) -> np.float32:
        both = (c1 != BAD) & (c2 != BAD)
        count = int(both.sum())
        val = F32((c1[both] == c2[both]).sum())
        with np.errstate(invalid="ignore", divide="ignore"):
            return val / F32(count)

    # the search half of findFrame, Ferns.cpp:182-241
    def search(self, rgba, vert4, norm4, time):
        rgb, vert, norm = resize(rgba)[..., :3], resize(vert4), resize(norm4)
        codes, good = self.encode(rgb, vert)
        dis, _ = self.dissimilarities(codes, good)
        minimum, minId = np.finfo(np.float32).max, -1
        for i, d in enumerate(dis):
            if d < minimum and time - self.frames[i]["srcTime"] > 300:
                minimum, minId = d, i
        hd = self.blockHDAware(codes, self.frames[minId]["codes"]) if minId != -1 else F32(0)
        return dict(codes=codes, goodCodes=good, dissim=dis, minId=minId, minimum=minimum, blockHDAware=hd, rgb=rgb, vert=vert, norm=norm)

    # Ferns.cpp:309-357, fp32 throughout
    def photometricCheck(self, vertSmall, imgSmall, estPose, fernPose, fernRgb):
        est, fern = np.asarray(estPose, np.float32).reshape(4, 4), np.asarray(fernPose, np.float32).reshape(4, 4)
        diff = (np.linalg.inv(fern.astype(np.float64)) @ est.astype(np.float64)).astype(np.float32)
        photoSum, photoCount = F32(0), 0
        fxr, fyr = F32(1) / self.invfx, F32(1) / self.invfy
        for (x, y, _r, _g, _b, _d) in self.table:
            v = vertSmall[y, x]
            z = F32(v[2])
            if z > 0 and int(F32(z * F32(1000.0))) < self.maxDepth:
                p = [F32(F32(F32(F32(diff[r, 0] * v[0]) + F32(diff[r, 1] * v[1])) + F32(diff[r, 2] * v[2])) + diff[r, 3]) for r in range(3)]
                with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
                    fu = F32(F32(F32(p[0] * fxr) / p[2]) + self.cx)
                    fv = F32(F32(F32(p[1] * fyr) / p[2]) + self.cy)
                if not (np.isfinite(fu) and np.isfinite(fv)):
                    continue
                u, w = int(fu), int(fv)
                if 0 <= u < self.width and 0 <= w < self.height and fernRgb[w, u].any():
                    photoSum += F32(np.abs(fernRgb[w, u].astype(np.int32) - imgSmall[y, x].astype(np.int32)).sum())
                    photoCount += 1
        with np.errstate(invalid="ignore", divide="ignore"):
            return photoSum / F32(photoCount), photoCount
