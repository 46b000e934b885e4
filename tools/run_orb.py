"""runs pslam_orb_describe a few times on a 640x480 frame with 1000 keypoints (for ncu launch lists)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api
ctx = api.Context(0)
rng = np.random.default_rng(77)
img = rng.integers(0, 256, (480, 640), dtype=np.uint8)
n = 1000
xy = np.stack([rng.uniform(35, 604, n), rng.uniform(35, 444, n)], 1).astype(np.float32)
octave = rng.integers(0, 8, n).astype(np.int32); angle = rng.uniform(0, 360, n).astype(np.float32)
for i in range(5):
    order, desc = ctx.orb_describe(img, xy, octave, angle)
print("ok", order.size, int(desc.sum()))
