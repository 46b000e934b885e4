"""runs pslam_orb_detect + pslam_orb_describe a few times on a 640x480 frame (for ncu launch lists)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from putslam_b200 import api
ctx = api.Context(0)
img = bench.orb_bench_image(np.random.default_rng(77))
for i in range(5):
    det = ctx.orb_detect(img, 500)
    order, desc = ctx.orb_describe(img, det["xy"], det["octave"], det["angle"])
print("ok", det["octave"].size, order.size, int(desc.sum()))
