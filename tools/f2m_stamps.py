"""host-side phase breakdown of pslam_frame_to_map_features at C3 (us): python tools/f2m_stamps.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api, synth
ctx = api.Context(0)
mf = synth.map_frame(M=5000, N=1000, seed=0)
args = (mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"], mf["cur_detdist"])
for pin in (False, True):
    if pin:
        for a in args[:4]: ctx.host_register(np.ascontiguousarray(a))
    acc = np.zeros(8); wall = 0.0
    for i in range(60):
        t0 = time.perf_counter()
        r = ctx.frame_to_map_features(*args, 0.12, 0.55, 0, seed=i, num_hyp=4096, match_cap=2048)
        dt = time.perf_counter() - t0
        if i >= 10:
            acc += np.array(ctx.host_stamps()); wall += dt
    acc /= 50
    print("pinned map" if pin else "staged", "python wall us", round(wall / 50 * 1e6, 1), "stamps(us): packed %.1f h2d-enq %.1f kernels-enq %.1f d2h-enq %.1f synced %.1f unpacked %.1f" % tuple(acc[1:7]))
