"""phase clocks of ransac_select_kernel (debug library: make -C putslam_b200/csrc dbg)"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api, host, synth
api.LIB_PATH = os.path.join(os.path.dirname(api.LIB_PATH), "libpslam_b200_dbg.so")
ctx = api.Context(0)
mf = synth.map_frame(M=5000, N=1000, seed=0)
ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
names = ["winner", "inlier list", "stage inliers", "means", "sigma", "svd+refit", "recount", "(end)"]
for num_hyp in (4096, 0):
    acc = np.zeros(7)
    for i in range(8):
        r = ctx.frame_to_map(mf["map_xyz"].astype(np.float32), mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55,
                             0, seed=i, num_hyp=num_hyp, match_cap=4096)
        clk = (C.c_longlong * 16)()
        ctx.lib.pslam_debug_select_clocks(clk)
        c = np.array(clk[:8], dtype=np.float64)
        if i >= 3:
            acc += np.diff(c)
    acc /= 5
    print("num_hyp", num_hyp, "inliers", r["inliers"].size, "total cycles", int(acc.sum()))
    for n, v in zip(names, acc):
        print(f"   {n:14s} {v:9.0f} cycles  {v / 1965:6.2f} us")
