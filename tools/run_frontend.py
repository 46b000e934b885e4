"""runs the C3 frame-to-map and C2 frame-to-frame pipelines a few times (for ncu launch lists)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api, host, synth
ctx = api.Context(0)
mf = synth.map_frame(M=5000, N=1000, seed=0)
ml = host.map_levels(mf["map_xyz"], mf["map_octave"], mf["map_detdist"])
cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
for i in range(6):
    r = ctx.frame_to_map(mf["map_xyz"].astype(np.float32), mf["map_desc"], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, 0,
                         seed=i, num_hyp=4096, match_cap=4096)
seq = synth.Sequence(n_frames=1000, n_kp=1000, seed=42)
f0, f1 = seq.frame(0), seq.frame(1)
p = ctx.frame_to_frame(None, None, f0["desc"], f0["uv"], f0["depth"])
for i in range(6):
    c = ctx.frame_to_frame(f0["desc"], p["xyz"], f1["desc"], f1["uv"], f1["depth"], seed=i)
print("ok", r["inliers"].size, c["inliers"].size)
