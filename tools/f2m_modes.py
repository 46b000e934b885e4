"""frame-to-map (C3) latency through the C ABI for the launch modes: graphs with programmatic edges / plain edges / no graphs.
Each mode runs in its own process (the switches are read once): python tools/f2m_modes.py"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, time
sys.path.insert(0, %r)
import numpy as np, torch
from putslam_b200 import api, synth
ctx = api.Context(0)
mf = synth.map_frame(M=5000, N=1000, seed=0)
args = (mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"], mf["cur_detdist"])
acc = np.zeros(8)
for i in range(110):
    r = ctx.frame_to_map_features(*args, 0.12, 0.55, 0, seed=i, num_hyp=4096, match_cap=2048)
    if i >= 10: acc += np.array(ctx.host_stamps())
acc /= 100
st = torch.cuda.ExternalStream(ctx.stream)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
ctx.frame_to_map_resident(); ctx.sync()
e0.record(st)
for i in range(50): ctx.frame_to_map_resident()
e1.record(st); ctx.sync()
print("packed %%.1f h2d %%.1f kernels-enq %%.1f d2h-enq %%.1f synced %%.1f done %%.1f us | device chain (resident replays) %%.1f us" %% (tuple(acc[1:7]) + (e0.elapsed_time(e1) / 50 * 1e3,)))
''' % ROOT
for name, env in (("graph, plain edges (default)", {}), ("graph + programmatic edges", {"PSLAM_GRAPH_PDL": "1"}), ("plain launches (PDL)", {"PSLAM_GRAPHS": "0"})):
    for rep in range(2):
        out = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True)
        print(f"{name:28s}", out.stdout.strip() or out.stderr[-300:], flush=True)
