"""runs pslam_klt_perform_tracking at the reference's configuration (640x480x3 frames, 1000 points, window 7, 3 levels);
prints the bench numbers as JSON, or with --once just a few calls (for ncu launch lists / captures)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from putslam_b200 import api
ctx = api.Context(0)
if "--once" in sys.argv:
    rng = np.random.default_rng(78)
    g = bench.orb_bench_image(rng)
    f0 = np.stack([g, np.roll(g, 3, 1), 255 - np.roll(g, 2, 0)], 2).copy()
    f1 = np.roll(np.roll(f0, 2, 1), 1, 0).copy()
    pts = np.stack([rng.uniform(8, 631, 1000), rng.uniform(8, 471, 1000)], 1).astype(np.float32)
    for _ in range(3):
        r = ctx.klt_track(f0, f1, pts, min_eig_threshold=0.0, prune=(25.0, 3.0))
    print("ok", int(r["status"].sum()), r["kept"].size)
else:
    print(json.dumps(bench.klt_numbers(ctx, 50, 5)))
