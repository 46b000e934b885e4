"""times the sweep (resident query) for the three work-unit forms at several map sizes: python tools/sweep_modes.py [n_kf ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from putslam_b200 import api, synth
sizes = [int(a) for a in sys.argv[1:]] or [10000, 1250]
for n_kf in sizes:
    db = synth.keyframe_db(n_kf=n_kf, per_kf=1000, n_query=1000, n_planted=min(20, n_kf // 2), shared=400, seed=7)
    ctx = api.Context(0)
    ctx.lc_append(db["db"], db["kf_off"])
    ref = None
    for mode in (0, 1, 2):
        ctx.lc_set_work_unit(mode)
        ids, sc = ctx.lc_query(db["query"], tau=64, k=16)
        if ref is None: ref = (ids, sc)
        same = bool(np.array_equal(ids, ref[0]) and np.array_equal(sc, ref[1]))
        ms = []
        for _ in range(6):
            ctx.lc_query_resident(64, 16); ctx.sync(); ms.append(ctx.lc_last_sweep_ms())
        t0 = time.perf_counter()
        for _ in range(10):
            ctx.lc_query_resident(64, 16)
        ctx.sync()
        wall = (time.perf_counter() - t0) / 10 * 1e3
        print(f"n_kf={n_kf} mode={mode} sweep_ms={np.median(ms):.3f} step_ms={wall:.3f} Gcmp/s={1000.0 * n_kf * 1000 / (wall * 1e-3) / 1e9:.1f} same={same}", flush=True)
    ctx.close()
