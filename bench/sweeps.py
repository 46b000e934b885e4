#!/usr/bin/env python
"""C5 sweeps (BASELINE.json configs[4]): RANSAC hypotheses 256..65536 at m = 1000 matches, and descriptor
database 1e5..1e8 (Q = 1000) for both sweep variants, at 1..8 GPUs (run under torchrun for N > 1).
Prints one JSON object per line; rank 0 only.  Device times are CUDA events on the ctx stream.

  python bench/sweeps.py [--max-db 1e8] [--out profiles/sweeps_r1.jsonl]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-db", type=float, default=1e8)
    ap.add_argument("--out", default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--only-ransac", action="store_true", help="hypothesis sweep only")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    from putslam_b200 import api, synth
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = api.Context(lr)
    if world > 1:
        uid = [api.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rank, world)
    st = torch.cuda.ExternalStream(ctx.stream)
    lines = []

    def emit(d):
        d["n_gpus"] = world
        if rank == 0:
            print(json.dumps(d), flush=True)
            lines.append(d)

    def dev_ms(fn, reps):
        fn(); ctx.sync()
        if world > 1:
            dist.barrier()
        best = []
        for _ in range(reps):
            if world > 1:
                dist.barrier(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); ctx.sync()
            best.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(best))], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- RANSAC hypothesis sweep (single GPU work; replicas on the other ranks) ----------------
    if rank == 0:
        from oracle import oracle as O
        mc = synth.matched_clouds(m=1000, inlier_frac=0.55, seed=3)
        for H in [256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536]:
            ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=1, num_hyp=H)
            t0 = time.perf_counter()
            for i in range(a.reps):
                r = ctx.ransac_estimate(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=i, num_hyp=H)
            e2e = (time.perf_counter() - t0) / a.reps * 1e3
            d = {"sweep": "ransac_hypotheses", "H": H, "matches": 1000, "e2e_ms": e2e, "hyp_per_s": H / (e2e * 1e-3),
                 "inliers": int(r["inliers"].size)}
            if not a.no_cpu and H <= 16384:
                t0 = time.perf_counter()
                o = O.ransac(mc["prev"], mc["cur"], mc["mq"], mc["mt"], seed=a.reps - 1, num_hyp=H)
                d["cpu_port_ms"] = (time.perf_counter() - t0) * 1e3
                d["parity_inliers_equal"] = bool(np.array_equal(o["inliers"], r["inliers"]))
            emit(d)
    if world > 1:
        dist.barrier()

    # ---------------- database size sweep ----------------
    if a.only_ransac:
        a.max_db = 0
    rng = np.random.default_rng(1000 + rank)
    q = np.random.default_rng(5).integers(0, 256, (1000, 32), dtype=np.uint8)
    for n_db in [1e5, 1e6, 1e7, 1e8]:
        if n_db > a.max_db:
            continue
        n_db = int(n_db)
        n_local = n_db // world
        ctx.lc_clear()
        ctx.lc_reserve(n_local, n_local // 1000 + 1)
        step = 1000 * 1000
        for s in range(0, n_local, step):
            e = min(n_local, s + step)
            blk = rng.integers(0, 256, (e - s, 32), dtype=np.uint8)
            off = np.arange(0, e - s + 1, 1000, dtype=np.int64)
            if off[-1] != e - s:
                off = np.append(off, e - s)
            ctx.lc_append(blk, off)
        ctx.lc_set_id_base(rank * (n_local // 1000)); ctx.lc_set_desc_base(rank * n_local)
        sharded = world > 1
        # upload the query through the host-pointer entry points once, then time the resident kernels
        if sharded:
            ctx.lc_query_sharded(q, root=-1, tau=64, k=16)
            ms1 = dev_ms(lambda: ctx.lc_query_sharded_resident(64, 16), a.reps)
            ctx.lc_knn2(q, sharded=True, root=-1)
            ms2 = dev_ms(lambda: ctx.lc_knn2_resident(True), a.reps)
        else:
            ctx.lc_query(q, tau=64, k=16)
            ms1 = dev_ms(lambda: ctx.lc_query_resident(64, 16), a.reps)
            ctx.lc_knn2(q)
            ms2 = dev_ms(lambda: ctx.lc_knn2_resident(False), a.reps)
        cm = 1000.0 * n_local * world
        emit({"sweep": "db_size", "n_db": n_db, "queries": 1000, "resident_bytes_per_gpu": n_local * 32,
              "v1_mutual_topk_ms": ms1, "v1_gcmps": cm / (ms1 * 1e-3) / 1e9,
              "v2_knn2_ms": ms2, "v2_gcmps": cm / (ms2 * 1e-3) / 1e9,
              "hbm_gbs_v2": 32.0 * n_local / (ms2 * 1e-3) / 1e9})
    if rank == 0 and not a.no_cpu:
        from oracle import oracle as O
        dbs = np.random.default_rng(2).integers(0, 256, (200000, 32), dtype=np.uint8)
        t0 = time.perf_counter(); O.knn2(q, dbs); dt = time.perf_counter() - t0
        emit({"sweep": "db_size_cpu_port", "n_db": 200000, "queries": 1000, "cores": 1, "v2_gcmps": 1000 * 200000 / dt / 1e9,
              "note": "oracle knn2, single thread (bounded sample)"})
    if a.out and rank == 0:
        with open(a.out, "a") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")
    if world > 1:
        dist.barrier(); ctx.comm_destroy(); dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
