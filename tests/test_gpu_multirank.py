"""GPU test (-m gpu) of the keyframe-sharded sweep over NCCL inside the library: one process per GPU,
communicator built from a ncclUniqueId exchanged through a file (no torch.distributed involved), query
broadcast from rank 0, all-gather of local top-k, merge.  Skipped on a single-GPU box."""
import os
import sys
import time

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, tmp):
    sys.path.insert(0, ROOT)
    from putslam_b200 import api, host, synth
    ctx = api.Context(rank)
    db = synth.keyframe_db(n_kf=203, per_kf=400, n_query=700, n_planted=9, shared=200, seed=31, ragged=True)
    b = host.shard_keyframes(203, world)
    lo, hi = b[rank], b[rank + 1]
    off = db["kf_off"][lo:hi + 1]
    ctx.lc_append(db["db"][off[0]:off[-1]], off - off[0])
    ctx.lc_set_id_base(lo)
    uid_path = os.path.join(tmp, "uid.bin")
    if rank == 0:
        uid = api.comm_unique_id()
        with open(uid_path + ".tmp", "wb") as f:
            f.write(uid)
        os.rename(uid_path + ".tmp", uid_path)
    else:
        t0 = time.time()
        while not os.path.exists(uid_path):
            assert time.time() - t0 < 60
            time.sleep(0.01)
        uid = open(uid_path, "rb").read()
    ctx.comm_init(uid, rank, world)
    res = []
    for k in (16, 5):
        ids, sc = ctx.lc_query_sharded(db["query"] if rank == 0 else None, root=0, tau=64, k=k, nq=700)
        res.append((ids.tolist(), sc.tolist()))
    ids, sc = ctx.lc_query_sharded(db["query"], root=-1, tau=40, k=16)      # every rank supplies the query
    res.append((ids.tolist(), sc.tolist()))
    ctx.lc_set_desc_base(int(off[0]))                                       # V2: global descriptor indices
    idx, dist = ctx.lc_knn2(db["query"] if rank == 0 else None, sharded=True, root=0, nq=700)
    np.save(os.path.join(tmp, f"knn{rank}.npy"), np.concatenate([idx, dist.astype(np.int64)], 1))
    np.save(os.path.join(tmp, f"res{rank}.npy"), np.array(res, dtype=object), allow_pickle=True)
    ctx.comm_destroy()
    ctx.close()


def test_sharded_sweep_nccl(tmp_path, O):
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    from putslam_b200 import synth
    mp.spawn(_worker, args=(world, str(tmp_path)), nprocs=world, join=True)
    db = synth.keyframe_db(n_kf=203, per_kf=400, n_query=700, n_planted=9, shared=200, seed=31, ragged=True)
    s64 = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64, threads=4)
    s40 = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=40, threads=4)
    exp = [O.topk(s64, 16), O.topk(s64, 5), O.topk(s40, 16)]
    for r in range(world):
        res = np.load(tmp_path / f"res{r}.npy", allow_pickle=True)
        for (ids, sc), (eid, esc) in zip(res, exp):
            assert list(ids) == eid.tolist() and list(sc) == esc.tolist(), f"rank {r}"
        knn = np.load(tmp_path / f"knn{r}.npy")
        oi, od = O.knn2(db["query"], db["db"])
        assert np.array_equal(knn[:, :2], oi.astype(np.int64)) and np.array_equal(knn[:, 2:], od.astype(np.int64))
