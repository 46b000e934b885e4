"""N>1 path on CPU: world_size-2 gloo run of the keyframe-sharded sweep's host logic (shard bounds, id
bases, all-gather of local top-k, merge).  The per-shard scores come from the oracle here (no GPU); on
the GPU the same plumbing is exercised by tests/test_gpu_sweep.py with NCCL inside the library."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from putslam_b200 import host, synth
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    db = synth.keyframe_db(n_kf=37, per_kf=90, n_query=80, n_planted=5, shared=40, seed=5, ragged=True)
    k = 6
    b = host.shard_keyframes(37, world)
    lo, hi = b[rank], b[rank + 1]
    off = db["kf_off"][lo:hi + 1]
    local = O.lc_scores(db["query"], db["db"][off[0]:off[-1]], off - off[0], tau=64)
    ids, sc = O.topk(local, k)
    pairs = torch.tensor(np.stack([sc, np.where(ids >= 0, ids + lo, -1)], 1).astype(np.int32))
    gathered = [torch.zeros_like(pairs) for _ in range(world)]
    dist.all_gather(gathered, pairs)
    allp = torch.cat(gathered).numpy()
    gid, gsc = host.merge_topk([(s, i) for s, i in allp], k)
    full = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64)
    eid, esc = O.topk(full, k)
    ok = np.array_equal(gid, eid) and np.array_equal(gsc, esc)
    open(os.path.join(out_dir, f"rank{rank}.txt"), "w").write("ok" if ok else f"bad {gid} {eid}")
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_topk_world2(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert open(tmp_path / f"rank{r}.txt").read() == "ok"
