"""GPU tests (-m gpu) of the loop-closure sweep: per-keyframe scores and top-k against the oracle,
ragged / empty keyframes, incremental appends, and a size-independent check at the full C4 size."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _fresh(ctx):
    ctx.lc_clear()
    ctx.lc_set_id_base(0)


@pytest.mark.parametrize("n_kf,per_kf,nq,ragged", [(64, 1000, 1000, False), (200, 300, 500, True), (50, 1000, 200, True),
                                                   (300, 128, 1000, False), (7, 4096, 1000, False)])
def test_lc_scores_vs_oracle(ctx, O, n_kf, per_kf, nq, ragged):
    from putslam_b200 import synth
    db = synth.keyframe_db(n_kf=n_kf, per_kf=per_kf, n_query=nq, n_planted=min(5, n_kf), shared=min(nq, per_kf) // 3,
                           seed=n_kf, ragged=ragged)
    _fresh(ctx)
    ctx.lc_append(db["db"], db["kf_off"])
    ref = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64, threads=8)
    eid, esc = O.topk(ref, 16)
    for unit in (1, 2, 3, 0):     # keyframe work units, tile work units, popcount range form, automatic (tensor cores)
        ctx.lc_set_work_unit(unit)
        ids, sc, scores = ctx.lc_query(db["query"], tau=64, k=16, want_scores=True)
        assert np.array_equal(scores, ref), unit
        assert np.array_equal(ids, eid) and np.array_equal(sc, esc), unit
        assert ctx.lc_tensor_status() == (unit == 0, 0), unit      # the automatic form IS the tcgen05 kernel, and it never timed out
    for tau in (0, 30, 256):
        _, _, s2 = ctx.lc_query(db["query"], tau=tau, k=4, want_scores=True)
        assert np.array_equal(s2, O.lc_scores(db["query"], db["db"], db["kf_off"], tau=tau, threads=8))


def test_lc_empty_keyframes_and_incremental_append(ctx, O):
    from putslam_b200 import synth
    db = synth.keyframe_db(n_kf=30, per_kf=200, n_query=300, n_planted=3, shared=100, seed=1)
    off = db["kf_off"].copy()
    # make keyframes 4 and 17 empty by collapsing their ranges
    counts = np.diff(off); counts[4] = 0; counts[17] = 0
    keep = np.concatenate([np.arange(off[k], off[k] + counts[k]) for k in range(30)])
    d2 = db["db"][keep]; off2 = np.concatenate([[0], np.cumsum(counts)])
    _fresh(ctx)
    for lo in range(0, 30, 7):   # append in pieces
        hi = min(30, lo + 7)
        ctx.lc_append(d2, off2[lo:hi + 1])
    assert ctx.lc_size() == (30, int(off2[-1]))
    ids, sc, scores = ctx.lc_query(db["query"], tau=64, k=8, want_scores=True)
    ref = O.lc_scores(db["query"], d2, off2, tau=64)
    assert np.array_equal(scores, ref) and scores[4] == 0 and scores[17] == 0
    assert np.array_equal(ids, O.topk(ref, 8)[0])
    _fresh(ctx)
    ids, sc = ctx.lc_query(db["query"], k=4)   # empty database
    assert (ids == -1).all()


def test_lc_full_size_planted_recall(ctx, O):
    """C4 size (10k keyframes x 1000): the 20 planted keyframes must be the top-20, random ones score ~0; the scores of a
    random sample of 512 keyframes (plus the planted ones) equal the oracle's per-keyframe cross-check count
    (matcher.cpp:835 + matcherOpenCV.cpp:198-206 per keyframe) exactly."""
    from putslam_b200 import synth
    db = synth.keyframe_db(n_kf=10000, per_kf=1000, n_query=1000, n_planted=20, shared=400, seed=7)
    _fresh(ctx)
    ctx.lc_reserve(db["db"].shape[0], 10000)
    ctx.lc_append(db["db"], db["kf_off"])
    ids, sc, scores = ctx.lc_query(db["query"], tau=64, k=20, want_scores=True)
    assert sorted(ids.tolist()) == db["planted"].tolist()
    assert sc.min() > 300 and np.delete(scores, db["planted"]).max() < 5
    sample = np.unique(np.concatenate([np.random.default_rng(3).choice(10000, 512, replace=False), db["planted"]]))
    sub_off = np.concatenate([[0], np.cumsum(db["kf_off"][sample + 1] - db["kf_off"][sample])]).astype(np.int64)
    sub_db = np.concatenate([db["db"][db["kf_off"][k]: db["kf_off"][k + 1]] for k in sample])
    assert np.array_equal(scores[sample], O.lc_scores(db["query"], sub_db, sub_off, tau=64, threads=4))
    # idempotence: the resident replay gives the same answer
    ctx.lc_query_resident(tau=64, k=20); ctx.sync()
    ids2, sc2 = ctx.lc_query(db["query"], tau=64, k=20)
    assert np.array_equal(ids, ids2) and np.array_equal(sc, sc2)
    # V2 at the same size (1e7 descriptors): the returned distances are the true Hamming distances of the returned
    # indices, sorted, and every strong match lies inside a planted keyframe
    idx, dist = ctx.lc_knn2(db["query"])
    assert (idx >= 0).all() and (dist[:, 0] <= dist[:, 1]).all()
    for col in (0, 1):
        true = np.unpackbits(db["db"][idx[:, col]] ^ db["query"], axis=1).sum(1)
        assert np.array_equal(true.astype(np.float32), dist[:, col])
    strong = dist[:, 0] <= 40
    assert strong.sum() >= 300 and np.isin(idx[strong, 0] // 1000, db["planted"]).all()
    assert dist[:, 1].max() <= 100          # second neighbours are planted copies or the best of 1e7 random rows
    if (~strong).any():
        assert dist[~strong, 0].min() > 60  # random 256-bit descriptors: nothing closer by chance in 1e7
    _fresh(ctx)


@pytest.mark.parametrize("n_desc,nq", [(50000, 300), (4097, 1000), (2048, 77), (130, 1000), (1, 5)])
def test_lc_knn2_whole_db_vs_oracle(ctx, O, n_desc, nq):
    """V2: two nearest database descriptors per query descriptor (oracle = cv2-pinned knnMatch restatement)."""
    rng = np.random.default_rng(n_desc + nq)
    db = rng.integers(0, 256, (n_desc, 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    if n_desc > 1000:   # exact duplicates and near-duplicates: ties must resolve to the lowest index
        db[rng.choice(n_desc, 200, replace=False)] = q[rng.integers(0, nq, 200)]
        db[7] = db[4100 % n_desc] = q[0]
    _fresh(ctx)
    ctx.lc_set_desc_base(0)
    ctx.lc_append(db, np.array([0, n_desc], np.int64) if n_desc <= 4096 else
                  np.arange(0, n_desc + 1, 1000).tolist() + ([n_desc] if n_desc % 1000 else []))
    oi, od = O.knn2(q, db)
    for unit in (3, 0):          # popcount kernel, tensor-core kernel
        ctx.lc_set_work_unit(unit)
        idx, dist = ctx.lc_knn2(q)
        assert np.array_equal(idx, oi.astype(np.int64)) and np.array_equal(dist, od.astype(np.float32)), unit
        assert ctx.lc_tensor_status() == (unit == 0, 0), unit
    ctx.lc_set_desc_base(10 ** 10)   # global ids beyond 32 bits
    idx2, _ = ctx.lc_knn2(q)
    assert np.array_equal(idx2[oi >= 0], oi[oi >= 0].astype(np.int64) + 10 ** 10) and (idx2[oi < 0] == -1).all()
    ctx.lc_set_desc_base(0)
    _fresh(ctx)


@pytest.mark.parametrize("nq", [1025, 1500, 2048])
def test_lc_wide_query_sets(ctx, O, nq):
    """More than 1024 query descriptors: 512-thread kernels with the 11/11-bit index split (keyframes <= 2048 rows)."""
    from putslam_b200 import api, synth
    db = synth.keyframe_db(n_kf=40, per_kf=1200, n_query=nq, n_planted=4, shared=500, seed=nq, ragged=True)
    _fresh(ctx)
    ctx.lc_append(db["db"], db["kf_off"])
    ref = O.lc_scores(db["query"], db["db"], db["kf_off"], tau=64, threads=8)
    for unit in (1, 2, 0):
        ctx.lc_set_work_unit(unit)
        ids, sc, scores = ctx.lc_query(db["query"], tau=64, k=8, want_scores=True)
        assert np.array_equal(scores, ref), unit
        assert np.array_equal(ids, O.topk(ref, 8)[0])
        assert ctx.lc_tensor_status() == (False, 0)      # beyond 1024 queries the popcount form runs
    idx, dist = ctx.lc_knn2(db["query"])
    oi, od = O.knn2(db["query"], db["db"])
    assert np.array_equal(idx, oi.astype(np.int64)) and np.array_equal(dist, od.astype(np.float32))
    # a keyframe above 2048 rows cannot be swept with a wide query set: loud error, not a wrong answer
    big = np.random.default_rng(0).integers(0, 256, (3000, 32), dtype=np.uint8)
    ctx.lc_append(big, np.array([0, 3000], np.int64))
    with pytest.raises(api.PslamError) as ei:
        ctx.lc_query(db["query"], tau=64, k=8)
    assert ei.value.code == api.ERR_UNSUPPORTED
    ids, sc = ctx.lc_query(db["query"][:1000], tau=64, k=8)      # still fine with <= 1024 queries
    _fresh(ctx)


@pytest.mark.parametrize("nq", [1, 255, 256, 257, 512, 513, 768, 1023, 1024])
def test_lc_tensor_form_boundaries(ctx, O, nq):
    """Tensor-core form at the edges of its tiling: query counts around the 256-row quarters, keyframes of 1 / 255 / 256 / 257 /
    511 / 512 / 513 / 4096 rows and empty ones (a pair of tiles holds 256 rows), fewer keyframes than CTA groups, duplicated
    rows (ties), tau at both ends -- scores and 2-NN against the oracle, with the popcount form as a second opinion."""
    rng = np.random.default_rng(1000 + nq)
    counts = np.array([1, 255, 256, 257, 0, 511, 512, 513, 4096, 3, 0, 1000], np.int64)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    db = rng.integers(0, 256, (int(off[-1]), 32), dtype=np.uint8)
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    plant = rng.choice(int(off[-1]), min(400, int(off[-1])), replace=False)
    db[plant] = q[rng.integers(0, nq, plant.size)]                 # exact duplicates: distance 0 and many ties
    db[off[3]:off[3] + 2] = q[0]; db[off[8] + 4095] = q[nq - 1]    # first rows of a keyframe, last row of a 4096-row keyframe
    _fresh(ctx)
    ctx.lc_append(db, off)
    for tau in (0, 64, 256):
        ref = O.lc_scores(q, db, off, tau=tau, threads=4)
        for unit in (0, 3):
            ctx.lc_set_work_unit(unit)
            ids, sc, scores = ctx.lc_query(q, tau=tau, k=5, want_scores=True)
            assert np.array_equal(scores, ref), (tau, unit)
            assert np.array_equal(ids, O.topk(ref, 5)[0]) and np.array_equal(sc, O.topk(ref, 5)[1])
            assert ctx.lc_tensor_status() == (unit == 0, 0)
    oi, od = O.knn2(q, db)
    for unit in (0, 3):
        ctx.lc_set_work_unit(unit)
        idx, dist = ctx.lc_knn2(q)
        assert np.array_equal(idx, oi.astype(np.int64)) and np.array_equal(dist, od.astype(np.float32)), unit
    ctx.lc_set_work_unit(0)
    _fresh(ctx)
