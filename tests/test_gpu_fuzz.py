"""GPU fuzz tests (-m gpu): seeded random shapes and contents around the kernels' tiling boundaries (256-query
tiles, 128-row train tiles, odd counts, heavy ties), all bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_desc(rng, n, ties):
    if ties:   # few distinct values -> many equal distances
        d = np.zeros((n, 32), np.uint8)
        d[:, : rng.integers(1, 4)] = rng.integers(0, rng.integers(2, 9), (n, 1))
        return d
    return rng.integers(0, 256, (n, 32), dtype=np.uint8)


def test_fuzz_bf_mutual_and_knn2(ctx, O):
    rng = np.random.default_rng(2026)
    sizes = [(1, 1), (1, 300), (300, 1), (255, 257), (256, 256), (257, 127), (511, 129), (512, 128), (513, 1000),
             (1023, 17), (1024, 1025), (1025, 33), (2049, 100)]
    sizes += [(int(rng.integers(1, 2200)), int(rng.integers(1, 2200))) for _ in range(12)]
    for k, (nq, nt) in enumerate(sizes):
        q = _rand_desc(rng, nq, ties=(k % 3 == 0)); t = _rand_desc(rng, nt, ties=(k % 3 == 0))
        if k % 3 == 1 and min(nq, nt) > 4:
            t[rng.choice(nt, min(nq, nt) // 2, replace=False)] = q[rng.choice(nq, min(nq, nt) // 2, replace=False)]
        a = ctx.match_bf_mutual(q, t); b = O.bf_mutual(q, t)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (nq, nt)
        idx, dist = ctx.match_knn2(q, t)
        oi, od = O.knn2(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od.astype(np.float32)), (nq, nt)


def test_fuzz_lc_sweep(ctx, O):
    rng = np.random.default_rng(7)
    for k in range(8):
        n_kf = int(rng.integers(1, 60)); nq = int(rng.integers(1, 1025))
        counts = rng.integers(0, 700, n_kf)
        if k == 0:
            counts[:] = 0
        off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        db = _rand_desc(rng, int(off[-1]), ties=(k % 2 == 0)) if off[-1] else np.zeros((0, 32), np.uint8)
        q = _rand_desc(rng, nq, ties=(k % 2 == 0))
        if off[-1] > 10 and k % 2 == 1:
            db[rng.choice(int(off[-1]), min(nq, int(off[-1])) // 3, replace=False)] = q[rng.choice(nq, min(nq, int(off[-1])) // 3, replace=False)]
        ctx.lc_clear(); ctx.lc_set_id_base(0); ctx.lc_set_desc_base(0)
        if off[-1] or n_kf:
            ctx.lc_append(db if off[-1] else np.zeros((1, 32), np.uint8), off)
        ref = O.lc_scores(q, db if off[-1] else np.zeros((1, 32), np.uint8), off, tau=70, threads=4)
        for unit in (1, 2, 3, 0):
            ctx.lc_set_work_unit(unit)
            ids, sc, scores = ctx.lc_query(q, tau=70, k=5, want_scores=True)
            assert np.array_equal(scores, ref), (k, unit)
            assert ctx.lc_tensor_status()[1] == 0
            assert np.array_equal(ids, O.topk(ref, 5)[0]) and np.array_equal(sc, O.topk(ref, 5)[1])
        ctx.lc_set_work_unit(0)
        if off[-1] >= 1:
            oi, od = O.knn2(q, db)
            for unit in (3, 0):
                ctx.lc_set_work_unit(unit)
                idx, dist = ctx.lc_knn2(q)
                assert np.array_equal(idx, oi.astype(np.int64)) and np.array_equal(dist, od.astype(np.float32)), (k, unit)
                assert ctx.lc_tensor_status()[1] == 0
    ctx.lc_clear()


def test_fuzz_ransac(ctx, O):
    from putslam_b200 import api, synth
    rng = np.random.default_rng(99)
    for k in range(12):
        m = int(rng.choice([3, 14, 15, 16, 31, 32, 33, 100, 777, 1024, 1025, 3000]))
        mc = synth.matched_clouds(m=m, inlier_frac=float(rng.uniform(0.1, 0.9)), seed=k)
        ev = int(rng.choice([0, 4, 1, 2]))
        num_hyp = int(rng.choice([0, 1, 7, 300]))
        prev = mc["prev"].copy()
        if k % 4 == 0:
            prev[rng.random(m) < 0.1, 2] = np.nan
        r = ctx.ransac_estimate(prev, mc["cur"], mc["mq"], mc["mt"], params=api.default_ransac_params(ev), seed=k,
                                num_hyp=num_hyp)
        o = O.ransac(prev, mc["cur"], mc["mq"], mc["mt"], params=O.default_ransac_params(ev), seed=k, num_hyp=num_hyp)
        assert r["hyp_used"] == o["hyp_used"] and r["best_ratio"] == o["best_ratio"], (k, m, ev, num_hyp)
        assert np.array_equal(r["inliers"], o["inliers"]), (k, m, ev, num_hyp)
        assert np.allclose(r["T"], o["T"], atol=1e-5, rtol=0, equal_nan=True)


def test_ransac_degenerate_geometry(ctx, O):
    """Collinear / coplanar / coincident samples make the 3-point cross-covariance rank 1 or 0 (SURVEY 8c degenerate
    cases): the per-hypothesis counts must still be bit-identical to the oracle."""
    rng = np.random.default_rng(5)
    m = 300
    t = rng.uniform(0.5, 3.0, m)
    line = np.stack([0.1 * t, 0.2 * t - 0.3, 0.8 + t], 1)                       # all points on one line
    plane = np.stack([rng.uniform(-1, 1, m), rng.uniform(-1, 1, m), np.full(m, 2.0)], 1)
    same = np.tile([[0.3, -0.2, 1.5]], (m, 1))                                   # all points coincide
    grid = np.round(rng.uniform(0.5, 3, (m, 3)) * 4) / 4                         # many exactly repeated coordinates
    for name, pts in (("line", line), ("plane", plane), ("same", same), ("grid", grid)):
        prev = pts.astype(np.float32)
        cur = (pts + np.array([0.01, -0.02, 0.03])).astype(np.float32)
        if name == "plane":
            cur[::3] += rng.normal(0, 0.2, (len(cur[::3]), 3)).astype(np.float32)
        mq = np.arange(m, dtype=np.int32); mt = rng.permutation(m).astype(np.int32)
        cur_p = np.empty_like(cur); cur_p[mt] = cur
        r = ctx.ransac_estimate(prev, cur_p, mq, mt, seed=17, num_hyp=512, want_counts=True)
        o = O.ransac(prev, cur_p, mq, mt, seed=17, num_hyp=512, want_counts=True)
        assert np.array_equal(r["counts"][:512], o["counts"][:512]), name
        assert np.array_equal(r["inliers"], o["inliers"]) and r["best_ratio"] == o["best_ratio"], name
        assert np.allclose(r["T"], o["T"], atol=1e-5, rtol=0, equal_nan=True), name
