"""The KLT device routine (putslam_b200/csrc/klt_point.cuh) run on the CPU with its lanes as a loop (tests/klt_emul.py),
against the cv2 golden vectors and the oracle.  This checks the kernel SOURCE without a GPU; the parity tests of the
kernel as it runs on the B200 are tests/test_gpu_klt.py."""
import numpy as np

from conftest import bits
import klt_emul as E


def test_klt_kernel_source_matches_cv2_golden(golden):
    g = golden["klt_cv2"]
    a, b, pts = g["a"], g["b"], g["pts"]
    cases = {"colour": (a, b, {}), "gray": (a[..., 0], b[..., 0], {}), "mineig": (a, b, {"min_eig_err": True}),
             "initflow": (a, b, {"init": g["init"]})}
    for spec in (True, False):                       # the compile-time 7 x 7 instantiations and the generic one
        for name in g["names"]:
            ia, ib, kw = cases[str(name)]
            nxt, st, err, levels = E.track(ia, ib, pts, specialised=spec, **kw)
            assert levels == 4
            assert np.array_equal(st, g[f"{name}_status"]), name
            ok = st == 1
            assert np.array_equal(bits(nxt[ok]), bits(g[f"{name}_next"][ok])), name
            assert np.array_equal(bits(err[ok]), bits(g[f"{name}_err"][ok])), name


def test_klt_kernel_source_matches_oracle_everywhere():
    """every output of every point, lost ones included (cv2 only defines the tracked ones): hard-edged frames, windows
    other than the reference's, points outside the frame, pyramid cut short by a small frame, clamped criteria"""
    from oracle import klt_oracle as K
    rng = np.random.default_rng(71)
    up = lambda m: np.repeat(np.repeat(m, 2, 0), 2, 1)
    for win, lev, cn, shape in ((9, 2, 1, (40, 52)), (13, 1, 3, (40, 52)), (7, 3, 3, (38, 50)), (4, 2, 3, (30, 30)), (7, 5, 1, (9, 20))):
        a = up(rng.integers(0, 2, shape + ((cn,) if cn == 3 else ()), dtype=np.uint8) * 255)
        b = np.roll(a, 1, 1); b[::3] = np.roll(b[::3], 1, 0)
        b = ((a.astype(np.int32) + b) // 2).astype(np.uint8)
        H, W = a.shape[:2]
        pts = np.stack([rng.uniform(-3, W + 3, 36), rng.uniform(-3, H + 3, 36)], 1).astype(np.float32)
        for kw in ({}, {"min_eig_err": True, "max_iter": 200, "eps": 0.0}, {"init": pts + np.float32(0.7)}):
            o_n, o_s, o_e = K.lk_pyr(a, b, pts, win=win, max_level=lev, **kw)
            e_n, e_s, e_e, _ = E.track(a, b, pts, win=win, max_level=lev, **kw)
            assert np.array_equal(e_s, o_s), (win, cn, kw.keys())
            assert np.array_equal(bits(e_n), bits(o_n)), (win, cn, kw.keys())
            assert np.array_equal(bits(e_e), bits(o_e)), (win, cn, kw.keys())
            assert 0 < o_s.sum()


def test_klt_pyrdown_source_matches_oracle():
    from oracle import klt_oracle as K
    rng = np.random.default_rng(72)
    for shape in ((97, 131, 3), (5, 7), (1, 9, 3), (2, 2), (33, 1)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        ref = K.pyr_down(img if img.ndim == 3 else img[..., None])
        assert np.array_equal(E.pyr_down(img).reshape(ref.shape), ref), shape


def test_klt_prune_source_matches_oracle():
    from oracle import klt_oracle as K
    rng = np.random.default_rng(73)
    pts = np.array([[10, 10], [10.5, 10], [50, 50], [50, 50.4], [90, 90], [10.2, 10.1]], np.float32)
    err = np.array([1.0, 2.0, 3.0, 3.0, 9.0, 0.5], np.float32)
    status = np.array([1, 1, 1, 1, 1, 0], np.uint8)
    assert E.prune(pts, err, status, 5.0, 1.0).tolist() == [2]
    for n, d in ((300, 3.0), (500, 0.0), (200, 1e9), (64, 5.0)):
        xy = rng.uniform(0, 60, (n, 2)).astype(np.float32)
        xy[1] = xy[0]; xy[3] = xy[2] + np.float32(3.0) * np.array([0.6, 0.8], np.float32)   # coincident / on the threshold
        e = rng.choice(np.arange(0, 40, dtype=np.float32), n)                               # many ties
        e[5] = np.nan
        st = (rng.uniform(size=n) < 0.9).astype(np.uint8)
        assert np.array_equal(E.prune(xy, e, st, 25.0, d), K.perform_tracking(e, st, xy, 25.0, d)), (n, d)


def test_klt_kernel_source_fuzz_against_live_cv2():
    """250 random configurations -- frame sizes 8 .. 260 (pyramids cut short), windows 3 .. 21, 0 .. 5 levels, gray and
    colour, noise / blurred / hard-edged / striped frames, both flags, thresholds, clamped criteria, points outside the
    frame -- every tracked point bit-identical to cv2 and every status equal"""
    import cv2
    rng = np.random.default_rng(2026)
    checked = 0
    for it in range(250):
        H = int(rng.integers(8, 200)); W = int(rng.integers(8, 260)); cn = int(rng.choice([1, 3]))
        win = int(rng.integers(3, 22)); lev = int(rng.integers(0, 6)); kind = int(rng.integers(0, 4))
        shape = (H, W, 3) if cn == 3 else (H, W)
        if kind == 0:
            a = rng.integers(0, 256, shape, dtype=np.uint8)
        elif kind == 1:
            a = cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (0, 0), float(rng.uniform(0.6, 3)))
        elif kind == 2:
            small = rng.integers(0, 2, ((H + 3) // 4, (W + 3) // 4) + ((3,) if cn == 3 else ()), dtype=np.uint8) * 255
            a = np.ascontiguousarray(np.repeat(np.repeat(small, 4, 0), 4, 1)[:H, :W])
        else:
            a = np.full(shape, int(rng.integers(0, 256)), np.uint8); a[::5] = 255 - a[::5]
        M = np.float32([[1 + rng.normal(0, 0.01), rng.normal(0, 0.01), rng.normal(0, 2)],
                        [rng.normal(0, 0.01), 1 + rng.normal(0, 0.01), rng.normal(0, 2)]])
        b = cv2.warpAffine(a, M, (W, H), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
        if rng.uniform() < 0.5:
            b = np.clip(b.astype(np.int32) + rng.integers(-4, 5, b.shape), 0, 255).astype(np.uint8)
        pts = np.stack([rng.uniform(-5, W + 5, 40), rng.uniform(-5, H + 5, 40)], 1).astype(np.float32)
        flags = int(rng.integers(0, 4)); thr = float(rng.choice([0.0, 1e-4, 1e-3, 0.05]))
        max_iter = int(rng.choice([0, 1, 5, 30, 100, 150])); eps = float(rng.choice([0.0, 0.01, 0.3, 20.0]))
        init = (pts + rng.normal(0, 1.5, pts.shape)).astype(np.float32) if flags & 1 else None
        cvf = (cv2.OPTFLOW_USE_INITIAL_FLOW if flags & 1 else 0) | (cv2.OPTFLOW_LK_GET_MIN_EIGENVALS if flags & 2 else 0)
        p1, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None if init is None else init.reshape(-1, 1, 2).copy(),
                                              winSize=(win, win), maxLevel=lev, criteria=(3, max_iter, eps), flags=cvf, minEigThreshold=thr)
        nxt, ms, me, _ = E.track(a, b, pts, win=win, max_level=lev, max_iter=max_iter, eps=eps, init=init,
                                 min_eig_err=bool(flags & 2), min_eig_thr=thr)
        st = st.ravel(); ok = st == 1
        where = (it, H, W, cn, win, lev, kind, flags)
        assert np.array_equal(ms, st), where
        assert np.array_equal(bits(nxt[ok]), bits(p1.reshape(-1, 2)[ok])), where
        assert np.array_equal(bits(me[ok]), bits(er.ravel()[ok])), where
        checked += int(ok.sum())
    assert checked > 3000


def test_klt_workspace_fits_the_default_shared_memory_for_every_supported_window():
    """launch_klt_track gives every warp klt_work_bytes(win, cn) of shared memory and fails above 48 KB per CTA: the whole
    supported range (window 3 .. 21, 1 or 3 channels) must fit with at least one warp; the reference's 7 x 7 x 3 with four"""
    L = E.load()
    for cn in (1, 3):
        for win in range(3, 22):
            assert L.klt_emul_work_bytes(win, cn) <= 48 * 1024, (win, cn)
    assert 4 * L.klt_emul_work_bytes(7, 3) <= 48 * 1024


def test_klt_kernel_source_non_finite_points_are_lost_like_in_cv2():
    """NaN / infinite / absurdly large positions and initial guesses: status 0 and err 0, as cv2 reports them (its float ->
    int conversion puts them outside every image; klt_floor does the same on host and device)"""
    import cv2
    rng = np.random.default_rng(1)
    a = cv2.GaussianBlur(rng.integers(0, 256, (60, 80, 3), dtype=np.uint8), (0, 0), 1.5); b = np.roll(a, 1, 1)
    pts = np.array([[np.nan, 10], [20, np.nan], [np.inf, 5], [-np.inf, 5], [1e30, 1e30], [-1e30, 3], [30, 30]], np.float32)
    crit = (3, 30, 0.01)
    _, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=(7, 7), maxLevel=3, criteria=crit)
    _, s, e, _ = E.track(a, b, pts)
    assert np.array_equal(s, st.ravel()) and s.tolist() == [0, 0, 0, 0, 0, 0, 1] and np.array_equal(bits(e), bits(er.ravel()))
    init = pts.copy(); init[6] = [np.nan, np.nan]
    good = np.tile(np.float32([[30, 30]]), (7, 1))
    _, st, _ = cv2.calcOpticalFlowPyrLK(a, b, good.reshape(-1, 1, 2), init.reshape(-1, 1, 2).copy(), winSize=(7, 7), maxLevel=3,
                                        criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
    _, s, _, _ = E.track(a, b, good, init=init)
    assert np.array_equal(s, st.ravel()) and not s.any()
