"""GPU parity tests (-m gpu) of the resident feature map: pslam_map_* + pslam_frame_to_resident_map.
The fused call must equal, bit for bit, the two-step path it replaces (pslam_map_prepare then
pslam_frame_to_map_features on the kept features) and the oracle's composition of the same reference functions
(PUTSLAM::getAndFilterFeaturesFromMap, src/PUTSLAM/PUTSLAM.cpp:624-674, then Matcher::matchXYZ, matcher.cpp:606-798)."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def global_map(seed, M=5000, N=1000, extra=1500, n_reobs=None):
    """synth.map_frame's camera-frame map moved to a global frame, plus features the filters must drop
    (behind the view-angle limit, beyond 5 m)."""
    from putslam_b200 import synth
    rng = np.random.default_rng(1000 + seed)
    mf = synth.map_frame(M=M, N=N, seed=seed, n_reobs=n_reobs or (7 * N) // 10)
    pose = np.eye(4)
    pose[:3, :3] = synth.rot_from_rotvec(rng.normal(0, 0.4, 3)); pose[:3, 3] = rng.uniform(-2, 2, 3)
    far = np.stack([rng.uniform(-2, 2, extra), rng.uniform(-1.5, 1.5, extra), rng.uniform(0.5, 8.0, extra)], 1)
    local = np.concatenate([mf["map_xyz"], far])
    order = rng.permutation(M + extra)                       # interleave droppable and matchable features
    local = local[order]
    desc = np.concatenate([mf["map_desc"], rng.integers(0, 256, (extra, 32), dtype=np.uint8)])[order]
    octv = np.concatenate([mf["map_octave"], rng.integers(0, 8, extra).astype(np.int32)])[order]
    det = np.concatenate([mf["map_detdist"], np.linalg.norm(far, axis=1) * rng.uniform(0.75, 1.25, extra)])[order]
    glob = local @ pose[:3, :3].T + pose[:3, 3]
    # optical axis of the view that described each feature: near the current axis for most, far off for some
    axes = np.stack([synth.rot_from_rotvec(rng.normal(0, 0.25, 3))[:, 2] for _ in range(M + extra)]) @ pose[:3, :3].T
    axes = axes.astype(np.float32)
    return dict(mf=mf, pose=pose, glob=glob, desc=np.ascontiguousarray(desc), oct=octv, det=det, axes=axes, local=local)


def prep_params():
    from putslam_b200 import api, synth
    return api.MapPrepareParams(synth.FX, synth.FY, synth.CX, synth.CY, 640, 480, 0.6, 5.0)


def two_step(ctx, g, glob, seed, num_hyp, mode=0):
    mf = g["mf"]
    kept, xl, uv, _ = ctx.map_prepare(glob, g["axes"], g["pose"], prep_params())
    r = ctx.frame_to_map_features(xl, g["desc"][kept], g["oct"][kept], g["det"][kept], mf["cur_xyz"], mf["cur_desc"],
                                  mf["cur_octave"], mf["cur_detdist"], 0.12, 0.55, mode, seed=seed, num_hyp=num_hyp)
    return kept, xl, uv, r


def assert_same(a, b):
    for k in ("mq", "mt", "md", "inliers"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(bits(a["T"]), bits(b["T"]))
    assert a["hyp_used"] == b["hyp_used"] and a["n_filtered"] == b["n_filtered"]
    assert a["best_ratio"] == b["best_ratio"] and a["inlier_ratio"] == b["inlier_ratio"]


@pytest.mark.parametrize("mode", [0, 1])
def test_resident_map_equals_two_step_and_oracle(O, mode):
    from putslam_b200 import api, host, synth
    ctx = api.Context(0)
    try:
        for seed in range(2):
            g = global_map(seed)
            mf = g["mf"]
            ctx.map_truncate(0)
            ctx.map_write(0, g["glob"], g["desc"], g["oct"], g["det"], g["axes"])
            assert ctx.map_size() == g["glob"].shape[0]
            for num_hyp in (0, 4096):
                kept, xl, uv, ref = two_step(ctx, g, g["glob"], 11 + seed, num_hyp, mode)
                out = ctx.frame_to_resident_map(g["pose"], prep_params(), mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"],
                                                mf["cur_detdist"], 0.12, 0.55, mode, seed=11 + seed, num_hyp=num_hyp,
                                                want_local=True)
                assert np.array_equal(out["kept"], kept) and 3000 < kept.size < g["glob"].shape[0]
                assert np.array_equal(bits(out["xyz_local"]), bits(xl)) and np.array_equal(bits(out["uv"]), bits(uv))
                assert_same(out, ref)
                assert out["inliers"].size > 300
                # without the optional camera-frame outputs the answer is the same
                lean = ctx.frame_to_resident_map(g["pose"], prep_params(), mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"],
                                                 mf["cur_detdist"], 0.12, 0.55, mode, seed=11 + seed, num_hyp=num_hyp)
                assert_same(lean, ref) and np.array_equal(lean["kept"], kept)
            # oracle composition of the reference functions
            ek, exl, _, _ = O.map_prepare(g["glob"], g["axes"], g["pose"], synth.FX, synth.FY, synth.CX, synth.CY, 640, 480,
                                          0.6, 5.0)
            assert np.array_equal(ek, kept)
            ml = host.map_levels(exl, g["oct"][ek], g["det"][ek])
            cl = host.current_levels(mf["cur_xyz"], mf["cur_octave"], mf["cur_detdist"])
            q, t, d, _ = O.guided_match(exl, g["desc"][ek], ml, mf["cur_xyz"], mf["cur_desc"], cl, 0.12, 0.55, mode)
            oref = O.ransac(exl.astype(np.float32), mf["cur_xyz"], q, t, seed=11 + seed, num_hyp=4096)
            assert np.array_equal(out["mq"], q) and np.array_equal(out["mt"], t) and np.array_equal(out["md"], d)
            assert np.array_equal(out["inliers"], oref["inliers"])
            assert np.abs(out["T"].astype(np.float64) - oref["T"]).max() <= 1e-5
            # the recovered motion is the planted pose error, and matched slots are real map slots
            assert np.abs(out["T"] - mf["T_gt"]).max() < 0.01
            assert (kept[out["mq"]] < g["glob"].shape[0]).all()
    finally:
        ctx.close()


def test_resident_map_incremental_updates(O):
    """Appending, overwriting a range (positions only, as after a pose-graph update), growing past the reserved
    capacity and truncating leave the map identical to one written in a single call."""
    from putslam_b200 import api
    ctx = api.Context(0)
    try:
        g = global_map(5, M=3000, N=800, extra=700)
        mf = g["mf"]
        n = g["glob"].shape[0]
        ctx.map_reserve(1000)                                     # smaller than the map: must grow and keep contents
        for a, b in [(0, 900), (900, 2500), (2500, n)]:
            ctx.map_write(a, g["glob"][a:b], g["desc"][a:b], g["oct"][a:b], g["det"][a:b], g["axes"][a:b])
        assert ctx.map_size() == n

        def run(glob, seed=3):
            out = ctx.frame_to_resident_map(g["pose"], prep_params(), mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"],
                                            mf["cur_detdist"], seed=seed, num_hyp=1024)
            kept, _, _, ref = two_step(ctx, g, glob, seed, 1024)
            assert np.array_equal(out["kept"], kept)
            assert_same(out, ref)
            return out

        base = run(g["glob"])
        # move a block of features (xyz only): the other attributes keep their values
        moved = g["glob"].copy()
        moved[1000:1800] += np.array([0.02, -0.01, 0.015])
        ctx.map_write(1000, xyz=moved[1000:1800])
        upd = run(moved)
        assert not np.array_equal(upd["md"], base["md"]) or not np.array_equal(bits(upd["T"]), bits(base["T"]))
        # new descriptors for a range
        nd = g["desc"].copy(); nd[200:260] ^= 0xFF
        ctx.map_write(200, desc=nd[200:260])
        g2 = dict(g); g2["desc"] = nd
        out = ctx.frame_to_resident_map(g["pose"], prep_params(), mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"],
                                        mf["cur_detdist"], seed=4, num_hyp=1024)
        kept, _, _, ref = two_step(ctx, g2, moved, 4, 1024)
        assert np.array_equal(out["kept"], kept)
        assert_same(out, ref)
        # truncate = the first slots only
        ctx.map_truncate(2000)
        assert ctx.map_size() == 2000
        g3 = {k: (v[:2000] if k in ("desc", "oct", "det", "axes") else v) for k, v in g2.items()}
        out = ctx.frame_to_resident_map(g["pose"], prep_params(), mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"],
                                        mf["cur_detdist"], seed=5, num_hyp=1024)
        kept, _, _, ref = two_step(ctx, g3, moved[:2000], 5, 1024)
        assert np.array_equal(out["kept"], kept) and kept.max() < 2000
        assert_same(out, ref)
    finally:
        ctx.close()


def test_resident_map_edge_cases():
    from putslam_b200 import api
    ctx = api.Context(0)
    try:
        g = global_map(7, M=600, N=200, extra=100)
        mf = g["mf"]
        cur = (mf["cur_xyz"], mf["cur_desc"], mf["cur_octave"], mf["cur_detdist"])
        # empty map: nothing kept, matchXYZ's "no matches" value
        out = ctx.frame_to_resident_map(g["pose"], prep_params(), *cur, seed=1, num_hyp=256)
        assert out["kept"].size == 0 and out["mq"].size == 0 and out["inlier_ratio"] == -1.0
        assert np.array_equal(out["T"], np.eye(4, dtype=np.float32))
        ctx.map_write(0, g["glob"], g["desc"], g["oct"], g["det"], g["axes"])
        # every feature behind the camera's view-angle limit: kept is empty again
        flipped = g["pose"].copy(); flipped[:3, :3] = flipped[:3, :3] @ np.diag([1.0, -1.0, -1.0])
        out = ctx.frame_to_resident_map(flipped, prep_params(), *cur, seed=1, num_hyp=256)
        assert out["kept"].size == 0 and out["mq"].size == 0 and out["inlier_ratio"] == -1.0
        # no current keypoints
        out = ctx.frame_to_resident_map(g["pose"], prep_params(), np.zeros((0, 3), np.float32), np.zeros((0, 32), np.uint8),
                                        np.zeros(0, np.int32), np.zeros(0), seed=1, num_hyp=256)
        assert out["kept"].size > 100 and out["mq"].size == 0 and out["inlier_ratio"] == -1.0
        # argument errors are reported, not masked
        with pytest.raises(api.PslamError):
            ctx.map_write(ctx.map_size() + 5, g["glob"][:3], g["desc"][:3], g["oct"][:3], g["det"][:3], g["axes"][:3])
        with pytest.raises(api.PslamError):
            ctx.map_write(ctx.map_size(), xyz=g["glob"][:3])          # extending needs every attribute
        with pytest.raises(api.PslamError):
            ctx.map_truncate(ctx.map_size() + 1)
    finally:
        ctx.close()


@pytest.mark.parametrize("M", [1, 255, 257, 40000, 150001])
def test_map_prepare_grid_compaction(ctx, O, M):
    """The filter's ordered compaction runs over the whole grid (per-CTA counts + look-back): sizes around the CTA
    size, beyond one CTA per SM (ranges of several passes), repeated launches (epoch-stamped counts, never reset)."""
    from putslam_b200 import synth
    rng = np.random.default_rng(M)
    pose = np.eye(4); pose[:3, :3] = synth.rot_from_rotvec([0.2, -0.4, 0.1]); pose[:3, 3] = [0.3, 0.1, -0.7]
    local = np.stack([rng.uniform(-3, 3, M), rng.uniform(-2, 2, M), rng.uniform(0.3, 7.0, M)], 1)
    glob = local @ pose[:3, :3].T + pose[:3, 3]
    th = rng.normal(0, 0.45, (M, 2))
    axes_l = np.stack([np.sin(th[:, 0]), np.sin(th[:, 1]) * np.cos(th[:, 0]), np.cos(th[:, 1]) * np.cos(th[:, 0])], 1)
    axes = (axes_l @ pose[:3, :3].T).astype(np.float32)
    ek, exl, euv, eang = O.map_prepare(glob, axes, pose, synth.FX, synth.FY, synth.CX, synth.CY, 640, 480, 0.6, 5.0)
    for _ in range(3):
        kept, xl, uv, ang = ctx.map_prepare(glob, axes, pose, prep_params())
        assert np.array_equal(kept, ek)
        assert np.allclose(xl, exl, rtol=1e-12, atol=1e-13) and np.allclose(uv, euv, rtol=1e-12, atol=1e-10)
        assert np.abs(ang - eang).max(initial=0) < 1e-6
    if M > 1000:
        assert 0.2 * M < kept.size < 0.9 * M
