"""How much of the bit-exact inlier-set claim rests on what cannot be pinned here (Eigen's internal arithmetic order)?

The reference's RANSAC.cpp / matcher.cpp are compiled against the Eigen stand-in (oracle/ref_shim) in seven variants
(oracle/Makefile): the default model (Eigen 3.3) and one perturbed choice each --
    _seq       product coefficients summed left to right (Eigen 3.2) instead of 3.3's (row . col).sum() halving: R*x + t
    _scalelhs  umeyama: (1/n * dst_demean) * src_demean^T instead of 1/n * (dst_demean * src_demean^T)
    _jac32     JacobiSVD with Eigen 3.2's threshold / no input scaling
    _sweep     JacobiSVD visiting the index pairs in the opposite order
    _f64       umeyama evaluated in float64 and rounded once
    _allseq    EVERY fixed-size reduction left to right, norm() included -- no Eigen version does that; it shows how sharp
               the level gate of matchXYZ is: a key point matched in the frame that detected it has detDist / |x| = 1 +- 1 ulp
               (matcher.cpp:53-55 sums left to right in float, Vector3f::norm() halves), and ceil(log(1.2^k * that) / log 1.2)
               is k or k + 1 on that ulp
For synthetic frames of the configurations C1-C3 this script runs every variant on identical inputs and the
identical sample stream and reports the fraction of frames whose FINAL INLIER SET, number of hypotheses drawn (hyp_used) or
pose (beyond 1e-5 m / 1e-5 rad) differ from the default build, and how far the poses move.  Needs oracle/_ref (built where
/root/reference exists); writes profiles/umeyama_sensitivity.json.

    python tools/umeyama_sensitivity.py [--c1 200] [--c2 100] [--c3 30]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O, ref_build as R   # noqa: E402
from putslam_b200 import synth                    # noqa: E402

CAM = (synth.FX, synth.FY, synth.CX, synth.CY)


def rot_angle(Ta, Tb):
    # angle of Ra Rb^T through the chord ||Ra - Rb||_F = 2 sqrt(2) sin(angle / 2): well conditioned near zero, and the
    # float32 rotations' own departure from orthogonality (1e-7) does not show up as 1e-3 rad the way arccos(trace) does
    d = np.linalg.norm(Ta[:3, :3].astype(np.float64) - Tb[:3, :3].astype(np.float64))
    return float(2 * np.arcsin(min(1.0, d / (2 * np.sqrt(2)))))


def frames_c12(n_kp, count, base_seed):
    for s in range(count):
        fp = synth.frame_pair(n=n_kp, seed=base_seed + s, outlier_frac=(0.25, 0.4, 0.55)[s % 3])
        x1, _ = O.backproject(fp["uv1"], fp["depth1"], *CAM, synth.DEPTH_SCALE)
        x2, _ = O.backproject(fp["uv2"], fp["depth2"], *CAM, synth.DEPTH_SCALE)
        q, t, _ = O.bf_mutual(fp["desc1"], fp["desc2"])
        yield s, x1, x2, q, t


def compare(stats, ref, out):
    stats["frames"] += 1
    stats["inlier_set_differs"] += int(not np.array_equal(ref["inl"], out["inl"]))
    stats["hyp_used_differs"] += int(ref["hyp"] != out["hyp"])
    dt = float(np.abs(ref["T"][:3, 3].astype(np.float64) - out["T"][:3, 3]).max()); dr = rot_angle(ref["T"], out["T"])
    stats["pose_beyond_1e-5"] += int(dt > 1e-5 or dr > 1e-5)
    stats["pose_bits_differ"] += int(not np.array_equal(ref["T"].view(np.uint32), out["T"].view(np.uint32)))
    stats["max_dt_m"] = max(stats["max_dt_m"], dt); stats["max_drot_rad"] = max(stats["max_drot_rad"], dr)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c1", type=int, default=200); ap.add_argument("--c2", type=int, default=100); ap.add_argument("--c3", type=int, default=30)
    a = ap.parse_args()
    variants = [v for v in R.VARIANTS if R.available(v)]
    res = {"variants": {v or "default": R.shim_model(v) for v in variants}, "configs": {}}
    t0 = time.time()
    for name, n_kp, count, base in (("C1_frame_pair_500", 500, a.c1, 0), ("C2_frame_pair_1000", 1000, a.c2, 5000)):
        st = {v: dict(frames=0, inlier_set_differs=0, hyp_used_differs=0, pose_bits_differ=0, max_dt_m=0.0, max_drot_rad=0.0,
                      **{"pose_beyond_1e-5": 0}) for v in variants[1:] + ["oracle.c"]}
        inl_mean = 0
        for s, x1, x2, q, t in frames_c12(n_kp, count, base):
            for ev in (0, 2):
                p = O.default_ransac_params(ev)
                outs = {}
                for v in variants:
                    r = R.ransac(x1, x2, q, t, args=R.from_oracle_params(p), seed=s, variant=v)
                    outs[v] = dict(inl=r["inliers"], hyp=r["hyp_used"], T=r["T"])
                o = O.ransac(x1, x2, q, t, params=p, seed=s)
                outs["oracle.c"] = dict(inl=o["inliers"], hyp=o["hyp_used"], T=o["T"])
                inl_mean += len(o["inliers"])
                for v in st:
                    compare(st[v], outs[""], outs[v])
        res["configs"][name] = dict(frames=2 * count, error_versions=[0, 2], mean_inliers=inl_mean / max(1, 2 * count), vs_default_build=st)
        print(name, json.dumps(st), f"{time.time() - t0:.0f}s", flush=True)
    st = {v: dict(frames=0, inlier_set_differs=0, hyp_used_differs=0, pose_bits_differ=0, max_dt_m=0.0, max_drot_rad=0.0,
                  matches_differ=0, **{"pose_beyond_1e-5": 0}) for v in variants[1:]}
    for s in range(a.c3):
        mf = synth.map_frame(M=5000, N=1000, seed=100 + s)
        outs = {}
        for v in variants:
            r = R.match_xyz(mf["map_xyz"], mf["map_desc"], mf["map_octave"], mf["map_detdist"], mf["cur_xyz"], mf["cur_desc"],
                            mf["cur_octave"], mf["cur_detdist"], seed=s, variant=v)
            outs[v] = dict(inl=r["pairs"], hyp=r["hyp_used"], T=r["T"], nm=r["n_matches"])
        for v in st:
            compare(st[v], outs[""], outs[v])
            st[v]["matches_differ"] += int(outs[v]["nm"] != outs[""]["nm"])
    res["configs"]["C3_frame_to_map_1000x5000"] = dict(frames=a.c3, vs_default_build=st,
                                                      note="whole Matcher::matchXYZ (level prediction uses Vector3f::norm, so the guided match list is included)")
    print("C3", json.dumps(st), f"{time.time() - t0:.0f}s", flush=True)
    res["reading"] = ("inlier_set_differs / hyp_used_differs = frames (out of `frames`) where the variant's final inlier set / number of "
                      "hypotheses differs from the default stand-in model; pose_beyond_1e-5 = frames whose pose moves by more than 1e-5 m or "
                      "1e-5 rad (north_star's tolerance).  oracle.c is the restatement the CUDA path is held to bit for bit.")
    json.dump(res, open(os.path.join(ROOT, "profiles", "umeyama_sensitivity.json"), "w"), indent=1)
    print("wrote profiles/umeyama_sensitivity.json")


if __name__ == "__main__":
    main()
