#!/usr/bin/env python
"""C2 (BASELINE.json configs[1]): synthetic TUM-shaped sequence, 1000 keypoints per frame, frame-to-frame tracking on
one B200 -- the loop of PUTSLAM::startProcessing (src/PUTSLAM/PUTSLAM.cpp:677-930) reduced to its VO step:
Matcher::match per frame (cross-check matching + back-projection + adaptive RANSAC), pose accumulated as
VOPoseEstimate *= increment with the |t| > 0.1 m guard (PUTSLAM.cpp:735-740).

Reports ms/frame (end to end from host buffers: descriptors, keypoints, depth image in; matches, inliers, pose out),
the absolute trajectory error against the ground-truth helix, and bit-parity of a sample of frames against the oracle.

  python bench/sequence.py [--frames 1000] [--check 20]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(frames=1000, check=20, verbose=False):
    from putslam_b200 import api, synth
    ctx = api.Context(0)
    cam = api.make_camera()
    seq = synth.Sequence(n_frames=1000, n_kp=1000, seed=42)
    f = seq.frame(0)
    prev = ctx.frame_to_frame(None, None, f["desc"], f["uv"], f["depth"], cam=cam)
    prev_desc, prev_frame = f["desc"], f
    pose = np.eye(4)
    gt0 = f["T_wc"]
    est, gts, times, inl, nmatch = [np.eye(4)], [np.eye(4)], [], [], []
    rng = np.random.default_rng(0)
    to_check = set(rng.choice(np.arange(1, frames), size=min(check, frames - 1), replace=False).tolist()) if check else set()
    parity_ok, parity_n = True, 0
    O = None
    for i in range(1, frames):
        f = seq.frame(i)
        t0 = time.perf_counter()
        cur = ctx.frame_to_frame(prev_desc, prev["xyz"], f["desc"], f["uv"], f["depth"], cam=cam, seed=i, num_hyp=0)
        times.append(time.perf_counter() - t0)
        T = cur["T"].astype(np.float64)
        if np.linalg.norm(T[:3, 3]) > 0.1:       # PUTSLAM.cpp:735-737
            T = np.eye(4)
        pose = pose @ T
        est.append(pose.copy()); gts.append(np.linalg.inv(gt0) @ f["T_wc"])
        inl.append(int(cur["inliers"].size)); nmatch.append(int(cur["mq"].size))
        if i in to_check:
            if O is None:
                from oracle import oracle as O_
                O = O_
            oq, ot, od = O.bf_mutual(prev_desc, f["desc"])
            x1, _ = O.backproject(prev_frame["uv"], prev_frame["depth"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
            x2, _ = O.backproject(f["uv"], f["depth"], synth.FX, synth.FY, synth.CX, synth.CY, 5000.0)
            ref = O.ransac(x1, x2, oq, ot, seed=i)
            ok = (np.array_equal(cur["mq"], oq) and np.array_equal(cur["mt"], ot) and np.array_equal(cur["md"], od) and
                  np.array_equal(cur["xyz"].view(np.uint32), x2.view(np.uint32)) and
                  np.array_equal(cur["inliers"], ref["inliers"]) and np.abs(cur["T"] - ref["T"]).max() <= 1e-5)
            parity_ok &= bool(ok); parity_n += 1
        prev, prev_desc, prev_frame = cur, f["desc"], f
    est = np.array(est); gts = np.array(gts)
    ate = float(np.sqrt(np.mean(np.sum((est[:, :3, 3] - gts[:, :3, 3]) ** 2, axis=1))))
    path = float(np.sum(np.linalg.norm(np.diff(gts[:, :3, 3], axis=0), axis=1)))
    t = np.array(times) * 1e3
    ctx.close()
    return {"config": "C2: 1000-keypoint frames along a helix (one turn over 1000 frames), frame-to-frame VO, adaptive RANSAC",
            "frames": frames, "e2e_ms_per_frame_mean": float(t.mean()), "e2e_ms_per_frame_median": float(np.median(t)),
            "e2e_ms_per_frame_p99": float(np.percentile(t, 99)), "fps": float(1e3 / t.mean()),
            "mean_matches": float(np.mean(nmatch)), "mean_inliers": float(np.mean(inl)),
            "ate_rmse_m": ate, "path_length_m": path, "parity_frames_checked": parity_n, "parity_bit_exact": parity_ok}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--check", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(a.frames, a.check)))
