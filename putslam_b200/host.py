"""Host-side logic of the hot path that deliberately stays on the CPU.

* predicted ORB pyramid levels: computed with the host libm in double exactly as the reference does
  (src/Matcher/matcher.cpp:639-651, 682-692) so that no device/host ulp difference can move a level;
* keyframe sharding and top-k merge for the multi-GPU loop-closure sweep.
"""
import math

import numpy as np

SCALE_FACTOR = 1.2           # include/putslam/Matcher/matcher.h:26
N_LEVELS = 8                 # matcher.h:27
LOG_SCALE_FACTOR = math.log(SCALE_FACTOR)   # matcher.h:28


def predicted_level(octave, det_dist, cur_dist):
    s = math.pow(SCALE_FACTOR, int(octave)) * float(det_dist) / float(cur_dist)
    lvl = int(math.ceil(math.log(s) / LOG_SCALE_FACTOR))
    return min(N_LEVELS - 1, max(0, lvl))


def current_levels(cur_xyz_f32, octaves, det_dists):
    """matcher.cpp:639-651: curDist = Eigen float .norm() widened to double."""
    p = np.asarray(cur_xyz_f32, np.float32).reshape(-1, 3)
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    nrm = np.sqrt((x * x + (y * y + z * z)).astype(np.float32)).astype(np.float64)
    with np.errstate(all="ignore"):
        return np.array([predicted_level(o, d, c) if c > 0 and d > 0 else 0
                         for o, d, c in zip(octaves, det_dists, nrm)], np.int32)


def map_levels(map_xyz_f64, octaves, det_dists):
    """matcher.cpp:682-692: curDist from the double position, ((x*x + y*y) + z*z) then sqrt."""
    p = np.asarray(map_xyz_f64, np.float64).reshape(-1, 3)
    nrm = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2])
    return np.array([predicted_level(o, d, c) if c > 0 and d > 0 else 0
                     for o, d, c in zip(octaves, det_dists, nrm)], np.int32)


def retry_gates(radius, accept_ratio, computation_number):
    """matcher.cpp:617-622: gates widen with the retry number (PUTSLAM.cpp:791-798 calls up to 10 times)."""
    if computation_number > 1:
        radius = radius + 0.02 * (computation_number - 1)
        accept_ratio = max(0.1, accept_ratio - 0.05 * (computation_number - 1))
    return radius, accept_ratio


def shard_keyframes(n_kf, world):
    """Contiguous keyframe ranges per rank: rank r owns [bounds[r], bounds[r+1])."""
    return [(n_kf * r) // world for r in range(world + 1)]


def merge_topk(pairs, k):
    """pairs: iterable of (score, kf_id); -> top-k by score desc, kf id asc (ids < 0 are empty slots)."""
    good = [(int(s), int(i)) for s, i in pairs if int(i) >= 0]
    good.sort(key=lambda p: (-p[0], p[1]))
    good = good[:k]
    ids = np.full(k, -1, np.int32); sc = np.full(k, -1, np.int32)
    for a, (s, i) in enumerate(good):
        ids[a], sc[a] = i, s
    return ids, sc
