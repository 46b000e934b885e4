"""putslam_b200 -- B200-native (sm_100a) implementation of PUTSLAM's front-end hot path.

csrc/      hand-written CUDA kernels + the C ABI (include/pslam_b200.h) -> libpslam_b200.so
api.py     ctypes binding used by tests / bench
host.py    host-side pieces of the reference interface that stay on the CPU (predicted pyramid levels,
           keyframe sharding, top-k merge)
synth.py   seeded synthetic inputs for the BASELINE configs
"""
from . import api, host, synth  # noqa: F401
from .api import Context, PslamError  # noqa: F401
