"""CPU test (-m "not gpu") of the tensor-core sweep's operand encoding: the product's own row expansion (expand_row /
decode_row in putslam_b200/csrc/lc_tc.cuh, host-callable) is run on the CPU by tests/tc_encode_host.cu and the int8 dot
products are evaluated straight from the canonical K-major tile layout the tcgen05 descriptors describe.  Asserted there:
every accumulator equals 512 (128 - Ham) + (255 - t) + (255 - q), its fields decode, padding rows (beyond the keyframe /
beyond nq) can never be a maximum, the row maximum breaks ties towards the lowest index, and decode_row inverts the
re-encoding of the resident map (ham256_encode).  The kernel itself is checked against the oracle in the GPU suite."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not present")
def test_tensor_core_operand_encoding_on_the_host():
    out_dir = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "tc_encode_host")
    cc = subprocess.run([NVCC, "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe,
                         os.path.join(ROOT, "tests", "tc_encode_host.cu")], capture_output=True, text=True, timeout=600)
    assert cc.returncode == 0, cc.stderr[-2000:]
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0 and run.stdout.startswith("ok"), run.stdout + run.stderr
    lo_valid, hi_padding = (int(x) for x in run.stdout.split()[1:3])
    assert hi_padding < 512 * (128 - 256) <= lo_valid          # below the worst possible real pair (distance 256)
