// unc_emul.cpp -- TEST INFRASTRUCTURE ONLY (built into tests/_build/, never linked into the product): the device source
// putslam_b200/csrc/unc_point.cuh compiled for the CPU and summed over the points sequentially, so that the kernel's
// arithmetic can be checked against the reference-derived golden vectors without a GPU.
//   g++ -O2 -ffp-contract=off -fPIC -shared -o tests/_build/libunc_emul.so tests/unc_emul.cpp
#include "../putslam_b200/csrc/unc_point.cuh"

using namespace pslam;

// A, B n x 3 row-major; CA, CB n x 9 row-major; T 12 doubles column-major 3 x 4; U 36 row-major.  Returns 1 when H inverts.
extern "C" int unc_emul(const double* A, const double* B, const double* CA, const double* CB, int n, const double* T, int mode,
                        double* U) {
    double Rm[9], tr[3], q[4];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Rm[3 * i + j] = T[3 * j + i];
        tr[i] = T[9 + i];
    }
    unc_quaternion(Rm, q);
    UncRot rot;
    if (mode == kUncEuler) unc_rot_euler(q, rot); else unc_rot_quat(q, rot);
    double acc[kUncAcc];
    for (int e = 0; e < kUncAcc; ++e) acc[e] = 0.0;
    for (int i = 0; i < n; ++i) unc_point(rot, tr, A + 3 * i, B + 3 * i, CA + 9 * i, CB + 9 * i, acc);
    return unc_finish(acc, n, U) ? 1 : 0;
}
