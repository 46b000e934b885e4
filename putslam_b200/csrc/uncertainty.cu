// uncertainty.cu -- K13: batched covariance of estimated transforms, TransformEst::computeUncertainty /
// computeUncertaintyG2O (reference include/putslam/TransformEst/transformEst.h:29-272; demos/demoKabsch.cpp:1028).
// The reference fills a dense 6n x 6n covariance and a 6n x 6 Jacobian per call (n = 100: 2.9 MB of mostly zeros, two
// dense products).  Only the 3 x 3 diagonal blocks are non-zero, so the product collapses to a sum over points of
// 6 x 6 terms (unc_point.cuh): one CTA per problem, a thread per point, a fixed-order reduction of 57 doubles, and one
// thread inverting the 6 x 6 Hessian.  Double precision throughout; latency-bound (a few us for a batch).
#include "common.cuh"
#include "kernels.h"
#include "unc_point.cuh"

namespace pslam {

constexpr int kUncThreads = 128;

// A, B: concatenated n_i x 3 row-major; CA, CB: n_i x 9 row-major; off[batch + 1]; T: batch x 12 COLUMN-major 3 x 4 (what
// pslam_kabsch_batch returns); U: batch x 36 row-major; ok[batch]: 0 when the Hessian is singular or the set is empty
__global__ void __launch_bounds__(kUncThreads)
uncertainty_batch_kernel(const double* __restrict__ A, const double* __restrict__ B, const double* __restrict__ CA,
                         const double* __restrict__ CB, const int* __restrict__ off, const double* __restrict__ T, int mode,
                         double* __restrict__ U, int* __restrict__ ok) {
    __shared__ UncRot rot;
    __shared__ double tr[3];
    __shared__ double part[kUncThreads / 32][kUncAcc];
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int o = off[p], n = off[p + 1] - o;
    if (n <= 0) {
        if (tid < 36) U[36 * (size_t)p + tid] = 0.0;
        if (tid == 0) ok[p] = 0;
        return;
    }
    if (tid == 0) {
        const double* Tp = T + 12 * (size_t)p;
        double Rm[9], q[4];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Rm[3 * i + j] = Tp[3 * j + i];
            tr[i] = Tp[9 + i];
        }
        unc_quaternion(Rm, q);
        if (mode == kUncEuler) unc_rot_euler(q, rot); else unc_rot_quat(q, rot);
    }
    __syncthreads();
    double acc[kUncAcc];
#pragma unroll
    for (int e = 0; e < kUncAcc; ++e) acc[e] = 0.0;
    for (int i = tid; i < n; i += kUncThreads) {
        const size_t g = (size_t)(o + i);
        unc_point(rot, tr, A + 3 * g, B + 3 * g, CA + 9 * g, CB + 9 * g, acc);
    }
    // fixed-order reduction: butterfly inside the warp, then the four warp sums in order
#pragma unroll
    for (int e = 0; e < kUncAcc; ++e) {
        double v = acc[e];
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
        if (lane == 0) part[warp][e] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double tot[kUncAcc];
        for (int e = 0; e < kUncAcc; ++e) {
            double v = part[0][e];
            for (int w = 1; w < kUncThreads / 32; ++w) v += part[w][e];
            tot[e] = v;
        }
        double Ur[36];
        for (int e = 0; e < 36; ++e) Ur[e] = 0.0;
        const bool good = unc_finish(tot, n, Ur);
        for (int e = 0; e < 36; ++e) U[36 * (size_t)p + e] = Ur[e];
        ok[p] = good ? 1 : 0;
    }
}

cudaError_t launch_uncertainty_batch(const double* d_A, const double* d_B, const double* d_CA, const double* d_CB,
                                     const int* d_off, const double* d_T, int batch, int mode, double* d_U, int* d_ok,
                                     cudaStream_t st, int* launches) {
    if (batch <= 0) return cudaSuccess;
    uncertainty_batch_kernel<<<batch, kUncThreads, 0, st>>>(d_A, d_B, d_CA, d_CB, d_off, d_T, mode, d_U, d_ok);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
