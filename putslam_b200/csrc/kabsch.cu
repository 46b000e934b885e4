// kabsch.cu -- K6: batched double-precision Kabsch, KabschEst::computeTransformation
// (reference src/TransformEst/kabschEst.cpp:24-68): centroids, H = sum (a-cA)(b-cB)^T, Jacobi SVD,
// R = V diag(1,1,sgn det H) U^T, t = cB - R cA, so that B ~= R A + t.  One CTA per point-set pair;
// every sum is a sequential chain owned by one thread (6 centroid chains, then 9 for H), which keeps
// the result independent of the launch shape.
#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

namespace pslam {

// A, B: concatenated n_i x 3 row-major point sets; off[batch+1] point offsets; T: batch x 12, row-major 3x4.
__global__ void __launch_bounds__(32)
kabsch_batch_kernel(const double* __restrict__ A, const double* __restrict__ B, const int* __restrict__ off,
                    double* __restrict__ T) {
    __shared__ double c[6];
    __shared__ double Hs[9];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int o = off[b], n = off[b + 1] - o;
    double* Tb = T + 12 * (size_t)b;
    if (n == 0) {
        if (tid < 12) Tb[tid] = (tid % 5 == 0) ? 1.0 : 0.0;
        return;
    }
    if (tid < 6) {
        const double* P = (tid < 3 ? A : B) + 3 * (size_t)o + (tid % 3);
        double s = 0.0;
        for (int k = 0; k < n; ++k) s = s + P[3 * (size_t)k];
        c[tid] = __ddiv_rn(s, (double)n);
    }
    __syncwarp();
    if (tid < 9) {
        const int i = tid / 3, j = tid % 3;
        const double* Pa = A + 3 * (size_t)o + i;
        const double* Pb = B + 3 * (size_t)o + j;
        const double ca = c[i], cb = c[3 + j];
        double s = 0.0;
        for (int k = 0; k < n; ++k) s = s + (Pa[3 * (size_t)k] - ca) * (Pb[3 * (size_t)k] - cb);
        Hs[tid] = s;
    }
    __syncwarp();
    if (tid == 0) {
        double H[9], Us[9], S[3], Vs[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = Hs[i];
        svd3<double>(H, Us, S, Vs);
        const double det = det3_lu(H);
        const double d = (det != 0.0) ? det : 1.0;
        const double sg = (double)((d > 0.0) - (d < 0.0));
        double R[9];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                // fixed inner size 3 -> c0 + (c1 + c2) (Eigen 3.3 unrolled redux; checked against the reference build)
                const double a = (Vs[3 * i + 0] * 1.0) * Us[3 * j + 0];
                const double b2 = (Vs[3 * i + 1] * 1.0) * Us[3 * j + 1];
                const double c2 = (Vs[3 * i + 2] * sg) * Us[3 * j + 2];
                R[3 * i + j] = a + (b2 + c2);
            }
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const double a = R[3 * i + 0] * (-c[0]);
            const double b2 = R[3 * i + 1] * (-c[1]);
            const double c2 = R[3 * i + 2] * (-c[2]);
            const double s = (a + (b2 + c2)) + c[3 + i];
#pragma unroll
            for (int j = 0; j < 3; ++j) Tb[4 * i + j] = R[3 * i + j];
            Tb[4 * i + 3] = s;
        }
    }
}

cudaError_t launch_kabsch_batch(const double* d_A, const double* d_B, const int* d_off, int batch, double* d_T,
                                cudaStream_t st, int* launches) {
    if (batch <= 0) return cudaSuccess;
    kabsch_batch_kernel<<<batch, 32, 0, st>>>(d_A, d_B, d_off, d_T);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
