// geometry.cuh -- small-matrix device math for stage 3 (RANSAC / Umeyama / Kabsch), sm_100a.
//
// Rigid alignment in the reference is Eigen::umeyama(..., with_scaling=false) in float32
// (reference src/TransformEst/RANSAC.cpp:225-226) and a JacobiSVD-based Kabsch in float64
// (reference src/TransformEst/kabschEst.cpp:24-68).  Both reduce to: 3x3 cross-covariance -> two-sided
// Jacobi SVD -> R = U diag(1,1,s) V^T.  The routines below implement that with every operation a
// single IEEE rounding (file is compiled -fmad=false -prec-div=true -prec-sqrt=true) in the operation
// order documented in DESIGN.md, so that hypothesis models -- and therefore inlier sets -- are
// reproducible bit for bit on any IEEE machine.
#pragma once
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

namespace pslam {

template <typename T> struct Real;
template <> struct Real<float> {
    __device__ static __forceinline__ float sqrt(float x) { return __fsqrt_rn(x); }
    __device__ static __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
    __device__ static __forceinline__ float abs(float x) { return fabsf(x); }
    __device__ static __forceinline__ float eps() { return FLT_EPSILON; }
    __device__ static __forceinline__ float tiny() { return FLT_MIN; }
};
template <> struct Real<double> {
    __device__ static __forceinline__ double sqrt(double x) { return __dsqrt_rn(x); }
    __device__ static __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
    __device__ static __forceinline__ double abs(double x) { return fabs(x); }
    __device__ static __forceinline__ double eps() { return DBL_EPSILON; }
    __device__ static __forceinline__ double tiny() { return DBL_MIN; }
};

// In-plane rotation of rows p,q of a row-major 3x3: x' = c x + s y ; y' = -s x + c y.
template <typename T, int P, int Q>
__device__ __forceinline__ void rot_rows(T (&M)[9], T c, T s) {
    if (c == T(1) && s == T(0)) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T xi = M[3 * P + i], yi = M[3 * Q + i];
        M[3 * P + i] = c * xi + s * yi;
        M[3 * Q + i] = (-s) * xi + c * yi;
    }
}
// Columns p,q rotated by the transpose of (c, s): x' = c x - s y ; y' = s x + c y.
template <typename T, int P, int Q>
__device__ __forceinline__ void rot_cols(T (&M)[9], T c, T s) {
    const T st = -s;
    if (c == T(1) && st == T(0)) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T xi = M[3 * i + P], yi = M[3 * i + Q];
        M[3 * i + P] = c * xi + st * yi;
        M[3 * i + Q] = (-st) * xi + c * yi;
    }
}

// One two-sided Jacobi step on the (P,Q) 2x2 block; returns true if a rotation was applied.
template <typename T, int P, int Q>
__device__ __forceinline__ bool jacobi_pair(T (&W)[9], T (&U)[9], T (&V)[9], T& maxDiag) {
    using R = Real<T>;
    T thr = (T(2) * R::eps()) * maxDiag;
    if (R::tiny() > thr) thr = R::tiny();
    if (!(R::abs(W[3 * P + Q]) > thr || R::abs(W[3 * Q + P]) > thr)) return false;
    T m00 = W[3 * P + P], m01 = W[3 * P + Q], m10 = W[3 * Q + P], m11 = W[3 * Q + Q];
    // first rotation: makes the block symmetric
    T r1c, r1s;
    const T t = m00 + m11, d = m10 - m01;
    if (R::abs(d) < R::tiny()) {
        r1s = T(0); r1c = T(1);
    } else {
        const T u = R::div(t, d);
        const T tmp = R::sqrt(T(1) + u * u);
        r1s = R::div(T(1), tmp);
        r1c = R::div(u, tmp);
    }
    if (!(r1c == T(1) && r1s == T(0))) {
        const T x0 = m00, y0 = m10, x1 = m01, y1 = m11;
        m00 = r1c * x0 + r1s * y0; m10 = (-r1s) * x0 + r1c * y0;
        m01 = r1c * x1 + r1s * y1; m11 = (-r1s) * x1 + r1c * y1;
    }
    // second rotation: diagonalises the symmetric block
    T jrc, jrs;
    const T deno = T(2) * R::abs(m01);
    if (deno < R::tiny()) {
        jrc = T(1); jrs = T(0);
    } else {
        const T tau = R::div(m00 - m11, deno);
        const T w = R::sqrt(tau * tau + T(1));
        T tt;
        if (tau > T(0)) tt = R::div(T(1), tau + w);
        else tt = R::div(T(1), tau - w);
        const T sign_t = tt > T(0) ? T(1) : T(-1);
        const T nn = R::div(T(1), R::sqrt(tt * tt + T(1)));
        jrs = (-sign_t) * R::div(m01, R::abs(m01)) * R::abs(tt) * nn;
        jrc = nn;
    }
    const T jtc = jrc, jts = -jrs;
    const T jlc = r1c * jtc - r1s * jts;
    const T jls = r1c * jts + r1s * jtc;
    rot_rows<T, P, Q>(W, jlc, jls);
    rot_cols<T, P, Q>(U, jlc, -jls);
    rot_cols<T, P, Q>(W, jrc, jrs);
    rot_cols<T, P, Q>(V, jrc, jrs);
    T a = R::abs(W[3 * P + P]);
    const T b = R::abs(W[3 * Q + Q]);
    if (b > a) a = b;
    if (a > maxDiag) maxDiag = a;
    return true;
}

template <typename T>
__device__ __forceinline__ void swap_cols(T (&M)[9], int a, int b) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const T va = (a == 0) ? M[3 * r] : (a == 1 ? M[3 * r + 1] : M[3 * r + 2]);
        const T vb = (b == 0) ? M[3 * r] : (b == 1 ? M[3 * r + 1] : M[3 * r + 2]);
        if (a == 0) M[3 * r] = vb; else if (a == 1) M[3 * r + 1] = vb; else M[3 * r + 2] = vb;
        if (b == 0) M[3 * r] = va; else if (b == 1) M[3 * r + 1] = va; else M[3 * r + 2] = va;
    }
}

// A = U diag(S) V^T, row-major, S sorted descending (first-max on ties), U/V orthogonal.
template <typename T>
__device__ __forceinline__ void svd3(const T (&A)[9], T (&U)[9], T (&S)[3], T (&V)[9]) {
    using R = Real<T>;
    T W[9];
    T scale = T(0);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const T a = R::abs(A[i]);
        if (a > scale) scale = a;
    }
    if (scale == T(0)) scale = T(1);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        W[i] = R::div(A[i], scale);
        U[i] = V[i] = (i % 4 == 0) ? T(1) : T(0);
    }
    T maxDiag = R::abs(W[0]);
    if (R::abs(W[4]) > maxDiag) maxDiag = R::abs(W[4]);
    if (R::abs(W[8]) > maxDiag) maxDiag = R::abs(W[8]);
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
        if (jacobi_pair<T, 1, 0>(W, U, V, maxDiag)) finished = false;
        if (jacobi_pair<T, 2, 0>(W, U, V, maxDiag)) finished = false;
        if (jacobi_pair<T, 2, 1>(W, U, V, maxDiag)) finished = false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const T a = W[4 * i];
        S[i] = R::abs(a);
        if (a < T(0)) {
#pragma unroll
            for (int r = 0; r < 3; ++r) U[3 * r + i] = -U[3 * r + i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) S[i] = S[i] * scale;
    // selection sort, descending, first maximum wins ties
    {
        int pos = 0;
        T mx = S[0];
        if (S[1] > mx) { mx = S[1]; pos = 1; }
        if (S[2] > mx) { mx = S[2]; pos = 2; }
        if (pos == 1) { const T t = S[0]; S[0] = S[1]; S[1] = t; swap_cols(U, 0, 1); swap_cols(V, 0, 1); }
        else if (pos == 2) { const T t = S[0]; S[0] = S[2]; S[2] = t; swap_cols(U, 0, 2); swap_cols(V, 0, 2); }
        if (S[2] > S[1]) { const T t = S[1]; S[1] = S[2]; S[2] = t; swap_cols(U, 1, 2); swap_cols(V, 1, 2); }
    }
}

template <typename T>
__device__ __forceinline__ T det3(const T (&m)[9]) {
    const T a = m[0] * (m[4] * m[8] - m[5] * m[7]);
    const T b = m[1] * (m[3] * m[8] - m[5] * m[6]);
    const T c = m[2] * (m[3] * m[7] - m[4] * m[6]);
    return a - b + c;
}

// MatrixXd::determinant() of a dynamic-size matrix = partialPivLu().determinant() (what kabschEst.cpp:53 evaluates on its
// MatrixXd A): first-max partial pivoting, column divided by the pivot, rank-1 update, sign * ((u00 * u11) * u22).
__device__ __forceinline__ double det3_lu(const double (&m)[9]) {
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = m[i];
    double sign = 1.0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int p = k;
        double best = fabs(a[3 * k + k]);
#pragma unroll
        for (int i = k + 1; i < 3; ++i) if (fabs(a[3 * i + k]) > best) { best = fabs(a[3 * i + k]); p = i; }
        if (p != k) {
#pragma unroll
            for (int j = 0; j < 3; ++j) { const double t = a[3 * k + j]; a[3 * k + j] = a[3 * p + j]; a[3 * p + j] = t; }
            sign = -sign;
        }
        if (a[3 * k + k] == 0.0) continue;
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            a[3 * i + k] = __ddiv_rn(a[3 * i + k], a[3 * k + k]);
#pragma unroll
            for (int j = k + 1; j < 3; ++j) a[3 * i + j] = a[3 * i + j] - a[3 * i + k] * a[3 * k + j];
        }
    }
    return ((sign * a[0]) * a[4]) * a[8];
}

// Rigid transform stored as R (row-major 3x3) and t.
struct Rigid3f {
    float R[9];
    float t[3];
    bool ok;
};

// Umeyama eq. 39-43 from means and the (already 1/n-scaled) cross-covariance sigma = E[(d-md)(s-ms)^T].
__device__ __forceinline__ Rigid3f umeyama_from_sigma(const float (&sig)[9], const float (&sm)[3], const float (&dm)[3]) {
    float U[9], S[3], V[9];
    svd3<float>(sig, U, S, V);
    float sg2 = 1.f;
    if (det3<float>(U) * det3<float>(V) < 0.f) sg2 = -1.f;
    Rigid3f out;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float s = (U[3 * i + 0] * 1.f) * V[3 * j + 0];
            s = s + (U[3 * i + 1] * 1.f) * V[3 * j + 1];
            s = s + (U[3 * i + 2] * sg2) * V[3 * j + 2];
            out.R[3 * i + j] = s;
        }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float s = out.R[3 * i + 0] * sm[0];
        s = s + out.R[3 * i + 1] * sm[1];
        s = s + out.R[3 * i + 2] * sm[2];
        out.t[i] = dm[i] - s;
    }
    out.ok = !isnan(out.R[0]);
    return out;
}

// Three-point model: src = current-frame points, dst = previous-frame points (dst ~= R src + t).
__device__ __forceinline__ Rigid3f umeyama3(const float (&src)[3][3], const float (&dst)[3][3]) {
    const float one_over_n = __fdiv_rn(1.f, 3.f);
    float sm[3], dm[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int k = 0; k < 3; ++k) { a = a + src[k][c]; b = b + dst[k][c]; }
        sm[c] = a * one_over_n;
        dm[c] = b * one_over_n;
    }
    float sig[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) sig[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float sd[3], dd[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) { sd[c] = src[k][c] - sm[c]; dd[c] = dst[k][c] - dm[c]; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) sig[3 * i + j] = sig[3 * i + j] + dd[i] * sd[j];
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) sig[i] = one_over_n * sig[i];
    return umeyama_from_sigma(sig, sm, dm);
}

// (r0*x0 + (r1*x1 + r2*x2)) + t : Eigen 3.3's Matrix3f * Vector3f coefficient is a fixed-size sum of three products,
// i.e. the unrolled binary split c0 + (c1 + c2) (the reference needs Eigen >= 3.3; checked against the reference build)
__device__ __forceinline__ void rigid_apply(const float (&R)[9], const float (&t)[3], float x, float y, float z,
                                            float& ox, float& oy, float& oz) {
    ox = (R[0] * x + (R[1] * y + R[2] * z)) + t[0];
    oy = (R[3] * x + (R[4] * y + R[5] * z)) + t[1];
    oz = (R[6] * x + (R[7] * y + R[8] * z)) + t[2];
}

// squared norm in the x^2 + (y^2 + z^2) association, then IEEE sqrt
__device__ __forceinline__ float norm3(float x, float y, float z) {
    const float yy = y * y, zz = z * z;
    return __fsqrt_rn(x * x + (yy + zz));
}

// General 4x4 inverse by cofactors (reference's reprojection metrics call Matrix4f::inverse(),
// src/TransformEst/RANSAC.cpp:337-338).  m, r row-major.
__device__ __forceinline__ float det3h(const float (&m)[16], int i1, int i2, int i3, int j1, int j2, int j3) {
    return m[4 * i1 + j1] * (m[4 * i2 + j2] * m[4 * i3 + j3] - m[4 * i2 + j3] * m[4 * i3 + j2]);
}
__device__ __forceinline__ void inverse4(const float (&m)[16], float (&r)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4;
            const int j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
            const float c = det3h(m, i1, i2, i3, j1, j2, j3) + det3h(m, i2, i3, i1, j1, j2, j3) +
                            det3h(m, i3, i1, i2, j1, j2, j3);
            r[4 * j + i] = ((i + j) & 1) ? -c : c;
        }
    const float det = (m[0] * r[0] + m[4] * r[1]) + (m[8] * r[2] + m[12] * r[3]);
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __fdiv_rn(r[i], det);
}

// Philox4x32-10 (Salmon et al. 2011), key = 64-bit seed, counter = {block, hypothesis, 0, 0}.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// usedPairs = 3 distinct indices in [0, m) by r % m with rejection
// (the counter-based stand-in for rand() % m in reference RANSAC.cpp:180-205).
__host__ __device__ __forceinline__ void sample3(uint32_t seed_lo, uint32_t seed_hi, uint32_t h, uint32_t m,
                                                 int (&out)[3]) {
    int got = 0;
    out[0] = out[1] = out[2] = -1;
    for (uint32_t blk = 0; got < 3; ++blk) {
        uint32_t r[4];
        philox4x32_10(blk, h, 0u, 0u, seed_lo, seed_hi, r);
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (got < 3) {
                const int idx = (int)(r[w] % m);
                const bool dup = (idx == out[0]) || (idx == out[1]);
                if (!dup) {
                    if (got == 0) out[0] = idx; else if (got == 1) out[1] = idx; else out[2] = idx;
                    ++got;
                }
            }
        }
    }
}

}  // namespace pslam
