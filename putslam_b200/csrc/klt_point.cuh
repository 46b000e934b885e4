// klt_point.cuh -- K11: pyramidal Lucas-Kanade for ONE point, executed by ONE WARP (klt.cu).
//
// Replaces the cv::calcOpticalFlowPyrLK call of MatcherOpenCV::performTracking (reference
// src/Matcher/matcherOpenCV.cpp:209-241; window / level / criteria parameters from
// resources/putslammatcherOpenCVParameters.xml:64).  The arithmetic is OpenCV's (un-vendored dependency); the result is
// required to equal cv2's bit for bit (positions, status, err), so every float operation below is a single IEEE
// rounding in OpenCV's operation order, and the window sums run in the accumulator order OpenCV's vector loop leaves
// behind (klt_chain / klt_combine).
//
// Work split of a warp: the window (win x win x channels values, 147 for the reference's 7 x 7 colour window) is spread
// over the lanes for the loads, the Scharr gradient, and the fixed-point bilinear interpolation; the five ordered
// partial sums of each of A11, A12, A22 (then b1, b2 per Newton step) are five independent dependent-add chains, one
// lane each; the 2 x 2 solve is done redundantly by every lane, so all control flow is warp-uniform.  Lanes exchange
// data only through the warp's shared-memory workspace between KLT_LANES sections -- which is also what lets the very
// same source run as a loop over 32 "lanes" on a CPU for the kernel's unit test (tests/klt_emul.cpp; test
// infrastructure, not a product path -- the library has no CPU route).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define KLT_HD __host__ __device__ __forceinline__
#else
#define KLT_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define KLT_LANES_BEGIN { const int lane = (int)(threadIdx.x & 31u);
#define KLT_LANES_END } __syncwarp();
#define KLT_UNROLL _Pragma("unroll")
#else
#define KLT_LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define KLT_LANES_END }
#define KLT_UNROLL
#endif

namespace pslam {

constexpr int kKltMaxLevels = 8;     // base + 7 (a 640 x 480 frame with a 7 x 7 window stops at level 6 anyway)
constexpr int kKltMaxWin = 21;
constexpr int kKltWBits = 14;        // bilinear weights in 2^14 fixed point

struct KltLevel {
    const uint8_t* I;                // previous frame, level image, tight rows of w * cn bytes
    const uint8_t* J;                // current frame
    int w, h;
};
struct KltParams {
    int n_levels;                    // max level + 1
    int win, cn;
    int max_iter;
    double eps_sq;                   // criteria.epsilon squared (double, like OpenCV)
    double min_eig_thr;
    int use_initial_flow;            // OPTFLOW_USE_INITIAL_FLOW
    int min_eig_err;                 // OPTFLOW_LK_GET_MIN_EIGENVALS
    KltLevel lv[kKltMaxLevels];
};

struct KltPlan {   // pyramid of one frame: level l is w[l] x h[l] x cn bytes (tight rows) at off[l]
    int n_levels;
    int w[kKltMaxLevels], h[kKltMaxLevels];
    size_t off[kKltMaxLevels];
    size_t bytes;
};
// cv::buildOpticalFlowPyramid: the pyramid ends at the first level whose successor would not exceed the window.
// Returns the number of levels used for `max_level`.
inline int klt_plan(int W, int H, int cn, int win, int max_level, KltPlan* P) {
    int w = W, h = H, n = 0;
    size_t off = 0;
    for (int level = 0; level <= max_level && level < kKltMaxLevels; ++level) {
        P->w[level] = w; P->h[level] = h; P->off[level] = off;
        off = (off + (size_t)w * h * cn + 255) & ~(size_t)255;
        n = level + 1;
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= win || h <= win) break;
    }
    P->n_levels = n;
    P->bytes = off;
    return n;
}
// the termination criteria as calcOpticalFlowPyrLK normalises them: count clamped to [0, 100], epsilon to [0, 10] and
// squared; a criterion that is not selected gets OpenCV's substitute (30 iterations / 0.01)
inline void klt_criteria(int use_count, int max_iter, int use_eps, double eps, int* iters, double* eps_sq) {
    int c = use_count ? max_iter : 30;
    c = c < 0 ? 0 : (c > 100 ? 100 : c);
    double e = use_eps ? eps : 0.01;
    e = e < 0. ? 0. : (e > 10. ? 10. : e);
    *iters = c;
    *eps_sq = e * e;
}

// per-warp workspace (shared memory on the device)
struct KltWork {
    uint8_t* pI;                     // (win + 3)^2 * cn : previous-frame patch, one pixel of margin for the gradient
    uint8_t* pJ;                     // (win + 1)^2 * cn : current-frame patch
    int16_t *gx, *gy;                // (win + 1)^2 * cn : Scharr gradient of the previous frame
    int16_t *Iw, *Ix, *Iy;           // win^2 * cn       : interpolated window and its gradient
    float* prod;                     // 3 * win^2 * cn   : float-converted products, laid out chain by chain (klt_slot)
    float* chain;                    // 16
    int* part;                       // 32
};
KLT_HD size_t klt_work_bytes(int win, int cn) {
    const size_t nP = (size_t)(win + 3) * (win + 3) * cn, nG = (size_t)(win + 1) * (win + 1) * cn, nW = (size_t)win * win * cn;
    size_t b = ((nP + 3) & ~(size_t)3) + ((nG + 3) & ~(size_t)3);
    b += 2 * 2 * ((nG + 1) & ~(size_t)1) + 3 * 2 * ((nW + 1) & ~(size_t)1);
    b += 3 * nW * sizeof(float) + 16 * sizeof(float) + 32 * sizeof(int);
    return (b + 15) & ~(size_t)15;
}
KLT_HD KltWork klt_carve(uint8_t* base, int win, int cn) {
    const size_t nP = (size_t)(win + 3) * (win + 3) * cn, nG = (size_t)(win + 1) * (win + 1) * cn, nW = (size_t)win * win * cn;
    KltWork w;
    w.pI = base; base += (nP + 3) & ~(size_t)3;
    w.pJ = base; base += (nG + 3) & ~(size_t)3;
    const size_t g2 = (nG + 1) & ~(size_t)1, w2 = (nW + 1) & ~(size_t)1;
    w.gx = (int16_t*)base; base += 2 * g2;
    w.gy = (int16_t*)base; base += 2 * g2;
    w.Iw = (int16_t*)base; base += 2 * w2;
    w.Ix = (int16_t*)base; base += 2 * w2;
    w.Iy = (int16_t*)base; base += 2 * w2;
    w.prod = (float*)base; base += 3 * nW * sizeof(float);
    w.chain = (float*)base; base += 16 * sizeof(float);
    w.part = (int*)base;
    return w;
}

// cv::borderInterpolate(p, n, BORDER_REFLECT_101)
KLT_HD int klt_reflect101(int p, int n) {
    if ((unsigned)p < (unsigned)n) return p;
    if (n == 1) return 0;
    do {
        p = p < 0 ? -p : 2 * n - 2 - p;
    } while ((unsigned)p >= (unsigned)n);
    return p;
}
KLT_HD int klt_round(float v) {       // cvRound: nearest, ties to even
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}
// cvFloor.  A NaN coordinate must end up outside every image, as it does in OpenCV (its conversion yields INT_MIN); the C
// cast of a NaN is undefined on the host and 0 on the device.  Values beyond the int range are out of the image either way.
KLT_HD int klt_floor(float v) { return v != v ? (int)0x80000000 : (int)floorf(v); }
KLT_HD int klt_descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }

// cv::pyrDown, 8-bit: 5 x 5 kernel [1 4 6 4 1] (x) [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8; integer, so the
// order of the 25 terms is free.  (ox, oy) = output pixel, c = channel.
KLT_HD uint8_t klt_pyrdown_px(const uint8_t* src, int w, int h, int cn, int ox, int oy, int c) {
    const int k[5] = {1, 4, 6, 4, 1};
    int xs[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 5; ++j) xs[j] = klt_reflect101(2 * ox - 2 + j, w) * cn + c;
    int v = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 5; ++i) {
        const uint8_t* row = src + (size_t)klt_reflect101(2 * oy - 2 + i, h) * w * cn;
        const int r = (int)row[xs[0]] + (int)row[xs[4]] + 4 * ((int)row[xs[1]] + (int)row[xs[3]]) + 6 * (int)row[xs[2]];
        v += k[i] * r;
    }
    return (uint8_t)((v + 128) >> 8);
}

struct KltWeights { int w00, w01, w10, w11; };
KLT_HD KltWeights klt_weights(float a, float b) {
    const float one = 1.f, s = (float)(1 << kKltWBits);
    KltWeights w;
    w.w00 = klt_round(((one - a) * (one - b)) * s);
    w.w01 = klt_round((a * (one - b)) * s);
    w.w10 = klt_round(((one - a) * b) * s);
    w.w11 = (1 << kKltWBits) - w.w00 - w.w01 - w.w10;
    return w;
}

// The window sums.  OpenCV's 128-bit loop leaves every sum  sum_i a[i] * b[i]  over a window of `rows` rows of `cols`
// values as five ordered partial sums ("chains"): chain k = 0..3 takes the values x < n8 = 8 * floor(cols / 8) of every
// row with x mod 4 == k, chain 4 the remaining values of every row, each in raster order, float accumulation of the
// float-converted integer products; the result is  chain4 + ((chain0 + chain2) + (chain1 + chain3))  (klt_combine).
//   A11 / A12 / A22: every product is converted and added on its own        -> klt_slot
//   b1 / b2 (the Newton steps): within each group of 8 the products x and x + 4 are first added as integers (a 16-bit
//   multiply-add instruction yields both), then converted and added           -> klt_slot_paired
// The lanes that compute the products store them as floats at the position they have in their chain, so that the lane
// that owns a chain reads it front to back: `len` dependent float adds and nothing else on the critical path.
KLT_HD int klt_slot(int y, int x, int rows, int cols) {
    const int n8 = (cols / 8) * 8, q = n8 >> 2;
    return x < n8 ? (x & 3) * rows * q + y * q + (x >> 2) : 4 * rows * q + y * (cols - n8) + (x - n8);
}
// unit g of row y: g < n8 / 2 is the pair (x, x + 4) with x = 8 * (g / 4) + g % 4, the rest are the single tail values
KLT_HD int klt_units_per_row(int cols) { return ((cols / 8) * 8) / 2 + cols - (cols / 8) * 8; }
KLT_HD int klt_slot_paired(int y, int g, int rows, int cols) {
    const int n8 = (cols / 8) * 8, qp = n8 >> 3, h8 = n8 >> 1;
    return g < h8 ? (g & 3) * rows * qp + y * qp + (g >> 2) : 4 * rows * qp + y * (cols - n8) + (g - h8);
}
KLT_HD float klt_chain_sum(const float* p, int len) {
    float s = 0.f;
    KLT_UNROLL
    for (int t = 0; t < len; ++t) s = s + p[t];
    return s;
}
// chain k of a sum stored with klt_slot (paired = 0) or klt_slot_paired (1)
KLT_HD float klt_chain(const float* base, int k, int rows, int cols, int paired) {
    const int n8 = (cols / 8) * 8, q = paired ? n8 >> 3 : n8 >> 2;
    if (k < 4) return klt_chain_sum(base + k * rows * q, rows * q);
    return klt_chain_sum(base + 4 * rows * q, rows * (cols - n8));
}
KLT_HD float klt_combine(const float* c) { return c[4] + ((c[0] + c[2]) + (c[1] + c[3])); }

// performTracking's pairwise rule (matcherOpenCV.cpp:254-266), seen from feature i: does the pair (i, j) remove i?
// The reference visits pairs a < b and drops a when err[a] > err[b], otherwise b (ties and NaNs included).  Closeness:
// sqrt(dx^2 + dy^2) < d in double on float differences (cv::norm(Point2f)); sq_thr is the smallest double whose square
// root is >= d, which turns it into an exact comparison of the squared distance.  lim = the smallest float >= d: a pair
// with |dx| >= lim or |dy| >= lim cannot be close (the rounded sum of squares is never below either square), so the
// double arithmetic is only done for the few candidates.
KLT_HD bool klt_pair_removes(int i, int j, float xi, float yi, float ei, float xj, float yj, float ej, double sq_thr, float lim) {
    const float dx = xi - xj, dy = yi - yj;    // the square does not see which way round the reference subtracts
    if (fabsf(dx) >= lim || fabsf(dy) >= lim || i == j) return false;
    const double s = (double)dx * (double)dx + (double)dy * (double)dy;
    if (!(s < sq_thr)) return false;
    return j > i ? (ei > ej) : !(ej > ei);
}
// lane's share of feature i's pair tests (features lane, lane + 32, ...): true when one of them removes i.
// Feature i goes when any lane says so.
KLT_HD bool klt_prune_lane(int i, int lane, int n, const float* xy, const float* err, double sq_thr, float lim) {
    const float xi = xy[2 * i], yi = xy[2 * i + 1], ei = err[i];
    bool removed = false;
    for (int j = lane; j < n; j += 32)
        removed = removed || klt_pair_removes(i, j, xi, yi, ei, xy[2 * j], xy[2 * j + 1], err[j], sq_thr, lim);
    return removed;
}

// Tracks one point through all levels.  Everything outside the KLT_LANES sections is warp-uniform.
//   (ptx, pty)  position in the previous frame (level 0)
//   (nxx, nxy)  in: initial guess when P.use_initial_flow; out: tracked position
//   WIN_T, CN_T window size / channel count known at compile time (0: taken from P) -- the index arithmetic of the lane
//               loops then compiles to constant divisions
template <int WIN_T, int CN_T>
KLT_HD void klt_track_point(const KltParams& P, const KltWork& W, float ptx, float pty, float& nxx, float& nxy,
                            uint8_t& status, float& err) {
    const int win = WIN_T ? WIN_T : P.win, cn = CN_T ? CN_T : P.cn, cols = win * cn, nW = win * cols;
    const int gcols = (win + 1) * cn, nG = (win + 1) * gcols;
    const int pcols = (win + 3) * cn, nP = (win + 3) * pcols;
    const int n8 = (cols / 8) * 8, h8 = n8 >> 1, upr = klt_units_per_row(cols), nU = win * upr;
    const int max_level = P.n_levels - 1;
    const float half = (float)((win - 1) * 0.5);
    const float flt_scale = 1.f / (float)(1 << 20);
    status = 1;
    err = 0.f;
    if (!P.use_initial_flow) { nxx = 0.f; nxy = 0.f; }

    for (int level = max_level; level >= 0; --level) {
        const uint8_t* I = P.lv[level].I;
        const uint8_t* J = P.lv[level].J;
        const int w = P.lv[level].w, h = P.lv[level].h;
        const float scale = (float)(1.0 / (double)(1 << level));
        float px = ptx * scale, py = pty * scale;
        float nx, ny;
        if (level == max_level) {
            if (P.use_initial_flow) { nx = nxx * scale; ny = nxy * scale; }
            else { nx = px; ny = py; }
        } else {
            nx = nxx * 2.f; ny = nxy * 2.f;
        }
        nxx = nx; nxy = ny;
        px = px - half; py = py - half;
        const int ix = klt_floor(px), iy = klt_floor(py);
        if (ix < -win || ix >= w || iy < -win || iy >= h) {
            if (level == 0) { status = 0; err = 0.f; }
            continue;
        }
        const KltWeights wi = klt_weights(px - (float)ix, py - (float)iy);

        // previous-frame patch, pixels (ix - 1 .. ix + win + 1) x (iy - 1 .. iy + win + 1), reflected at the borders
        KLT_LANES_BEGIN
        KLT_UNROLL
        for (int i = lane; i < nP; i += 32) {
            const int yy = i / pcols, r = i - yy * pcols, xx = r / cn, c = r - xx * cn;
            const int sx = klt_reflect101(ix - 1 + xx, w), sy = klt_reflect101(iy - 1 + yy, h);
            W.pI[i] = I[((size_t)sy * w + sx) * cn + c];
        }
        KLT_LANES_END
        // Scharr gradient at the (win + 1)^2 pixels the window touches; zero outside the image
        KLT_LANES_BEGIN
        KLT_UNROLL
        for (int i = lane; i < nG; i += 32) {
            const int yy = i / gcols, r = i - yy * gcols, xx = r / cn;
            int dx = 0, dy = 0;
            if ((unsigned)(ix + xx) < (unsigned)w && (unsigned)(iy + yy) < (unsigned)h) {
                const uint8_t* p = W.pI + (yy + 1) * pcols + cn + r;   // centre: patch (yy + 1, xx + 1), channel c
                const int a00 = p[-pcols - cn], a01 = p[-pcols], a02 = p[-pcols + cn];
                const int a10 = p[-cn], a12 = p[cn];
                const int a20 = p[pcols - cn], a21 = p[pcols], a22 = p[pcols + cn];
                dx = (3 * (a02 + a22) + 10 * a12) - (3 * (a00 + a20) + 10 * a10);
                dy = 3 * ((a20 - a00) + (a22 - a02)) + 10 * (a21 - a01);
            }
            W.gx[i] = (int16_t)dx;
            W.gy[i] = (int16_t)dy;
        }
        KLT_LANES_END
        // window of I (5 fractional bits kept) and of its gradient at the sub-pixel position
        KLT_LANES_BEGIN
        KLT_UNROLL
        for (int i = lane; i < nW; i += 32) {
            const int y = i / cols, r = i - y * cols;
            const uint8_t* p = W.pI + (y + 1) * pcols + cn + r;
            W.Iw[i] = (int16_t)klt_descale(p[0] * wi.w00 + p[cn] * wi.w01 + p[pcols] * wi.w10 + p[pcols + cn] * wi.w11, kKltWBits - 5);
            const int g = y * gcols + r;
            W.Ix[i] = (int16_t)klt_descale(W.gx[g] * wi.w00 + W.gx[g + cn] * wi.w01 + W.gx[g + gcols] * wi.w10 + W.gx[g + gcols + cn] * wi.w11, kKltWBits);
            W.Iy[i] = (int16_t)klt_descale(W.gy[g] * wi.w00 + W.gy[g + cn] * wi.w01 + W.gy[g + gcols] * wi.w10 + W.gy[g + gcols + cn] * wi.w11, kKltWBits);
            const int vx = W.Ix[i], vy = W.Iy[i], slot = klt_slot(y, r, win, cols);
            W.prod[slot] = (float)(vx * vx);
            W.prod[nW + slot] = (float)(vx * vy);
            W.prod[2 * nW + slot] = (float)(vy * vy);
        }
        KLT_LANES_END
        KLT_LANES_BEGIN
        if (lane < 15) {
            const int s = lane / 5, k = lane - 5 * s;
            W.chain[lane] = klt_chain(W.prod + s * nW, k, win, cols, 0);
        }
        KLT_LANES_END
        const float A11 = klt_combine(W.chain) * flt_scale, A12 = klt_combine(W.chain + 5) * flt_scale,
                    A22 = klt_combine(W.chain + 10) * flt_scale;
        float D = A11 * A22 - A12 * A12;
        const float dA = A11 - A22;
        const float min_eig = ((A22 + A11) - sqrtf(dA * dA + (4.f * A12) * A12)) / (float)(2 * win * win);
        if (P.min_eig_err) err = min_eig;
        if ((double)min_eig < P.min_eig_thr || D < 1.1920928955078125e-07f) {
            if (level == 0) status = 0;
            continue;
        }
        D = 1.f / D;
        nx = nx - half; ny = ny - half;
        float pdx = 0.f, pdy = 0.f;
        for (int j = 0; j < P.max_iter; ++j) {
            const int jx = klt_floor(nx), jy = klt_floor(ny);
            if (jx < -win || jx >= w || jy < -win || jy >= h) {
                if (level == 0) status = 0;
                break;
            }
            const KltWeights wj = klt_weights(nx - (float)jx, ny - (float)jy);
            KLT_LANES_BEGIN
            KLT_UNROLL
            for (int i = lane; i < nG; i += 32) {
                const int yy = i / gcols, r = i - yy * gcols, xx = r / cn, c = r - xx * cn;
                const int sx = klt_reflect101(jx + xx, w), sy = klt_reflect101(jy + yy, h);
                W.pJ[i] = J[((size_t)sy * w + sx) * cn + c];
            }
            KLT_LANES_END
            // J - I at the window positions times the gradient, unit by unit (a pair x, x + 4 or a single tail value)
            KLT_LANES_BEGIN
            KLT_UNROLL
            for (int u = lane; u < nU; u += 32) {
                const int y = u / upr, g = u - y * upr;
                const bool pair = g < h8;
                const int x0 = pair ? ((g >> 2) << 3) + (g & 3) : n8 + (g - h8);
                const uint8_t* p = W.pJ + y * gcols + x0;
                const int i0 = y * cols + x0;
                const int d0 = klt_descale(p[0] * wj.w00 + p[cn] * wj.w01 + p[gcols] * wj.w10 + p[gcols + cn] * wj.w11, kKltWBits - 5) - (int)W.Iw[i0];
                int s1 = d0 * (int)W.Ix[i0], s2 = d0 * (int)W.Iy[i0];
                if (pair) {
                    const int d1 = klt_descale(p[4] * wj.w00 + p[4 + cn] * wj.w01 + p[4 + gcols] * wj.w10 + p[4 + gcols + cn] * wj.w11, kKltWBits - 5) - (int)W.Iw[i0 + 4];
                    s1 += d1 * (int)W.Ix[i0 + 4]; s2 += d1 * (int)W.Iy[i0 + 4];
                }
                const int slot = klt_slot_paired(y, g, win, cols);
                W.prod[slot] = (float)s1;
                W.prod[nW + slot] = (float)s2;
            }
            KLT_LANES_END
            KLT_LANES_BEGIN
            if (lane < 10) {
                const int s = lane / 5, k = lane - 5 * s;
                W.chain[lane] = klt_chain(W.prod + s * nW, k, win, cols, 1);
            }
            KLT_LANES_END
            const float b1 = klt_combine(W.chain) * flt_scale, b2 = klt_combine(W.chain + 5) * flt_scale;
            const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
            nx = nx + dx; ny = ny + dy;
            nxx = nx + half; nxy = ny + half;
            if ((double)dx * (double)dx + (double)dy * (double)dy <= P.eps_sq) break;
            if (j > 0 && fabs((double)(dx + pdx)) < 0.01 && fabs((double)(dy + pdy)) < 0.01) {
                nxx = nxx - dx * 0.5f; nxy = nxy - dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0 && !P.min_eig_err) {
            // err = mean |J - I| over the window at the final position, in 1/32 grey levels -> grey levels
            const float qx = nxx - half, qy = nxy - half;
            const int jx = klt_floor(qx), jy = klt_floor(qy);
            if (jx < -win || jx >= w || jy < -win || jy >= h) { status = 0; continue; }
            const KltWeights wj = klt_weights(qx - (float)jx, qy - (float)jy);
            KLT_LANES_BEGIN
            KLT_UNROLL
            for (int i = lane; i < nG; i += 32) {
                const int yy = i / gcols, r = i - yy * gcols, xx = r / cn, c = r - xx * cn;
                const int sx = klt_reflect101(jx + xx, w), sy = klt_reflect101(jy + yy, h);
                W.pJ[i] = J[((size_t)sy * w + sx) * cn + c];
            }
            KLT_LANES_END
            KLT_LANES_BEGIN
            int e = 0;   // sum of |diff| <= 21 * 21 * 3 * 8160 < 2^24: exact as an integer, equal to OpenCV's float sum
            for (int i = lane; i < nW; i += 32) {
                const int y = i / cols, r = i - y * cols;
                const uint8_t* p = W.pJ + y * gcols + r;
                const int d = klt_descale(p[0] * wj.w00 + p[cn] * wj.w01 + p[gcols] * wj.w10 + p[gcols + cn] * wj.w11, kKltWBits - 5) - (int)W.Iw[i];
                e += d < 0 ? -d : d;
            }
            W.part[lane] = e;
            KLT_LANES_END
            int e = 0;
            for (int l = 0; l < 32; ++l) e += W.part[l];
            err = ((float)e * 1.f) / (float)(32 * win * cn * win);
        }
    }
}

}  // namespace pslam
