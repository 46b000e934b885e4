// hamming.cu -- brute-force Hamming matching kernels (sm_100a).
//
//   K2  bf_tile_kernel + bf_finalize_kernel : cv::BFMatcher(NORM_HAMMING, crossCheck=true).match
//        (reference src/Matcher/matcherOpenCV.cpp:97-106,198-206; semantics SURVEY A.2)
//   K2' knn2_tile_kernel + knn2_merge_kernel : knnMatch(k=2) extension (north_star ratio test)
//   K7  lc_sweep_kernel + lc_topk_kernel     : query frame vs every keyframe of the resident map DB,
//        per-keyframe mutual-NN count with distance <= tau, local top-k
//        (generalises Matcher::matchFeatureLoopClosure, reference src/Matcher/matcher.cpp:802-861)
//
// Common shape: a thread owns RQ query descriptors in registers (8 words each); train descriptors
// are staged global->shared with TMA bulk copies and read back as warp-broadcast LDS.128; distance =
// CSA-compressed popcount (common.cuh); argmin = unsigned min over (dist << 16 | index), which
// resolves ties to the lowest index exactly like OpenCV's first-argmin.
#include <stdlib.h>

#include "common.cuh"
#include "hamming_tile.cuh"
#include "kernels.h"

namespace pslam {

// In-place re-encoding of descriptor rows for ham256_key_enc (run once when keyframe descriptors are appended).
__global__ void lc_encode_rows_kernel(uint4* __restrict__ rows, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 a = rows[2 * i], b = rows[2 * i + 1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    ham256_encode(w);
    rows[2 * i] = make_uint4(w[0], w[1], w[2], w[3]);
    rows[2 * i + 1] = make_uint4(w[4], w[5], w[6], w[7]);
}

// ------------------------------------------------------------------------------------------------
// K2: one query set vs one train set, spread over the grid by train range (x) and query tile (y).
// Keys use CTA-local train indices; results go to global memory as (dist << 16 | global index).
// ------------------------------------------------------------------------------------------------
template <int RQ, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
bf_tile_kernel(const uint4* __restrict__ query, int nq, const uint4* __restrict__ train, int nt, int t_per_cta,
               uint32_t* __restrict__ rowmin_g, uint32_t* __restrict__ colmin_g) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* tile = reinterpret_cast<uint4*>(smem_raw);                                   // t_per_cta * 32 B
    uint32_t* partial = reinterpret_cast<uint32_t*>(smem_raw + (size_t)t_per_cta * 32);  // NW * kTT
    uint64_t* bar = reinterpret_cast<uint64_t*>(partial + NW * kTT);

    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int t0 = blockIdx.x * t_per_cta;
    const int tcnt = min(t_per_cta, nt - t0);
    const int qbase = blockIdx.y * (NT * RQ);

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)tcnt * 32u);
        tma_load_1d(tile, train + 2 * (size_t)t0, (uint32_t)tcnt * 32u, bar);
    }
    QueryRegs<RQ, NT> Q;
    Q.load(query, nq, qbase, tid);
    uint32_t rowmin[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) rowmin[j] = 0xffffffffu;
    __syncthreads();  // barrier init visible
    mbar_wait(bar, 0);

    for (int s = 0; s < tcnt; s += kTT) {
        const int cnt = min(kTT, tcnt - s);
        tile_compute<RQ, NT>(Q, rowmin, tile + 2 * s, cnt, (uint32_t)s, partial + warp * kTT, lane);
        __syncthreads();
        if (tid < cnt) {
            uint32_t m = partial[tid];
#pragma unroll
            for (int w = 1; w < NW; ++w) m = min(m, partial[w * kTT + tid]);
            if (m < kKeyInvalid) {  // a real query won this column
                const uint32_t g = (key_dist(m) << 16) | (uint32_t)(qbase + (int)key_qoff(m));
                if (gridDim.y == 1) colmin_g[t0 + s + tid] = g;
                else atomicMin(colmin_g + t0 + s + tid, g);
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < RQ; ++j) {
        const int q = qbase + j * NT + tid;
        if (q < nq) atomicMin(rowmin_g + q, (key_dist(rowmin[j]) << 16) | (uint32_t)(t0 + (int)key_tidx(rowmin[j])));
    }
}

// Cross-check + ordered compaction (single CTA).  out layout: [0] = n, then q[cap], t[cap], dist[cap].
__global__ void __launch_bounds__(1024, 1)
bf_finalize_kernel(const uint32_t* __restrict__ rowmin_g, const uint32_t* __restrict__ colmin_g, int nq, int cap,
                   int* __restrict__ out) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    chain_begin();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry = 0;
    __syncthreads();
    int* oq = out + 1;
    int* ot = out + 1 + cap;
    float* od = reinterpret_cast<float*>(out + 1 + 2 * cap);
    for (int base = 0; base < nq; base += 1024) {
        const int q = base + tid;
        bool keep = false;
        uint32_t rp = 0;
        if (q < nq) {
            rp = rowmin_g[q];
            const uint32_t t = rp & 0xffffu;
            keep = (rp != 0xffffffffu) && ((colmin_g[t] & 0xffffu) == (uint32_t)q);
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, keep);
        const int wpre = __popc(bal & ((1u << lane) - 1u));
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const int c = warp_tot[w];
            if (w < warp) woff += c;
            tot += c;
        }
        const int pos = carry + woff + wpre;
        if (keep && pos < cap) {
            oq[pos] = q;
            ot[pos] = (int)(rp & 0xffffu);
            od[pos] = (float)(rp >> 16);
        }
        __syncthreads();
        if (tid == 0) carry += tot;
        __syncthreads();
    }
    if (tid == 0) out[0] = carry;
}

// ------------------------------------------------------------------------------------------------
// K2': two nearest neighbours per query (ascending distance, lowest train index first on ties).
// Each CTA handles one train range; per-range top-2 go to a partial buffer and are merged per query.
// ------------------------------------------------------------------------------------------------
template <int RQ, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
knn2_tile_kernel(const uint4* __restrict__ query, int nq, const uint4* __restrict__ train, int nt, int t_per_cta,
                 uint2* __restrict__ partial_g /* [gridDim.x][nq] */) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* tile = reinterpret_cast<uint4*>(smem_raw);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)t_per_cta * 32);
    const int tid = threadIdx.x;
    const int t0 = blockIdx.x * t_per_cta;
    const int tcnt = min(t_per_cta, nt - t0);
    const int qbase = blockIdx.y * (NT * RQ);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_expect_tx(bar, (uint32_t)tcnt * 32u);
        tma_load_1d(tile, train + 2 * (size_t)t0, (uint32_t)tcnt * 32u, bar);
    }
    QueryRegs<RQ, NT> Q;
    Q.load(query, nq, qbase, tid);
    uint32_t m1[RQ], m2[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) m1[j] = m2[j] = 0xffffffffu;
    __syncthreads();
    mbar_wait(bar, 0);
#pragma unroll 2
    for (int tt = 0; tt < tcnt; ++tt) {
        const uint4 a = tile[2 * tt], b = tile[2 * tt + 1];
#pragma unroll
        for (int j = 0; j < RQ; ++j) {
            const uint32_t k = ham256_key(Q.v[j], a, b, (uint32_t)tt << kKeyQBits);
            m2[j] = min(m2[j], max(m1[j], k));
            m1[j] = min(m1[j], k);
        }
    }
#pragma unroll
    for (int j = 0; j < RQ; ++j) {
        const int q = qbase + j * NT + tid;
        if (q < nq) {
            const uint32_t g1 = (key_dist(m1[j]) << 16) | (uint32_t)(t0 + (int)key_tidx(m1[j]));
            const uint32_t g2 = (m2[j] == 0xffffffffu) ? 0xffffffffu
                                                       : ((key_dist(m2[j]) << 16) | (uint32_t)(t0 + (int)key_tidx(m2[j])));
            partial_g[(size_t)blockIdx.x * nq + q] = make_uint2(g1, g2);
        }
    }
}

// out_idx/out_dist: nq x 2 (int / float); -1 where fewer than two train descriptors exist.
__global__ void knn2_merge_kernel(const uint2* __restrict__ partial_g, int nparts, int nq, int* __restrict__ out_idx,
                                  float* __restrict__ out_dist) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
    for (int p = 0; p < nparts; ++p) {
        const uint2 v = partial_g[(size_t)p * nq + q];
        // merge two sorted pairs
        m2 = min(m2, max(m1, v.x));
        m1 = min(m1, v.x);
        m2 = min(m2, max(m1, v.y));  // v.y >= v.x, so it can only land in slot 2
    }
    out_idx[2 * q] = (m1 == 0xffffffffu) ? -1 : (int)(m1 & 0xffffu);
    out_dist[2 * q] = (m1 == 0xffffffffu) ? -1.f : (float)(m1 >> 16);
    out_idx[2 * q + 1] = (m2 == 0xffffffffu) ? -1 : (int)(m2 & 0xffffu);
    out_dist[2 * q + 1] = (m2 == 0xffffffffu) ? -1.f : (float)(m2 >> 16);
}

// ------------------------------------------------------------------------------------------------
// K7: loop-closure sweep.  Persistent CTAs walk keyframes (round-robin); the keyframe's descriptors
// stream through a kStages-deep TMA/mbarrier ring of kTT-row tiles.  Per keyframe: row minima in
// registers, column minima reduced per sub-tile into shared memory, then the cross-check count.
// ------------------------------------------------------------------------------------------------

struct SweepCursor {  // walks (keyframe, tile) items of this CTA in order
    int kf, tile, ntiles;
    int64_t off;  // descriptor offset of the keyframe
    int cnt;      // descriptors in the keyframe
};

__device__ __forceinline__ void cursor_load(SweepCursor& c, const int64_t* __restrict__ kf_off, int n_kf) {
    while (c.kf < n_kf) {  // skip empty keyframes (they score 0, written up front)
        c.off = kf_off[c.kf];
        c.cnt = (int)(kf_off[c.kf + 1] - c.off);
        c.ntiles = (c.cnt + kTT - 1) / kTT;
        if (c.ntiles > 0) break;
        c.kf += gridDim.x;
    }
    c.tile = 0;
}
__device__ __forceinline__ void cursor_next(SweepCursor& c, const int64_t* __restrict__ kf_off, int n_kf) {
    if (++c.tile >= c.ntiles) {
        c.kf += gridDim.x;
        cursor_load(c, kf_off, n_kf);
    }
}

template <int RQ, int NT, int QB = kKeyQBits>
__global__ void __launch_bounds__(NT, 512 / NT)
lc_sweep_kernel(const uint4* __restrict__ query, int nq, const uint4* __restrict__ db,
                const int64_t* __restrict__ kf_off, int n_kf, int tau, int* __restrict__ scores) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* stages = reinterpret_cast<uint4*>(smem_raw);                                        // kStages*kTT*32
    uint32_t* partial = reinterpret_cast<uint32_t*>(smem_raw + kStages * kTT * 32);            // 2*NW*kTT
    uint32_t* colmin = partial + 2 * NW * kTT;                                                 // kMaxKfDesc
    uint64_t* bars = reinterpret_cast<uint64_t*>(colmin + kMaxKfDesc);                         // kStages
    int* score_s = reinterpret_cast<int*>(bars + kStages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
        *score_s = 0;
    }
    QueryRegs<RQ, NT> Q;
    Q.template load<true>(query, nq, 0, tid);         // the resident map is stored re-encoded (ham256_key_enc)
    uint32_t rowmin[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) rowmin[j] = 0xffffffffu;
    __syncthreads();

    // keyframes with no descriptors never enter the tile stream: score them here
    for (int kf = blockIdx.x * NT + tid; kf < n_kf; kf += gridDim.x * NT)
        if (kf_off[kf + 1] == kf_off[kf]) scores[kf] = 0;

    SweepCursor prod, cons;
    prod.kf = cons.kf = blockIdx.x;
    cursor_load(cons, kf_off, n_kf);
    prod = cons;
    int issued = 0;
    if (tid == 0) {
        for (; issued < kStages - 1 && prod.kf < n_kf; ++issued) {
            const int cnt = min(kTT, prod.cnt - prod.tile * kTT);
            uint64_t* bar = bars + (issued % kStages);
            mbar_expect_tx(bar, (uint32_t)cnt * 32u);
            tma_load_1d(stages + (size_t)(issued % kStages) * kTT * 2, db + 2 * (prod.off + (int64_t)prod.tile * kTT),
                        (uint32_t)cnt * 32u, bar);
            cursor_next(prod, kf_off, n_kf);
        }
    }

    for (int it = 0; cons.kf < n_kf; ++it) {
        const int stage = it % kStages;
        const uint32_t phase = (uint32_t)(it / kStages) & 1u;
        if (tid == 0 && prod.kf < n_kf) {  // refill the stage freed by the previous iteration
            const int cnt = min(kTT, prod.cnt - prod.tile * kTT);
            const int st = issued % kStages;
            mbar_expect_tx(bars + st, (uint32_t)cnt * 32u);
            tma_load_1d(stages + (size_t)st * kTT * 2, db + 2 * (prod.off + (int64_t)prod.tile * kTT),
                        (uint32_t)cnt * 32u, bars + st);
            ++issued;
            cursor_next(prod, kf_off, n_kf);
        }
        const int tbase = cons.tile * kTT;
        const int cnt = min(kTT, cons.cnt - tbase);
        uint32_t* pbuf = partial + (it & 1) * (NW * kTT);
        mbar_wait(bars + stage, phase);
        tile_compute<RQ, NT, QB, true>(Q, rowmin, stages + (size_t)stage * kTT * 2, cnt, (uint32_t)tbase, pbuf + warp * kTT, lane);
        __syncthreads();
        if (tid < cnt) {
            uint32_t m = pbuf[tid];
#pragma unroll
            for (int w = 1; w < NW; ++w) m = min(m, pbuf[w * kTT + tid]);
            colmin[tbase + tid] = m;
        }
        const bool last = (cons.tile + 1 == cons.ntiles);
        if (last) {
            __syncthreads();
            int c = 0;
#pragma unroll
            for (int j = 0; j < RQ; ++j) {
                const uint32_t rk = rowmin[j];
                // the winner of the column this row points at must be this very (dist, t, q) key
                if (Q.off[j] < kKeyInvalid && colmin[key_tidx<QB>(rk)] == rk && (int)key_dist(rk) <= tau) ++c;
                rowmin[j] = 0xffffffffu;
            }
            c = (int)warp_add_u32((uint32_t)c);
            if (lane == 0 && c) atomicAdd(score_s, c);
            __syncthreads();
            if (tid == 0) {
                scores[cons.kf] = *score_s;
                *score_s = 0;
            }
        }
        cursor_next(cons, kf_off, n_kf);
    }
}

// ------------------------------------------------------------------------------------------------
// K7 (split form) for small maps / small shards: with fewer than a few keyframes per CTA the keyframe is too
// coarse a work unit (100 keyframes keep 100 of 296 CTAs busy).  Here the unit is one 128-row TILE:
// persistent CTAs walk all tiles of all keyframes; per tile they write the per-query row keys of that tile
// (4 KB, coalesced) and the finished column keys of its rows; lc_sweep_finalize_kernel then merges the row
// partials of each keyframe and does the cross-check count.  Extra traffic: 4 KB per tile -- only used when
// the map is small enough for that to be noise (launcher decides).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_item(int item, const int* __restrict__ tile_start, const int64_t* __restrict__ kf_off,
                                           int n_kf, int& kf, int& tile, int64_t& row0, int& cnt) {
    int lo = 0, hi = n_kf;  // largest kf with tile_start[kf] <= item
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid] <= item) lo = mid; else hi = mid;
    }
    kf = lo;
    tile = item - tile_start[kf];
    const int64_t o = kf_off[kf];
    const int kcnt = (int)(kf_off[kf + 1] - o);
    row0 = o + (int64_t)tile * kTT;
    cnt = min(kTT, kcnt - tile * kTT);
}

template <int RQ, int NT, int QB = kKeyQBits>
__global__ void __launch_bounds__(NT, 512 / NT)
lc_sweep_split_kernel(const uint4* __restrict__ query, int nq, const uint4* __restrict__ db,
                      const int64_t* __restrict__ kf_off, const int* __restrict__ tile_start, int n_kf, int n_tiles,
                      uint32_t* __restrict__ rowpart /* [n_tiles][RQ*NT] */, uint32_t* __restrict__ colmin_g /* [n_kf][4096] */) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* stages = reinterpret_cast<uint4*>(smem_raw);
    uint32_t* partial = reinterpret_cast<uint32_t*>(smem_raw + kStages * kTT * 32);  // 2*NW*kTT
    uint64_t* bars = reinterpret_cast<uint64_t*>(partial + 2 * NW * kTT);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    QueryRegs<RQ, NT> Q;
    Q.template load<true>(query, nq, 0, tid);
    __syncthreads();
    const int my_items = n_tiles > (int)blockIdx.x ? (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    auto issue = [&](int a) {
        int kf, tile, cnt; int64_t row0;
        split_item((int)blockIdx.x + a * (int)gridDim.x, tile_start, kf_off, n_kf, kf, tile, row0, cnt);
        const int st = a % kStages;
        mbar_expect_tx(bars + st, (uint32_t)cnt * 32u);
        tma_load_1d(stages + (size_t)st * kTT * 2, db + 2 * row0, (uint32_t)cnt * 32u, bars + st);
    };
    if (tid == 0)
        for (int a = 0; a < kStages - 1 && a < my_items; ++a) issue(a);
    for (int it = 0; it < my_items; ++it) {
        const int stage = it % kStages;
        const uint32_t phase = (uint32_t)(it / kStages) & 1u;
        if (tid == 0 && it + kStages - 1 < my_items) issue(it + kStages - 1);
        const int item = (int)blockIdx.x + it * (int)gridDim.x;
        int kf, tile, cnt; int64_t row0;
        split_item(item, tile_start, kf_off, n_kf, kf, tile, row0, cnt);
        const int tbase = tile * kTT;
        uint32_t rowmin[RQ];
#pragma unroll
        for (int j = 0; j < RQ; ++j) rowmin[j] = 0xffffffffu;
        uint32_t* pbuf = partial + (it & 1) * (NW * kTT);
        mbar_wait(bars + stage, phase);
        tile_compute<RQ, NT, QB, true>(Q, rowmin, stages + (size_t)stage * kTT * 2, cnt, (uint32_t)tbase, pbuf + warp * kTT, lane);
#pragma unroll
        for (int j = 0; j < RQ; ++j) rowpart[(size_t)item * (RQ * NT) + j * NT + tid] = rowmin[j];
        __syncthreads();
        if (tid < cnt) {
            uint32_t m = pbuf[tid];
#pragma unroll
            for (int w = 1; w < NW; ++w) m = min(m, pbuf[w * kTT + tid]);
            colmin_g[(size_t)kf * kMaxKfDesc + tbase + tid] = m;
        }
    }
}

// One CTA per keyframe: merge the row partials of its tiles, cross-check against the column keys, count.
__global__ void __launch_bounds__(256)
lc_sweep_finalize_kernel(const int* __restrict__ tile_start, const int64_t* __restrict__ kf_off, int nq, int row_stride, int qb,
                         const uint32_t* __restrict__ rowpart, const uint32_t* __restrict__ colmin_g, int tau,
                         int* __restrict__ scores) {
    __shared__ int s_cnt;
    const int kf = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    const int t0 = tile_start[kf], t1 = tile_start[kf + 1];
    int c = 0;
    if (t1 > t0) {
        for (int q = tid; q < nq; q += 256) {
            uint32_t rk = 0xffffffffu;
            for (int t = t0; t < t1; ++t) rk = min(rk, rowpart[(size_t)t * row_stride + q]);
            const uint32_t t = (rk >> qb) & ((1u << (kKeyDShift - qb)) - 1u);
            if (rk != 0xffffffffu && colmin_g[(size_t)kf * kMaxKfDesc + t] == rk && (int)key_dist(rk) <= tau) ++c;
        }
    }
    c = (int)warp_add_u32((uint32_t)c);
    if ((tid & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (tid == 0) scores[kf] = s_cnt;
}

// ------------------------------------------------------------------------------------------------
// V2 sweep: two nearest database descriptors per query descriptor over the WHOLE resident database
// (keyframe boundaries ignored).  Persistent CTAs walk 2048-row chunks (round-robin) through the same
// TMA ring; inside a chunk the running top-2 are 32-bit keys with chunk-local indices, at chunk end they
// are folded into 64-bit (dist << 40 | global index) keys.  Per-CTA results are merged per query.
// ------------------------------------------------------------------------------------------------
// chunk = 128 << shift rows, at most 2048: the chunk-local row index must fit the 12-bit train field of the key
__device__ __forceinline__ unsigned long long key64(uint32_t k, long long base) {
    return ((unsigned long long)key_dist(k) << 40) | (unsigned long long)(base + (long long)key_tidx(k));
}
__device__ __forceinline__ void top2_insert64(unsigned long long& g1, unsigned long long& g2, unsigned long long v) {
    const unsigned long long hi = v > g1 ? v : g1;
    g1 = v < g1 ? v : g1;
    g2 = hi < g2 ? hi : g2;
}

template <int RQ, int NT>
__global__ void __launch_bounds__(NT, 512 / NT)
lc_knn2_kernel(const uint4* __restrict__ query, int nq, const uint4* __restrict__ db, long long n_desc,
               long long desc_id_base, int tpc_shift /* log2(tiles per chunk) */,
               ulonglong2* __restrict__ partial /* [gridDim.x][nq] */) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* stages = reinterpret_cast<uint4*>(smem_raw);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kStages * kTT * 32);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
    }
    QueryRegs<RQ, NT> Q;
    Q.template load<true>(query, nq, 0, tid);
    unsigned long long g1[RQ], g2[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) g1[j] = g2[j] = ~0ull;
    __syncthreads();

    const long long n_tiles_total = (n_desc + kTT - 1) / kTT;
    const int kTilesPerChunk = 1 << tpc_shift;
    const int chunk_rows = kTilesPerChunk * kTT;
    const long long n_chunks = (n_desc + chunk_rows - 1) / chunk_rows;
    // flat list of this CTA's tiles: chunk c = blockIdx.x + i*gridDim.x, tiles c*16 .. c*16+15
    auto tile_of = [&](long long item) -> long long {
        const long long c = (long long)blockIdx.x + (item >> tpc_shift) * gridDim.x;
        return (c << tpc_shift) + (item & (kTilesPerChunk - 1));
    };
    long long my_chunks = n_chunks > blockIdx.x ? (n_chunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // number of items: full chunks have 16 tiles, the globally last chunk may have fewer
    long long n_items = 0;
    for (long long i = 0; i < my_chunks; ++i) {
        const long long c = (long long)blockIdx.x + i * gridDim.x;
        const long long first = c * kTilesPerChunk;
        const long long cnt = n_tiles_total - first;
        n_items += cnt < kTilesPerChunk ? cnt : kTilesPerChunk;
    }
    auto issue = [&](long long item) {
        const long long t = tile_of(item);
        const long long row0 = t * kTT;
        const int cnt = (int)((n_desc - row0) < kTT ? (n_desc - row0) : kTT);
        const int st = (int)(item % kStages);
        mbar_expect_tx(bars + st, (uint32_t)cnt * 32u);
        tma_load_1d(stages + (size_t)st * kTT * 2, db + 2 * row0, (uint32_t)cnt * 32u, bars + st);
    };
    if (tid == 0)
        for (long long i = 0; i < kStages - 1 && i < n_items; ++i) issue(i);

    uint32_t m1[RQ], m2[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) m1[j] = m2[j] = 0xffffffffu;
    for (long long it = 0; it < n_items; ++it) {
        const int stage = (int)(it % kStages);
        const uint32_t phase = (uint32_t)(it / kStages) & 1u;
        if (tid == 0 && it + kStages - 1 < n_items) issue(it + kStages - 1);
        const long long t = tile_of(it);
        const long long row0 = t * kTT;
        const int cnt = (int)((n_desc - row0) < kTT ? (n_desc - row0) : kTT);
        const uint32_t tbase = (uint32_t)((t & (kTilesPerChunk - 1)) * kTT);  // chunk-local row of the tile
        mbar_wait(bars + stage, phase);
        const uint4* tile = stages + (size_t)stage * kTT * 2;
        int tt = 0;
#pragma unroll 1
        for (; tt + 2 <= cnt; tt += 2) {
            const uint4 a0 = tile[2 * tt], b0 = tile[2 * tt + 1], a1 = tile[2 * tt + 2], b1 = tile[2 * tt + 3];
            const uint32_t t0 = (tbase + (uint32_t)tt) << kKeyQBits, t1 = t0 + (1u << kKeyQBits);
#pragma unroll
            for (int j = 0; j < RQ; ++j) {
                const uint32_t k0 = ham256_key_enc(Q.v[j], a0, b0, t0);
                const uint32_t k1 = ham256_key_enc(Q.v[j], a1, b1, t1);
                const uint32_t lo = min(k0, k1), hi = max(k0, k1);
                const uint32_t mid = max(m1[j], lo);
                m1[j] = min(m1[j], lo);
                m2[j] = __vimin3_u32(m2[j], mid, hi);
            }
        }
        if (tt < cnt) {
            const uint4 a0 = tile[2 * tt], b0 = tile[2 * tt + 1];
            const uint32_t t0 = (tbase + (uint32_t)tt) << kKeyQBits;
#pragma unroll
            for (int j = 0; j < RQ; ++j) {
                const uint32_t k0 = ham256_key_enc(Q.v[j], a0, b0, t0);
                m2[j] = min(m2[j], max(m1[j], k0));
                m1[j] = min(m1[j], k0);
            }
        }
        const bool chunk_end = ((t & (kTilesPerChunk - 1)) == kTilesPerChunk - 1) || (t == n_tiles_total - 1);
        if (chunk_end) {
            const long long base = desc_id_base + (t >> tpc_shift) * (long long)chunk_rows;
#pragma unroll
            for (int j = 0; j < RQ; ++j) {
                if (m1[j] != 0xffffffffu) top2_insert64(g1[j], g2[j], key64(m1[j], base));
                if (m2[j] != 0xffffffffu) top2_insert64(g1[j], g2[j], key64(m2[j], base));
                m1[j] = m2[j] = 0xffffffffu;
            }
        }
        __syncthreads();  // every warp is done with this stage before it is refilled
    }
#pragma unroll
    for (int j = 0; j < RQ; ++j) {
        const int q = j * NT + tid;
        if (q < nq) partial[(size_t)blockIdx.x * nq + q] = make_ulonglong2(g1[j], g2[j]);
    }
}

// Merge nparts sorted pairs per query.  out_keys (nullable): nq x 2 merged 64-bit keys (for the all-gather);
// out_idx / out_dist (nullable): nq x 2 global descriptor index (int64, -1 = none) and distance (float).
__global__ void lc_knn2_merge_kernel(const ulonglong2* __restrict__ partial, int nparts, int nq,
                                     unsigned long long* __restrict__ out_keys, long long* __restrict__ out_idx,
                                     float* __restrict__ out_dist) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    unsigned long long g1 = ~0ull, g2 = ~0ull;
    for (int p = 0; p < nparts; ++p) {
        const ulonglong2 v = partial[(size_t)p * nq + q];
        top2_insert64(g1, g2, v.x);
        top2_insert64(g1, g2, v.y);
    }
    if (out_keys) { out_keys[2 * q] = g1; out_keys[2 * q + 1] = g2; }
    if (out_idx) {
        out_idx[2 * q] = g1 == ~0ull ? -1 : (long long)(g1 & ((1ull << 40) - 1));
        out_idx[2 * q + 1] = g2 == ~0ull ? -1 : (long long)(g2 & ((1ull << 40) - 1));
        out_dist[2 * q] = g1 == ~0ull ? -1.f : (float)(g1 >> 40);
        out_dist[2 * q + 1] = g2 == ~0ull ? -1.f : (float)(g2 >> 40);
    }
}

// Local top-k over per-keyframe scores: score descending, keyframe id ascending on ties.  Single CTA.
// Scores are bounded by the query count (<= 1024), so a shared-memory histogram finds the cut score s*
// exactly: everything above s* is taken, ties at s* are taken in id order (ordered compaction), and the
// <= 64 selected keys are placed by rank counting.  A handful of block barriers instead of k passes.
// out_pairs: k x {score, global keyframe id}; unused slots {-1, -1}.
__global__ void __launch_bounds__(1024, 1)
lc_topk_kernel(const int* __restrict__ scores, int n_kf, int kf_id_base, int k, int* __restrict__ out_pairs) {
    __shared__ int hist[kTopkMaxScore + 2];
    __shared__ unsigned long long sel[64];
    __shared__ int warp_tot[32];
    __shared__ int s_cut, s_above, s_nsel, s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kTopkMaxScore + 2; i += 1024) hist[i] = 0;
    if (tid < 64) sel[tid] = 0ull;
    if (tid == 0) { s_nsel = 0; s_carry = 0; }
    __syncthreads();
    // scores are heavily repeated (most keyframes score ~0): aggregate equal scores inside the warp first so that
    // the shared-memory atomics do not serialise on a handful of bins
    for (int base = 0; base < n_kf; base += 1024) {
        const int i = base + tid;
        const int sc = i < n_kf ? min(max(scores[i], 0), kTopkMaxScore + 1) : -1;
        const uint32_t peers = __match_any_sync(0xffffffffu, sc);
        if (sc >= 0 && lane == __ffs(peers) - 1) atomicAdd(hist + sc, __popc(peers));
    }
    __syncthreads();
    if (warp == 0) {  // cut = largest s with count(score >= s) >= k (or 0); above = count(score > cut)
        int acc = 0, cut = 0, above = 0;
        bool found = false;
        for (int hi = kTopkMaxScore + 1; hi >= 0 && !found; hi -= 32) {
            const int sidx = hi - lane;
            const int c = sidx >= 0 ? hist[sidx] : 0;
            int incl = c;  // inclusive prefix over lanes (descending scores)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += o;
            }
            const uint32_t hit = __ballot_sync(0xffffffffu, sidx >= 0 && acc + incl >= k);
            if (hit) {
                const int l = __ffs(hit) - 1;
                cut = hi - l;
                above = acc + __shfl_sync(0xffffffffu, incl - c, l);
                found = true;
            } else {
                acc += __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        if (!found) { cut = 0; above = acc - hist[0]; }
        if (lane == 0) { s_cut = cut; s_above = above; }
    }
    __syncthreads();
    const int cut = s_cut, above = s_above;
    const int need_eq = max(0, min(k, n_kf) - above);  // ties at the cut to take, lowest ids first
    for (int base = 0; base < n_kf; base += 1024) {
        const int i = base + tid;
        const int sc = i < n_kf ? min(max(scores[i], 0), kTopkMaxScore + 1) : -1;
        if (sc > cut) {
            const int slot = atomicAdd(&s_nsel, 1);
            if (slot < 64) sel[slot] = ((unsigned long long)(uint32_t)sc << 32) | (0xffffffffu - (uint32_t)i);
        }
        if (s_carry < need_eq) {  // uniform: s_carry only changes between the barriers below
            const bool eq = (sc == cut);
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_tot[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int w = 0; w < 32; ++w) {
                const int cw = warp_tot[w];
                if (w < warp) woff += cw;
                tot += cw;
            }
            const int pos = s_carry + woff + wpre;
            if (eq && pos < need_eq) {
                const int slot = atomicAdd(&s_nsel, 1);
                if (slot < 64) sel[slot] = ((unsigned long long)(uint32_t)sc << 32) | (0xffffffffu - (uint32_t)i);
            }
            __syncthreads();
            if (tid == 0) s_carry += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    if (tid < 2 * k) out_pairs[tid] = -1;
    __syncthreads();
    if (tid < 64) {
        const unsigned long long key = sel[tid];
        if (key) {
            int rank = 0;
            for (int j = 0; j < 64; ++j) rank += (sel[j] > key) ? 1 : 0;
            if (rank < k) {
                out_pairs[2 * rank] = (int)(key >> 32);
                out_pairs[2 * rank + 1] = kf_id_base + (int)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
            }
        }
    }
}

// Merge world*k gathered {score, id} pairs (<= 1024) into the global top-k by rank counting.
__global__ void __launch_bounds__(1024, 1)
lc_merge_topk_kernel(const int* __restrict__ gathered, int n_pairs, int k, int* __restrict__ out_pairs) {
    __shared__ unsigned long long keys[1024];
    const int tid = threadIdx.x;
    unsigned long long key = 0ull;
    if (tid < n_pairs && gathered[2 * tid + 1] >= 0)
        key = ((unsigned long long)(uint32_t)gathered[2 * tid] << 32) | (0xffffffffu - (uint32_t)gathered[2 * tid + 1]);
    keys[tid] = key;
    if (tid < 2 * k) out_pairs[tid] = -1;
    __syncthreads();
    if (key) {
        int rank = 0;
        for (int j = 0; j < n_pairs; ++j) rank += (keys[j] > key) ? 1 : 0;
        if (rank < k) {
            out_pairs[2 * rank] = (int)(key >> 32);
            out_pairs[2 * rank + 1] = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
constexpr int kNT = 256;   // threads per CTA of the matching kernels
static int pick_rq(int nq) { return nq <= kNT ? 1 : (nq <= 2 * kNT ? 2 : 4); }

static int pick_t_per_cta(int nt, int qtiles, int sm_count) {
    // aim for about two CTAs per SM, at least 16 and at most 256 train descriptors per CTA
    int want = (2 * sm_count + qtiles - 1) / qtiles;
    int tpc = (nt + want - 1) / want;
    tpc = (tpc + 15) / 16 * 16;
    if (tpc < 16) tpc = 16;
    if (tpc > 256) tpc = 256;
    return tpc;
}

cudaError_t launch_bf_mutual(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, uint32_t* d_rowmin,
                             uint32_t* d_colmin, int* d_out, int cap, int sm_count, cudaStream_t st, int* launches) {
    cudaError_t e;
    if ((e = cudaMemsetAsync(d_rowmin, 0xff, sizeof(uint32_t) * (size_t)nq, st)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(d_colmin, 0xff, sizeof(uint32_t) * (size_t)nt, st)) != cudaSuccess) return e;
    const int rq = pick_rq(nq);
    const int qtiles = (nq + kNT * rq - 1) / (kNT * rq);
    const int tpc = pick_t_per_cta(nt, qtiles, sm_count);
    dim3 grid((nt + tpc - 1) / tpc, qtiles);
    const size_t smem = (size_t)tpc * 32 + sizeof(uint32_t) * (kNT / 32) * kTT + 16;
    const uint4* q4 = reinterpret_cast<const uint4*>(d_query);
    const uint4* t4 = reinterpret_cast<const uint4*>(d_train);
    if (rq == 1) e = launch_chained(bf_tile_kernel<1, kNT>, grid, dim3(kNT), smem, st, q4, nq, t4, nt, tpc, d_rowmin, d_colmin);
    else if (rq == 2) e = launch_chained(bf_tile_kernel<2, kNT>, grid, dim3(kNT), smem, st, q4, nq, t4, nt, tpc, d_rowmin, d_colmin);
    else e = launch_chained(bf_tile_kernel<4, kNT>, grid, dim3(kNT), smem, st, q4, nq, t4, nt, tpc, d_rowmin, d_colmin);
    if (e != cudaSuccess) return e;
    if ((e = launch_chained(bf_finalize_kernel, dim3(1), dim3(1024), 0, st, (const uint32_t*)d_rowmin,
                            (const uint32_t*)d_colmin, nq, cap, d_out)) != cudaSuccess) return e;
    if (launches) *launches += 2;
    return cudaGetLastError();
}

int knn2_parts(int nq, int nt, int sm_count) {
    const int rq = pick_rq(nq);
    const int qtiles = (nq + kNT * rq - 1) / (kNT * rq);
    const int tpc = pick_t_per_cta(nt, qtiles, sm_count);
    return (nt + tpc - 1) / tpc;
}

cudaError_t launch_knn2(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, uint2* d_partial, int* d_idx,
                        float* d_dist, int sm_count, cudaStream_t st, int* launches) {
    const int rq = pick_rq(nq);
    const int qtiles = (nq + kNT * rq - 1) / (kNT * rq);
    const int tpc = pick_t_per_cta(nt, qtiles, sm_count);
    dim3 grid((nt + tpc - 1) / tpc, qtiles);
    const size_t smem = (size_t)tpc * 32 + 16;
    const uint4* q4 = reinterpret_cast<const uint4*>(d_query);
    const uint4* t4 = reinterpret_cast<const uint4*>(d_train);
    if (nt > 0) {
        if (rq == 1) knn2_tile_kernel<1, kNT><<<grid, kNT, smem, st>>>(q4, nq, t4, nt, tpc, d_partial);
        else if (rq == 2) knn2_tile_kernel<2, kNT><<<grid, kNT, smem, st>>>(q4, nq, t4, nt, tpc, d_partial);
        else knn2_tile_kernel<4, kNT><<<grid, kNT, smem, st>>>(q4, nq, t4, nt, tpc, d_partial);
        if (launches) *launches += 1;
    }
    knn2_merge_kernel<<<(nq + 255) / 256, 256, 0, st>>>(d_partial, nt > 0 ? (int)grid.x : 0, nq, d_idx, d_dist);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// Shapes tried for the sweep on B200 (C4, 1 GPU): 4 queries/thread x 256 threads x 2 CTAs/SM = 775 Gcmp/s (kept);
// 8 x 128 x 4 CTAs/SM = 753; 2 x 512 x 2 CTAs/SM (32 warps/SM) = 778.  The LOP3 pipe, not latency, is the limit.
template <int NT>
static size_t lc_sweep_smem_nt() {
    return (size_t)kStages * kTT * 32 + sizeof(uint32_t) * (2 * (NT / 32) * kTT + kMaxKfDesc) + sizeof(uint64_t) * kStages + 16;
}
int lc_max_kf_desc() { return kMaxKfDesc; }
int lc_max_query() { return 2048; }
int lc_max_kf_desc_wide() { return 1 << (kKeyDShift - 11); }   // 2048 descriptors per keyframe when nq > 1024

cudaError_t lc_sweep_configure() {
    cudaError_t e;
#define CFG(RQ, NT)                                                                                              \
    if ((e = cudaFuncSetAttribute(lc_sweep_kernel<RQ, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                                  (int)lc_sweep_smem_nt<NT>())) != cudaSuccess) return e;
    CFG(1, 256) CFG(2, 256) CFG(4, 256)
#undef CFG
    if ((e = cudaFuncSetAttribute(lc_sweep_kernel<4, 512, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)lc_sweep_smem_nt<512>())) != cudaSuccess) return e;
    return cudaSuccess;
}

cudaError_t launch_lc_sweep(const uint8_t* d_query, int nq, const uint8_t* d_db, const int64_t* d_kf_off, int n_kf,
                            int tau, int* d_scores, int sm_count, cudaStream_t st, int* launches) {
    if (n_kf <= 0) return cudaSuccess;
    const int rq = pick_rq(nq);
    const uint4* q4 = reinterpret_cast<const uint4*>(d_query);
    const uint4* db4 = reinterpret_cast<const uint4*>(d_db);
    if (nq > 1024) {   // up to 2048 queries: 512 threads x 4 queries, one CTA per SM, 11/11-bit index fields
        int grid1 = sm_count < n_kf ? sm_count : n_kf;
        lc_sweep_kernel<4, 512, 11><<<grid1, 512, lc_sweep_smem_nt<512>(), st>>>(q4, nq, db4, d_kf_off, n_kf, tau, d_scores);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    int grid = 2 * sm_count;
    if (grid > n_kf) grid = n_kf;
    const size_t smem = lc_sweep_smem_nt<256>();
    if (rq == 1) lc_sweep_kernel<1, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, n_kf, tau, d_scores);
    else if (rq == 2) lc_sweep_kernel<2, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, n_kf, tau, d_scores);
    else lc_sweep_kernel<4, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, n_kf, tau, d_scores);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

size_t lc_split_rowpart_bytes(int n_tiles) { return sizeof(uint32_t) * 2048 * (size_t)(n_tiles > 0 ? n_tiles : 1); }
size_t lc_split_colmin_bytes(int n_kf) { return sizeof(uint32_t) * (size_t)kMaxKfDesc * (size_t)(n_kf > 0 ? n_kf : 1); }

cudaError_t launch_lc_sweep_split(const uint8_t* d_query, int nq, const uint8_t* d_db, const int64_t* d_kf_off,
                                  const int* d_tile_start, int n_kf, int n_tiles, int tau, uint32_t* d_rowpart,
                                  uint32_t* d_colmin, int* d_scores, int sm_count, cudaStream_t st, int* launches) {
    if (n_kf <= 0) return cudaSuccess;
    const int rq = pick_rq(nq);
    const uint4* q4 = reinterpret_cast<const uint4*>(d_query);
    const uint4* db4 = reinterpret_cast<const uint4*>(d_db);
    if (nq > 1024) {
        int grid1 = sm_count < n_tiles ? sm_count : n_tiles;
        if (n_tiles > 0) {
            const size_t smem1 = (size_t)kStages * kTT * 32 + sizeof(uint32_t) * 2 * (512 / 32) * kTT + sizeof(uint64_t) * kStages + 16;
            lc_sweep_split_kernel<4, 512, 11><<<grid1, 512, smem1, st>>>(q4, nq, db4, d_kf_off, d_tile_start, n_kf, n_tiles, d_rowpart, d_colmin);
            if (launches) *launches += 1;
        }
        lc_sweep_finalize_kernel<<<n_kf, 256, 0, st>>>(d_tile_start, d_kf_off, nq, 2048, 11, d_rowpart, d_colmin, tau, d_scores);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    int grid = 2 * sm_count;
    if (grid > n_tiles) grid = n_tiles;
    if (n_tiles > 0) {
        const size_t smem = (size_t)kStages * kTT * 32 + sizeof(uint32_t) * 2 * (256 / 32) * kTT + sizeof(uint64_t) * kStages + 16;
        if (rq == 1) lc_sweep_split_kernel<1, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, d_tile_start, n_kf, n_tiles, d_rowpart, d_colmin);
        else if (rq == 2) lc_sweep_split_kernel<2, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, d_tile_start, n_kf, n_tiles, d_rowpart, d_colmin);
        else lc_sweep_split_kernel<4, 256><<<grid, 256, smem, st>>>(q4, nq, db4, d_kf_off, d_tile_start, n_kf, n_tiles, d_rowpart, d_colmin);
        if (launches) *launches += 1;
    }
    lc_sweep_finalize_kernel<<<n_kf, 256, 0, st>>>(d_tile_start, d_kf_off, nq, rq * 256, kKeyQBits, d_rowpart, d_colmin, tau, d_scores);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lc_encode_rows(uint8_t* d_rows, long long n, cudaStream_t st, int* launches) {
    if (n <= 0) return cudaSuccess;
    lc_encode_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<uint4*>(d_rows), n);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lc_topk(const int* d_scores, int n_kf, int kf_id_base, int k, int* d_out_pairs, cudaStream_t st,
                           int* launches) {
    lc_topk_kernel<<<1, 1024, 0, st>>>(d_scores, n_kf, kf_id_base, k, d_out_pairs);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lc_merge_topk(const int* d_gathered, int n_pairs, int k, int* d_out_pairs, cudaStream_t st,
                                 int* launches) {
    lc_merge_topk_kernel<<<1, 1024, 0, st>>>(d_gathered, n_pairs, k, d_out_pairs);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// rows per chunk: large enough to amortise the 64-bit fold, small enough to give every CTA work
static int lc_knn2_shift(long long n_desc, int sm_count) {   // log2(tiles per chunk), chunk = 128 << shift rows
    const long long want = n_desc / (2LL * sm_count);
    int sh = 0;
    while (sh < 4 && ((long long)kTT << (sh + 1)) <= want) ++sh;   // at most 2048 rows (12-bit local index)
    return sh;
}
static int lc_knn2_chunk(long long n_desc, int sm_count) { return kTT << lc_knn2_shift(n_desc, sm_count); }
int lc_knn2_grid(long long n_desc, int sm_count) {
    const int chunk = lc_knn2_chunk(n_desc, sm_count);
    long long chunks = (n_desc + chunk - 1) / chunk;
    long long g = 2LL * sm_count;
    if (g > chunks) g = chunks;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_lc_knn2(const uint8_t* d_query, int nq, const uint8_t* d_db, long long n_desc, long long desc_id_base,
                           void* d_partial, int grid, int sm_count, cudaStream_t st, int* launches) {
    const int chunk = lc_knn2_shift(n_desc, sm_count);
    const uint4* q4 = reinterpret_cast<const uint4*>(d_query);
    const uint4* db4 = reinterpret_cast<const uint4*>(d_db);
    const size_t smem = (size_t)kStages * kTT * 32 + sizeof(uint64_t) * kStages + 16;
    ulonglong2* part = reinterpret_cast<ulonglong2*>(d_partial);
    if (nq > 1024) {
        lc_knn2_kernel<4, 512><<<grid, 512, smem, st>>>(q4, nq, db4, n_desc, desc_id_base, chunk, part);
        if (launches) *launches += 1;
        return cudaGetLastError();
    }
    const int rq = pick_rq(nq);
    if (rq == 1) lc_knn2_kernel<1, 256><<<grid, 256, smem, st>>>(q4, nq, db4, n_desc, desc_id_base, chunk, part);
    else if (rq == 2) lc_knn2_kernel<2, 256><<<grid, 256, smem, st>>>(q4, nq, db4, n_desc, desc_id_base, chunk, part);
    else lc_knn2_kernel<4, 256><<<grid, 256, smem, st>>>(q4, nq, db4, n_desc, desc_id_base, chunk, part);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_lc_knn2_merge(const void* d_partial, int nparts, int nq, unsigned long long* d_keys, long long* d_idx,
                                 float* d_dist, cudaStream_t st, int* launches) {
    lc_knn2_merge_kernel<<<(nq + 127) / 128, 128, 0, st>>>(reinterpret_cast<const ulonglong2*>(d_partial), nparts, nq,
                                                         d_keys, d_idx, d_dist);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
