// lc_sweep.cu -- K7: the loop-closure / place-recognition sweep (sm_100a), range form with a fused tail.
//
//   One query frame (<= 2048 descriptors in registers) against every keyframe of the resident, re-encoded map: per
//   keyframe the mutual-nearest-neighbour count with distance <= tau (cv::BFMatcher(NORM_HAMMING, crossCheck) per
//   keyframe, the generalisation of Matcher::matchFeatureLoopClosure, reference src/Matcher/matcher.cpp:802-861), then
//   the local top-k and -- on a sharded map -- the exchange and merge of the per-rank top-k, all in ONE launch.
//
// Work decomposition.  The map is a list of 128-row tiles in keyframe order (tile_start[] = prefix count per keyframe).
// CTA c of G owns the CONTIGUOUS tile range [c*T/G, (c+1)*T/G): load is balanced to one tile (0.3 % at C4) whatever the
// number of keyframes per GPU, and consecutive tiles belong to the same keyframe, so row minima stay in registers and
// column minima in shared memory exactly as if the CTA owned whole keyframes.  Only the (at most two) keyframes cut by
// the ends of a CTA's range go through global memory: each CTA writes the row keys and column keys of its piece, adds
// its tile count to a per-keyframe counter, and the CTA that completes the count merges the pieces and scores the
// keyframe.  Nobody ever waits.  (Round 1 had two kernels for this: whole keyframes per CTA, which leaves 20 % of the
// machine idle at 1250 keyframes per GPU, and a tile form that wrote 4 KB of row keys per tile plus a finalize launch.)
//
// Tail.  The CTA that finishes last (a grid-wide counter) computes the local top-k from the scores -- no second launch --
// and, when the map is sharded over several GPUs with peer access, writes its k {score, id} pairs into slot [rank] of every
// peer's exchange buffer over NVLink, raises a flag there, waits for the flags of all peers in its own buffer and merges
// the world*k pairs: the all-gather and the merge of the NCCL path (kept as the fallback) without leaving the kernel.
#include <stdlib.h>

#include "common.cuh"
#include "hamming_tile.cuh"
#include "kernels.h"
#include "lc_tc.cuh"

namespace pslam {

__device__ __forceinline__ int range_begin(int c, int grid, int n_tiles) { return (int)(((long long)c * n_tiles) / grid); }
// CTA that owns tile g: the largest c with range_begin(c) <= g
__device__ __forceinline__ int range_owner(int g, int grid, int n_tiles) {
    int c = (int)(((long long)g * grid) / n_tiles);
    if (c >= grid) c = grid - 1;
    while (c + 1 < grid && range_begin(c + 1, grid, n_tiles) <= g) ++c;
    while (c > 0 && range_begin(c, grid, n_tiles) > g) --c;
    return c;
}

struct RangeCursor {  // walks the tiles [g, g_end) of one CTA in order
    int kf, tile, ntiles, cnt, g;
    int64_t off;
};
__device__ __forceinline__ void rc_load_kf(RangeCursor& c, const int64_t* __restrict__ kf_off, int n_kf) {
    while (c.kf < n_kf) {
        c.off = kf_off[c.kf];
        c.cnt = (int)(kf_off[c.kf + 1] - c.off);
        c.ntiles = (c.cnt + kTT - 1) / kTT;
        if (c.ntiles > 0) break;
        ++c.kf;                                    // empty keyframes own no tile
    }
}
__device__ __forceinline__ void rc_seek(RangeCursor& c, int g, const int* __restrict__ tile_start,
                                        const int64_t* __restrict__ kf_off, int n_kf) {
    int lo = 0, hi = n_kf;                         // largest kf with tile_start[kf] <= g (the last such one is non-empty)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid] <= g) lo = mid; else hi = mid;
    }
    c.kf = lo;
    c.g = g;
    rc_load_kf(c, kf_off, n_kf);
    c.tile = g - tile_start[c.kf];
}
__device__ __forceinline__ void rc_next(RangeCursor& c, const int64_t* __restrict__ kf_off, int n_kf) {
    ++c.g;
    if (++c.tile >= c.ntiles) {
        ++c.kf;
        c.tile = 0;
        rc_load_kf(c, kf_off, n_kf);
    }
}

// ---- block-wide top-k over per-keyframe scores (score descending, keyframe id ascending on ties) ----------------------
// Scores are bounded by the query count, so a shared-memory histogram finds the cut score exactly: everything above it is
// taken, ties at the cut are taken in id order, and the <= 64 selected keys are placed by rank counting.
// smem: hist[kTopkMaxScore + 2] ints, then 64 x 8 B keys, then 32 + 4 ints.  pairs_out (shared or global): k x {score, id}.
constexpr size_t kTopkSmemBytes = sizeof(int) * (kTopkMaxScore + 2 + 32 + 4) + sizeof(unsigned long long) * 64 + 16;
template <int NT>
__device__ void block_topk(const int* __restrict__ scores, int n_kf, int kf_id_base, int k, uint8_t* smem, int* pairs_out) {
    constexpr int NW = NT / 32;
    int* hist = reinterpret_cast<int*>(smem);
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(smem + ((sizeof(int) * (kTopkMaxScore + 2) + 15) & ~(size_t)15));
    int* warp_tot = reinterpret_cast<int*>(sel + 64);
    int* s_vars = warp_tot + 32;                   // cut, above, nsel, carry
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < kTopkMaxScore + 2; i += NT) hist[i] = 0;
    if (tid < 64) sel[tid] = 0ull;
    int* s_max = warp_tot + 31;                    // highest score present (warp_tot itself uses NW <= 16 slots)
    if (tid == 0) { s_vars[2] = 0; s_vars[3] = 0; *s_max = 0; }
    __syncthreads();
    // most keyframes score ~0: aggregate equal scores inside the warp first so the shared atomics do not serialise
    for (int base = 0; base < n_kf; base += NT) {
        const int i = base + tid;
        const int sc = i < n_kf ? min(max(__ldcg(scores + i), 0), kTopkMaxScore + 1) : -1;
        const uint32_t peers = __match_any_sync(0xffffffffu, sc);
        if (sc >= 0 && lane == __ffs(peers) - 1) atomicAdd(hist + sc, __popc(peers));
        const int wm = __reduce_max_sync(0xffffffffu, sc);
        if (lane == 0 && wm > 0) atomicMax(s_max, wm);
    }
    __syncthreads();
    if (warp == 0) {  // cut = largest s with count(score >= s) >= k (or 0); above = count(score > cut)
        int acc = 0, cut = 0, above = 0;
        bool found = false;
        for (int hi = *s_max; hi >= 0 && !found; hi -= 32) {      // nothing scores above *s_max
            const int sidx = hi - lane;
            const int c = sidx >= 0 ? hist[sidx] : 0;
            int incl = c;
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
        if (lane == 0) { s_vars[0] = cut; s_vars[1] = above; }
    }
    __syncthreads();
    const int cut = s_vars[0], above = s_vars[1];
    const int need_eq = max(0, min(k, n_kf) - above);  // ties at the cut to take, lowest ids first
    for (int base = 0; base < n_kf; base += NT) {
        const int i = base + tid;
        const int sc = i < n_kf ? min(max(__ldcg(scores + i), 0), kTopkMaxScore + 1) : -1;
        if (sc > cut) {
            const int slot = atomicAdd(&s_vars[2], 1);
            if (slot < 64) sel[slot] = ((unsigned long long)(uint32_t)sc << 32) | (0xffffffffu - (uint32_t)i);
        }
        if (s_vars[3] < need_eq) {  // uniform: the carry only changes between the barriers below
            const bool eq = (sc == cut);
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            const int wpre = __popc(bal & ((1u << lane) - 1u));
            if (lane == 0) warp_tot[warp] = __popc(bal);
            __syncthreads();
            int woff = 0, tot = 0;
            for (int w = 0; w < NW; ++w) {
                const int cw = warp_tot[w];
                if (w < warp) woff += cw;
                tot += cw;
            }
            const int pos = s_vars[3] + woff + wpre;
            if (eq && pos < need_eq) {
                const int slot = atomicAdd(&s_vars[2], 1);
                if (slot < 64) sel[slot] = ((unsigned long long)(uint32_t)sc << 32) | (0xffffffffu - (uint32_t)i);
            }
            __syncthreads();
            if (tid == 0) s_vars[3] += tot;
            __syncthreads();
        }
    }
    __syncthreads();
    for (int i = tid; i < 2 * k; i += NT) pairs_out[i] = -1;
    __syncthreads();
    if (tid < 64) {
        const unsigned long long key = sel[tid];
        if (key) {
            int rank = 0;
            for (int j = 0; j < 64; ++j) rank += (sel[j] > key) ? 1 : 0;
            if (rank < k) {
                pairs_out[2 * rank] = (int)(key >> 32);
                pairs_out[2 * rank + 1] = kf_id_base + (int)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
            }
        }
    }
    __syncthreads();
}

// ---- the peer exchange of the per-rank top-k (NVLink stores + flags) -----------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// pairs: this rank's k x {score, id} (global memory).  x.local / x.peer[r]: exchange buffers (layout in kernels.h).
// On return merged_out (global) holds the same global top-k on every rank.
template <int NT>
__device__ void exchange_and_merge(const LcExchange& x, const int* pairs, int k, uint8_t* smem, int* merged_out) {
    const int tid = threadIdx.x;
    const int bank = (int)(x.epoch & 1u);
    const size_t slot = ((size_t)bank * kLcMaxRanks + (size_t)x.rank) * (2 * kLcMaxTopk);
    for (int i = tid; i < x.world * 2 * k; i += NT) {                         // my pairs -> slot [bank][rank] of every peer
        const int r = i / (2 * k), e = i % (2 * k);
        x.peer[r][kLcXchgPairsOff + slot + e] = pairs[e];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < x.world)                                                         // publish: flag [bank][rank] on every peer
        st_release_sys(reinterpret_cast<uint32_t*>(x.peer[tid]) + kLcXchgFlagsOff + bank * kLcMaxRanks + x.rank, x.epoch);
    if (tid < x.world) {                                                       // wait for everybody's pairs in MY buffer
        const uint32_t* f = reinterpret_cast<const uint32_t*>(x.local) + kLcXchgFlagsOff + bank * kLcMaxRanks + tid;
        while (ld_acquire_sys(f) != x.epoch) __nanosleep(64);
    }
    __syncthreads();
    // merge world*k <= 1024 pairs by rank counting (score descending, id ascending)
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem);
    const int n_pairs = x.world * k;
    for (int i = tid; i < n_pairs; i += NT) {
        const int r = i / k, e = i % k;
        const int* p = x.local + kLcXchgPairsOff + ((size_t)bank * kLcMaxRanks + (size_t)r) * (2 * kLcMaxTopk) + 2 * e;
        const int sc = __ldcg(p), id = __ldcg(p + 1);
        keys[i] = id >= 0 ? (((unsigned long long)(uint32_t)sc << 32) | (0xffffffffu - (uint32_t)id)) : 0ull;
    }
    for (int i = tid; i < 2 * k; i += NT) merged_out[i] = -1;
    __syncthreads();
    for (int i = tid; i < n_pairs; i += NT) {
        const unsigned long long key = keys[i];
        if (!key) continue;
        int rank = 0;
        for (int j = 0; j < n_pairs; ++j) rank += (keys[j] > key) ? 1 : 0;
        if (rank < k) {
            merged_out[2 * rank] = (int)(key >> 32);
            merged_out[2 * rank + 1] = (int)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
        }
    }
}

// ---- the sweep -----------------------------------------------------------------------------------------------------------
template <int RQ, int NT, int QB = kKeyQBits>
__global__ void __launch_bounds__(NT, 512 / NT)
lc_sweep_range_kernel(const LcSweepArgs a) {
    constexpr int NW = NT / 32;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint4* stages = reinterpret_cast<uint4*>(smem_raw);                                        // kStages*kTT*32
    uint32_t* partial = reinterpret_cast<uint32_t*>(smem_raw + kStages * kTT * 32);            // 2*NW*kTT
    uint32_t* colmin = partial + 2 * NW * kTT;                                                 // kMaxKfDesc
    uint64_t* bars = reinterpret_cast<uint64_t*>(colmin + kMaxKfDesc);                         // kStages
    int* s_int = reinterpret_cast<int*>(bars + kStages);                                       // score, flag, -, -, producer cursor

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = (int)gridDim.x, b = (int)blockIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(bars + s, 1);
        fence_mbar_init();
        s_int[0] = 0; s_int[1] = 0;
    }
    if (a.qflag) {   // the query is pushed by the root rank over NVLink: wait until this query's copy has landed
        if (tid == 0) while (ld_acquire_sys(a.qflag) != a.qepoch) __nanosleep(32);
        __syncthreads();
    }
    QueryRegs<RQ, NT> Q;
    Q.template load<true>(a.query, a.nq, 0, tid);         // the resident map is stored re-encoded (ham256_key_enc)
    uint32_t rowmin[RQ];
#pragma unroll
    for (int j = 0; j < RQ; ++j) rowmin[j] = 0xffffffffu;
    __syncthreads();

    // keyframes with no descriptors own no tile: score them here
    for (int kf = b * NT + tid; kf < a.n_kf; kf += G * NT)
        if (a.kf_off[kf + 1] == a.kf_off[kf]) a.scores[kf] = 0;

    const int g0 = range_begin(b, G, a.n_tiles), g1 = range_begin(b + 1, G, a.n_tiles);
    RangeCursor cons;
    RangeCursor& prod = *reinterpret_cast<RangeCursor*>(s_int + 4);   // only thread 0 walks it: kept out of everybody's registers
    if (g0 < g1) rc_seek(cons, g0, a.tile_start, a.kf_off, a.n_kf);
    else { cons.g = g1; cons.kf = a.n_kf; cons.tile = cons.ntiles = cons.cnt = 0; cons.off = 0; }
    int issued = 0;
    auto issue = [&]() {
        const int cnt = min(kTT, prod.cnt - prod.tile * kTT);
        const int st = issued % kStages;
        mbar_expect_tx(bars + st, (uint32_t)cnt * 32u);
        tma_load_1d(stages + (size_t)st * kTT * 2, a.db + 2 * (prod.off + (int64_t)prod.tile * kTT), (uint32_t)cnt * 32u, bars + st);
        ++issued;
        rc_next(prod, a.kf_off, a.n_kf);
    };
    if (tid == 0) {
        prod = cons;
        for (int s = 0; s < kStages - 1 && prod.g < g1; ++s) issue();
    }

    int seg_tile0 = cons.tile;          // first tile (within its keyframe) of the piece being accumulated
    bool first_piece = true;            // the piece that starts at g0 (slot 2b); a later cut piece is the last one (slot 2b + 1)
    for (int it = 0; cons.g < g1; ++it) {
        const int stage = it % kStages;
        const uint32_t phase = (uint32_t)(it / kStages) & 1u;
        if (tid == 0 && prod.g < g1) issue();   // refill the stage freed by the previous iteration
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
        const bool kf_end = (cons.tile + 1 == cons.ntiles);
        if (kf_end || cons.g + 1 == g1) {                     // the piece [seg_tile0, cons.tile] of keyframe cons.kf is complete
            __syncthreads();
            const bool whole = kf_end && seg_tile0 == 0;
            if (whole) {
                int c = 0;
#pragma unroll
                for (int j = 0; j < RQ; ++j) {
                    const uint32_t rk = rowmin[j];
                    // the winner of the column this row points at must be this very (dist, t, q) key
                    if (Q.off[j] < kKeyInvalid && colmin[key_tidx<QB>(rk)] == rk && (int)key_dist(rk) <= a.tau) ++c;
                }
                c = (int)warp_add_u32((uint32_t)c);
                if (lane == 0 && c) atomicAdd(&s_int[0], c);
                __syncthreads();
                if (tid == 0) { a.scores[cons.kf] = s_int[0]; s_int[0] = 0; }
            } else {
                // a keyframe cut by the range: publish this piece, the CTA that completes the keyframe scores it
                const int slot = 2 * b + (first_piece ? 0 : 1);
                uint32_t* rp = a.rowpart + (size_t)slot * (RQ * NT);
                uint32_t* cp = a.colpart + (size_t)slot * kMaxKfDesc;
#pragma unroll
                for (int j = 0; j < RQ; ++j) rp[j * NT + tid] = rowmin[j];
                const int r0 = seg_tile0 * kTT, r1 = min(cons.cnt, (cons.tile + 1) * kTT);
                for (int r = r0 + tid; r < r1; r += NT) cp[r - r0] = colmin[r];
                __threadfence();
                __syncthreads();
                const int mine = cons.tile + 1 - seg_tile0;
                if (tid == 0) s_int[1] = (atomicAdd(a.kf_done + cons.kf, mine) + mine == cons.ntiles) ? 1 : 0;
                __syncthreads();
                if (s_int[1]) {
                    __threadfence();
                    const int T0 = a.tile_start[cons.kf], T1 = T0 + cons.ntiles;
                    const int c_lo = range_owner(T0, G, a.n_tiles), c_hi = range_owner(T1 - 1, G, a.n_tiles);
                    int c = 0;
#pragma unroll
                    for (int j = 0; j < RQ; ++j) {
                        uint32_t rk = 0xffffffffu;
                        for (int cc = c_lo; cc <= c_hi; ++cc) {
                            const int sl = 2 * cc + (T0 <= range_begin(cc, G, a.n_tiles) ? 0 : 1);
                            rk = min(rk, __ldcg(a.rowpart + (size_t)sl * (RQ * NT) + j * NT + tid));
                        }
                        if (Q.off[j] < kKeyInvalid && rk != 0xffffffffu && (int)key_dist(rk) <= a.tau) {
                            const int t = (int)key_tidx<QB>(rk);
                            const int cc = range_owner(T0 + t / kTT, G, a.n_tiles);
                            const int sb = max(T0, range_begin(cc, G, a.n_tiles));
                            const int sl = 2 * cc + (T0 <= range_begin(cc, G, a.n_tiles) ? 0 : 1);
                            if (__ldcg(a.colpart + (size_t)sl * kMaxKfDesc + (t - (sb - T0) * kTT)) == rk) ++c;
                        }
                    }
                    c = (int)warp_add_u32((uint32_t)c);
                    if (lane == 0 && c) atomicAdd(&s_int[0], c);
                    __syncthreads();
                    if (tid == 0) { a.scores[cons.kf] = s_int[0]; s_int[0] = 0; a.kf_done[cons.kf] = 0; }
                }
            }
#pragma unroll
            for (int j = 0; j < RQ; ++j) rowmin[j] = 0xffffffffu;
            seg_tile0 = 0;               // a piece only ends at a keyframe end or at the end of the range
            first_piece = false;
        }
        rc_next(cons, a.kf_off, a.n_kf);
    }

    // ---- tail: the CTA that finishes last turns the scores into the (global) top-k ----
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_int[1] = (atomicAdd(a.cta_done, 1u) == (unsigned)(G - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (!s_int[1]) return;
    __threadfence();
    if (tid == 0) *a.cta_done = 0u;
    block_topk<NT>(a.scores, a.n_kf, a.kf_id_base, a.k, smem_raw, a.out_pairs);
    if (a.x.world > 1 && a.x.local) {
        __threadfence();
        __syncthreads();
        exchange_and_merge<NT>(a.x, a.out_pairs, a.k, smem_raw, a.out_merged);
    }
}

template <int NT>
static size_t lc_range_smem() {
    size_t s = (size_t)kStages * kTT * 32 + sizeof(uint32_t) * (2 * (NT / 32) * kTT + kMaxKfDesc) + sizeof(uint64_t) * kStages + 16 + 48;
    const size_t tail = kTopkSmemBytes > sizeof(unsigned long long) * 1024 ? kTopkSmemBytes : sizeof(unsigned long long) * 1024;
    return s > tail ? s : tail;
}

cudaError_t lc_sweep_range_configure() {
    cudaError_t e;
#define CFG(RQ, NT)                                                                                              \
    if ((e = cudaFuncSetAttribute(lc_sweep_range_kernel<RQ, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                  (int)lc_range_smem<NT>())) != cudaSuccess) return e;
    CFG(1, 256) CFG(2, 256) CFG(4, 256)
#undef CFG
    if ((e = cudaFuncSetAttribute(lc_sweep_range_kernel<4, 512, 11>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)lc_range_smem<512>())) != cudaSuccess) return e;
    return cudaSuccess;
}

int lc_range_grid(int nq, int n_tiles, int sm_count) {
    int grid = nq > 1024 ? sm_count : 2 * sm_count;
    if (grid > n_tiles) grid = n_tiles;
    return grid < 1 ? 1 : grid;
}
size_t lc_range_rowpart_bytes(int grid) { return sizeof(uint32_t) * 2048 * 2 * (size_t)grid; }
size_t lc_range_colpart_bytes(int grid) { return sizeof(uint32_t) * (size_t)kMaxKfDesc * 2 * (size_t)grid; }

cudaError_t launch_lc_sweep_range(const LcSweepArgs& a, int grid, cudaStream_t st, int* launches) {
    const int nq = a.nq;
    const int rq = nq <= 256 ? 1 : (nq <= 512 ? 2 : 4);
    if (nq > 1024) lc_sweep_range_kernel<4, 512, 11><<<grid, 512, lc_range_smem<512>(), st>>>(a);
    else if (rq == 1) lc_sweep_range_kernel<1, 256><<<grid, 256, lc_range_smem<256>(), st>>>(a);
    else if (rq == 2) lc_sweep_range_kernel<2, 256><<<grid, 256, lc_range_smem<256>(), st>>>(a);
    else lc_sweep_range_kernel<4, 256><<<grid, 256, lc_range_smem<256>(), st>>>(a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// ---- tensor-core form (lc_tc.cuh): the sweep (per-keyframe finalize fused into its epilogue warps), then -- in the CTA
//      that finishes last -- the same tail as the range form: top-k, peer exchange, merge.  One launch per query.
__global__ void __launch_bounds__(tc::kThreads, 1) lc_tc_sweep_tail_kernel(const tc::SweepArgs A, const LcSweepArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ int s_last;
    tc::sweep_body(A);                      // ends with a CTA-wide barrier; shared memory is free from here on
    const int tid = threadIdx.x;
    if (tid == 0) {
        __threadfence();
        s_last = (atomicAdd(a.cta_done, 1u) == (unsigned)(gridDim.x - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid == 0) *a.cta_done = 0u;
    // keyframes beyond the reach of the grid's groups do not exist (every group strides the whole list); empty maps: n_kf == 0
    block_topk<tc::kThreads>(a.scores, a.n_kf, a.kf_id_base, a.k, smem, a.out_pairs);
    if (a.x.world > 1 && a.x.local) {
        __threadfence();
        __syncthreads();
        exchange_and_merge<tc::kThreads>(a.x, a.out_pairs, a.k, smem, a.out_merged);
    }
}
constexpr int kKnn2ProdWarps = 4;     // eight (one row per thread) measured 3 % slower
cudaError_t lc_sweep_tc_configure() {
    cudaError_t e = cudaFuncSetAttribute(lc_tc_sweep_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc::lc_tc_knn2_kernel<kKnn2ProdWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::kSmemBytes);
}
// V2 on the tensor cores: number of partial results per query (= CTA groups) and the launch
int lc_knn2_tc_parts(int nq, int sm_count) {
    const int splits = (nq + tc::kQRows - 1) / tc::kQRows;
    return sm_count / splits > 0 ? sm_count / splits : 1;
}
cudaError_t launch_lc_knn2_tc(const uint8_t* d_query, int nq, const uint8_t* d_db, long long n_desc, long long desc_id_base,
                              void* d_partial, int* d_status, int sm_count, cudaStream_t st, int* launches) {
    tc::Knn2Args A;
    A.db = reinterpret_cast<const uint32_t*>(d_db); A.n_desc = n_desc; A.desc_id_base = desc_id_base; A.db_encoded = 1;
    A.query = reinterpret_cast<const uint32_t*>(d_query); A.nq = nq;
    A.partial = reinterpret_cast<ulonglong2*>(d_partial); A.status = d_status;
    A.n_splits = (nq + tc::kQRows - 1) / tc::kQRows;
    tc::lc_tc_knn2_kernel<kKnn2ProdWarps><<<A.n_splits * lc_knn2_tc_parts(nq, sm_count), (kKnn2ProdWarps + 9) * 32, tc::kSmemBytes, st>>>(A);
    if (launches) *launches += 1;
    return cudaGetLastError();
}
int lc_tc_max_query() { return tc::kMaxQueries; }
size_t lc_tc_rowbest_bytes(int n_kf) { return sizeof(uint32_t) * (size_t)tc::kMaxQueries * (size_t)(n_kf > 0 ? n_kf : 1); }
size_t lc_tc_colbest_bytes(long long n_desc, int nq) {
    const int splits = (nq + tc::kQRows - 1) / tc::kQRows;
    return sizeof(uint32_t) * (size_t)splits * (size_t)(n_desc > 0 ? n_desc : 1);
}
cudaError_t launch_lc_sweep_tc(const LcSweepArgs& a, long long n_desc, uint32_t* d_rowbest, uint32_t* d_colbest, int* d_status,
                               int sm_count, cudaStream_t st, int* launches) {
    static_assert(tc::kSmemBytes >= (int)kTopkSmemBytes && tc::kSmemBytes >= (int)(sizeof(unsigned long long) * 1024), "tail scratch fits the tile buffers");
    tc::SweepArgs A;
    A.db = reinterpret_cast<const uint32_t*>(a.db);
    A.kf_off = reinterpret_cast<const long long*>(a.kf_off);
    A.n_kf = a.n_kf;
    A.db_encoded = 1;                       // the resident map is stored re-encoded (lc_encode_rows_kernel)
    A.query = reinterpret_cast<const uint32_t*>(a.query);
    A.nq = a.nq;
    A.n_desc = n_desc;
    A.row_best = d_rowbest; A.col_best = d_colbest; A.status = d_status;
    A.tau = a.tau; A.scores = a.scores; A.kf_done = a.kf_done;
    A.n_splits = (a.nq + tc::kQRows - 1) / tc::kQRows;
    A.qflag = a.qflag; A.qepoch = a.qepoch; A.stamps = nullptr;
    const int groups = sm_count / A.n_splits > 0 ? sm_count / A.n_splits : 1;
    lc_tc_sweep_tail_kernel<<<A.n_splits * groups, tc::kThreads, tc::kSmemBytes, st>>>(A, a);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// the root rank's side of the query broadcast over NVLink: CTA p copies the query into peer p's query buffer and
// raises that peer's query flag (st.release.sys after a system-scope fence)
__global__ void lc_push_query_kernel(const uint4* __restrict__ query, int n16, LcExchange x, uint32_t qepoch) {
    const int p = blockIdx.x;
    uint4* dst = reinterpret_cast<uint4*>(x.peer[p] + kLcXchgQueryOff);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = query[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(reinterpret_cast<uint32_t*>(x.peer[p]) + kLcXchgQFlagOff, qepoch);
}
cudaError_t launch_lc_push_query(const uint8_t* d_query, int nq, const LcExchange& x, uint32_t qepoch, cudaStream_t st, int* launches) {
    lc_push_query_kernel<<<x.world, 256, 0, st>>>(reinterpret_cast<const uint4*>(d_query), nq * 2, x, qepoch);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace pslam
