// hamming_tile.cuh -- the inner building blocks shared by the brute-force Hamming kernels (hamming.cu) and the loop-closure
// sweep (lc_sweep.cu): per-thread query registers and the sub-tile compare with packed (dist, train, query) keys.
#pragma once
#include "common.cuh"

namespace pslam {

constexpr int kTT = 128;  // train descriptors per column-reduce sub-tile

// Per-thread query registers.  Query q = qtile_base + j*NT + tid; `off` is the query field of the packed
// key (offset inside the q-tile, < 1024); rows beyond nq carry kKeyInvalid and can never win a column.
template <int RQ, int NT>
struct QueryRegs {
    uint32_t v[RQ][8];
    uint32_t off[RQ];
    // ENC: re-encode the rows for ham256_key_enc (the train side must be encoded too: the resident keyframe map is)
    template <bool ENC = false>
    __device__ __forceinline__ void load(const uint4* __restrict__ query, int nq, int qbase, int tid) {
#pragma unroll
        for (int j = 0; j < RQ; ++j) {
            const int q = qbase + j * NT + tid;
            uint4 a = make_uint4(0, 0, 0, 0), b = a;
            if (q < nq) {
                a = __ldg(query + 2 * (size_t)q);
                b = __ldg(query + 2 * (size_t)q + 1);
            }
            v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
            v[j][4] = b.x; v[j][5] = b.y; v[j][6] = b.z; v[j][7] = b.w;
            if (ENC) ham256_encode(v[j]);
            off[j] = (uint32_t)(j * NT + tid) | ((q < nq) ? 0u : kKeyInvalid);
        }
    }
};

// One sub-tile: cnt (<= kTT) train descriptors in shared memory vs this thread's RQ queries, two train
// descriptors per iteration so that the row update is a single 3-input min; the column keys of a train fold two
// queries per 3-input min as well (1 min per pair in all).
//   rowmin[j] : running key for query j (min over train)   partial_w[tt] : this warp's min over its queries
template <int RQ, int NT, int QB = kKeyQBits, bool ENC = false>
__device__ __forceinline__ void tile_compute(const QueryRegs<RQ, NT>& Q, uint32_t (&rowmin)[RQ],
                                             const uint4* __restrict__ tile, int cnt, uint32_t tbase,
                                             uint32_t* __restrict__ partial_w, int lane) {
    int tt = 0;
#pragma unroll 1
    for (; tt + 2 <= cnt; tt += 2) {
        const uint4 a0 = tile[2 * tt], b0 = tile[2 * tt + 1], a1 = tile[2 * tt + 2], b1 = tile[2 * tt + 3];
        const uint32_t t0 = (tbase + (uint32_t)tt) << QB, t1 = t0 + (1u << QB);
        uint32_t k0[RQ], k1[RQ];
#pragma unroll
        for (int j = 0; j < RQ; ++j) {
            k0[j] = ham256_key_t<ENC>(Q.v[j], a0, b0, Q.off[j] + t0);
            k1[j] = ham256_key_t<ENC>(Q.v[j], a1, b1, Q.off[j] + t1);
            rowmin[j] = __vimin3_u32(rowmin[j], k0[j], k1[j]);
        }
        uint32_t c0 = k0[0], c1 = k1[0];
        if (RQ % 2 == 0) { c0 = min(c0, k0[1]); c1 = min(c1, k1[1]); }
#pragma unroll
        for (int j = 2 - (RQ & 1); j + 1 < RQ; j += 2) {
            c0 = __vimin3_u32(c0, k0[j], k0[j + 1]);
            c1 = __vimin3_u32(c1, k1[j], k1[j + 1]);
        }
        c0 = warp_min_u32(c0);
        c1 = warp_min_u32(c1);
        if (lane == 0) *reinterpret_cast<uint2*>(partial_w + tt) = make_uint2(c0, c1);
    }
    if (tt < cnt) {
        const uint4 a0 = tile[2 * tt], b0 = tile[2 * tt + 1];
        const uint32_t t0 = (tbase + (uint32_t)tt) << QB;
        uint32_t c0 = 0xffffffffu;
#pragma unroll
        for (int j = 0; j < RQ; ++j) {
            const uint32_t k0 = ham256_key_t<ENC>(Q.v[j], a0, b0, Q.off[j] + t0);
            rowmin[j] = min(rowmin[j], k0);
            c0 = min(c0, k0);
        }
        c0 = warp_min_u32(c0);
        if (lane == 0) partial_w[tt] = c0;
    }
}

constexpr int kStages = 4;
constexpr int kMaxKfDesc = 1 << kKeyTBits;  // 4096 descriptors per keyframe (train field of the key)
constexpr int kTopkMaxScore = 2048;

}  // namespace pslam
