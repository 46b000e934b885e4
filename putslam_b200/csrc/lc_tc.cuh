// lc_tc.cuh -- the loop-closure sweep (K7) on the 5th-generation tensor cores (tcgen05 / TMEM), bit-exact.
//
// What it computes is what lc_sweep_range_kernel computes: for every keyframe of the map, the number of query descriptors
// q whose nearest keyframe descriptor t* (Hamming, lowest index on ties) has q as ITS nearest query (lowest index on
// ties) at a distance <= tau -- cv::BFMatcher(NORM_HAMMING, crossCheck = true).match per keyframe
// (reference: src/Matcher/matcherOpenCV.cpp:198-206 called from src/Matcher/matcher.cpp:835).
//
// Arithmetic.  Descriptor rows are expanded to signed bytes: query rows bit 0 -> +4, bit 1 -> -4, map rows +-64.  Then
//     sum_k q_k t_k = 256 (256 - 2 Ham) = 512 (128 - Ham)                       (int32 accumulation, exact)
// and 20 of the 32 slots that pad K from 256 to 288 (kind::i8 takes K in steps of 32) carry
//     + (255 - t_local) + (255 - q_local)          the index fields: the raw accumulator IS the comparison key
//     - 130 048 if the target row or the query row is padding (beyond the keyframe / beyond nq): never the maximum
// so that  max over a row    (fixed q)  = lowest distance, then lowest t
//          max over a column (fixed t)  = lowest distance, then lowest q
// with no ALU instruction per pair beyond the 3-input maximum itself.  Both maxima are wanted along the direction a
// thread owns after tcgen05.ld (one TMEM lane = one accumulator row), so every block of pairs goes through the tensor
// cores twice: queries as A / targets as B (rows = queries), and targets as A / queries as B (rows = targets) -- the
// tensor pipe has the room (the integer pipes, which bound the popcount kernel, are nearly idle here).
//
// Work split.  A CTA keeps 256 expanded queries resident in shared memory (72 KB), so a keyframe is swept by four CTAs
// (query quarters, "splits"); each streams the keyframe's rows through a ring of four 128-row tiles (144 KB), expanding
// them on the fly from the 32-byte rows in HBM (the map's layout does not change).  Per 256-row pair of tiles a CTA
// issues four accumulation groups of 9 MMAs (M = 128, N = 256, K = 32): rows = query block 0 / 1 against the pair, and
// rows = tile 0 / 1 of the pair against the 256 queries; two 256-column TMEM buffers alternate between the MMA warp and
// two sets of four epilogue warps.  Per-query results are complete inside a CTA; per-target results are partial (one per
// split): the finalize warp of the split that finishes a keyframe last merges them, cross-checks and counts.  The kernel
// around sweep_body (lc_sweep.cu: lc_tc_sweep_tail_kernel) then turns the scores into the top-k in its last CTA.
// lc_tc_knn2_kernel below is the V2 sweep (two nearest map descriptors per query) on the same pipeline.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pslam {
namespace tc {

constexpr int kRowBytes = 288;                    // 256 data slots + 32 extra
constexpr int kChunks = kRowBytes / 16;           // 16-byte K chunks per row
constexpr int kLBO = 128;                         // bytes between K-adjacent core matrices (8 rows x 16 B)
constexpr int kSBO = kChunks * 128;               // bytes between 8-row groups
constexpr int kTileRows = 128;
constexpr int kTileBytes = kTileRows * kRowBytes; // 36 864
constexpr int kPairRows = 256;
constexpr int kQRows = 256;                       // resident queries per CTA
constexpr int kSplits = 4;                        // query quarters: up to 1024 queries
constexpr int kMaxQueries = kQRows * kSplits;
constexpr int kQVal = 4, kTVal = 64;               // |q_k|, |t_k|: product 256 per agreeing slot, -256 per differing one
constexpr int kStepShift = 9;                     // accumulator = (128 - Ham) << 9 | index fields (< 512)
constexpr int kProdWarps = 4;                     // two rows of a pair per thread (8 warps, one row each, measured slower)
constexpr int kEpiWarp0 = kProdWarps;             // epilogue sets: warps kEpiWarp0 .. +3 and +4 .. +7 (lane quarter = warp & 3)
constexpr int kMmaWarp = kProdWarps + 8, kFinWarp = kProdWarps + 9;
constexpr int kThreads = (kProdWarps + 10) * 32;  // producers | 8 epilogue | MMA issuer | per-keyframe finalize
constexpr int kSmemBytes = 2 * kTileBytes + 4 * kTileBytes + 256;
constexpr long long kSpinLimit = 2000000000ll;    // ~1 s of clocks: a wrong barrier shows up as an error, not as a hang

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle: canonical layout ((8,n),2):((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((kLBO >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((kSBO >> 4) & 0x3fff) << 32;
    d |= 1ull << 46;                               // descriptor version (Blackwell)
    return d;                                      // base offset 0, layout type 0 = SWIZZLE_NONE
}
// instruction descriptor: D = s32, A = B = signed 8 bit, both K-major, dense
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}\n" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// false on time-out or when another role has given up (abort flag in shared memory)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, volatile int* abort_flag) {
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    int spins = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 1023) == 0) {
            if (*abort_flag) return false;
            if (clock64() - t0 > kSpinLimit) { *abort_flag = 1; return false; }
        }
    }
    return true;
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ int max3(int a, int b, int c) { return max(max(a, b), c); }   // ptxas fuses the pair into VIMNMX3

// 4 descriptor bits -> 4 signed bytes.  Query rows (expanded once per CTA): bit 0 -> +4, bit 1 -> -4, two multiply-adds
// and an AND (the second adds 0x04 + bit * 0xF8 per byte: 0x04 or 0xFC, no carry between bytes).
__host__ __device__ __forceinline__ uint32_t spread4_query(uint32_t nib) {
    const uint32_t bits = (nib * 0x00204081u) & 0x01010101u;            // bit j of the nibble -> byte j
    return bits * (uint32_t)(256 - 2 * kQVal) + 0x01010101u * (uint32_t)kQVal;
}
// Map rows (expanded for every sweep -- the producers' inner loop): bit 0 -> +64 = 0x40, bit 1 -> -64 = 0xC0, which differ
// in the sign bit only: one multiply puts bit j of the field at position 8 j + 7, one LOP3 masks and ORs 0x40 in.  `field`
// may carry up to three bits of garbage above the nibble (7 bits in all land on distinct positions, none of them a sign
// position, so nothing carries); the same holds for bits 4-10 with the multiplier that takes bit 4 + j to 8 j + 7, so one
// shifted word serves both nibbles of a byte.
template <uint32_t kMul = 0x10204080u>
__host__ __device__ __forceinline__ uint32_t spread4_map(uint32_t field7) {
#ifdef __CUDA_ARCH__
    // one LOP3 ((y & m) | c, LUT 0xEA) with both constants in registers: with two immediates the compiler emits two
    uint32_t r;
    const uint32_t m = 0x80808080u, c = 0x40404040u;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(field7 * kMul), "r"(m), "r"(c));
    return r;
#else
    return ((field7 * kMul) & 0x80808080u) | 0x40404040u;
#endif
}
// the two extra chunks of a row.  Slots (bytes of chunk 16, then chunk 17):
//   0,1   target index:  target rows (idx & 1, idx >> 1)            query rows (1, 2)
//   2,3   query index:   query rows  (idx & 1, idx >> 1)            target rows (1, 2)
//   4-11  padding target: target rows 0 / -128 when padding         query rows +127
//   12-19 padding query:  query rows  0 / -128 when padding         target rows +127
__host__ __device__ __forceinline__ void extra_chunks(bool is_query, int idx, bool valid, uint4& c16, uint4& c17) {
    const uint32_t own = (uint32_t)(idx & 1) | ((uint32_t)(idx >> 1) << 8), wts = 1u | (2u << 8);
    const uint32_t pad = valid ? 0u : 0x80808080u, full = 0x7f7f7f7fu;
    if (is_query) { c16 = make_uint4(wts | (own << 16), full, full, pad); c17 = make_uint4(pad, 0u, 0u, 0u); }
    else          { c16 = make_uint4(own | (wts << 16), pad, pad, full);  c17 = make_uint4(full, 0u, 0u, 0u); }
}
// one 32-byte descriptor row (8 words, already decoded to the plain bit order) -> the 288-byte operand row r of a tile
__host__ __device__ __forceinline__ void expand_row(const uint32_t (&w)[8], bool valid, uint8_t* tile, int r, int idx, bool is_query) {
    uint8_t* base = tile + (r >> 3) * kSBO + (r & 7) * 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t x = w[i];
        uint4 lo, hi;
        if (valid) {
            if (is_query) {
                lo.x = spread4_query(x & 15u);         lo.y = spread4_query((x >> 4) & 15u);
                lo.z = spread4_query((x >> 8) & 15u);  lo.w = spread4_query((x >> 12) & 15u);
                hi.x = spread4_query((x >> 16) & 15u); hi.y = spread4_query((x >> 20) & 15u);
                hi.z = spread4_query((x >> 24) & 15u); hi.w = spread4_query(x >> 28);
            } else {
                // per byte: one shift, two masks (bits 0-6 and 4-10 of the shifted word: seven consecutive bits each), and
                // for the high nibble a multiplier that takes bit 4 + j to position 8 j + 7
                constexpr uint32_t kHi = 0x01020408u;      // 2^3 + 2^10 + 2^17 + 2^24
                const uint32_t x1 = x >> 8, x2 = x >> 16, x3 = x >> 24;
                lo.x = spread4_map(x & 0x7fu);   lo.y = spread4_map<kHi>(x & 0x7f0u);
                lo.z = spread4_map(x1 & 0x7fu);  lo.w = spread4_map<kHi>(x1 & 0x7f0u);
                hi.x = spread4_map(x2 & 0x7fu);  hi.y = spread4_map<kHi>(x2 & 0x7f0u);
                hi.z = spread4_map(x3 & 0x7fu);  hi.w = spread4_map<kHi>(x3 & 0xf0u);
            }
        } else {
            lo = make_uint4(0, 0, 0, 0); hi = lo;
        }
        *reinterpret_cast<uint4*>(base + (2 * i) * kLBO) = lo;
        *reinterpret_cast<uint4*>(base + (2 * i + 1) * kLBO) = hi;
    }
    uint4 c16, c17;
    extra_chunks(is_query, idx, valid, c16, c17);
    *reinterpret_cast<uint4*>(base + 16 * kLBO) = c16;
    *reinterpret_cast<uint4*>(base + 17 * kLBO) = c17;
}
// rows of the map are stored re-encoded for the popcount kernel (ham256_encode, common.cuh); undo it
__host__ __device__ __forceinline__ void decode_row(uint32_t (&w)[8]) {
    const uint32_t e2 = w[2], e5 = w[5];
    w[2] = e2 ^ w[0] ^ w[1];
    w[5] = e5 ^ w[3] ^ w[4];
    w[6] = w[6] ^ e2 ^ e5;
}

// producer helper: this thread's rows of the pair that starts at `row0` (rows beyond `row_end` are padding)
template <int PW = kProdWarps>
struct PairRows {
    static constexpr int R = 256 / (PW * 32);
    uint32_t w[R][8];
    bool valid[R];
};
template <int PW>
__device__ __forceinline__ void load_pair_rows(const uint32_t* __restrict__ db, long long row0, long long row_end, int tid, PairRows<PW>& r) {
#pragma unroll
    for (int h = 0; h < PairRows<PW>::R; ++h) {
        const long long row = row0 + h * (PW * 32) + tid;
        r.valid[h] = row < row_end;
        if (r.valid[h]) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(db + (size_t)row * 8));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(db + (size_t)row * 8) + 1);
            r.w[h][0] = a.x; r.w[h][1] = a.y; r.w[h][2] = a.z; r.w[h][3] = a.w; r.w[h][4] = b.x; r.w[h][5] = b.y; r.w[h][6] = b.z; r.w[h][7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) r.w[h][i] = 0;
        }
    }
}
template <int PW>
__device__ __forceinline__ void expand_pair_rows(PairRows<PW>& r, bool db_encoded, uint8_t* pair_tiles, int tid) {
#pragma unroll
    for (int h = 0; h < PairRows<PW>::R; ++h) {
        const int rp = h * (PW * 32) + tid;               // row inside the pair
        if (db_encoded && r.valid[h]) decode_row(r.w[h]);
        expand_row(r.w[h], r.valid[h], pair_tiles + (rp >> 7) * kTileBytes, rp & 127, 255 - rp, false);
    }
}

// maximum of the 256 accumulators of this thread's TMEM lane in buffer `taddr` (lane already folded into the address).
// (Measured: 2 or 4 independent chains instead of one, eight producer warps instead of four, and all eight epilogue warps
// on every group with the halves merged through shared memory are each 2-6 % SLOWER: neither the epilogue nor the
// producers' instruction count is the limiter; what did help was hiding the producers' load latency and making the
// expansion cheaper.  The tensor pipe is 92 % busy.)
__device__ __forceinline__ int row_max_256(uint32_t taddr) {
    int m = -0x7fffffff;
#pragma unroll
    for (int c = 0; c < 256; c += 64) {
        int a[32], b[32];
        tmem_ld32(taddr + (uint32_t)c, a);
        tmem_ld32(taddr + (uint32_t)(c + 32), b);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) m = max3(m, a[j], a[j + 1]);
#pragma unroll
        for (int j = 0; j < 32; j += 2) m = max3(m, b[j], b[j + 1]);
    }
    return m;
}

struct SweepArgs {
    const uint32_t* db;          // n_desc x 8 words
    const long long* kf_off;     // n_kf + 1
    int n_kf;
    int db_encoded;              // rows stored with ham256_encode
    const uint32_t* query;       // nq x 8 words, plain
    int nq;
    long long n_desc;            // stride of col_best between splits
    uint32_t* row_best;          // [n_kf][kMaxQueries]: Ham << 16 | t (index inside the keyframe)
    uint32_t* col_best;          // [n_splits][n_desc]:  Ham << 16 | q
    int* status;                 // set non-zero when a wait timed out
    int tau;                     // fused finalize: score[kf] = #{cross-check matches with Ham <= tau}
    int* scores;                 // [n_kf]; written by the CTA that finishes a keyframe last (null: no fused finalize, see finalize_keyframe)
    int* kf_done;                // [n_kf] arrival counters of the splits, zero between launches
    int n_splits;                // ceil(nq / kQRows): query quarters in use; the grid is n_splits x groups
    const uint32_t* qflag;       // wait until *qflag == qepoch before reading the query (pushed by a peer over NVLink); null: it is here
    uint32_t qepoch;
    unsigned long long* stamps;  // null, or 8 x %globaltimer of CTA 0 (bench/tc_probe.cu): entry, roles start, first / last accumulator group read, exit
};

__device__ __forceinline__ unsigned long long global_ns() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void sweep_body(const SweepArgs& A) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;                                 // 2 tiles: the CTA's 256 queries
    uint8_t* st = smem + 2 * kTileBytes;                // ring of 4 tiles = 2 pairs
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);   // full[2], empty[2], tfull[2], tempty[2], fin_full, fin_free
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = (int)blockIdx.x % A.n_splits, group = (int)blockIdx.x / A.n_splits, n_groups = (int)gridDim.x / A.n_splits;
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4), bar_tempty = smem_u32(bars + 6);
    const uint32_t bar_fin_full = smem_u32(bars + 8), bar_fin_free = smem_u32(bars + 9);
    const bool stamp = A.stamps && blockIdx.x == 0;
    if (stamp && tid == 0) A.stamps[0] = global_ns();

    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, kProdWarps);   // one arrival per producer warp
            mbar_init(bar_empty + 8 * i, 1);      // tcgen05.commit
            mbar_init(bar_tfull + 8 * i, 1);      // tcgen05.commit
            mbar_init(bar_tempty + 8 * i, 4);     // one arrival per epilogue warp of the set
        }
        mbar_init(bar_fin_full, 8);               // the eight epilogue warps have written a keyframe's results
        mbar_init(bar_fin_free, 1);               // the finalize warp is done with the previous keyframe
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_before();
    __syncthreads();                                     // barriers initialised, TMEM address published
    fence_after();
    // The producers go straight to the map (their first pair takes longer than the query takes to arrive and expand);
    // everybody else expands the CTA's queries first and meets at a named barrier of their own.
    if (warp >= kProdWarps) {
        constexpr int kRest = kThreads - kProdWarps * 32;
        const int t2 = tid - kProdWarps * 32;
        if (A.qflag) {        // the query is pushed by the root rank over NVLink: wait until this query's copy has landed
            if (t2 == 0) {
                uint32_t v;
                do { asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(A.qflag) : "memory"); if (v != A.qepoch) __nanosleep(32); } while (v != A.qepoch);
            }
            asm volatile("bar.sync 2, %0;" :: "n"(kRest) : "memory");
        }
        if (t2 < kQRows) {
            const int q = split * kQRows + t2;
            uint32_t w[8];
            const bool valid = q < A.nq;
            if (valid) {      // L2 loads: with a pushed query the bytes were written by a peer while this kernel was already running
                const uint4 a = __ldcg(reinterpret_cast<const uint4*>(A.query + (size_t)q * 8));
                const uint4 b = __ldcg(reinterpret_cast<const uint4*>(A.query + (size_t)q * 8) + 1);
                w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = 0;
            }
            expand_row(w, valid, sq + (t2 >> 7) * kTileBytes, t2 & 127, 255 - t2, true);
        }
        fence_async_smem();
        asm volatile("bar.sync 2, %0;" :: "n"(kRest) : "memory");
    }
    const uint32_t tm = s_tmem;
    volatile int* abort_flag = &s_abort;
    if (stamp && (tid == 0 || tid == kMmaWarp * 32)) A.stamps[tid == 0 ? 1 : 2] = global_ns();   // producers / the rest past the prologue

    if (warp < kProdWarps) {
        // ===== producers: two rows per thread and pair.  The loads of the NEXT pair are issued before this pair is
        //       expanded, so their latency (an L2 miss is ~2000 clk, as long as a pair's MMAs) never sits between two pairs. =====
        uint32_t it = 0;
        int kf = group, p = 0, pairs = 0;
        long long r0 = 0, r1 = 0;
        auto seek = [&]() {          // (kf, p) -> the next pair that exists; kf >= n_kf when there is none
            while (kf < A.n_kf) {
                r0 = A.kf_off[kf]; r1 = A.kf_off[kf + 1];
                pairs = (int)((r1 - r0 + kPairRows - 1) / kPairRows);
                if (p < pairs) return;
                kf += n_groups; p = 0;
            }
        };
        PairRows<kProdWarps> cur, nxt;
        seek();
        if (kf < A.n_kf) load_pair_rows(A.db, r0 + (long long)p * kPairRows, r1, tid, nxt);
        while (kf < A.n_kf) {
            cur = nxt;
            ++p; seek();
            if (kf < A.n_kf) load_pair_rows(A.db, r0 + (long long)p * kPairRows, r1, tid, nxt);
            const int s = it & 1;
            if (!mbar_wait(bar_empty + 8 * s, ((it >> 1) & 1) ^ 1, abort_flag)) goto done;
            expand_pair_rows(cur, A.db_encoded != 0, st + 2 * s * kTileBytes, tid);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
            ++it;
        }
    } else if (warp == kMmaWarp) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint64_t dq = make_desc(smem_u32(sq));
            const uint32_t idesc = make_idesc(128, 256);
            uint32_t it = 0, g = 0;
            for (int kf = group; kf < A.n_kf; kf += n_groups) {
                const long long r0 = A.kf_off[kf], r1 = A.kf_off[kf + 1];
                const int pairs = (int)((r1 - r0 + kPairRows - 1) / kPairRows);
                for (int p = 0; p < pairs; ++p, ++it) {
                    const int s = it & 1;
                    if (!mbar_wait(bar_full + 8 * s, (it >> 1) & 1, abort_flag)) goto done;
                    fence_after();
                    const uint64_t dt = make_desc(smem_u32(st + 2 * s * kTileBytes));
#pragma unroll 1
                    for (int grp = 0; grp < 4; ++grp, ++g) {
                        const int buf = grp & 1;      // = g & 1: four groups per pair
                        // rows = query block `buf` vs the pair (grp 0, 1); rows = tile `buf` of the pair vs the queries (grp 2, 3)
                        const uint64_t ad = (grp < 2 ? dq : dt) + (uint64_t)((buf * kTileBytes) >> 4);
                        const uint64_t bd = grp < 2 ? dt : dq;
                        if (!mbar_wait(bar_tempty + 8 * buf, ((g >> 1) & 1) ^ 1, abort_flag)) goto done;
                        fence_after();
#pragma unroll
                        for (int k = 0; k < kRowBytes / 32; ++k)
                            mma_i8(tm + (uint32_t)(buf * 256), ad + (uint64_t)((k * 2 * kLBO) >> 4), bd + (uint64_t)((k * 2 * kLBO) >> 4), idesc, k > 0);
                        mma_commit(bar_tfull + 8 * buf);
                    }
                    mma_commit(bar_empty + 8 * s);
                }
            }
        }
    } else if (warp == kFinWarp) {
        // ===== per-keyframe finalize.  The split that finishes a keyframe last merges the partial column results,
        //       cross-checks and counts -- off the critical path: the epilogue warps never wait for a fence or an atomic.
        if (A.scores) {
            uint32_t kfi = 0;
            for (int kf = group; kf < A.n_kf; kf += n_groups, ++kfi) {
                const long long r0 = A.kf_off[kf];
                const int n_t = (int)(A.kf_off[kf + 1] - r0);
                if (!mbar_wait(bar_fin_full, kfi & 1, abort_flag)) goto done;
                int last = 0;
                if (lane == 0) {
                    __threadfence();                                      // the epilogue warps' results, observed through the barrier
                    last = atomicAdd(A.kf_done + kf, 1) == A.n_splits - 1;
                    if (last) A.kf_done[kf] = 0;
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence();
                    int cnt = 0;
                    if (n_t > 0)
                        for (int q0 = 0; q0 < A.nq; q0 += 256) {              // 8 queries per lane in flight
                            uint32_t rb[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) { const int q = q0 + 32 * i + lane; rb[i] = q < A.nq ? __ldcg(A.row_best + (size_t)kf * kMaxQueries + q) : 0xffffffffu; }
                            // all column loads of the 8 queries issued before the first use: one L2 round trip, not 32
                            uint32_t c[8][kSplits];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int ham = (int)(rb[i] >> 16), t = (int)(rb[i] & 0xffffu);
                                const bool ok = ham <= A.tau && t < n_t;
#pragma unroll
                                for (int sp = 0; sp < kSplits; ++sp)
                                    c[i][sp] = (ok && sp < A.n_splits) ? __ldcg(A.col_best + (size_t)sp * (size_t)A.n_desc + (size_t)(r0 + t)) : 0xffffffffu;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t m = min(min(c[i][0], c[i][1]), min(c[i][2], c[i][3]));
                                cnt += (m != 0xffffffffu && (int)(m & 0xffffu) == q0 + 32 * i + lane) ? 1 : 0;
                            }
                        }
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                    if (lane == 0) A.scores[kf] = cnt;
                }
                if (stamp && lane == 0) A.stamps[5] = global_ns();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_fin_free);
            }
        }
    } else {
        // ===== epilogue: set b = (warp - 4) >> 2 owns TMEM buffer b; lane quarter = warp & 3 =====
        const int b = (warp - kEpiWarp0) >> 2, lq = warp & 3;
        const uint32_t taddr = tm + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * 256);
        const int row_in_blk = lq * 32 + lane;                 // accumulator row of this thread
        const int q_local = b * kTileRows + row_in_blk;        // its query in the O1 groups
        const int iq = 255 - q_local;
        uint32_t n = 0;                                        // groups seen on this buffer
        uint32_t kfi = 0;                                      // keyframes finished by this CTA
        for (int kf = group; kf < A.n_kf; kf += n_groups) {
            const long long r0 = A.kf_off[kf], r1 = A.kf_off[kf + 1];
            const int n_t = (int)(r1 - r0);
            const int pairs = (n_t + kPairRows - 1) / kPairRows;
            int best_h = -0x7fffffff, best_m = 0, best_p = 0;
            for (int p = 0; p < pairs; ++p) {
                // --- rows = queries: best target of the pair for this thread's query ---
                if (!mbar_wait(bar_tfull + 8 * b, n & 1, abort_flag)) goto done;
                fence_after();
                int m = row_max_256(taddr);
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
                if (stamp && tid == kEpiWarp0 * 32 && n == 0) A.stamps[3] = global_ns();
                ++n;
                const int h = m >> kStepShift;                 // 128 - Ham (the index fields sum to < 512)
                if (h > best_h) { best_h = h; best_m = m; best_p = p; }     // later pairs hold higher t: strictly better only
                // --- rows = targets: best query (of this split) for this thread's target ---
                if (!mbar_wait(bar_tfull + 8 * b, n & 1, abort_flag)) goto done;
                fence_after();
                m = row_max_256(taddr);
                fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
                ++n;
                const int t_local = b * kTileRows + row_in_blk;            // row inside the pair
                const int t_kf = p * kPairRows + t_local;
                if (t_kf < n_t) {
                    const int ham = 128 - (m >> kStepShift);
                    const int itv = 255 - t_local;
                    const int ql = 255 - ((m & 511) - itv);
                    A.col_best[(size_t)split * (size_t)A.n_desc + (size_t)(r0 + t_kf)] = ((uint32_t)ham << 16) | (uint32_t)(split * kQRows + ql);
                }
            }
            {
                uint32_t rb = 0xffffffffu;
                if (pairs > 0) {
                    const int ham = 128 - best_h;
                    const int tl = 255 - ((best_m & 511) - iq);
                    rb = ((uint32_t)ham << 16) | (uint32_t)(best_p * kPairRows + tl);
                }
                A.row_best[(size_t)kf * kMaxQueries + split * kQRows + q_local] = rb;
            }
            if (stamp && tid == kEpiWarp0 * 32) A.stamps[4] = global_ns();      // overwritten per keyframe: the last one stays
            if (A.scores) {      // hand the keyframe to the finalize warp (which must be done with the previous one)
                if (!mbar_wait(bar_fin_free, (kfi & 1) ^ 1, abort_flag)) goto done;
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_fin_full);
            }
            ++kfi;
        }
    }
done:
    __syncwarp();
    fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_free(tm, 512);
    if (tid == 0 && s_abort) *A.status = 1;
    if (stamp && tid == 0) A.stamps[6] = global_ns();
}
__global__ void __launch_bounds__(kThreads, 1) lc_tc_sweep_kernel(const SweepArgs A) { sweep_body(A); }

// ---------------------------------------------------------------------------------------------------------------------
// V2: the two nearest map descriptors of every query descriptor over the WHOLE map (knnMatch(query, map, 2): the ratio
// test's input).  Same operands and pipeline as the sweep, one orientation only (rows = queries): a pair of tiles costs
// two accumulation groups instead of four.  The epilogue keeps a running (best, second) per query row; a batch of 64
// accumulators is reduced to its maximum (the same 3-input maxima as the sweep) and only looked at value by value when
// that maximum beats the current SECOND best -- which, rows being visited in ascending index order and ties going to the
// lower index, happens O(log n) times per query over the whole map (but 32 x that per warp: the second look is therefore
// branch-free for the whole warp; the first version, a per-value scan with divergent branches, ran at a third of the speed).
// ---------------------------------------------------------------------------------------------------------------------
struct Knn2Args {
    const uint32_t* db;          // n_desc x 8 words
    long long n_desc;
    long long desc_id_base;      // global index of row 0 (sharded maps)
    int db_encoded;
    const uint32_t* query;       // nq x 8 words, plain
    int nq;
    ulonglong2* partial;         // [groups][nq]: the two smallest keys  Ham << 40 | global index  (~0 = none)
    int* status;
    int n_splits;
};

__device__ __forceinline__ void top2_push(int h, int p, int t, int& h1, int& p1, int& t1, int& h2, int& p2, int& t2) {
    if (h > h1) { h2 = h1; p2 = p1; t2 = t1; h1 = h; p1 = p; t1 = t; }
    else if (h > h2) { h2 = h; p2 = p; t2 = t; }
}

// PW producer warps (4: two rows of a pair per thread; 8 was tried -- V2 issues half the MMAs per pair -- and is slower),
// then 8 epilogue warps and the MMA warp
template <int PW>
__global__ void __launch_bounds__((PW + 9) * 32, 1) lc_tc_knn2_kernel(const Knn2Args A) {
    constexpr int kEpi0 = PW, kMma = PW + 8;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sq = smem;
    uint8_t* st = smem + 2 * kTileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTileBytes);   // full[2], empty[2], tfull[2], tempty[2]
    __shared__ uint32_t s_tmem;
    __shared__ int s_abort;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int split = (int)blockIdx.x % A.n_splits, group = (int)blockIdx.x / A.n_splits, n_groups = (int)gridDim.x / A.n_splits;
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + 2), bar_tfull = smem_u32(bars + 4), bar_tempty = smem_u32(bars + 6);
    const long long n_pairs = (A.n_desc + kPairRows - 1) / kPairRows;
    if (tid == 0) {
        s_abort = 0;
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_full + 8 * i, PW);
            mbar_init(bar_empty + 8 * i, 1);
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMma) tmem_alloc(smem_u32(&s_tmem), 512);
    if (tid < kQRows) {
        const int q = split * kQRows + tid;
        uint32_t w[8];
        const bool valid = q < A.nq;
        if (valid) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(A.query + (size_t)q * 8));
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(A.query + (size_t)q * 8) + 1);
            w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = 0;
        }
        expand_row(w, valid, sq + (tid >> 7) * kTileBytes, tid & 127, 255 - tid, true);
    }
    fence_async_smem();
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tm = s_tmem;
    volatile int* abort_flag = &s_abort;

    if (warp < PW) {
        uint32_t it = 0;
        PairRows<PW> cur, nxt;
        if (group < n_pairs) load_pair_rows(A.db, (long long)group * kPairRows, A.n_desc, tid, nxt);
        for (long long p = group; p < n_pairs; p += n_groups, ++it) {
            const int s = it & 1;
            cur = nxt;
            if (p + n_groups < n_pairs) load_pair_rows(A.db, (p + n_groups) * kPairRows, A.n_desc, tid, nxt);   // next pair's loads in flight
            if (!mbar_wait(bar_empty + 8 * s, ((it >> 1) & 1) ^ 1, abort_flag)) goto done;
            expand_pair_rows(cur, A.db_encoded != 0, st + 2 * s * kTileBytes, tid);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * s);
        }
    } else if (warp == kMma) {
        if (lane == 0) {
            const uint64_t dq = make_desc(smem_u32(sq));
            const uint32_t idesc = make_idesc(128, 256);
            uint32_t it = 0;
            for (long long p = group; p < n_pairs; p += n_groups, ++it) {
                const int s = it & 1;
                if (!mbar_wait(bar_full + 8 * s, (it >> 1) & 1, abort_flag)) goto done;
                fence_after();
                const uint64_t dt = make_desc(smem_u32(st + 2 * s * kTileBytes));
#pragma unroll 1
                for (int buf = 0; buf < 2; ++buf) {                  // rows = query block `buf` against the pair
                    if (!mbar_wait(bar_tempty + 8 * buf, (it & 1) ^ 1, abort_flag)) goto done;
                    fence_after();
                    const uint64_t ad = dq + (uint64_t)((buf * kTileBytes) >> 4);
#pragma unroll
                    for (int k = 0; k < kRowBytes / 32; ++k)
                        mma_i8(tm + (uint32_t)(buf * 256), ad + (uint64_t)((k * 2 * kLBO) >> 4), dt + (uint64_t)((k * 2 * kLBO) >> 4), idesc, k > 0);
                    mma_commit(bar_tfull + 8 * buf);
                }
                mma_commit(bar_empty + 8 * s);
            }
        }
    } else {
        const int b = (warp - kEpi0) >> 2, lq = warp & 3;
        const uint32_t taddr = tm + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * 256);
        const int q_local = b * kTileRows + lq * 32 + lane;
        const int iq = 255 - q_local;
        constexpr int kNone = -0x7fffffff;
        int h1 = kNone, p1 = 0, t1 = 0, h2 = kNone, p2 = 0, t2 = 0;
        if (split * kQRows + q_local >= A.nq) h1 = h2 = 0x7fffffff;    // padding query rows never ask for the slow path
        uint32_t it = 0;
        int pi = 0;                                               // pair counter of this CTA (global pair = group + pi * n_groups)
        for (long long p = group; p < n_pairs; p += n_groups, ++it, ++pi) {
            if (!mbar_wait(bar_tfull + 8 * b, it & 1, abort_flag)) goto done;
            fence_after();
#pragma unroll
            for (int c = 0; c < 256; c += 64) {
                int a[32], bb[32];
                tmem_ld32(taddr + (uint32_t)c, a);
                tmem_ld32(taddr + (uint32_t)(c + 32), bb);
                tmem_ld_wait();
                int m = kNone;
#pragma unroll
                for (int j = 0; j < 32; j += 2) m = max3(m, a[j], a[j + 1]);
#pragma unroll
                for (int j = 0; j < 32; j += 2) m = max3(m, bb[j], bb[j + 1]);
                // Rare per lane, not per warp (32 lanes x 64 values): when any lane needs it the whole warp takes the exact top
                // two of the batch WITHOUT branches (the raw accumulators of a row are all distinct and ordered like
                // (distance, index)), then the lanes that asked push them; rows come in ascending index order, so a new
                // entry must be strictly better than what is already there.
                if (__any_sync(0xffffffffu, (m >> kStepShift) > h2)) {
                    int b1 = kNone, b2 = kNone;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { b2 = max(b2, min(b1, a[j])); b1 = max(b1, a[j]); }
#pragma unroll
                    for (int j = 0; j < 32; ++j) { b2 = max(b2, min(b1, bb[j])); b1 = max(b1, bb[j]); }
                    const int g1 = b1 >> kStepShift, g2 = b2 >> kStepShift;
                    if (g1 > h2 && g1 >= -128) top2_push(g1, pi, 255 - ((b1 & 511) - iq), h1, p1, t1, h2, p2, t2);
                    if (g2 > h2 && g2 >= -128) top2_push(g2, pi, 255 - ((b2 & 511) - iq), h1, p1, t1, h2, p2, t2);
                }
                __syncwarp();
            }
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * b);
        }
        const int q = split * kQRows + q_local;
        if (q < A.nq) {
            unsigned long long k1 = ~0ull, k2 = ~0ull;
            if (h1 != kNone) k1 = ((unsigned long long)(128 - h1) << 40) | (unsigned long long)(A.desc_id_base + ((long long)group + (long long)p1 * n_groups) * kPairRows + t1);
            if (h2 != kNone) k2 = ((unsigned long long)(128 - h2) << 40) | (unsigned long long)(A.desc_id_base + ((long long)group + (long long)p2 * n_groups) * kPairRows + t2);
            A.partial[(size_t)group * A.nq + q] = make_ulonglong2(k1, k2);
        }
    }
done:
    __syncwarp();
    fence_before();
    __syncthreads();
    if (warp == kMma) tmem_free(tm, 512);
    if (tid == 0 && s_abort) *A.status = 1;
}

// cross-check + count of one keyframe (whole CTA): merges the per-split column results, then
// score[kf] = #{q : colbest[t*(q)].q == q, Ham <= tau}.  s_col: 4096 words, s_cnt: 1 int of shared memory.
__device__ __forceinline__ void finalize_keyframe(int kf, const long long* __restrict__ kf_off, int nq, long long n_desc,
                                                  const uint32_t* __restrict__ row_best, const uint32_t* __restrict__ col_best,
                                                  int tau, int* __restrict__ scores, uint32_t* s_col, int* s_cnt) {
    const long long r0 = kf_off[kf];
    const int n_t = (int)(kf_off[kf + 1] - r0);
    const int splits = (nq + kQRows - 1) / kQRows;
    if (threadIdx.x == 0) *s_cnt = 0;
    for (int t = threadIdx.x; t < n_t; t += blockDim.x) {
        uint32_t m = 0xffffffffu;
        for (int s = 0; s < splits; ++s) m = min(m, __ldcs(col_best + (size_t)s * (size_t)n_desc + (size_t)(r0 + t)));
        s_col[t] = m;
    }
    __syncthreads();
    int cnt = 0;
    if (n_t > 0)
        for (int q = threadIdx.x; q < nq; q += blockDim.x) {
            const uint32_t rb = __ldcs(row_best + (size_t)kf * kMaxQueries + q);
            const int ham = (int)(rb >> 16), t = (int)(rb & 0xffffu);
            if (ham <= tau && (int)(s_col[t] & 0xffffu) == q) ++cnt;
        }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(s_cnt, cnt);
    __syncthreads();
    if (threadIdx.x == 0) scores[kf] = *s_cnt;
    __syncthreads();
}

}  // namespace tc
}  // namespace pslam
