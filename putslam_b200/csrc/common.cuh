// common.cuh -- shared device helpers for the sm_100a kernels of putslam_b200.
// Compiled with -fmad=false -prec-div=true -prec-sqrt=true: every float/double operation is a
// single IEEE-754 rounding, which is what the CPU reference build does (no FMA on baseline x86-64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PSLAM_SM_COUNT_HINT 148

// ----------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol, sm_90+).  A frame is a chain of short, latency-bound kernels on
// one stream; launched with launch_chained(), kernel k+1 is scheduled while kernel k still runs and parks in
// chain_begin() until k has completed and its writes are visible, so the launch latency (2-3 us per boundary) is
// hidden.  Every chained kernel calls chain_begin() as its first statement in every thread: nothing is read or
// written before the predecessor is done, and "k+1 complete" implies "k complete" along the whole chain.
// After a kernel that was not launched this way (or a copy) the wait returns immediately.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void chain_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // let the next kernel in the stream be scheduled
    asm volatile("griddepcontrol.wait;" ::: "memory");                // predecessor complete, memory visible
}
// A frame's chain can also be RECORDED instead of launched: while a ChainRecorder is active on the calling thread every
// launch_chained() appends {function, grid, block, shared memory, a copy of the arguments} to it, and the owner replays the
// list as one CUDA graph (ctx.cu: one cudaGraphLaunch per frame instead of one launch call per kernel; the executable graph
// is kept while the chain's shape -- functions, grids, blocks -- stays the same and only the node arguments are refreshed).
#ifndef __CUDACC_RTC__
#include <cstring>
#include <vector>
struct ChainRecorder {
    struct Node {
        const void* func; dim3 grid, block; size_t smem;
        std::vector<unsigned char> blob;       // argument values, each at an 16-byte aligned offset
        std::vector<size_t> offs;
    };
    std::vector<Node> nodes;
    size_t used = 0;                           // nodes of the current recording (the vector's entries are reused)
    bool active = false;
    Node& next() {
        if (used == nodes.size()) nodes.emplace_back();
        Node& n = nodes[used++];
        n.offs.clear();
        return n;
    }
};
extern thread_local ChainRecorder* g_chain_recorder;   // defined in ctx.cu
template <typename T>
inline void chain_record_arg(ChainRecorder::Node& n, size_t& at, const T& v) {
    at = (at + 15) & ~(size_t)15;
    if (n.blob.size() < at + sizeof(T)) n.blob.resize(at + sizeof(T) + 256);
    std::memcpy(n.blob.data() + at, &v, sizeof(T));
    n.offs.push_back(at);
    at += sizeof(T);
}
#endif
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                  Args&&... args) {
    if (g_chain_recorder && g_chain_recorder->active) {
        ChainRecorder::Node& n = g_chain_recorder->next();
        n.func = (const void*)kernel; n.grid = grid; n.block = block; n.smem = smem;
        size_t at = 0;
        (void)at;
        int dummy[] = {0, (chain_record_arg<KArgs>(n, at, static_cast<KArgs>(args)), 0)...};
        (void)dummy;
        return cudaSuccess;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk, SASS: UBLKCP).  Descriptor tiles are contiguous runs of
// 32-byte rows, so the 1-D bulk form is the natural TMA shape: no tensor map needed.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra.uni WAIT_DONE;\n\t"
        "bra.uni WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ----------------------------------------------------------------------------------------------
// 256-bit Hamming distance.
//   Algorithmic definition (cv::norm NORM_HAMMING): sum over 8 words of popc(q ^ t) = 8 XOR + 8 POPC.
//   POPC issues at a quarter of the LOP3 rate, so the 8 XORed words are first compressed with four
//   3:2 carry-save adders (2 LOP3 each): weight-1 {s2, x7}, weight-2 {s3}, weight-4 {c3}
//   => 16 LOP3 + 4 POPC, result identical bit for bit.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lop3_xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t lop3_maj(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}

// Packed reduction key:  dist << 22 | train index (12 bits) << 10 | query offset (10 bits).
// Within one query row the query field is constant, within one train column the train field is
// constant, so ONE unsigned min over this key is simultaneously the first-argmin over train indices
// (row reduction) and the first-argmin over query indices (column reduction) -- lowest index wins ties,
// exactly like OpenCV.  `low` carries the two index fields; the popcount weights are folded into the
// shifts of an IMAD chain so the adds stay off the LOP3 pipe.
constexpr int kKeyQBits = 10;
constexpr int kKeyTBits = 12;
constexpr int kKeyDShift = kKeyQBits + kKeyTBits;  // 22
constexpr uint32_t kKeyInvalid = 0x80000000u;      // set in the query field of padding rows: never wins

__device__ __forceinline__ uint32_t ham256_key(const uint32_t (&q)[8], const uint4& ta, const uint4& tb, uint32_t low) {
    uint32_t x0 = q[0] ^ ta.x, x1 = q[1] ^ ta.y, x2 = q[2] ^ ta.z, x3 = q[3] ^ ta.w;
    uint32_t x4 = q[4] ^ tb.x, x5 = q[5] ^ tb.y, x6 = q[6] ^ tb.z, x7 = q[7] ^ tb.w;
    uint32_t s0 = lop3_xor3(x0, x1, x2), c0 = lop3_maj(x0, x1, x2);
    uint32_t s1 = lop3_xor3(x3, x4, x5), c1 = lop3_maj(x3, x4, x5);
    uint32_t s2 = lop3_xor3(s0, s1, x6), c2 = lop3_maj(s0, s1, x6);
    uint32_t s3 = lop3_xor3(c0, c1, c2), c3 = lop3_maj(c0, c1, c2);
    uint32_t r = low;
    r = __popc(s2) * (1u << kKeyDShift) + r;
    r = __popc(x7) * (1u << kKeyDShift) + r;
    r = __popc(s3) * (2u << kKeyDShift) + r;
    r = __popc(c3) * (4u << kKeyDShift) + r;
    return r;
}
// ----------------------------------------------------------------------------------------------
// The same distance on RE-ENCODED rows: 13 LOP3 + 4 POPC, result identical bit for bit.
//   A carry-save adder over three XORed words x0, x1, x2 needs sum = x0^x1^x2 and carry = maj(x0, x1, x2).  XOR is
//   linear, so sum = (q0^q1^q2) ^ (t0^t1^t2): if every descriptor row stores e2 = w0^w1^w2 in place of w2, the sum
//   is ONE xor of two stored words and x2 itself is never formed; the carry is a 3-input function of (x0, x1, sum)
//   because x2 = sum^x0^x1: maj(a, b, a^b^c) = LUT 0xD4.  The same holds one level up: s2 = s0^s1^x6 =
//   (q0^..^q6) ^ (t0^..^t6), and c2 = maj(s0, s1, x6) = LUT_0xD4(s0, s1, s2).  Encoded row (an invertible linear map
//   of the original 8 words, so still 32 bytes):
//       e = ( w0, w1, w0^w1^w2, w3, w4, w3^w4^w5, w0^w1^w2^w3^w4^w5^w6, w7 )
//   per pair: x0 x1 s0 | c0 | x3 x4 s1 | c1 | s2 | c2 | x7 | s3 c3  = 13 LOP3 (16 before), 4 POPC, weights 1,1,2,4.
//   The resident keyframe map is encoded once when descriptors are appended; a query is encoded when it is loaded
//   into registers.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lop3_maj_sum(uint32_t a, uint32_t b, uint32_t sum) {   // maj(a, b, a^b^sum)
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(a), "r"(b), "r"(sum));
    return r;
}
__host__ __device__ __forceinline__ void ham256_encode(uint32_t (&w)[8]) {
    const uint32_t e2 = w[0] ^ w[1] ^ w[2], e5 = w[3] ^ w[4] ^ w[5];
    w[6] = e2 ^ e5 ^ w[6];
    w[2] = e2;
    w[5] = e5;
}
__device__ __forceinline__ uint32_t ham256_key_enc(const uint32_t (&q)[8], const uint4& ta, const uint4& tb, uint32_t low) {
    const uint32_t x0 = q[0] ^ ta.x, x1 = q[1] ^ ta.y, s0 = q[2] ^ ta.z;
    const uint32_t x3 = q[3] ^ ta.w, x4 = q[4] ^ tb.x, s1 = q[5] ^ tb.y;
    const uint32_t s2 = q[6] ^ tb.z, x7 = q[7] ^ tb.w;
    const uint32_t c0 = lop3_maj_sum(x0, x1, s0);
    const uint32_t c1 = lop3_maj_sum(x3, x4, s1);
    const uint32_t c2 = lop3_maj_sum(s0, s1, s2);
    const uint32_t s3 = lop3_xor3(c0, c1, c2), c3 = lop3_maj(c0, c1, c2);
    uint32_t r = low;
    r = __popc(s2) * (1u << kKeyDShift) + r;
    r = __popc(x7) * (1u << kKeyDShift) + r;
    r = __popc(s3) * (2u << kKeyDShift) + r;
    r = __popc(c3) * (4u << kKeyDShift) + r;
    return r;
}
template <bool ENC>
__device__ __forceinline__ uint32_t ham256_key_t(const uint32_t (&q)[8], const uint4& ta, const uint4& tb, uint32_t low) {
    return ENC ? ham256_key_enc(q, ta, tb, low) : ham256_key(q, ta, tb, low);
}
__device__ __forceinline__ uint32_t key_dist(uint32_t k) { return k >> kKeyDShift; }
// QB = bits of the query field; the train field takes the remaining kKeyDShift - QB bits (10/12 by default,
// 11/11 for query sets of up to 2048 descriptors)
template <int QB = kKeyQBits>
__device__ __forceinline__ uint32_t key_tidx(uint32_t k) { return (k >> QB) & ((1u << (kKeyDShift - QB)) - 1u); }
template <int QB = kKeyQBits>
__device__ __forceinline__ uint32_t key_qoff(uint32_t k) { return k & ((1u << QB) - 1u); }

__device__ __forceinline__ uint32_t ham256(const uint32_t (&q)[8], const uint4& ta, const uint4& tb) {
    return ham256_key(q, ta, tb, 0u) >> kKeyDShift;
}

// saturating-subtract "Hamming" of Matcher::matchXYZ (reference src/Matcher/matcher.cpp:719-721):
// popcount of per-byte max(a - b, 0).
__device__ __forceinline__ uint32_t satsub_popc32(uint32_t a, uint32_t b) {
    return __popc(__vsubus4(a, b));
}

// Logical CTA index for grid-wide ordered compactions with look-back: the order in which the CTAs of this launch arrive
// here, not blockIdx.x.  A CTA only ever waits on lower logical indices, i.e. on CTAs that are already running, so the
// look-back cannot hang when other streams or processes keep part of the grid from being co-resident or the hardware hands
// out block indices in another order.  The ticket word is {epoch : 32 | arrivals : 32}: the first CTA of a launch (a new
// epoch) restarts the count, so nothing has to be reset between launches.
__device__ __forceinline__ unsigned int take_cta_ticket(unsigned long long* word, unsigned int epoch) {
    // two contention-free atomics instead of a compare-and-swap loop (which serialises the CTAs of a launch that arrive
    // together): epochs only grow (ctx.cu next_epoch clears the words when the 32-bit epoch wraps), so a 64-bit max moves
    // the word to {epoch, 0} exactly once per launch, and the add hands out 0, 1, 2, ...
    atomicMax(word, (unsigned long long)epoch << 32);
    return (unsigned int)(atomicAdd(word, 1ull) & 0xffffffffull);
}

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t warp_add_u32(uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }
