// ctx.cu -- the C ABI of libpslam_b200.so (include/pslam_b200.h): context, staging, stage entry
// points, fused per-frame pipelines, the resident loop-closure database and the NCCL glue.
//
// Data movement pattern of every host-pointer entry point: caller buffers -> one pinned arena ->
// ONE cudaMemcpyAsync H2D -> kernels on the ctx stream -> ONE cudaMemcpyAsync D2H of the output arena
// -> one stream synchronise -> caller buffers.  No CPU implementation of any stage exists in this
// library: if the device or a launch fails the call returns an error.
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <mutex>
#include <algorithm>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "geometry.cuh"
#include "kernels.h"

using namespace pslam;

namespace {

struct DevBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
};
struct HostBuf {
    uint8_t* p = nullptr;
    size_t cap = 0;
};

struct Arena {
    size_t off = 0;
    size_t take(size_t bytes) {
        const size_t o = off;
        off = (off + bytes + 255) & ~(size_t)255;
        return o;
    }
};

// NCCL is resolved at run time (dlopen) so that the library has no link-time dependency on it and
// shares whatever libnccl.so.2 the host process already loaded.
typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_ncclGetUniqueId)(nccl_uid*);
typedef int (*fn_ncclCommInitRank)(void**, int, nccl_uid, int);
typedef int (*fn_ncclCommDestroy)(void*);
typedef int (*fn_ncclAllGather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_ncclBroadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_ncclGetErrorString)(int);
struct NcclApi;
void nccl_load(NcclApi* out);
struct NcclApi {
    void* handle = nullptr;
    fn_ncclGetUniqueId GetUniqueId = nullptr;
    fn_ncclCommInitRank CommInitRank = nullptr;
    fn_ncclCommDestroy CommDestroy = nullptr;
    fn_ncclAllGather AllGather = nullptr;
    fn_ncclBroadcast Broadcast = nullptr;
    fn_ncclGetErrorString GetErrorString = nullptr;
    bool ok = false;
};
NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { nccl_load(&api); });
    return &api;
}
void nccl_load(NcclApi* out) {
    NcclApi& api = *out;
    const char* names[] = {"libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) return;
    api.GetUniqueId = (fn_ncclGetUniqueId)dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (fn_ncclCommInitRank)dlsym(api.handle, "ncclCommInitRank");
    api.CommDestroy = (fn_ncclCommDestroy)dlsym(api.handle, "ncclCommDestroy");
    api.AllGather = (fn_ncclAllGather)dlsym(api.handle, "ncclAllGather");
    api.Broadcast = (fn_ncclBroadcast)dlsym(api.handle, "ncclBroadcast");
    api.GetErrorString = (fn_ncclGetErrorString)dlsym(api.handle, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.Broadcast;
}
constexpr int kNcclUint8 = 1, kNcclInt32 = 2;
// below this many keyframes per GPU the sweep uses 128-row tiles as work units (see hamming.cu, split form)

// what a *_resident re-run needs to replay a pipeline on the buffers already in HBM
struct F2MState {
    bool valid = false;
    // device-side level prediction (pslam_frame_to_map_features): raw attributes -> map_xyz / levels work buffers
    bool device_levels = false;
    const double* map_xyz_d; const int* map_oct; const double* map_det; const int* cur_oct; const double* cur_det;
    float* map_xyz_w; int* map_level_w; int* cur_level_w;
    const float* map_xyz; const uint8_t* map_desc; const int* map_level; int M;
    const float* cur_xyz; const uint8_t* cur_desc; const int* cur_level; int N;
    float sq_radius_f; double ratio; int mode; int cap;
    int* count; int* best; void* cache; int* gout;
    RansacDeviceParams rp; RansacWorkspace ws;
};
struct F2FState {
    bool valid = false;
    const uint8_t* prev_desc; const float* prev_xyz; int n_prev;
    const uint8_t* cur_desc; const float* cur_uv; int n_cur;
    const uint16_t* depth; int W, H, stride; pslam_camera cam; int undistort; double scale;
    float* cur_xyz; float* cur_uv_und; double* det_dist;
    uint32_t* rowmin; uint32_t* colmin; int* mout; int cap;
    RansacDeviceParams rp; RansacWorkspace ws;
};

}  // namespace

struct pslam_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    char err[512] = {0};
    uint64_t launches = 0;
    HostBuf h_in, h_out;
    // caller memory page-locked through pslam_host_register: copied to the device straight from where it lies
    struct PinnedRange { const uint8_t* p; size_t bytes; };
    std::vector<PinnedRange> pinned;
    double stamps[8] = {0};   // host-side phase times (us) of the last fused frame call
    // fused frame chains replayed as CUDA graphs (one per chain kind: frame-to-map, frame-to-frame)
    ChainRecorder recorder;
    struct ChainGraph {
        cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
        std::vector<cudaGraphNode_t> nodes;
        struct Sig { const void* func; dim3 grid, block; size_t smem; };
        std::vector<Sig> sig;
    } graph_f2m, graph_f2f;
    int use_graphs = 1;       // PSLAM_GRAPHS=0 in the environment: plain launches
    DevBuf d_in, d_out, d_work;
    // loop-closure database
    uint8_t* d_db = nullptr;
    int64_t db_cap = 0, db_n = 0;
    int64_t* d_kf_off = nullptr;
    std::vector<int64_t> h_kf_off;  // host mirror (n_kf + 1)
    // split (tile-granular) sweep for small maps: prefix tile counts + scratch
    DevBuf d_tile_start, d_split;
    int n_tiles = 0;
    bool tiles_dirty = true;
    int max_kf_desc = 0;        // largest keyframe appended so far
    int lc_work_unit = 0;       // 0 fused-tail forms (tensor-core sweep when it applies, else the range form), 1 keyframes, 2 tiles (round-1 forms), 3 range form only
    DevBuf d_tc_row, d_tc_col;  // tensor-core form: per-query results per keyframe | per-split per-target results
    int* d_tc_status = nullptr; // non-zero after a barrier time-out inside the tensor-core sweep
    bool lc_tensor = true;      // PSLAM_LC_TENSOR=0 keeps the popcount kernels
    bool lc_last_tensor = false;
    DevBuf d_range;             // range form: row pieces | column pieces of the keyframes cut by CTA ranges
    int* d_kf_done = nullptr;   // per-keyframe tile counters of those keyframes (zero between launches)
    unsigned int* d_cta_done = nullptr;
    // peer exchange of the sharded sweep (CUDA IPC over NVLink); p2p false -> NCCL all-gather + merge kernel
    int* d_xchg = nullptr;
    int* peer_xchg[64] = {nullptr};
    bool p2p = false;
    uint32_t lc_epoch = 0, lc_qepoch = 0;
    bool lc_query_in_xchg = false;   // the resident query of this rank was pushed into the exchange buffer by the root
    int kf_cap = 0, n_kf = 0, kf_id_base = 0;
    long long desc_id_base = 0;
    DevBuf d_knn;  // V2 sweep scratch: per-CTA partials | merged keys | gathered keys | idx | dist
    int* d_scores = nullptr;
    uint8_t* d_lc_query = nullptr;
    int lc_nq = 0;
    int* d_lc_pairs = nullptr;  // local k pairs | gathered world*k pairs | merged k pairs
    bool lc_configured = false;
    cudaEvent_t ev_sweep0 = nullptr, ev_sweep1 = nullptr;  // bracket the sweep kernel of the last query
    // NCCL
    void* comm = nullptr;
    int rank = 0, world = 1;
    int stop_rule = 0;          // 0 reference RANSAC rule, 1 USAC standard stopping
    double usac_conf = 0.99;
    // last RANSAC (for pslam_ransac_last_counts)
    int* d_last_counts = nullptr;
    int last_H = 0;
    F2MState f2m;
    F2FState f2f;
    // resident feature map (pslam_map_*): SoA, slot = caller's feature index
    double* d_map_xyz = nullptr; uint8_t* d_map_desc = nullptr; int* d_map_oct = nullptr; double* d_map_det = nullptr;
    float* d_map_axis = nullptr;
    int map_cap = 0, map_n = 0;
    // ORB descriptor path: resize coefficient tables + sampling pattern, cached per image size / level count
    DevBuf d_orb_tab;
    int orb_W = 0, orb_H = 0, orb_levels = 0;
    bool orb_constants = false;
    // the frame of the last pslam_orb_detect / pslam_orb_describe, as uploaded (tight rows): pslam_orb_describe with a
    // NULL image works on it, so that detect -> describe of one frame uploads it once
    DevBuf d_orb_frame;
    int orbf_W = 0, orbf_H = 0, orbf_ch = 0;
    bool orbf_valid = false;
    // KLT: the pyramids of the two frames of the last pslam_klt_* call; klt_cur = the buffer holding its current frame
    // (the previous frame of the next call), -1 = none
    DevBuf d_klt_pyr[2];
    int klt_cur = -1, klt_W = 0, klt_H = 0, klt_cn = 0, klt_levels = 0;
    // map_prepare_kernel's per-CTA counts (stamped with prep_epoch, so they are never reset)
    unsigned long long* d_prep_counts = nullptr;
    unsigned int prep_epoch = 0;
};

namespace {

int fail(pslam_ctx* c, int code, const char* fmt, ...) {
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(ctx, PSLAM_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

// Stamped per-CTA sums of the grid-wide ordered compactions (mapprep.cu, guided.cu): slots [0, sm) belong to
// map_prepare_kernel, [sm, 3 sm) to guided_emit_kernel.  Zeroed once; every launch gets a fresh non-zero epoch.
int next_epoch(pslam_ctx* ctx, unsigned int* epoch);
// look-back slot arrays, each preceded by its ticket word (common.cuh take_cta_ticket): [t][prep: sm][t][emit: 2 sm]
unsigned long long* prep_slots(pslam_ctx* ctx) { return ctx->d_prep_counts + 1; }
unsigned long long* emit_slots(pslam_ctx* ctx) { return ctx->d_prep_counts + 2 + (ctx->sm_count > 0 ? ctx->sm_count : 1); }

int ensure_dev(pslam_ctx* ctx, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return PSLAM_OK;
    size_t want = bytes + bytes / 4 + 4096;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
    }
    CK(cudaMalloc((void**)&b.p, want));
    b.cap = want;
    // buffers moved: resident replays no longer point at valid inputs
    ctx->f2m.valid = false;
    ctx->f2f.valid = false;
    return PSLAM_OK;
}
int ensure_host(pslam_ctx* ctx, HostBuf& b, size_t bytes) {
    // every entry point that stages inputs passes through here with the input arena before it overwrites d_in / d_out /
    // d_work: whatever resident frame chain pointed into them is stale from now on (the frame entry points set their own
    // flag again once their inputs are in place)
    if (&b == &ctx->h_in) { ctx->f2m.valid = false; ctx->f2f.valid = false; }
    if (bytes <= b.cap) return PSLAM_OK;
    size_t want = bytes + bytes / 4 + 4096;
    if (b.p) {
        CK(cudaStreamSynchronize(ctx->stream));
        CK(cudaFreeHost(b.p));
        b.p = nullptr; b.cap = 0;
    }
    CK(cudaMallocHost((void**)&b.p, want));
    b.cap = want;
    // the result buffer moved: the resident chains write into it from the device (RansacWorkspace::out_host)
    ctx->f2m.valid = false;
    ctx->f2f.valid = false;
    return PSLAM_OK;
}
#define TRY(x)                     \
    do {                           \
        int r__ = (x);             \
        if (r__ != PSLAM_OK) return r__; \
    } while (0)

int next_epoch(pslam_ctx* ctx, unsigned int* epoch) {
    if (!ctx->d_prep_counts) {
        const size_t bytes = sizeof(unsigned long long) * (3 * (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 1) + 2);
        CK(cudaMalloc((void**)&ctx->d_prep_counts, bytes));
        CK(cudaMemsetAsync(ctx->d_prep_counts, 0, bytes, ctx->stream));
    }
    if (++ctx->prep_epoch == 0) {   // 0 is the value of a never-written slot; the ticket words need growing epochs
        ctx->prep_epoch = 1;
        const size_t bytes = sizeof(unsigned long long) * (3 * (size_t)(ctx->sm_count > 0 ? ctx->sm_count : 1) + 2);
        CK(cudaMemsetAsync(ctx->d_prep_counts, 0, bytes, ctx->stream));
    }
    *epoch = ctx->prep_epoch;
    return PSLAM_OK;
}


float float_at_least(double v) {  // smallest float >= v, so that (double)f < v  <=>  f < result
    float f = (float)v;
    if ((double)f < v) f = nextafterf(f, INFINITY);
    return f;
}

float sq_threshold(float thr_f) {  // smallest float T such that sqrtf(T) >= thr_f (sqrtf is correctly rounded, monotone)
    if (!(thr_f > 0.f)) return 0.f;          // sqrtf(s) < thr_f <= 0 never holds; s < 0 never holds for a sum of squares
    if (isinf(thr_f)) return INFINITY;
    float t = thr_f * thr_f;
    if (isinf(t)) t = 3.402823466e+38f;
    while (t > 0.f && sqrtf(t) >= thr_f) t = nextafterf(t, 0.f);
    while (sqrtf(t) < thr_f) t = nextafterf(t, INFINITY);
    return t;
}

// computeRANSACIteration (reference src/TransformEst/RANSAC.cpp:457-461) with the host libm; the conversion to int
// saturates (the reference's is undefined behaviour out of range -- DESIGN.md, deliberate divergences)
int ransac_iterations_host(double w) {
    const double v = log(1 - 0.98) / log(1 - pow(w, 3.0));
    if (v != v) return (int)0x80000000;
    if (v >= 2147483648.0) return 0x7fffffff;
    if (v <= -2147483649.0) return (int)0x80000000;
    return (int)v;
}

int make_ransac_params(pslam_ctx* ctx, const pslam_ransac_params* p, uint64_t seed, int num_hyp, RansacDeviceParams& o) {
    if (!p) return fail(ctx, PSLAM_ERR_ARG, "ransac params is NULL");
    if (p->used_pairs != 3) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "used_pairs must be 3 (got %d)", p->used_pairs);
    if (p->error_version == 3)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "MAHALANOBIS_ERROR is dead code in the reference (RANSAC.cpp:301-303)");
    if (p->error_version < 0 || p->error_version > 4) return fail(ctx, PSLAM_ERR_ARG, "bad error_version %d", p->error_version);
    if (num_hyp < 0 || num_hyp > (1 << 20)) return fail(ctx, PSLAM_ERR_ARG, "num_hyp out of range");
    o.error_version = p->error_version;
    o.thr_euclid = p->inlier_threshold_euclidean;
    o.thr_euclid_f = float_at_least(p->inlier_threshold_euclidean);
    o.sq_thr_euclid_f = sq_threshold(o.thr_euclid_f);
    o.thr_reproj = p->inlier_threshold_reprojection;
    o.min_inlier_ratio = p->minimal_inlier_ratio_threshold;
    o.min_matches = p->minimal_number_of_matches;
    o.fx = p->fx; o.fy = p->fy; o.cx = p->cx; o.cy = p->cy;
    o.seed_lo = (uint32_t)seed; o.seed_hi = (uint32_t)(seed >> 32);
    o.num_hyp = num_hyp;
    o.stop_rule = ctx->stop_rule;
    o.usac_conf = ctx->usac_conf;
    o.iters_min_ratio = ransac_iterations_host(p->minimal_inlier_ratio_threshold);
    return PSLAM_OK;
}

// carve a RANSAC workspace out of the work arena
struct RansacLayout {
    size_t pts, keep, nfil, counts, models, result;
    int m_cap, h_cap;
};
RansacLayout plan_ransac(Arena& A, int m_cap, const RansacDeviceParams& rp) {
    RansacLayout L;
    L.m_cap = m_cap > 0 ? m_cap : 1;
    L.h_cap = ransac_hypothesis_budget(rp);
    L.pts = A.take(sizeof(float) * 6 * (size_t)L.m_cap);
    L.keep = A.take(sizeof(int) * 2 * (size_t)L.m_cap);
    L.nfil = A.take(sizeof(int) * 4);
    L.counts = A.take(sizeof(int) * (size_t)L.h_cap);
    L.models = A.take(sizeof(float) * 12 * (size_t)L.h_cap);
    return L;
}
RansacWorkspace bind_ransac(const RansacLayout& L, uint8_t* work, int* result) {
    RansacWorkspace ws;
    ws.pts = reinterpret_cast<float*>(work + L.pts);
    ws.keep = reinterpret_cast<int*>(work + L.keep);
    ws.n_filtered = reinterpret_cast<int*>(work + L.nfil);
    ws.counts = reinterpret_cast<int*>(work + L.counts);
    ws.models = reinterpret_cast<float*>(work + L.models);
    ws.result = result;
    ws.m_cap = L.m_cap;
    ws.h_cap = L.h_cap;
    return ws;
}

double point_inlier_ratio(const int* inl_t, int n_inl, const int* all_t, int n_all) {
    int mx = -1;
    for (int i = 0; i < n_all; ++i) if (all_t[i] > mx) mx = all_t[i];
    for (int i = 0; i < n_inl; ++i) if (inl_t[i] > mx) mx = inl_t[i];
    std::vector<uint8_t> seen((size_t)(mx + 1), 0);
    int a = 0, b = 0;
    for (int i = 0; i < n_all; ++i) if (!seen[all_t[i]]) { seen[all_t[i]] = 1; ++b; }
    std::fill(seen.begin(), seen.end(), 0);
    for (int i = 0; i < n_inl; ++i) if (!seen[inl_t[i]]) { seen[inl_t[i]] = 1; ++a; }
    return (double)a / (double)b;  // 0/0 -> NaN exactly like the reference's double(0)/double(0)
}

void unpack_ransac_result(const int* res, float* T_out, int* inl_out, int* n_inl, double* best, int* used, int* nfil) {
    const int n = res[0];
    double ratio = 0.0;
    memcpy(&ratio, res + 20, sizeof(double));
    if (n_inl) *n_inl = n;
    if (used) {
        *used = res[1];
        if (res[1] < 0) {   // adaptive reference rule: the loop bound after the last improvement, finished here (kernels.h)
            const int b = ransac_iterations_host(ratio);
            int bound = res[24] < b ? res[24] : b;
            if (bound > res[26]) bound = res[26];   // hypotheses actually scored
            *used = res[25] + 1 > bound ? res[25] + 1 : bound;
        }
    }
    if (nfil) *nfil = res[2];
    if (T_out) memcpy(T_out, res + 4, sizeof(float) * 16);
    if (best) *best = ratio;
    if (inl_out && n > 0) memcpy(inl_out, res + kRansacHdrInts, sizeof(int) * (size_t)n);
}

}  // namespace

// =============================================================================================
// ---- frame chains as CUDA graphs -----------------------------------------------------------------------
thread_local ChainRecorder* g_chain_recorder = nullptr;

static void chain_graph_free(pslam_ctx::ChainGraph& g) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.graph) cudaGraphDestroy(g.graph);
    g.exec = nullptr; g.graph = nullptr; g.nodes.clear(); g.sig.clear();
}
// Replays the kernels recorded in ctx->recorder as one graph launch.  The executable graph is rebuilt when the chain's
// shape (functions, grid / block sizes, shared memory) changes; otherwise only the node arguments are refreshed.
static cudaError_t chain_graph_submit(pslam_ctx* ctx, pslam_ctx::ChainGraph& g) {
    ChainRecorder& R = ctx->recorder;
    const size_t n = R.used;
    if (n == 0) return cudaSuccess;
    std::vector<void*> argv;
    auto params_of = [&](const ChainRecorder::Node& nd, cudaKernelNodeParams& p) {
        argv.resize(nd.offs.size());
        for (size_t a = 0; a < nd.offs.size(); ++a) argv[a] = (void*)(nd.blob.data() + nd.offs[a]);
        p = cudaKernelNodeParams();
        p.func = const_cast<void*>(nd.func); p.gridDim = nd.grid; p.blockDim = nd.block;
        p.sharedMemBytes = (unsigned int)nd.smem; p.kernelParams = argv.data(); p.extra = nullptr;
    };
    bool same = g.exec && g.sig.size() == n;
    for (size_t i = 0; same && i < n; ++i) {
        const auto& a = g.sig[i]; const auto& b = R.nodes[i];
        same = a.func == b.func && a.smem == b.smem && a.grid.x == b.grid.x && a.grid.y == b.grid.y && a.grid.z == b.grid.z &&
               a.block.x == b.block.x && a.block.y == b.block.y && a.block.z == b.block.z;
    }
    cudaError_t e;
    if (!same) {
        chain_graph_free(g);
        if ((e = cudaGraphCreate(&g.graph, 0)) != cudaSuccess) return e;
        g.nodes.resize(n);
        // programmatic edges (the graph form of programmatic dependent launch): node k+1 is scheduled when node k executes
        // griddepcontrol.launch_dependents (first statement of every chained kernel) and parks in griddepcontrol.wait until
        // k has completed -- the same overlap of launch latency the plain launches get from launch_chained()
        static const int pdl_edges = []() { const char* e = getenv("PSLAM_GRAPH_PDL"); return (e && e[0] == '1') ? 1 : 0; }();   // measured: plain edges are faster end to end
        for (int programmatic = pdl_edges; programmatic >= 0; --programmatic) {
            bool ok = true;
            g.sig.clear();
            for (size_t i = 0; i < n && ok; ++i) {
                cudaKernelNodeParams p;
                params_of(R.nodes[i], p);
                if (programmatic) {
                    ok = cudaGraphAddKernelNode(&g.nodes[i], g.graph, nullptr, 0, &p) == cudaSuccess;
                    if (ok && i) {
                        cudaGraphEdgeData ed = {};
                        ed.from_port = cudaGraphKernelNodePortProgrammatic; ed.to_port = 0; ed.type = cudaGraphDependencyTypeProgrammatic;
                        ok = cudaGraphAddDependencies_v2(g.graph, &g.nodes[i - 1], &g.nodes[i], &ed, 1) == cudaSuccess;
                    }
                } else {
                    ok = cudaGraphAddKernelNode(&g.nodes[i], g.graph, i ? &g.nodes[i - 1] : nullptr, i ? 1 : 0, &p) == cudaSuccess;
                }
                g.sig.push_back({R.nodes[i].func, R.nodes[i].grid, R.nodes[i].block, R.nodes[i].smem});
            }
            if (ok && cudaGraphInstantiate(&g.exec, g.graph, 0) == cudaSuccess) break;
            cudaGetLastError();
            if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
            cudaGraphDestroy(g.graph); g.graph = nullptr;
            if (!programmatic) return cudaErrorUnknown;
            if ((e = cudaGraphCreate(&g.graph, 0)) != cudaSuccess) return e;
        }
    } else {
        for (size_t i = 0; i < n; ++i) {
            cudaKernelNodeParams p;
            params_of(R.nodes[i], p);
            if ((e = cudaGraphExecKernelNodeSetParams(g.exec, g.nodes[i], &p)) != cudaSuccess) return e;
        }
    }
    return cudaGraphLaunch(g.exec, ctx->stream);
}
// run `enqueue` (a function that only issues launch_chained calls on the ctx stream) either directly or recorded + replayed
template <typename F>
static int chain_run(pslam_ctx* ctx, pslam_ctx::ChainGraph& g, F enqueue) {
    if (!ctx->use_graphs) return enqueue();
    ctx->recorder.used = 0;
    ctx->recorder.active = true;
    g_chain_recorder = &ctx->recorder;
    const int r = enqueue();
    ctx->recorder.active = false;
    g_chain_recorder = nullptr;
    if (r != PSLAM_OK) return r;
    const cudaError_t e = chain_graph_submit(ctx, g);
    if (e != cudaSuccess) {   // graphs unavailable for this chain: fall back to plain launches for good
        cudaGetLastError();
        chain_graph_free(g);
        ctx->use_graphs = 0;
        return enqueue();
    }
    return PSLAM_OK;
}


extern "C" {

int pslam_version(void) { return 100; }

int pslam_ctx_create(int device, pslam_ctx** out) {
    if (!out) return PSLAM_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return PSLAM_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return PSLAM_ERR_ARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return PSLAM_ERR_CUDA;
    if (prop.major != 10) return PSLAM_ERR_NO_DEVICE;  // the kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return PSLAM_ERR_CUDA;
    pslam_ctx* ctx = new pslam_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    { const char* g = getenv("PSLAM_GRAPHS"); ctx->use_graphs = !(g && g[0] == '0'); }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return PSLAM_ERR_CUDA;
    }
    *out = ctx;
    return PSLAM_OK;
}

static void lc_peer_teardown(pslam_ctx* ctx);
void pslam_ctx_destroy(pslam_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    lc_peer_teardown(ctx);
    chain_graph_free(ctx->graph_f2m); chain_graph_free(ctx->graph_f2f);
    if (ctx->comm && nccl_api()->ok) nccl_api()->CommDestroy(ctx->comm);
    for (const auto& r : ctx->pinned) cudaHostUnregister(const_cast<uint8_t*>(r.p));
    cudaGetLastError();
    cudaFree(ctx->d_in.p); cudaFree(ctx->d_out.p); cudaFree(ctx->d_work.p); cudaFree(ctx->d_knn.p);
    cudaFree(ctx->d_tile_start.p); cudaFree(ctx->d_split.p); cudaFree(ctx->d_range.p); cudaFree(ctx->d_tc_row.p); cudaFree(ctx->d_tc_col.p); cudaFree(ctx->d_tc_status); cudaFree(ctx->d_kf_done); cudaFree(ctx->d_cta_done);
    cudaFreeHost(ctx->h_in.p); cudaFreeHost(ctx->h_out.p);
    cudaFree(ctx->d_db); cudaFree(ctx->d_kf_off); cudaFree(ctx->d_scores); cudaFree(ctx->d_lc_query);
    cudaFree(ctx->d_lc_pairs);
    cudaFree(ctx->d_map_xyz); cudaFree(ctx->d_map_desc); cudaFree(ctx->d_map_oct); cudaFree(ctx->d_map_det);
    cudaFree(ctx->d_map_axis); cudaFree(ctx->d_prep_counts); cudaFree(ctx->d_orb_tab.p); cudaFree(ctx->d_orb_frame.p);
    cudaFree(ctx->d_klt_pyr[0].p); cudaFree(ctx->d_klt_pyr[1].p);
    if (ctx->ev_sweep0) cudaEventDestroy(ctx->ev_sweep0);
    if (ctx->ev_sweep1) cudaEventDestroy(ctx->ev_sweep1);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* pslam_last_error(const pslam_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
void* pslam_ctx_stream(pslam_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t pslam_kernel_launches(const pslam_ctx* ctx) { return ctx ? ctx->launches : 0; }
int pslam_sm_count(const pslam_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int pslam_ctx_sync(pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return PSLAM_OK;
}

void pslam_ransac_sample(uint64_t seed, uint32_t hyp, int m, int out3[3]) {
    int s[3];
    sample3((uint32_t)seed, (uint32_t)(seed >> 32), hyp, (uint32_t)m, s);
    out3[0] = s[0]; out3[1] = s[1]; out3[2] = s[2];
}

double pslam_point_inlier_ratio(const int* inlier_train, int n_inliers, const int* all_train, int n_all) {
    return point_inlier_ratio(inlier_train, n_inliers, all_train, n_all);
}

// ---- stage 1 ----------------------------------------------------------------------------------
int pslam_backproject(pslam_ctx* ctx, const float* uv, int n, const uint16_t* depth, int W, int H, int row_stride,
                      const pslam_camera* cam, int undistort, double depth_scale, float* uv_undist_out, float* xyz_out,
                      double* det_dist_out, double* cov_out, const pslam_cov_params* cov) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n < 0 || (n > 0 && (!uv || !xyz_out)) || !depth || !cam || W <= 0 || H <= 0 || row_stride < W)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_backproject: bad argument");
    if (cov_out && !cov) return fail(ctx, PSLAM_ERR_ARG, "cov_out needs cov params");
    if (n == 0) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_uv = in.take(sizeof(float) * 2 * (size_t)n);
    const size_t o_depth = in.take(sizeof(uint16_t) * (size_t)H * row_stride);
    const size_t o_xyz = out.take(sizeof(float) * 3 * (size_t)n);
    const size_t o_und = out.take(sizeof(float) * 2 * (size_t)n);
    const size_t o_dd = out.take(sizeof(double) * (size_t)n);
    const size_t o_cov = out.take(cov_out ? sizeof(double) * 9 * (size_t)n : 8);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    memcpy(ctx->h_in.p + o_uv, uv, sizeof(float) * 2 * (size_t)n);
    memcpy(ctx->h_in.p + o_depth, depth, sizeof(uint16_t) * ((size_t)(H - 1) * row_stride + (size_t)W));   // not past the last row of an ROI
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_backproject((const float*)(ctx->d_in.p + o_uv), n, (const uint16_t*)(ctx->d_in.p + o_depth), W, H,
                          row_stride, *cam, undistort, depth_scale, (float*)(ctx->d_out.p + o_und),
                          (float*)(ctx->d_out.p + o_xyz), (double*)(ctx->d_out.p + o_dd),
                          cov_out ? (double*)(ctx->d_out.p + o_cov) : nullptr, cov, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(xyz_out, ctx->h_out.p + o_xyz, sizeof(float) * 3 * (size_t)n);
    if (uv_undist_out) memcpy(uv_undist_out, ctx->h_out.p + o_und, sizeof(float) * 2 * (size_t)n);
    if (det_dist_out) memcpy(det_dist_out, ctx->h_out.p + o_dd, sizeof(double) * (size_t)n);
    if (cov_out) memcpy(cov_out, ctx->h_out.p + o_cov, sizeof(double) * 9 * (size_t)n);
    return PSLAM_OK;
}

int pslam_normal_uncertainty(pslam_ctx* ctx, const int* px, int n, const uint16_t* depth, int W, int H, int row_stride,
                             const pslam_camera* cam, double depth_scale, double scale_uncertainty_normal,
                             double* normals_out, double* cov_out, double* info_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n < 0 || (n > 0 && !px) || !depth || !cam || W <= 0 || H <= 0 || row_stride < W)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_normal_uncertainty: bad argument");
    if (n == 0) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_px = in.take(8 * (size_t)n), o_depth = in.take(sizeof(uint16_t) * (size_t)H * row_stride);
    const size_t o_n = out.take(24 * (size_t)n), o_cov = out.take(72 * (size_t)n), o_info = out.take(72 * (size_t)n);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    memcpy(ctx->h_in.p + o_px, px, 8 * (size_t)n);
    memcpy(ctx->h_in.p + o_depth, depth, sizeof(uint16_t) * ((size_t)(H - 1) * row_stride + (size_t)W));   // not past the last row of an ROI
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_normal_cov((const int*)(ctx->d_in.p + o_px), n, (const uint16_t*)(ctx->d_in.p + o_depth), W, H, row_stride, *cam,
                         depth_scale, scale_uncertainty_normal, (double*)(ctx->d_out.p + o_n), (double*)(ctx->d_out.p + o_cov),
                         info_out ? (double*)(ctx->d_out.p + o_info) : nullptr, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, info_out ? out.off : o_info, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (normals_out) memcpy(normals_out, ctx->h_out.p + o_n, 24 * (size_t)n);
    if (cov_out) memcpy(cov_out, ctx->h_out.p + o_cov, 72 * (size_t)n);
    if (info_out) memcpy(info_out, ctx->h_out.p + o_info, 72 * (size_t)n);
    return PSLAM_OK;
}

// The libm-sensitive diagonals of RGBD::computeRGBGradient (reference src/RGBD/RGBD.cpp:164-166), evaluated once with
// the host libm so that the device kernel truncates exactly like host code on this machine would.
static void gradient_diag_table(int* out) {
    for (int q = 0; q < 4; ++q) {
        volatile double gx = (q & 1) ? -1.0 : 1.0, gy = (q & 2) ? -1.0 : 1.0;   // volatile: no compile-time folding
        const double angle = atan2(gy, gx) + (M_PI / 2.0);
        out[4 * q + 0] = (int)(sqrt(2.0) * sin(angle));
        out[4 * q + 1] = (int)(sqrt(2.0) * cos(angle));
        out[4 * q + 2] = (int)(sqrt(2.0) * sin(angle + M_PI));
        out[4 * q + 3] = (int)(sqrt(2.0) * cos(angle + M_PI));
    }
}

int pslam_gradient_uncertainty(pslam_ctx* ctx, const int* px, int n, const uint8_t* rgb, int rgb_row_bytes,
                               const uint16_t* depth, int W, int H, int row_stride, const pslam_camera* cam,
                               double depth_scale, double scale_uncertainty_gradient, double* grad_out, double* cov_out,
                               double* info_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n < 0 || (n > 0 && !px) || !rgb || !depth || !cam || W <= 0 || H <= 0 || row_stride < W || rgb_row_bytes < 3 * W)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_gradient_uncertainty: bad argument");
    if (n == 0) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_px = in.take(8 * (size_t)n), o_depth = in.take(sizeof(uint16_t) * (size_t)H * row_stride);
    const size_t o_rgb = in.take((size_t)H * rgb_row_bytes);
    const size_t o_g = out.take(24 * (size_t)n), o_cov = out.take(72 * (size_t)n), o_info = out.take(72 * (size_t)n);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    memcpy(ctx->h_in.p + o_px, px, 8 * (size_t)n);
    memcpy(ctx->h_in.p + o_depth, depth, sizeof(uint16_t) * ((size_t)(H - 1) * row_stride + (size_t)W));   // not past the last row of an ROI
    memcpy(ctx->h_in.p + o_rgb, rgb, (size_t)(H - 1) * rgb_row_bytes + 3 * (size_t)W);
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int diag[16];
    gradient_diag_table(diag);
    int l = 0;
    CK(launch_gradient_cov((const int*)(ctx->d_in.p + o_px), n, (const uint8_t*)(ctx->d_in.p + o_rgb), rgb_row_bytes,
                           (const uint16_t*)(ctx->d_in.p + o_depth), W, H, row_stride, *cam, depth_scale,
                           scale_uncertainty_gradient, diag, (double*)(ctx->d_out.p + o_g),
                           (cov_out || !info_out) ? (double*)(ctx->d_out.p + o_cov) : nullptr,
                           info_out ? (double*)(ctx->d_out.p + o_info) : nullptr, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, info_out ? out.off : o_info, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (grad_out) memcpy(grad_out, ctx->h_out.p + o_g, 24 * (size_t)n);
    if (cov_out) memcpy(cov_out, ctx->h_out.p + o_cov, 72 * (size_t)n);
    if (info_out) memcpy(info_out, ctx->h_out.p + o_info, 72 * (size_t)n);
    return PSLAM_OK;
}

int pslam_information_matrices(pslam_ctx* ctx, const double* uvz, int n, const pslam_cov_params* cov, double* info_out,
                               double* cov_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n < 0 || !cov || (n > 0 && (!uvz || !info_out))) return fail(ctx, PSLAM_ERR_ARG, "pslam_information_matrices: bad argument");
    if (n == 0) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_in = in.take(24 * (size_t)n);
    const size_t o_info = out.take(72 * (size_t)n), o_cov = out.take(72 * (size_t)n);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    memcpy(ctx->h_in.p + o_in, uvz, 24 * (size_t)n);
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_information((const double*)(ctx->d_in.p + o_in), n, *cov, cov_out ? (double*)(ctx->d_out.p + o_cov) : nullptr,
                          (double*)(ctx->d_out.p + o_info), ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(info_out, ctx->h_out.p + o_info, 72 * (size_t)n);
    if (cov_out) memcpy(cov_out, ctx->h_out.p + o_cov, 72 * (size_t)n);
    return PSLAM_OK;
}

// ---- stage 2 ----------------------------------------------------------------------------------
int pslam_match_bf_mutual(pslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int desc_bytes,
                          int* out_query_idx, int* out_train_idx, float* out_distance, int* n_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || nq < 0 || nt < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_match_bf_mutual: bad argument");
    *n_out = 0;
    if (desc_bytes != PSLAM_DESC_BYTES) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "descriptor width %d (only 32)", desc_bytes);
    if (nq > PSLAM_MAX_BF_ROWS || nt > PSLAM_MAX_BF_ROWS)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "nq/nt above %d: use the loop-closure database API", PSLAM_MAX_BF_ROWS);
    if (nq == 0 || nt == 0) return PSLAM_OK;  // BFMatcher on an empty set yields no matches
    if (!query || !train || !out_query_idx || !out_train_idx || !out_distance)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_match_bf_mutual: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int cap = nq < nt ? nq : nt;
    Arena in, out, work;
    const size_t o_q = in.take((size_t)nq * 32), o_t = in.take((size_t)nt * 32);
    const size_t o_out = out.take(sizeof(int) * (1 + 3 * (size_t)cap));
    const size_t o_row = work.take(sizeof(uint32_t) * (size_t)nq), o_col = work.take(sizeof(uint32_t) * (size_t)nt);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    memcpy(ctx->h_in.p + o_q, query, (size_t)nq * 32);
    memcpy(ctx->h_in.p + o_t, train, (size_t)nt * 32);
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_bf_mutual(ctx->d_in.p + o_q, nq, ctx->d_in.p + o_t, nt, (uint32_t*)(ctx->d_work.p + o_row),
                        (uint32_t*)(ctx->d_work.p + o_col), (int*)(ctx->d_out.p + o_out), cap, ctx->sm_count,
                        ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int* r = (const int*)(ctx->h_out.p + o_out);
    const int n = r[0];
    memcpy(out_query_idx, r + 1, sizeof(int) * (size_t)n);
    memcpy(out_train_idx, r + 1 + cap, sizeof(int) * (size_t)n);
    memcpy(out_distance, r + 1 + 2 * cap, sizeof(float) * (size_t)n);
    *n_out = n;
    return PSLAM_OK;
}

int pslam_match_knn2(pslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt, int desc_bytes,
                     int* out_idx, float* out_dist) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (nq < 0 || nt < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_match_knn2: bad argument");
    if (desc_bytes != PSLAM_DESC_BYTES) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "descriptor width %d (only 32)", desc_bytes);
    if (nq > PSLAM_MAX_BF_ROWS || nt > PSLAM_MAX_BF_ROWS)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "nq/nt above %d: use pslam_lc_knn2", PSLAM_MAX_BF_ROWS);
    if (nq == 0) return PSLAM_OK;
    if (!query || (nt > 0 && !train) || !out_idx || !out_dist) return fail(ctx, PSLAM_ERR_ARG, "pslam_match_knn2: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int parts = nt > 0 ? knn2_parts(nq, nt, ctx->sm_count) : 1;
    Arena in, out, work;
    const size_t o_q = in.take((size_t)nq * 32), o_t = in.take((size_t)(nt > 0 ? nt : 1) * 32);
    const size_t o_idx = out.take(sizeof(int) * 2 * (size_t)nq), o_dist = out.take(sizeof(float) * 2 * (size_t)nq);
    const size_t o_part = work.take(sizeof(uint2) * (size_t)parts * nq);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    memcpy(ctx->h_in.p + o_q, query, (size_t)nq * 32);
    if (nt > 0) memcpy(ctx->h_in.p + o_t, train, (size_t)nt * 32);
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_knn2(ctx->d_in.p + o_q, nq, ctx->d_in.p + o_t, nt, (uint2*)(ctx->d_work.p + o_part),
                   (int*)(ctx->d_out.p + o_idx), (float*)(ctx->d_out.p + o_dist), ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(out_idx, ctx->h_out.p + o_idx, sizeof(int) * 2 * (size_t)nq);
    memcpy(out_dist, ctx->h_out.p + o_dist, sizeof(float) * 2 * (size_t)nq);
    return PSLAM_OK;
}

int pslam_match_guided_xyz(pslam_ctx* ctx, const float* map_xyz, const uint8_t* map_desc, const int* map_level, int M,
                           const float* cur_xyz, const uint8_t* cur_desc, const int* cur_level, int N, int desc_bytes,
                           double radius, double accept_ratio, int distance_mode, int* out_query_idx,
                           int* out_train_idx, float* out_distance, int cap, int* n_out, int* perfect_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || M < 0 || N < 0 || cap < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_match_guided_xyz: bad argument");
    *n_out = 0;
    if (perfect_out) *perfect_out = 0;
    if (desc_bytes != PSLAM_DESC_BYTES) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "descriptor width %d (only 32)", desc_bytes);
    if (distance_mode != 0 && distance_mode != 1) return fail(ctx, PSLAM_ERR_ARG, "distance_mode must be 0 or 1");
    if (N > 12000) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "N above 12000 current keypoints");
    if (M == 0 || N == 0) return PSLAM_OK;
    if (!map_xyz || !map_desc || !map_level || !cur_xyz || !cur_desc || !cur_level || (cap > 0 && (!out_query_idx || !out_train_idx || !out_distance)))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_match_guided_xyz: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int dcap = cap > 0 ? cap : 1;
    Arena in, out, work;
    const size_t o_mx = in.take(12 * (size_t)M), o_md = in.take(32 * (size_t)M), o_ml = in.take(4 * (size_t)M);
    const size_t o_cx = in.take(12 * (size_t)N), o_cd = in.take(32 * (size_t)N), o_cl = in.take(4 * (size_t)N);
    const size_t o_out = out.take(sizeof(int) * (2 + 3 * (size_t)dcap));
    const size_t o_cnt = work.take(sizeof(int) * (2 * (size_t)M + 1)), o_best = work.take(sizeof(int) * (size_t)M);
    const size_t o_cache = work.take(guided_cache_bytes(M));
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* h = ctx->h_in.p;
    memcpy(h + o_mx, map_xyz, 12 * (size_t)M); memcpy(h + o_md, map_desc, 32 * (size_t)M); memcpy(h + o_ml, map_level, 4 * (size_t)M);
    memcpy(h + o_cx, cur_xyz, 12 * (size_t)N); memcpy(h + o_cd, cur_desc, 32 * (size_t)N); memcpy(h + o_cl, cur_level, 4 * (size_t)N);
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d = ctx->d_in.p;
    int l = 0;
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    CK(launch_guided_match((const float*)(d + o_mx), d + o_md, (const int*)(d + o_ml), M, (const float*)(d + o_cx),
                           d + o_cd, (const int*)(d + o_cl), N, sq_threshold(float_at_least(radius)), accept_ratio,
                           distance_mode, (int*)(ctx->d_work.p + o_cnt), (int*)(ctx->d_work.p + o_best),
                           ctx->d_work.p + o_cache, (int*)(ctx->d_out.p + o_out), dcap, emit_slots(ctx), epoch,
                           ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int* r = (const int*)(ctx->h_out.p + o_out);
    const int total = r[0];
    const int n = total < cap ? total : cap;
    if (n > 0) {
        memcpy(out_query_idx, r + 2, sizeof(int) * (size_t)n);
        memcpy(out_train_idx, r + 2 + dcap, sizeof(int) * (size_t)n);
        memcpy(out_distance, r + 2 + 2 * dcap, sizeof(float) * (size_t)n);
    }
    *n_out = total;
    if (perfect_out) *perfect_out = r[1];
    if (total > cap) return fail(ctx, PSLAM_ERR_CAPACITY, "guided matching produced %d matches, capacity %d", total, cap);
    return PSLAM_OK;
}

// ---- stage 3 ----------------------------------------------------------------------------------
int pslam_ransac_estimate(pslam_ctx* ctx, const float* prev, int n_prev, const float* cur, int n_cur,
                          const int* match_query, const int* match_train, int m, const pslam_ransac_params* params,
                          uint64_t seed, int num_hyp, float* T_out, int* inlier_idx_out, int* n_inliers_out,
                          double* best_ratio_out, int* hyp_used_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!T_out || !n_inliers_out || n_prev < 0 || n_cur < 0 || m < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_ransac_estimate: bad argument");
    RansacDeviceParams rp;
    TRY(make_ransac_params(ctx, params, seed, num_hyp, rp));
    for (int i = 0; i < 16; ++i) T_out[i] = (i % 5 == 0) ? 1.f : 0.f;
    *n_inliers_out = 0;
    if (best_ratio_out) *best_ratio_out = 0.0;
    if (hyp_used_out) *hyp_used_out = 0;
    if (m == 0 || m < params->minimal_number_of_matches) return PSLAM_OK;  // RANSAC.cpp:77-80 before any device work
    if (!prev || !cur || !match_query || !match_train || !inlier_idx_out) return fail(ctx, PSLAM_ERR_ARG, "pslam_ransac_estimate: null buffer");
    for (int k = 0; k < m; ++k)
        if (match_query[k] < 0 || match_query[k] >= n_prev || match_train[k] < 0 || match_train[k] >= n_cur)
            return fail(ctx, PSLAM_ERR_ARG, "match %d indexes outside the point sets", k);
    CK(cudaSetDevice(ctx->device));
    Arena in, out, work;
    const size_t o_p = in.take(12 * (size_t)n_prev), o_c = in.take(12 * (size_t)n_cur);
    const size_t o_mq = in.take(4 * (size_t)m), o_mt = in.take(4 * (size_t)m);
    const size_t o_res = out.take(sizeof(int) * ransac_result_ints(m));
    RansacLayout L = plan_ransac(work, m, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* h = ctx->h_in.p;
    memcpy(h + o_p, prev, 12 * (size_t)n_prev); memcpy(h + o_c, cur, 12 * (size_t)n_cur);
    memcpy(h + o_mq, match_query, 4 * (size_t)m); memcpy(h + o_mt, match_train, 4 * (size_t)m);
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d = ctx->d_in.p;
    RansacWorkspace ws = bind_ransac(L, ctx->d_work.p, (int*)(ctx->d_out.p + o_res));
    int l = 0;
    CK(cudaMemsetAsync(ws.counts, 0xff, sizeof(int) * (size_t)ws.h_cap, ctx->stream));
    CK(launch_ransac((const float*)(d + o_p), (const float*)(d + o_c), (const int*)(d + o_mq), (const int*)(d + o_mt),
                     nullptr, m, rp, ws, ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    ctx->d_last_counts = ws.counts;
    ctx->last_H = ws.h_cap;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unpack_ransac_result((const int*)(ctx->h_out.p + o_res), T_out, inlier_idx_out, n_inliers_out, best_ratio_out,
                         hyp_used_out, nullptr);
    return PSLAM_OK;
}

int pslam_ransac_set_stopping(pslam_ctx* ctx, int rule, double confidence) {
    if (!ctx) return PSLAM_ERR_ARG;
    if ((rule != 0 && rule != 1) || !(confidence > 0.0 && confidence < 1.0)) return fail(ctx, PSLAM_ERR_ARG, "bad stopping rule");
    ctx->stop_rule = rule;
    ctx->usac_conf = confidence;
    return PSLAM_OK;
}

int pslam_ransac_last_counts(pslam_ctx* ctx, int* counts_out, int cap, int* n_out) {
    if (!ctx || !counts_out || !n_out) return PSLAM_ERR_ARG;
    *n_out = 0;
    if (!ctx->d_last_counts) return fail(ctx, PSLAM_ERR_ARG, "no RANSAC run on this ctx yet");
    CK(cudaSetDevice(ctx->device));
    const int n = cap < ctx->last_H ? cap : ctx->last_H;
    CK(cudaMemcpyAsync(counts_out, ctx->d_last_counts, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *n_out = n;
    return PSLAM_OK;
}

int pslam_kabsch_batch(pslam_ctx* ctx, const double* A, const double* B, const int* offsets, int batch, double* T_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!offsets || !T_out))) return fail(ctx, PSLAM_ERR_ARG, "pslam_kabsch_batch: bad argument");
    if (batch == 0) return PSLAM_OK;
    const int total = offsets[batch];
    if (total < 0 || (total > 0 && (!A || !B))) return fail(ctx, PSLAM_ERR_ARG, "pslam_kabsch_batch: bad point buffers");
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_a = in.take(24 * (size_t)(total > 0 ? total : 1)), o_b = in.take(24 * (size_t)(total > 0 ? total : 1));
    const size_t o_off = in.take(4 * (size_t)(batch + 1));
    const size_t o_t = out.take(sizeof(double) * 12 * (size_t)batch);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    uint8_t* h = ctx->h_in.p;
    if (total > 0) { memcpy(h + o_a, A, 24 * (size_t)total); memcpy(h + o_b, B, 24 * (size_t)total); }
    memcpy(h + o_off, offsets, 4 * (size_t)(batch + 1));
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    CK(launch_kabsch_batch((const double*)(ctx->d_in.p + o_a), (const double*)(ctx->d_in.p + o_b),
                           (const int*)(ctx->d_in.p + o_off), batch, (double*)(ctx->d_out.p + o_t), ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // device result is row-major 3x4; the ABI is column-major (Eigen layout)
    const double* r = (const double*)(ctx->h_out.p + o_t);
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) T_out[12 * (size_t)b + 3 * j + i] = r[12 * (size_t)b + 4 * i + j];
    return PSLAM_OK;
}

int pslam_transform_uncertainty_batch(pslam_ctx* ctx, const double* A, const double* B, const double* covA, const double* covB,
                                      const int* offsets, const double* T, int batch, int parametrization, double* U_out,
                                      int* ok_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (batch < 0 || (batch > 0 && (!offsets || !T || !U_out)) || (parametrization != PSLAM_UNCERTAINTY_EULER &&
                                                                    parametrization != PSLAM_UNCERTAINTY_QUATERNION))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_transform_uncertainty_batch: bad argument");
    if (batch == 0) return PSLAM_OK;
    const int total = offsets[batch];
    for (int b = 0; b < batch; ++b)
        if (offsets[b] < 0 || offsets[b + 1] < offsets[b]) return fail(ctx, PSLAM_ERR_ARG, "pslam_transform_uncertainty_batch: offsets not ascending");
    if (total > 0 && (!A || !B || !covA || !covB)) return fail(ctx, PSLAM_ERR_ARG, "pslam_transform_uncertainty_batch: null point buffer");
    CK(cudaSetDevice(ctx->device));
    const size_t tot = (size_t)(total > 0 ? total : 1);
    Arena in, out;
    const size_t o_a = in.take(24 * tot), o_b = in.take(24 * tot), o_ca = in.take(72 * tot), o_cb = in.take(72 * tot);
    const size_t o_off = in.take(4 * (size_t)(batch + 1)), o_t = in.take(96 * (size_t)batch);
    const size_t o_u = out.take(288 * (size_t)batch), o_ok = out.take(4 * (size_t)batch);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    ctx->f2m.valid = false; ctx->f2f.valid = false;   // the arenas are reused
    uint8_t* h = ctx->h_in.p;
    if (total > 0) {
        memcpy(h + o_a, A, 24 * (size_t)total); memcpy(h + o_b, B, 24 * (size_t)total);
        memcpy(h + o_ca, covA, 72 * (size_t)total); memcpy(h + o_cb, covB, 72 * (size_t)total);
    }
    memcpy(h + o_off, offsets, 4 * (size_t)(batch + 1));
    memcpy(h + o_t, T, 96 * (size_t)batch);
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    const uint8_t* d = ctx->d_in.p;
    int l = 0;
    CK(launch_uncertainty_batch((const double*)(d + o_a), (const double*)(d + o_b), (const double*)(d + o_ca),
                                (const double*)(d + o_cb), (const int*)(d + o_off), (const double*)(d + o_t), batch,
                                parametrization, (double*)(ctx->d_out.p + o_u), (int*)(ctx->d_out.p + o_ok), ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // device result is row-major; the ABI is column-major (Eigen's Mat66 layout)
    const double* u = (const double*)(ctx->h_out.p + o_u);
    for (int b = 0; b < batch; ++b)
        for (int i = 0; i < 6; ++i)
            for (int j = 0; j < 6; ++j) U_out[36 * (size_t)b + 6 * j + i] = u[36 * (size_t)b + 6 * i + j];
    if (ok_out) memcpy(ok_out, ctx->h_out.p + o_ok, 4 * (size_t)batch);
    return PSLAM_OK;
}

// ---- caller-pinned input buffers ---------------------------------------------------------------------
int pslam_host_register(pslam_ctx* ctx, const void* ptr, size_t bytes) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!ptr || bytes == 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_host_register: empty range");
    CK(cudaSetDevice(ctx->device));
    for (const auto& r : ctx->pinned)
        if (r.p == (const uint8_t*)ptr && r.bytes == bytes) return PSLAM_OK;
    const cudaError_t e = cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, PSLAM_ERR_CUDA, "cudaHostRegister: %s", cudaGetErrorString(e)); }
    ctx->pinned.push_back({(const uint8_t*)ptr, bytes});
    return PSLAM_OK;
}
int pslam_host_unregister(pslam_ctx* ctx, const void* ptr) {
    if (!ctx) return PSLAM_ERR_ARG;
    for (size_t i = 0; i < ctx->pinned.size(); ++i)
        if (ctx->pinned[i].p == (const uint8_t*)ptr) {
            cudaSetDevice(ctx->device);
            cudaStreamSynchronize(ctx->stream);          // no copy from the range may still be in flight
            cudaHostUnregister(const_cast<void*>(ptr));
            cudaGetLastError();
            ctx->pinned.erase(ctx->pinned.begin() + (long)i);
            return PSLAM_OK;
        }
    return fail(ctx, PSLAM_ERR_ARG, "pslam_host_unregister: range was not registered on this ctx");
}
static bool host_is_pinned(const pslam_ctx* ctx, const void* p, size_t bytes) {
    const uint8_t* q = (const uint8_t*)p;
    for (const auto& r : ctx->pinned)
        if (q >= r.p && q + bytes <= r.p + r.bytes) return true;
    return false;
}
// Host -> device placement of the inputs of one call.  Inputs inside a registered range go with their own asynchronous
// copy from where they lie; the others are packed into the pinned arena and leave with one copy.
struct Uploader {
    pslam_ctx* ctx; uint8_t* d; uint8_t* h;
    size_t lo = (size_t)-1, hi = 0;
    cudaError_t err = cudaSuccess;
    void put(size_t off, const void* src, size_t bytes) {
        if (bytes == 0 || err != cudaSuccess) return;
        if (host_is_pinned(ctx, src, bytes)) { err = cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, ctx->stream); return; }
        memcpy(h + off, src, bytes);
        if (off < lo) lo = off;
        if (off + bytes > hi) hi = off + bytes;
    }
    cudaError_t flush() {
        if (err != cudaSuccess) return err;
        if (hi > lo) return cudaMemcpyAsync(d + lo, h + lo, hi - lo, cudaMemcpyHostToDevice, ctx->stream);
        return cudaSuccess;
    }
};

// host-side phase stamps of the last fused call (pslam_debug_host_stamps): nanoseconds since the call began
struct Stamp {
    pslam_ctx* c; std::chrono::steady_clock::time_point t0;
    explicit Stamp(pslam_ctx* ctx) : c(ctx), t0(std::chrono::steady_clock::now()) { for (double& v : c->stamps) v = 0; }
    void mark(int k) { c->stamps[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count(); }
};
int pslam_debug_host_stamps(const pslam_ctx* ctx, double out8[8]) {
    if (!ctx || !out8) return PSLAM_ERR_ARG;
    for (int i = 0; i < 8; ++i) out8[i] = ctx->stamps[i];
    return PSLAM_OK;
}

// ---- fused pipelines ----------------------------------------------------------------------------
// host-libm tables for the device level prediction (see guided.cu)
struct HostLevelTables {
    double pow_tab[16];
    int lvl_tab[16];
    double log_sf;
    HostLevelTables() {
        const double scaleFactor = 1.2;
        log_sf = std::log(scaleFactor);
        for (int k = 0; k < 16; ++k) {
            pow_tab[k] = pow(scaleFactor, k);
            lvl_tab[k] = (int)std::ceil(std::log(pow_tab[k]) / log_sf);
        }
    }
};
static const HostLevelTables& level_tables() {
    static const HostLevelTables t;
    return t;
}

static int enqueue_f2m_launches(pslam_ctx* ctx);
static int enqueue_f2m(pslam_ctx* ctx) {   // the chain of one frame-to-map frame, replayed as one CUDA graph
    return chain_run(ctx, ctx->graph_f2m, [&]() { return enqueue_f2m_launches(ctx); });
}
static int enqueue_f2m_launches(pslam_ctx* ctx) {
    F2MState& s = ctx->f2m;
    int l = 0;
    if (s.device_levels) {
        const HostLevelTables& t = level_tables();
        CK(launch_predict_levels(s.map_xyz_d, s.map_oct, s.map_det, s.M, s.cur_xyz, s.cur_oct, s.cur_det, s.N, t.pow_tab,
                                 t.lvl_tab, t.log_sf, s.map_xyz_w, s.map_level_w, s.cur_level_w, ctx->stream, &l));
    }
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    CK(launch_guided_match(s.map_xyz, s.map_desc, s.map_level, s.M, s.cur_xyz, s.cur_desc, s.cur_level, s.N, s.sq_radius_f,
                           s.ratio, s.mode, s.count, s.best, s.cache, s.gout, s.cap, emit_slots(ctx), epoch, ctx->sm_count,
                           ctx->stream, &l));
    CK(launch_ransac(s.map_xyz, s.cur_xyz, s.gout + 2, s.gout + 2 + s.cap, s.gout, 0, s.rp, s.ws, ctx->sm_count,
                     ctx->stream, &l));
    ctx->launches += l;
    return PSLAM_OK;
}

struct F2MRaw {   // raw feature attributes for the device-level path (NULL map_xyz_d = levels given by the caller)
    const double* map_xyz_d = nullptr; const int* map_oct = nullptr; const double* map_det = nullptr;
    const int* cur_oct = nullptr; const double* cur_det = nullptr;
};
static int frame_to_map_core(pslam_ctx* ctx, const float* map_xyz, const uint8_t* map_desc, const int* map_level, int M,
                             const float* cur_xyz, const uint8_t* cur_desc, const int* cur_level, int N, double radius,
                             double accept_ratio, int distance_mode, const pslam_ransac_params* params, uint64_t seed,
                             int num_hyp, int match_cap, int* match_query_out, int* match_train_out, float* match_dist_out,
                             int* inlier_idx_out, pslam_frame_result* result, const F2MRaw& raw) {
    const bool dev_levels = raw.map_xyz_d != nullptr;
    if (!ctx) return PSLAM_ERR_ARG;
    if (!result || M < 0 || N < 0 || match_cap <= 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_map: bad argument");
    Stamp stamp(ctx);
    memset(result, 0, sizeof(*result));
    for (int i = 0; i < 16; ++i) result->T[i] = (i % 5 == 0) ? 1.f : 0.f;
    result->inlier_ratio = -1.0;
    RansacDeviceParams rp;
    TRY(make_ransac_params(ctx, params, seed, num_hyp, rp));
    if (distance_mode != 0 && distance_mode != 1) return fail(ctx, PSLAM_ERR_ARG, "distance_mode must be 0 or 1");
    if (N > 12000) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "N above 12000 current keypoints");
    if (M == 0 || N == 0) return PSLAM_OK;  // no matches -> -1.0 (matcher.cpp:755)
    if ((!dev_levels && (!map_xyz || !map_level || !cur_level)) ||
        (dev_levels && (!raw.map_oct || !raw.map_det || !raw.cur_oct || !raw.cur_det)) || !map_desc || !cur_xyz || !cur_desc ||
        !match_query_out || !match_train_out || !match_dist_out || !inlier_idx_out)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_map: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int cap = match_cap;
    Arena in, out, work;
    // host-level path: map xyz (float) + levels are inputs; device-level path: map xyz (double), octaves, detDists
    const size_t o_mx = in.take((dev_levels ? 24 : 12) * (size_t)M), o_md = in.take(32 * (size_t)M), o_ml = in.take(4 * (size_t)M);
    const size_t o_mdet = in.take(dev_levels ? 8 * (size_t)M : 8);
    const size_t o_cx = in.take(12 * (size_t)N), o_cd = in.take(32 * (size_t)N), o_cl = in.take(4 * (size_t)N);
    const size_t o_cdet = in.take(dev_levels ? 8 * (size_t)N : 8);
    const size_t o_wx = work.take(12 * (size_t)M), o_wml = work.take(4 * (size_t)M), o_wcl = work.take(4 * (size_t)N);
    const size_t o_g = out.take(sizeof(int) * (2 + 3 * (size_t)cap));
    const size_t o_res = out.take(sizeof(int) * ransac_result_ints(cap));
    const size_t o_cnt = work.take(sizeof(int) * (2 * (size_t)M + 1)), o_best = work.take(sizeof(int) * (size_t)M);
    const size_t o_cache = work.take(guided_cache_bytes(M));
    RansacLayout L = plan_ransac(work, cap, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* d = ctx->d_in.p;
    Uploader up{ctx, d, ctx->h_in.p};
    // the current frame's (small) arrays sit together at the end of the arena so that they leave with one copy; the map
    // side goes straight from the caller's memory when that is registered (pslam_host_register)
    up.put(o_cx, cur_xyz, 12 * (size_t)N); up.put(o_cd, cur_desc, 32 * (size_t)N);
    if (dev_levels) { up.put(o_cl, raw.cur_oct, 4 * (size_t)N); up.put(o_cdet, raw.cur_det, 8 * (size_t)N); }
    else up.put(o_cl, cur_level, 4 * (size_t)N);
    up.put(o_md, map_desc, 32 * (size_t)M);
    if (dev_levels) {
        up.put(o_mx, raw.map_xyz_d, 24 * (size_t)M); up.put(o_ml, raw.map_oct, 4 * (size_t)M);
        up.put(o_mdet, raw.map_det, 8 * (size_t)M);
    } else {
        up.put(o_mx, map_xyz, 12 * (size_t)M); up.put(o_ml, map_level, 4 * (size_t)M);
    }
    stamp.mark(1);
    CK(up.flush());
    stamp.mark(2);
    F2MState& s = ctx->f2m;
    s.device_levels = dev_levels;
    s.M = M; s.N = N;
    s.map_desc = d + o_md; s.cur_xyz = (const float*)(d + o_cx); s.cur_desc = d + o_cd;
    if (dev_levels) {
        s.map_xyz_d = (const double*)(d + o_mx); s.map_oct = (const int*)(d + o_ml); s.map_det = (const double*)(d + o_mdet);
        s.cur_oct = (const int*)(d + o_cl); s.cur_det = (const double*)(d + o_cdet);
        s.map_xyz_w = (float*)(ctx->d_work.p + o_wx); s.map_level_w = (int*)(ctx->d_work.p + o_wml);
        s.cur_level_w = (int*)(ctx->d_work.p + o_wcl);
        s.map_xyz = s.map_xyz_w; s.map_level = s.map_level_w; s.cur_level = s.cur_level_w;
    } else {
        s.map_xyz = (const float*)(d + o_mx); s.map_level = (const int*)(d + o_ml); s.cur_level = (const int*)(d + o_cl);
    }
    s.sq_radius_f = sq_threshold(float_at_least(radius)); s.ratio = accept_ratio; s.mode = distance_mode; s.cap = cap;
    s.count = (int*)(ctx->d_work.p + o_cnt); s.best = (int*)(ctx->d_work.p + o_best); s.cache = ctx->d_work.p + o_cache;
    s.gout = (int*)(ctx->d_out.p + o_g);
    s.rp = rp; s.ws = bind_ransac(L, ctx->d_work.p, (int*)(ctx->d_out.p + o_res));
    // the last kernel of the chain writes the output arena into the page-locked result buffer itself: no copy-engine
    // transfer (and its scheduling latency) between the end of the chain and the host
    s.ws.out_host = ctx->h_out.p; s.ws.out_dev = ctx->d_out.p; s.ws.out_bytes = out.off;
    s.ws.out_cap = cap; s.ws.out_res_ints = (int)(o_res / sizeof(int));   // o_g is the start of the arena
    s.valid = true;
    TRY(enqueue_f2m(ctx));
    stamp.mark(3);
    ctx->d_last_counts = s.ws.counts; ctx->last_H = s.ws.h_cap;
    stamp.mark(4);
    CK(cudaStreamSynchronize(ctx->stream));
    stamp.mark(5);
    const int* g = (const int*)(ctx->h_out.p + o_g);
    const int total = g[0];
    const int n = total < cap ? total : cap;
    result->n_matches = total;
    if (n > 0) {
        memcpy(match_query_out, g + 2, 4 * (size_t)n);
        memcpy(match_train_out, g + 2 + cap, 4 * (size_t)n);
        memcpy(match_dist_out, g + 2 + 2 * cap, 4 * (size_t)n);
    }
    if (total > cap) return fail(ctx, PSLAM_ERR_CAPACITY, "guided matching produced %d matches, capacity %d", total, cap);
    if (total == 0) return PSLAM_OK;
    unpack_ransac_result((const int*)(ctx->h_out.p + o_res), result->T, inlier_idx_out, &result->n_inliers,
                         &result->best_ratio, &result->hyp_used, &result->n_filtered);
    std::vector<int> inl_t((size_t)result->n_inliers);
    for (int i = 0; i < result->n_inliers; ++i) inl_t[i] = match_train_out[inlier_idx_out[i]];
    result->inlier_ratio = point_inlier_ratio(inl_t.data(), result->n_inliers, match_train_out, n);
    stamp.mark(6);
    return PSLAM_OK;
}

int pslam_frame_to_map(pslam_ctx* ctx, const float* map_xyz, const uint8_t* map_desc, const int* map_level, int M,
                       const float* cur_xyz, const uint8_t* cur_desc, const int* cur_level, int N, double radius,
                       double accept_ratio, int distance_mode, const pslam_ransac_params* params, uint64_t seed,
                       int num_hyp, int match_cap, int* match_query_out, int* match_train_out, float* match_dist_out,
                       int* inlier_idx_out, pslam_frame_result* result) {
    if (!ctx) return PSLAM_ERR_ARG;
    return frame_to_map_core(ctx, map_xyz, map_desc, map_level, M, cur_xyz, cur_desc, cur_level, N, radius, accept_ratio,
                             distance_mode, params, seed, num_hyp, match_cap, match_query_out, match_train_out,
                             match_dist_out, inlier_idx_out, result, F2MRaw());
}

int pslam_frame_to_map_features(pslam_ctx* ctx, const double* map_xyz, const uint8_t* map_desc, const int* map_octave,
                                const double* map_det_dist, int M, const float* cur_xyz, const uint8_t* cur_desc,
                                const int* cur_octave, const double* cur_det_dist, int N, double radius, double accept_ratio,
                                int distance_mode, const pslam_ransac_params* params, uint64_t seed, int num_hyp,
                                int match_cap, int* match_query_out, int* match_train_out, float* match_dist_out,
                                int* inlier_idx_out, pslam_frame_result* result) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (M > 0 && !map_xyz) return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_map_features: null map positions");
    F2MRaw raw;
    raw.map_xyz_d = map_xyz; raw.map_oct = map_octave; raw.map_det = map_det_dist;
    raw.cur_oct = cur_octave; raw.cur_det = cur_det_dist;
    static const double dummy = 0.0;
    if (M == 0) raw.map_xyz_d = &dummy;   // keeps the "device levels" path selected; the core returns before any use
    return frame_to_map_core(ctx, nullptr, map_desc, nullptr, M, cur_xyz, cur_desc, nullptr, N, radius, accept_ratio,
                             distance_mode, params, seed, num_hyp, match_cap, match_query_out, match_train_out,
                             match_dist_out, inlier_idx_out, result, raw);
}

int pslam_frame_to_map_resident(pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!ctx->f2m.valid) return fail(ctx, PSLAM_ERR_ARG, "no resident frame-to-map inputs");
    CK(cudaSetDevice(ctx->device));
    return enqueue_f2m(ctx);
}

static int enqueue_f2f_launches(pslam_ctx* ctx);
static int enqueue_f2f(pslam_ctx* ctx) {
    return chain_run(ctx, ctx->graph_f2f, [&]() { return enqueue_f2f_launches(ctx); });
}
static int enqueue_f2f_launches(pslam_ctx* ctx) {
    F2FState& s = ctx->f2f;
    int l = 0;
    CK(launch_bf_mutual(s.prev_desc, s.n_prev, s.cur_desc, s.n_cur, s.rowmin, s.colmin, s.mout, s.cap, ctx->sm_count,
                        ctx->stream, &l));
    CK(launch_backproject(s.cur_uv, s.n_cur, s.depth, s.W, s.H, s.stride, s.cam, s.undistort, s.scale, s.cur_uv_und,
                          s.cur_xyz, s.det_dist, nullptr, nullptr, ctx->stream, &l));
    CK(launch_ransac(s.prev_xyz, s.cur_xyz, s.mout + 1, s.mout + 1 + s.cap, s.mout, 0, s.rp, s.ws, ctx->sm_count,
                     ctx->stream, &l));
    ctx->launches += l;
    return PSLAM_OK;
}

int pslam_frame_to_frame(pslam_ctx* ctx, const uint8_t* prev_desc, const float* prev_xyz, int n_prev,
                         const uint8_t* cur_desc, const float* cur_uv, int n_cur, const uint16_t* depth, int W, int H,
                         int row_stride, const pslam_camera* cam, int undistort, double depth_scale,
                         const pslam_ransac_params* params, uint64_t seed, int num_hyp, float* cur_xyz_out,
                         float* cur_uv_undist_out, double* cur_det_dist_out, int* match_query_out, int* match_train_out,
                         float* match_dist_out, int* inlier_idx_out, pslam_frame_result* result) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!result || n_prev < 0 || n_cur < 0 || !cam || !depth || W <= 0 || H <= 0 || row_stride < W)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_frame: bad argument");
    memset(result, 0, sizeof(*result));
    for (int i = 0; i < 16; ++i) result->T[i] = (i % 5 == 0) ? 1.f : 0.f;
    RansacDeviceParams rp;
    TRY(make_ransac_params(ctx, params, seed, num_hyp, rp));
    if (n_prev > PSLAM_MAX_BF_ROWS || n_cur > PSLAM_MAX_BF_ROWS) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "more than %d keypoints", PSLAM_MAX_BF_ROWS);
    if (n_cur == 0) return PSLAM_OK;
    if (!cur_desc || !cur_uv || !cur_xyz_out || (n_prev > 0 && (!prev_desc || !prev_xyz)) || !match_query_out ||
        !match_train_out || !match_dist_out || !inlier_idx_out)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_frame: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int np = n_prev > 0 ? n_prev : 1;
    const int cap = (n_prev < n_cur ? n_prev : n_cur) > 0 ? (n_prev < n_cur ? n_prev : n_cur) : 1;
    Arena in, out, work;
    const size_t o_pd = in.take(32 * (size_t)np), o_px = in.take(12 * (size_t)np);
    const size_t o_cd = in.take(32 * (size_t)n_cur), o_uv = in.take(8 * (size_t)n_cur);
    const size_t o_depth = in.take(sizeof(uint16_t) * (size_t)H * row_stride);
    const size_t o_xyz = out.take(12 * (size_t)n_cur), o_und = out.take(8 * (size_t)n_cur), o_dd = out.take(8 * (size_t)n_cur);
    const size_t o_m = out.take(sizeof(int) * (1 + 3 * (size_t)cap));
    const size_t o_res = out.take(sizeof(int) * ransac_result_ints(cap));
    const size_t o_row = work.take(4 * (size_t)np), o_col = work.take(4 * (size_t)n_cur);
    RansacLayout L = plan_ransac(work, cap, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* h = ctx->h_in.p;
    if (n_prev > 0) { memcpy(h + o_pd, prev_desc, 32 * (size_t)n_prev); memcpy(h + o_px, prev_xyz, 12 * (size_t)n_prev); }
    memcpy(h + o_cd, cur_desc, 32 * (size_t)n_cur); memcpy(h + o_uv, cur_uv, 8 * (size_t)n_cur);
    memcpy(h + o_depth, depth, sizeof(uint16_t) * ((size_t)(H - 1) * row_stride + (size_t)W));   // not past the last row of an ROI
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d = ctx->d_in.p;
    F2FState& s = ctx->f2f;
    s.prev_desc = d + o_pd; s.prev_xyz = (const float*)(d + o_px); s.n_prev = n_prev;
    s.cur_desc = d + o_cd; s.cur_uv = (const float*)(d + o_uv); s.n_cur = n_cur;
    s.depth = (const uint16_t*)(d + o_depth); s.W = W; s.H = H; s.stride = row_stride; s.cam = *cam;
    s.undistort = undistort; s.scale = depth_scale;
    s.cur_xyz = (float*)(ctx->d_out.p + o_xyz); s.cur_uv_und = (float*)(ctx->d_out.p + o_und);
    s.det_dist = (double*)(ctx->d_out.p + o_dd);
    s.rowmin = (uint32_t*)(ctx->d_work.p + o_row); s.colmin = (uint32_t*)(ctx->d_work.p + o_col);
    s.mout = (int*)(ctx->d_out.p + o_m); s.cap = cap;
    s.rp = rp; s.ws = bind_ransac(L, ctx->d_work.p, (int*)(ctx->d_out.p + o_res));
    s.valid = n_prev > 0;
    if (n_prev > 0) {
        TRY(enqueue_f2f(ctx));
        ctx->d_last_counts = s.ws.counts; ctx->last_H = s.ws.h_cap;
    } else {  // first frame: back-projection only (Matcher::detectInitFeatures, matcher.cpp:37-58)
        int l = 0;
        CK(cudaMemsetAsync(s.mout, 0, sizeof(int), ctx->stream));
        CK(launch_backproject(s.cur_uv, n_cur, s.depth, W, H, row_stride, *cam, undistort, depth_scale, s.cur_uv_und,
                              s.cur_xyz, s.det_dist, nullptr, nullptr, ctx->stream, &l));
        ctx->launches += l;
    }
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(cur_xyz_out, ctx->h_out.p + o_xyz, 12 * (size_t)n_cur);
    if (cur_uv_undist_out) memcpy(cur_uv_undist_out, ctx->h_out.p + o_und, 8 * (size_t)n_cur);
    if (cur_det_dist_out) memcpy(cur_det_dist_out, ctx->h_out.p + o_dd, 8 * (size_t)n_cur);
    const int* mo = (const int*)(ctx->h_out.p + o_m);
    const int n = mo[0];
    result->n_matches = n;
    if (n > 0) {
        memcpy(match_query_out, mo + 1, 4 * (size_t)n);
        memcpy(match_train_out, mo + 1 + cap, 4 * (size_t)n);
        memcpy(match_dist_out, mo + 1 + 2 * cap, 4 * (size_t)n);
    }
    if (n_prev == 0) return PSLAM_OK;
    unpack_ransac_result((const int*)(ctx->h_out.p + o_res), result->T, inlier_idx_out, &result->n_inliers,
                         &result->best_ratio, &result->hyp_used, &result->n_filtered);
    std::vector<int> inl_t((size_t)result->n_inliers);
    for (int i = 0; i < result->n_inliers; ++i) inl_t[i] = match_train_out[inlier_idx_out[i]];
    result->inlier_ratio = point_inlier_ratio(inl_t.data(), result->n_inliers, match_train_out, n);
    return PSLAM_OK;
}

int pslam_loop_closure_pair(pslam_ctx* ctx, const uint8_t* desc0, const float* xyz0, int n0, const uint8_t* desc1,
                            const float* xyz1, int n1, const pslam_ransac_params* params, uint64_t seed, int num_hyp,
                            int* match_query_out, int* match_train_out, float* match_dist_out, int* inlier_idx_out,
                            pslam_frame_result* result) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!result || n0 < 0 || n1 < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_loop_closure_pair: bad argument");
    memset(result, 0, sizeof(*result));
    for (int i = 0; i < 16; ++i) result->T[i] = (i % 5 == 0) ? 1.f : 0.f;
    RansacDeviceParams rp;
    TRY(make_ransac_params(ctx, params, seed, num_hyp, rp));
    if (n0 > PSLAM_MAX_BF_ROWS || n1 > PSLAM_MAX_BF_ROWS) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "more than %d features", PSLAM_MAX_BF_ROWS);
    if (n0 < 10 || n1 < 10) return PSLAM_OK;   // "Too few features" -> 0 (matcher.cpp:830-834)
    if (!desc0 || !xyz0 || !desc1 || !xyz1 || !match_query_out || !match_train_out || !match_dist_out || !inlier_idx_out)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_loop_closure_pair: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int cap = n0 < n1 ? n0 : n1;
    Arena in, out, work;
    const size_t o_d0 = in.take(32 * (size_t)n0), o_x0 = in.take(12 * (size_t)n0);
    const size_t o_d1 = in.take(32 * (size_t)n1), o_x1 = in.take(12 * (size_t)n1);
    const size_t o_m = out.take(sizeof(int) * (1 + 3 * (size_t)cap));
    const size_t o_res = out.take(sizeof(int) * ransac_result_ints(cap));
    const size_t o_row = work.take(4 * (size_t)n0), o_col = work.take(4 * (size_t)n1);
    RansacLayout L = plan_ransac(work, cap, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    ctx->f2m.valid = false; ctx->f2f.valid = false;   // the arenas are reused
    uint8_t* h = ctx->h_in.p;
    memcpy(h + o_d0, desc0, 32 * (size_t)n0); memcpy(h + o_x0, xyz0, 12 * (size_t)n0);
    memcpy(h + o_d1, desc1, 32 * (size_t)n1); memcpy(h + o_x1, xyz1, 12 * (size_t)n1);
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d = ctx->d_in.p;
    int* mout = (int*)(ctx->d_out.p + o_m);
    RansacWorkspace ws = bind_ransac(L, ctx->d_work.p, (int*)(ctx->d_out.p + o_res));
    int l = 0;
    CK(launch_bf_mutual(d + o_d0, n0, d + o_d1, n1, (uint32_t*)(ctx->d_work.p + o_row), (uint32_t*)(ctx->d_work.p + o_col),
                        mout, cap, ctx->sm_count, ctx->stream, &l));
    CK(launch_ransac((const float*)(d + o_x0), (const float*)(d + o_x1), mout + 1, mout + 1 + cap, mout, 0, rp, ws,
                     ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    ctx->d_last_counts = ws.counts; ctx->last_H = ws.h_cap;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int* mo = (const int*)(ctx->h_out.p + o_m);
    const int n = mo[0];
    result->n_matches = n;
    if (n <= 0) { result->inlier_ratio = -1.0; return PSLAM_OK; }   // matcher.cpp:838-839
    memcpy(match_query_out, mo + 1, 4 * (size_t)n);
    memcpy(match_train_out, mo + 1 + cap, 4 * (size_t)n);
    memcpy(match_dist_out, mo + 1 + 2 * cap, 4 * (size_t)n);
    unpack_ransac_result((const int*)(ctx->h_out.p + o_res), result->T, inlier_idx_out, &result->n_inliers,
                         &result->best_ratio, &result->hyp_used, &result->n_filtered);
    std::vector<int> inl_t((size_t)result->n_inliers);
    for (int i = 0; i < result->n_inliers; ++i) inl_t[i] = match_train_out[inlier_idx_out[i]];
    result->inlier_ratio = point_inlier_ratio(inl_t.data(), result->n_inliers, match_train_out, n);
    return PSLAM_OK;
}

int pslam_frame_to_frame_resident(pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!ctx->f2f.valid) return fail(ctx, PSLAM_ERR_ARG, "no resident frame-to-frame inputs");
    CK(cudaSetDevice(ctx->device));
    return enqueue_f2f(ctx);
}

// ---- map-side preparation -------------------------------------------------------------------------
int pslam_map_prepare(pslam_ctx* ctx, const double* map_xyz, const float* view_axis, int M, const double camera_pose[16],
                      const pslam_map_prepare_params* params, int* kept_idx, double* xyz_local, double* uv, double* angles,
                      int* n_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || M < 0 || !camera_pose || !params) return fail(ctx, PSLAM_ERR_ARG, "pslam_map_prepare: bad argument");
    *n_out = 0;
    if (M == 0) return PSLAM_OK;
    if (!map_xyz || !view_axis || !kept_idx || !xyz_local || !uv || !angles) return fail(ctx, PSLAM_ERR_ARG, "pslam_map_prepare: null buffer");
    CK(cudaSetDevice(ctx->device));
    Arena in, out;
    const size_t o_x = in.take(24 * (size_t)M), o_a = in.take(12 * (size_t)M);
    const size_t o_n = out.take(16), o_k = out.take(4 * (size_t)M), o_xl = out.take(24 * (size_t)M);
    const size_t o_uv = out.take(16 * (size_t)M), o_ang = out.take(8 * (size_t)M);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    memcpy(ctx->h_in.p + o_x, map_xyz, 24 * (size_t)M);
    memcpy(ctx->h_in.p + o_a, view_axis, 12 * (size_t)M);
    CK(cudaMemcpyAsync(ctx->d_in.p, ctx->h_in.p, in.off, cudaMemcpyHostToDevice, ctx->stream));
    int l = 0;
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    CK(launch_map_prepare((const double*)(ctx->d_in.p + o_x), (const float*)(ctx->d_in.p + o_a), M, camera_pose, params->fx,
                          params->fy, params->cx, params->cy, params->image_w, params->image_h, params->max_angle,
                          params->max_z, (int*)(ctx->d_out.p + o_k), (double*)(ctx->d_out.p + o_xl),
                          (double*)(ctx->d_out.p + o_uv), (double*)(ctx->d_out.p + o_ang), (int*)(ctx->d_out.p + o_n),
                          prep_slots(ctx), epoch, ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int n = *(const int*)(ctx->h_out.p + o_n);
    memcpy(kept_idx, ctx->h_out.p + o_k, 4 * (size_t)n);
    memcpy(xyz_local, ctx->h_out.p + o_xl, 24 * (size_t)n);
    memcpy(uv, ctx->h_out.p + o_uv, 16 * (size_t)n);
    memcpy(angles, ctx->h_out.p + o_ang, 8 * (size_t)n);
    *n_out = n;
    return PSLAM_OK;
}

// the caller's frame, packed to tight rows through the pinned staging area `stage`, into the context's resident frame
static int orb_upload_frame(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels, uint8_t* stage) {
    const size_t row = (size_t)channels * W, bytes = row * H;
    ctx->orbf_valid = false;
    TRY(ensure_dev(ctx, ctx->d_orb_frame, bytes));
    if ((size_t)row_bytes == row) memcpy(stage, image, bytes);
    else for (int y = 0; y < H; ++y) memcpy(stage + (size_t)y * row, image + (size_t)y * row_bytes, row);
    CK(cudaMemcpyAsync(ctx->d_orb_frame.p, stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->orbf_W = W; ctx->orbf_H = H; ctx->orbf_ch = channels; ctx->orbf_valid = true;
    return PSLAM_OK;
}

// ---- ORB descriptors ---------------------------------------------------------------------------------
// resize coefficient tables + sampling pattern + constants: rebuilt only when the image size or the level count changes
// (layout of d_orb_tab: 1024 bytes of pattern, then the tables)
static int orb_prepare_tables(pslam_ctx* ctx, int W, int H, int nlev, const OrbPlan& P) {
    if (ctx->orb_W == W && ctx->orb_H == H && ctx->orb_levels == nlev && ctx->orb_constants) return PSLAM_OK;
    Arena t;
    const size_t o_pat = t.take(1024), o_tab = t.take(4 * P.tab_ints);
    if (o_tab != 1024) return fail(ctx, PSLAM_ERR_CUDA, "orb table layout");
    TRY(ensure_dev(ctx, ctx->d_orb_tab, t.off));
    std::vector<int> tab(P.tab_ints > 0 ? P.tab_ints : 1);
    orb_fill_tables(P, tab.data());
    CK(orb_upload_constants(ctx->d_orb_tab.p + o_pat, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_orb_tab.p + o_tab, tab.data(), 4 * P.tab_ints, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // `tab` is pageable and dies here
    ctx->orb_W = W; ctx->orb_H = H; ctx->orb_levels = nlev; ctx->orb_constants = true;
    return PSLAM_OK;
}

int pslam_orb_describe(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels, const float* kp_xy,
                       const int* kp_octave, const float* kp_angle_deg, int n, int* order_out, int* n_out,
                       uint8_t* desc_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || n < 0 || W <= 0 || H <= 0 || (channels != 1 && channels != 3) || (image && row_bytes < channels * W))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_describe: bad argument");
    *n_out = 0;
    if (!image && !(ctx->orbf_valid && ctx->orbf_W == W && ctx->orbf_H == H && ctx->orbf_ch == channels))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_describe: no resident frame of %d x %d x %d (image == NULL needs a preceding "
                    "pslam_orb_detect / pslam_orb_describe of that frame on this context)", W, H, channels);
    if (n == 0) return PSLAM_OK;
    if (!kp_xy || !kp_octave || !kp_angle_deg || !order_out || !desc_out)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_describe: null buffer");
    // level count from all provided keypoints (cv::ORB::compute does this before it filters them)
    int nlev = 0;
    for (int i = 0; i < n; ++i) {
        if (kp_octave[i] < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_describe: negative octave");
        if (kp_octave[i] + 1 > nlev) nlev = kp_octave[i] + 1;
    }
    if (nlev > kOrbMaxLevels) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "pslam_orb_describe: octave above %d", kOrbMaxLevels - 1);
    OrbPlan P;
    orb_plan(W, H, nlev, &P);
    if (P.w[nlev - 1] <= 32 || P.h[nlev - 1] <= 32)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "pslam_orb_describe: pyramid level %d is %d x %d (must exceed 32 x 32)", nlev - 1,
                    P.w[nlev - 1], P.h[nlev - 1]);
    // KeyPointsFilter::runByImageBorder(edgeThreshold = 31) on the rounded position, then a stable regroup by octave
    std::vector<int> order((size_t)n);
    int start[kOrbMaxLevels + 1] = {0};
    auto inside = [&](int i) {
        const long ix = lrintf(kp_xy[2 * i]), iy = lrintf(kp_xy[2 * i + 1]);   // cvRound: half to even
        return ix >= 31 && ix < W - 31 && iy >= 31 && iy < H - 31;
    };
    for (int i = 0; i < n; ++i) order[(size_t)i] = inside(i) ? kp_octave[i] : -1;   // scratch: level or -1
    for (int i = 0; i < n; ++i) if (order[(size_t)i] >= 0) ++start[order[(size_t)i] + 1];
    for (int l = 0; l < nlev; ++l) start[l + 1] += start[l];
    const int m = start[nlev];
    {   // counting sort by level, input order kept inside a level
        std::vector<int> sorted((size_t)(m > 0 ? m : 1));
        int cur[kOrbMaxLevels];
        for (int l = 0; l < nlev; ++l) cur[l] = start[l];
        for (int i = 0; i < n; ++i) if (order[(size_t)i] >= 0) sorted[(size_t)cur[order[(size_t)i]]++] = i;
        order.swap(sorted);
    }
    *n_out = m;
    if (m == 0) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    Arena in, out, work;
    const size_t img_bytes = (size_t)channels * W * H;
    const size_t o_rec = in.take(20 * (size_t)m);
    const size_t o_desc = out.take(32 * (size_t)m);
    const size_t o_plain = work.take(P.plain_bytes), o_ext = work.take(P.ext_bytes), o_row = work.take(4 * P.row_floats);
    TRY(ensure_host(ctx, ctx->h_in, in.off + (image ? img_bytes + 256 : 0)));
    TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    TRY(orb_prepare_tables(ctx, W, H, nlev, P));
    const uint8_t* d_pat = ctx->d_orb_tab.p;
    const int* d_tab = (const int*)(ctx->d_orb_tab.p + 1024);
    // per-keypoint record: level pixel, level, a = (float)cos(angle), b = (float)sin(angle)
    uint8_t* h = ctx->h_in.p;
    int* rec = (int*)(h + o_rec);
    for (int k = 0; k < m; ++k) {
        const int i = order[(size_t)k], l = kp_octave[i];
        const float inv = 1.f / orb_level_scale(l);
        float angle = kp_angle_deg[i];
        angle *= (float)(3.141592653589793238462643383279502884 / 180.f);
        double sn, cs;
        sincos((double)angle, &sn, &cs);   // glibc: the same kernels as sin() and cos()
        const float a = (float)cs, b = (float)sn;
        rec[5 * k] = (int)lrintf(kp_xy[2 * i] * inv);
        rec[5 * k + 1] = (int)lrintf(kp_xy[2 * i + 1] * inv);
        rec[5 * k + 2] = l;
        memcpy(&rec[5 * k + 3], &a, 4); memcpy(&rec[5 * k + 4], &b, 4);
        order_out[k] = i;
    }
    uint8_t* d_plain = ctx->d_work.p + o_plain;
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    if (image) TRY(orb_upload_frame(ctx, image, W, H, row_bytes, channels, h + in.off));
    if (channels == 1)   // gray: the frame is level 0
        CK(cudaMemcpyAsync(d_plain, ctx->d_orb_frame.p, img_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    int l = 0;
    CK(launch_orb_describe(channels == 3 ? ctx->d_orb_frame.p : nullptr, W, H, 3 * W, P, d_plain, ctx->d_work.p + o_ext,
                           (float*)(ctx->d_work.p + o_row), d_tab, d_pat, (const int*)(ctx->d_in.p + o_rec), m,
                           ctx->d_out.p + o_desc, ctx->stream, &l));
    ctx->launches += l;
    ctx->f2m.valid = false; ctx->f2f.valid = false;   // the work arena of a previous frame call was reused
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, out.off, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(desc_out, ctx->h_out.p + o_desc, 32 * (size_t)m);
    return PSLAM_OK;
}

// KeyPointsFilter::retainBest (OpenCV features2d, keypoint.cpp) on the detector's working records: the n best by
// response, plus everything tied with the n-th; the order left behind is that of std::nth_element + std::partition,
// which is what OpenCV's own output order is made of
struct OrbKp {
    float x, y, response;   // level pixel, current response (FAST score, later Harris)
    float harris, angle;
};
static void orb_retain_best(std::vector<OrbKp>& k, int n) {
    if (n >= 0 && k.size() > (size_t)n) {
        if (n == 0) { k.clear(); return; }
        std::nth_element(k.begin(), k.begin() + n - 1, k.end(), [](const OrbKp& a, const OrbKp& b) { return a.response > b.response; });
        const float ambiguous = k[(size_t)n - 1].response;
        auto new_end = std::partition(k.begin() + n, k.end(), [ambiguous](const OrbKp& a) { return a.response >= ambiguous; });
        k.resize((size_t)(new_end - k.begin()));
    }
}

int pslam_orb_detect(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels, int colour_order,
                     int nfeatures, float* kp_xy, float* kp_size, float* kp_angle, float* kp_response, int* kp_octave, int cap,
                     int* n_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || !image || W <= 0 || H <= 0 || (channels != 1 && channels != 3) || row_bytes < channels * W || nfeatures < 0 ||
        cap < 0 || (colour_order != 0 && colour_order != 1))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_detect: bad argument");
    *n_out = 0;
    if (cap > 0 && (!kp_xy || !kp_size || !kp_angle || !kp_response || !kp_octave))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_orb_detect: null buffer");
    const int nlev = 8;
    OrbPlan P;
    orb_plan(W, H, nlev, &P);
    if (P.w[nlev - 1] < 1 || P.h[nlev - 1] < 1)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "pslam_orb_detect: image too small for an 8-level pyramid");
    CK(cudaSetDevice(ctx->device));
    // computeKeyPoints: nfeatures split over the levels by 1 / scaleFactor (float arithmetic, cvRound), rest on the last
    int per_level[8];
    {
        const float factor = (float)(1.0 / (double)1.2f);
        float nd = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlev));
        int sum = 0;
        for (int l = 0; l < nlev - 1; ++l) {
            per_level[l] = (int)lrintf(nd);
            sum += per_level[l];
            nd *= factor;
        }
        per_level[nlev - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    }
    // corners after non-max suppression: strict 3x3 maxima, so at most one per 2x2 block of any level
    const int ccap = (int)(P.row_floats / 4 + 64);
    const int first = 12288;                  // records fetched with the header; the rest only if there are more
    Arena out, work;
    const size_t img_bytes = (size_t)channels * W * H;
    const size_t o_hdr = out.take(16), o_cand = out.take(24 * (size_t)ccap);
    const size_t o_plain = work.take(P.plain_bytes), o_score = work.take(P.row_floats);
    TRY(ensure_host(ctx, ctx->h_in, img_bytes + 256));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    TRY(orb_prepare_tables(ctx, W, H, nlev, P));
    const int* d_tab = (const int*)(ctx->d_orb_tab.p + 1024);
    uint8_t* d_plain = ctx->d_work.p + o_plain;
    TRY(orb_upload_frame(ctx, image, W, H, row_bytes, channels, ctx->h_in.p));
    if (channels == 1)
        CK(cudaMemcpyAsync(d_plain, ctx->d_orb_frame.p, img_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    int l = 0;
    CK(launch_orb_detect(channels == 3 ? ctx->d_orb_frame.p : nullptr, colour_order, W, H, 3 * W, P, d_plain,
                         ctx->d_work.p + o_score, d_tab, 20, (int*)(ctx->d_out.p + o_cand), ccap, (int*)(ctx->d_out.p + o_hdr),
                         prep_slots(ctx), epoch, ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    ctx->f2m.valid = false; ctx->f2f.valid = false;
    // the buffers hold ccap records: a small image (ccap < first) must not fetch past them
    const int head_recs = first < ccap ? first : ccap;
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, o_cand + 24 * (size_t)head_recs, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int found = *(const int*)(ctx->h_out.p + o_hdr);
    if (found > ccap)
        return fail(ctx, PSLAM_ERR_CAPACITY, "pslam_orb_detect: %d FAST corners after suppression, device buffer holds %d", found, ccap);
    if (found > first) {
        CK(cudaMemcpyAsync(ctx->h_out.p + o_cand + 24 * (size_t)first, ctx->d_out.p + o_cand + 24 * (size_t)first,
                           24 * (size_t)(found - first), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    const int* rec = (const int*)(ctx->h_out.p + o_cand);
    int n = 0, pos = 0;
    std::vector<OrbKp> kp;
    for (int lev = 0; lev < nlev; ++lev) {
        kp.clear();
        for (; pos < found && rec[6 * pos] == lev; ++pos) {
            OrbKp k;
            k.x = (float)rec[6 * pos + 1]; k.y = (float)rec[6 * pos + 2]; k.response = (float)rec[6 * pos + 3];
            memcpy(&k.harris, &rec[6 * pos + 4], 4); memcpy(&k.angle, &rec[6 * pos + 5], 4);
            kp.push_back(k);
        }
        orb_retain_best(kp, 2 * per_level[lev]);          // "keep more points than necessary as FAST does not give amazing corners"
        for (OrbKp& k : kp) k.response = k.harris;
        orb_retain_best(kp, per_level[lev]);
        const float sf = orb_level_scale(lev);
        for (const OrbKp& k : kp) {
            if (n < cap) {
                kp_xy[2 * n] = k.x * sf; kp_xy[2 * n + 1] = k.y * sf;
                kp_size[n] = 31 * sf; kp_angle[n] = k.angle; kp_response[n] = k.response; kp_octave[n] = lev;
            }
            ++n;
        }
    }
    *n_out = n;
    if (n > cap) return fail(ctx, PSLAM_ERR_CAPACITY, "pslam_orb_detect: %d keypoints, capacity %d", n, cap);
    return PSLAM_OK;
}

int pslam_fast_detect(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels, int colour_order,
                      int threshold, float* kp_xy, float* kp_response, int cap, int* n_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!n_out || !image || W <= 0 || H <= 0 || (channels != 1 && channels != 3) || row_bytes < channels * W || cap < 0 ||
        threshold < 1 || threshold > 254 || (colour_order != 0 && colour_order != 1) || (cap > 0 && (!kp_xy || !kp_response)))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_fast_detect: bad argument");
    *n_out = 0;
    CK(cudaSetDevice(ctx->device));
    OrbPlan P;
    orb_plan(W, H, 1, &P);
    const int ccap = (int)(P.row_floats / 4 + 64);
    const int first = 16384;
    Arena out, work;
    const size_t img_bytes = (size_t)channels * W * H;
    const size_t o_hdr = out.take(16), o_cand = out.take(24 * (size_t)ccap);
    const size_t o_plain = work.take(P.plain_bytes), o_score = work.take(P.row_floats);
    TRY(ensure_host(ctx, ctx->h_in, img_bytes + 256));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* d_plain = ctx->d_work.p + o_plain;
    TRY(orb_upload_frame(ctx, image, W, H, row_bytes, channels, ctx->h_in.p));
    if (channels == 1)
        CK(cudaMemcpyAsync(d_plain, ctx->d_orb_frame.p, img_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    int l = 0;
    CK(launch_fast_detect(channels == 3 ? ctx->d_orb_frame.p : nullptr, colour_order, W, H, 3 * W, P, d_plain,
                          ctx->d_work.p + o_score, threshold, (int*)(ctx->d_out.p + o_cand), ccap, (int*)(ctx->d_out.p + o_hdr),
                          prep_slots(ctx), epoch, ctx->sm_count, ctx->stream, &l));
    ctx->launches += l;
    ctx->f2m.valid = false; ctx->f2f.valid = false;
    const size_t head = o_cand + 24 * (size_t)(first < ccap ? first : ccap);
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, head, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int found = *(const int*)(ctx->h_out.p + o_hdr);
    if (found > ccap)
        return fail(ctx, PSLAM_ERR_CAPACITY, "pslam_fast_detect: %d corners after suppression, device buffer holds %d", found, ccap);
    if (found > first) {
        CK(cudaMemcpyAsync(ctx->h_out.p + head, ctx->d_out.p + head, 24 * (size_t)(found - first), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    const int* rec = (const int*)(ctx->h_out.p + o_cand);
    const int n = found < cap ? found : cap;
    for (int k = 0; k < n; ++k) {
        kp_xy[2 * k] = (float)rec[6 * k + 1]; kp_xy[2 * k + 1] = (float)rec[6 * k + 2];
        kp_response[k] = (float)rec[6 * k + 3];
    }
    *n_out = found;
    if (found > cap) return fail(ctx, PSLAM_ERR_CAPACITY, "pslam_fast_detect: %d corners, capacity %d", found, cap);
    return PSLAM_OK;
}

// ---- KLT tracking (MatcherOpenCV::performTracking, reference src/Matcher/matcherOpenCV.cpp:209-300) ----------------
static double sq_threshold_d(double d) {   // smallest double T with sqrt(T) >= d: sqrt(s) < d  <=>  s < T
    if (!(d > 0.)) return 0.;              // nothing is closer than a non-positive (or NaN) distance
    if (isinf(d)) return INFINITY;
    double t = d * d;
    if (isinf(t)) t = 1.7976931348623157e308;
    while (t > 0. && sqrt(t) >= d) t = nextafter(t, 0.);
    while (sqrt(t) < d) t = nextafter(t, INFINITY);
    return t;
}
static float prune_limit(double d) {       // smallest float >= d (pre-test of klt_pair_removes); +inf disables the pre-test
    if (!(d > 0.)) return 0.f;             // nothing is close anyway (sq_threshold_d gives 0): everything is discarded early
    return float_at_least(d);
}
static void pack_rows(uint8_t* dst, const uint8_t* src, size_t row, int H, int row_bytes) {
    if ((size_t)row_bytes == row) memcpy(dst, src, row * H);
    else for (int y = 0; y < H; ++y) memcpy(dst + (size_t)y * row, src + (size_t)y * row_bytes, row);
}
// the steps Matcher::trackKLT runs on performTracking's survivors (matcher.cpp:160-207), for pslam_klt_frame
struct KltFrameArgs {
    const float* prev_xyz; const uint16_t* depth; int depth_stride; const pslam_camera* cam; int undistort; double depth_scale;
    const pslam_ransac_params* rparams; uint64_t seed; int num_hyp;
    float* uv_und_out; float* xyz_out; double* det_dist_out; int* inlier_idx_out; pslam_frame_result* result;
};
static int klt_run(pslam_ctx* ctx, const char* who, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H,
                   int row_bytes, int channels, const float* prev_xy, float* cur_xy, int n, int win, int max_level,
                   int criteria_type, int max_iter, double eps, int flags, double min_eig_threshold, bool prune,
                   double error_threshold, double min_distance, uint8_t* status, float* err, int* kept_idx_out,
                   int* n_kept_out, const KltFrameArgs* fr = nullptr) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n_kept_out) *n_kept_out = 0;
    RansacDeviceParams rp = {};
    if (fr) {
        if (!fr->result) return fail(ctx, PSLAM_ERR_ARG, "%s: bad argument", who);
        memset(fr->result, 0, sizeof(*fr->result));
        for (int i = 0; i < 16; ++i) fr->result->T[i] = (i % 5 == 0) ? 1.f : 0.f;
        if (!fr->cam || !fr->depth || fr->depth_stride < W || (n > 0 && (!fr->prev_xyz || !fr->xyz_out || !fr->inlier_idx_out)))
            return fail(ctx, PSLAM_ERR_ARG, "%s: bad argument", who);
        TRY(make_ransac_params(ctx, fr->rparams, fr->seed, fr->num_hyp, rp));
    }
    if (!cur_image || W <= 0 || H <= 0 || (channels != 1 && channels != 3) || row_bytes < channels * W || n < 0 ||
        (n > 0 && (!prev_xy || !cur_xy || !status || !err)) || (prune && (!n_kept_out || (n > 0 && !kept_idx_out))) ||
        (flags & ~(PSLAM_KLT_USE_INITIAL_FLOW | PSLAM_KLT_GET_MIN_EIGENVALS)) || (criteria_type & ~3))
        return fail(ctx, PSLAM_ERR_ARG, "%s: bad argument", who);
    if (win < 3 || win > kKltMaxWin || max_level < 0 || max_level >= kKltMaxLevels)
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "%s: window %d (3 .. %d) / max_level %d (0 .. %d)", who, win, kKltMaxWin, max_level,
                    kKltMaxLevels - 1);
    CK(cudaSetDevice(ctx->device));
    KltPlan plan, full;
    klt_plan(W, H, channels, win, max_level, &plan);
    klt_plan(W, H, channels, 0, kKltMaxLevels - 1, &full);       // buffer layout: independent of window and depth
    const bool same_shape = ctx->klt_cur >= 0 && ctx->klt_W == W && ctx->klt_H == H && ctx->klt_cn == channels;
    if (!prev_image && !(same_shape && ctx->klt_levels >= plan.n_levels))
        return fail(ctx, PSLAM_ERR_ARG, "%s: prev_image == NULL needs a preceding pslam_klt_* call on this context with a "
                    "%d x %d x %d frame and at least %d pyramid levels", who, W, H, channels, plan.n_levels);
    const int jbuf = prev_image ? 0 : 1 - ctx->klt_cur, ibuf = 1 - jbuf;
    const size_t row = (size_t)channels * W, img_bytes = row * H;
    Arena in, out;
    const size_t o_img_j = in.take(img_bytes), o_img_i = in.take(prev_image ? img_bytes : 0);
    const size_t o_prev = in.take(8 * (size_t)n), o_init = in.take(8 * (size_t)n);
    const size_t o_xy = out.take(8 * (size_t)n), o_err = out.take(4 * (size_t)n), o_st = out.take((size_t)n), o_keep = out.take((size_t)n);
    // fused frame: previous 3-D points and the depth image in; survivors' match list, compacted positions, their
    // undistorted / back-projected form and the RANSAC result out
    const int cap = n > 0 ? n : 1;
    const size_t depth_bytes = fr ? sizeof(uint16_t) * (size_t)H * fr->depth_stride : 0;
    const size_t o_pxyz = in.take(fr ? 12 * (size_t)n : 0), o_depth = in.take(depth_bytes);
    const size_t o_m = out.take(fr ? sizeof(int) * (1 + 3 * (size_t)cap) : 0), o_cxy = out.take(fr ? 8 * (size_t)cap : 0);
    const size_t o_und = out.take(fr ? 8 * (size_t)cap : 0), o_xyz = out.take(fr ? 12 * (size_t)cap : 0);
    const size_t o_dd = out.take(fr ? 8 * (size_t)cap : 0), o_res = out.take(fr ? sizeof(int) * ransac_result_ints(cap) : 0);
    Arena work;
    RansacLayout RL = {};
    if (fr) RL = plan_ransac(work, cap, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_host(ctx, ctx->h_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_in, in.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    if (fr) TRY(ensure_dev(ctx, ctx->d_work, work.off));
    ctx->klt_cur = -1;                                            // until this call has succeeded
    TRY(ensure_dev(ctx, ctx->d_klt_pyr[0], full.bytes)); TRY(ensure_dev(ctx, ctx->d_klt_pyr[1], full.bytes));
    uint8_t* h = ctx->h_in.p;
    pack_rows(h + o_img_j, cur_image, row, H, row_bytes);
    if (prev_image) pack_rows(h + o_img_i, prev_image, row, H, row_bytes);
    const bool init = (flags & PSLAM_KLT_USE_INITIAL_FLOW) != 0;
    if (n > 0) {
        memcpy(h + o_prev, prev_xy, 8 * (size_t)n);
        if (init) memcpy(h + o_init, cur_xy, 8 * (size_t)n);
    }
    uint8_t* pyrI = ctx->d_klt_pyr[ibuf].p;
    uint8_t* pyrJ = ctx->d_klt_pyr[jbuf].p;
    CK(cudaMemcpyAsync(pyrJ, h + o_img_j, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (prev_image) CK(cudaMemcpyAsync(pyrI, h + o_img_i, img_bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (n > 0) CK(cudaMemcpyAsync(ctx->d_in.p + o_prev, h + o_prev, (init ? o_init + 8 * (size_t)n : o_prev + 8 * (size_t)n) - o_prev,
                                  cudaMemcpyHostToDevice, ctx->stream));
    if (fr && n > 0) {
        memcpy(h + o_pxyz, fr->prev_xyz, 12 * (size_t)n);
        memcpy(h + o_depth, fr->depth, sizeof(uint16_t) * ((size_t)(H - 1) * fr->depth_stride + (size_t)W));   // not past the last row of an ROI
        CK(cudaMemcpyAsync(ctx->d_in.p + o_pxyz, h + o_pxyz, o_depth + depth_bytes - o_pxyz, cudaMemcpyHostToDevice, ctx->stream));
    }
    int l = 0;
    CK(launch_klt_pyramid(prev_image ? pyrI : nullptr, pyrJ, plan, channels, ctx->stream, &l));
    float* d_xy = (float*)(ctx->d_out.p + o_xy);
    if (n > 0) {
        if (init) CK(cudaMemcpyAsync(d_xy, ctx->d_in.p + o_init, 8 * (size_t)n, cudaMemcpyDeviceToDevice, ctx->stream));
        KltParams P;
        memset(&P, 0, sizeof(P));
        P.n_levels = plan.n_levels; P.win = win; P.cn = channels;
        klt_criteria(criteria_type & 1, max_iter, criteria_type & 2, eps, &P.max_iter, &P.eps_sq);
        P.min_eig_thr = min_eig_threshold;
        P.use_initial_flow = init ? 1 : 0;
        P.min_eig_err = (flags & PSLAM_KLT_GET_MIN_EIGENVALS) ? 1 : 0;
        for (int k = 0; k < plan.n_levels; ++k) {
            P.lv[k].I = pyrI + full.off[k]; P.lv[k].J = pyrJ + full.off[k];
            P.lv[k].w = plan.w[k]; P.lv[k].h = plan.h[k];
        }
        CK(launch_klt_track(P, (const float*)(ctx->d_in.p + o_prev), d_xy, n, ctx->d_out.p + o_st, (float*)(ctx->d_out.p + o_err),
                            ctx->stream, &l));
        if (prune)
            CK(launch_klt_prune(d_xy, (const float*)(ctx->d_out.p + o_err), ctx->d_out.p + o_st, n, error_threshold,
                                sq_threshold_d(min_distance), prune_limit(min_distance), ctx->d_out.p + o_keep, ctx->stream, &l));
        if (fr) {
            int* mout = (int*)(ctx->d_out.p + o_m);
            float* cxy = (float*)(ctx->d_out.p + o_cxy);
            float* xyz = (float*)(ctx->d_out.p + o_xyz);
            CK(launch_klt_compact(ctx->d_out.p + o_keep, d_xy, n, cap, mout, cxy, ctx->stream, &l));
            CK(launch_backproject(cxy, n, (const uint16_t*)(ctx->d_in.p + o_depth), W, H, fr->depth_stride, *fr->cam, fr->undistort,
                                  fr->depth_scale, (float*)(ctx->d_out.p + o_und), xyz, (double*)(ctx->d_out.p + o_dd), nullptr,
                                  nullptr, ctx->stream, &l));
            RansacWorkspace ws = bind_ransac(RL, ctx->d_work.p, (int*)(ctx->d_out.p + o_res));
            CK(launch_ransac((const float*)(ctx->d_in.p + o_pxyz), xyz, mout + 1, mout + 1 + cap, mout, 0, rp, ws, ctx->sm_count,
                             ctx->stream, &l));
            ctx->d_last_counts = ws.counts; ctx->last_H = ws.h_cap;
        }
        CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_out.p, prune ? out.off : o_keep, cudaMemcpyDeviceToHost, ctx->stream));
    }
    ctx->launches += l;
    ctx->f2m.valid = false; ctx->f2f.valid = false;
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->klt_cur = jbuf; ctx->klt_W = W; ctx->klt_H = H; ctx->klt_cn = channels;
    // the previous frame's buffer keeps the levels it was built with; the reusable depth is that of the current frame
    ctx->klt_levels = plan.n_levels;
    if (n > 0) {
        memcpy(cur_xy, ctx->h_out.p + o_xy, 8 * (size_t)n);
        memcpy(err, ctx->h_out.p + o_err, 4 * (size_t)n);
        memcpy(status, ctx->h_out.p + o_st, (size_t)n);
        if (prune) {
            const uint8_t* keep = ctx->h_out.p + o_keep;
            int m = 0;
            for (int i = 0; i < n; ++i) if (keep[i]) kept_idx_out[m++] = i;
            *n_kept_out = m;
        }
        if (fr) {
            const int* mo = (const int*)(ctx->h_out.p + o_m);
            const int m = mo[0];
            if (m != *n_kept_out || memcmp(mo + 1, kept_idx_out, 4 * (size_t)m) != 0)
                return fail(ctx, PSLAM_ERR_CUDA, "%s: device compaction disagrees with the keep flags", who);
            pslam_frame_result* R = fr->result;
            R->n_matches = m;
            if (fr->uv_und_out) memcpy(fr->uv_und_out, ctx->h_out.p + o_und, 8 * (size_t)m);
            memcpy(fr->xyz_out, ctx->h_out.p + o_xyz, 12 * (size_t)m);
            if (fr->det_dist_out) memcpy(fr->det_dist_out, ctx->h_out.p + o_dd, 8 * (size_t)m);
            unpack_ransac_result((const int*)(ctx->h_out.p + o_res), R->T, fr->inlier_idx_out, &R->n_inliers, &R->best_ratio,
                                 &R->hyp_used, &R->n_filtered);
            // pointInlierRatio over trainIdx (RANSAC.h:56-66): survivor j has trainIdx j, every one distinct
            R->inlier_ratio = m > 0 ? (double)R->n_inliers / (double)m : 0.0;
        }
    }
    return PSLAM_OK;
}

int pslam_klt_track(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H, int row_bytes,
                    int channels, const float* prev_xy, float* cur_xy, int n, int win, int max_level, int criteria_type,
                    int max_iter, double eps, int flags, double min_eig_threshold, uint8_t* status, float* err) {
    return klt_run(ctx, "pslam_klt_track", prev_image, cur_image, W, H, row_bytes, channels, prev_xy, cur_xy, n, win, max_level,
                   criteria_type, max_iter, eps, flags, min_eig_threshold, false, 0., 0., status, err, nullptr, nullptr);
}

int pslam_klt_perform_tracking(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H, int row_bytes,
                               int channels, const float* prev_xy, float* cur_xy, int n, int win, int max_level,
                               int criteria_type, int max_iter, double eps, int flags, double min_eig_threshold,
                               double error_threshold, double min_distance, uint8_t* status, float* err, int* kept_idx_out,
                               int* n_kept_out) {
    return klt_run(ctx, "pslam_klt_perform_tracking", prev_image, cur_image, W, H, row_bytes, channels, prev_xy, cur_xy, n, win,
                   max_level, criteria_type, max_iter, eps, flags, min_eig_threshold, true, error_threshold, min_distance, status,
                   err, kept_idx_out, n_kept_out);
}

int pslam_klt_frame(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H, int row_bytes, int channels,
                    const float* prev_xy, const float* prev_xyz, float* cur_xy, int n, int win, int max_level, int criteria_type,
                    int max_iter, double eps, int flags, double min_eig_threshold, double error_threshold, double min_distance,
                    const uint16_t* depth, int depth_row_stride, const pslam_camera* cam, int undistort, double depth_scale,
                    const pslam_ransac_params* params, uint64_t seed, int num_hyp, uint8_t* status, float* err,
                    int* kept_idx_out, int* n_kept_out, float* kept_uv_undist_out, float* kept_xyz_out,
                    double* kept_det_dist_out, int* inlier_idx_out, pslam_frame_result* result) {
    KltFrameArgs fr = {prev_xyz, depth, depth_row_stride, cam, undistort, depth_scale, params, seed, num_hyp,
                       kept_uv_undist_out, kept_xyz_out, kept_det_dist_out, inlier_idx_out, result};
    return klt_run(ctx, "pslam_klt_frame", prev_image, cur_image, W, H, row_bytes, channels, prev_xy, cur_xy, n, win, max_level,
                   criteria_type, max_iter, eps, flags, min_eig_threshold, true, error_threshold, min_distance, status, err,
                   kept_idx_out, n_kept_out, &fr);
}

// ---- resident feature map -------------------------------------------------------------------------
// The map side of Matcher::matchXYZ kept in HBM between frames: per frame only the camera pose and
// the current keypoints cross PCIe; the view-angle / depth filters, the move to the camera frame, the level
// prediction, guided matching and RANSAC run back to back on the device.
static int grow_map_array(pslam_ctx* ctx, void** p, size_t bytes_old, size_t bytes_new) {
    void* np_ = nullptr;
    CK(cudaMalloc(&np_, bytes_new));
    if (*p && bytes_old) CK(cudaMemcpyAsync(np_, *p, bytes_old, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*p) CK(cudaFree(*p));
    *p = np_;
    return PSLAM_OK;
}

int pslam_map_reserve(pslam_ctx* ctx, int max_features) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (max_features < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_map_reserve: negative size");
    if (max_features <= ctx->map_cap) return PSLAM_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t o = (size_t)ctx->map_n, n = (size_t)max_features;
    TRY(grow_map_array(ctx, (void**)&ctx->d_map_xyz, 24 * o, 24 * n));
    TRY(grow_map_array(ctx, (void**)&ctx->d_map_desc, 32 * o, 32 * n));
    TRY(grow_map_array(ctx, (void**)&ctx->d_map_oct, 4 * o, 4 * n));
    TRY(grow_map_array(ctx, (void**)&ctx->d_map_det, 8 * o, 8 * n));
    TRY(grow_map_array(ctx, (void**)&ctx->d_map_axis, 12 * o, 12 * n));
    ctx->map_cap = max_features;
    return PSLAM_OK;
}

int pslam_map_write(pslam_ctx* ctx, int first, int count, const double* xyz, const uint8_t* desc, const int* octave,
                    const double* det_dist, const float* view_axis) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (first < 0 || count < 0 || first > ctx->map_n) return fail(ctx, PSLAM_ERR_ARG, "pslam_map_write: range must start inside or at the end of the map");
    if (count == 0) return PSLAM_OK;
    const bool extends = first + count > ctx->map_n;
    if (extends && (!xyz || !desc || !octave || !det_dist || !view_axis))
        return fail(ctx, PSLAM_ERR_ARG, "pslam_map_write: new features need every attribute");
    if (first + count > ctx->map_cap) {
        int want = ctx->map_cap + ctx->map_cap / 2;
        if (want < first + count) want = first + count;
        if (want < 1024) want = 1024;
        TRY(pslam_map_reserve(ctx, want));
    }
    CK(cudaSetDevice(ctx->device));
    Arena in;
    const size_t c = (size_t)count;
    const size_t o_x = in.take(24 * c), o_d = in.take(32 * c), o_o = in.take(4 * c), o_t = in.take(8 * c), o_a = in.take(12 * c);
    TRY(ensure_host(ctx, ctx->h_in, in.off));
    uint8_t* h = ctx->h_in.p;
    const size_t f = (size_t)first;
    if (xyz) { memcpy(h + o_x, xyz, 24 * c); CK(cudaMemcpyAsync(ctx->d_map_xyz + 3 * f, h + o_x, 24 * c, cudaMemcpyHostToDevice, ctx->stream)); }
    if (desc) { memcpy(h + o_d, desc, 32 * c); CK(cudaMemcpyAsync(ctx->d_map_desc + 32 * f, h + o_d, 32 * c, cudaMemcpyHostToDevice, ctx->stream)); }
    if (octave) { memcpy(h + o_o, octave, 4 * c); CK(cudaMemcpyAsync(ctx->d_map_oct + f, h + o_o, 4 * c, cudaMemcpyHostToDevice, ctx->stream)); }
    if (det_dist) { memcpy(h + o_t, det_dist, 8 * c); CK(cudaMemcpyAsync(ctx->d_map_det + f, h + o_t, 8 * c, cudaMemcpyHostToDevice, ctx->stream)); }
    if (view_axis) { memcpy(h + o_a, view_axis, 12 * c); CK(cudaMemcpyAsync(ctx->d_map_axis + 3 * f, h + o_a, 12 * c, cudaMemcpyHostToDevice, ctx->stream)); }
    CK(cudaStreamSynchronize(ctx->stream));
    if (extends) ctx->map_n = first + count;
    return PSLAM_OK;
}

int pslam_map_truncate(pslam_ctx* ctx, int n) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n < 0 || n > ctx->map_n) return fail(ctx, PSLAM_ERR_ARG, "pslam_map_truncate: size outside [0, current size]");
    ctx->map_n = n;
    return PSLAM_OK;
}

int pslam_map_size(const pslam_ctx* ctx, int* n_features) {
    if (!ctx || !n_features) return PSLAM_ERR_ARG;
    *n_features = ctx->map_n;
    return PSLAM_OK;
}

int pslam_frame_to_resident_map(pslam_ctx* ctx, const double camera_pose[16], const pslam_map_prepare_params* prep,
                                const float* cur_xyz, const uint8_t* cur_desc, const int* cur_octave,
                                const double* cur_det_dist, int N, double radius, double accept_ratio, int distance_mode,
                                const pslam_ransac_params* params, uint64_t seed, int num_hyp, int match_cap,
                                int* kept_idx_out, int* n_kept_out, double* xyz_local_out, double* uv_out,
                                int* match_query_out, int* match_train_out, float* match_dist_out, int* inlier_idx_out,
                                pslam_frame_result* result) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!result || !camera_pose || !prep || !n_kept_out || N < 0 || match_cap <= 0)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_resident_map: bad argument");
    memset(result, 0, sizeof(*result));
    for (int i = 0; i < 16; ++i) result->T[i] = (i % 5 == 0) ? 1.f : 0.f;
    result->inlier_ratio = -1.0;
    *n_kept_out = 0;
    RansacDeviceParams rp;
    TRY(make_ransac_params(ctx, params, seed, num_hyp, rp));
    if (distance_mode != 0 && distance_mode != 1) return fail(ctx, PSLAM_ERR_ARG, "distance_mode must be 0 or 1");
    if (N > 12000) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "N above 12000 current keypoints");
    const int M = ctx->map_n;
    if (M == 0) return PSLAM_OK;
    if (!kept_idx_out || (N > 0 && (!cur_xyz || !cur_desc || !cur_octave || !cur_det_dist)) || !match_query_out ||
        !match_train_out || !match_dist_out || !inlier_idx_out)
        return fail(ctx, PSLAM_ERR_ARG, "pslam_frame_to_resident_map: null buffer");
    CK(cudaSetDevice(ctx->device));
    const int cap = match_cap;
    const int Nn = N > 0 ? N : 1;
    Arena in, out, work;
    const size_t o_cx = in.take(12 * (size_t)Nn), o_cd = in.take(32 * (size_t)Nn), o_co = in.take(4 * (size_t)Nn);
    const size_t o_cdet = in.take(8 * (size_t)Nn);
    const size_t o_n = out.take(16), o_g = out.take(sizeof(int) * (2 + 3 * (size_t)cap));
    const size_t o_res = out.take(sizeof(int) * ransac_result_ints(cap));
    const size_t o_k = out.take(4 * (size_t)M), o_xl = out.take(24 * (size_t)M), o_uv = out.take(16 * (size_t)M);
    const size_t o_ang = work.take(8 * (size_t)M), o_wd = work.take(32 * (size_t)M), o_wo = work.take(4 * (size_t)M);
    const size_t o_wt = work.take(8 * (size_t)M), o_wx = work.take(12 * (size_t)M), o_wml = work.take(4 * (size_t)M);
    const size_t o_wcl = work.take(4 * (size_t)Nn);
    const size_t o_cnt = work.take(sizeof(int) * (2 * (size_t)M + 1)), o_best = work.take(sizeof(int) * (size_t)M);
    const size_t o_cache = work.take(guided_cache_bytes(M));
    RansacLayout L = plan_ransac(work, cap, rp);
    TRY(ensure_host(ctx, ctx->h_in, in.off)); TRY(ensure_dev(ctx, ctx->d_in, in.off));
    TRY(ensure_host(ctx, ctx->h_out, out.off)); TRY(ensure_dev(ctx, ctx->d_out, out.off));
    TRY(ensure_dev(ctx, ctx->d_work, work.off));
    uint8_t* h = ctx->h_in.p;
    if (N > 0) {
        memcpy(h + o_cx, cur_xyz, 12 * (size_t)N); memcpy(h + o_cd, cur_desc, 32 * (size_t)N);
        memcpy(h + o_co, cur_octave, 4 * (size_t)N); memcpy(h + o_cdet, cur_det_dist, 8 * (size_t)N);
    }
    CK(cudaMemcpyAsync(ctx->d_in.p, h, in.off, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t* d = ctx->d_in.p; uint8_t* w = ctx->d_work.p; uint8_t* o = ctx->d_out.p;
    int* d_n = (int*)(o + o_n);
    int l = 0;
    unsigned int epoch = 0;
    TRY(next_epoch(ctx, &epoch));
    CK(launch_map_prepare(ctx->d_map_xyz, ctx->d_map_axis, M, camera_pose, prep->fx, prep->fy, prep->cx, prep->cy,
                          prep->image_w, prep->image_h, prep->max_angle, prep->max_z, (int*)(o + o_k), (double*)(o + o_xl),
                          (double*)(o + o_uv), (double*)(w + o_ang), d_n, prep_slots(ctx), epoch, ctx->sm_count,
                          ctx->stream, &l, ctx->d_map_desc, w + o_wd, ctx->d_map_oct, (int*)(w + o_wo), ctx->d_map_det,
                          (double*)(w + o_wt)));
    int* gout = (int*)(o + o_g);
    RansacWorkspace ws = bind_ransac(L, w, (int*)(o + o_res));
    if (N > 0) {
        const HostLevelTables& t = level_tables();
        unsigned int epoch2 = 0;
        TRY(next_epoch(ctx, &epoch2));
        CK(launch_predict_levels((const double*)(o + o_xl), (const int*)(w + o_wo), (const double*)(w + o_wt), M,
                                 (const float*)(d + o_cx), (const int*)(d + o_co), (const double*)(d + o_cdet), N, t.pow_tab,
                                 t.lvl_tab, t.log_sf, (float*)(w + o_wx), (int*)(w + o_wml), (int*)(w + o_wcl), ctx->stream,
                                 &l, d_n));
        CK(launch_guided_match((const float*)(w + o_wx), w + o_wd, (const int*)(w + o_wml), M, (const float*)(d + o_cx),
                               d + o_cd, (const int*)(w + o_wcl), N, sq_threshold(float_at_least(radius)), accept_ratio,
                               distance_mode, (int*)(w + o_cnt), (int*)(w + o_best), w + o_cache, gout, cap, emit_slots(ctx),
                               epoch2, ctx->sm_count, ctx->stream, &l, d_n));
        CK(launch_ransac((const float*)(w + o_wx), (const float*)(d + o_cx), gout + 2, gout + 2 + cap, gout, 0, rp, ws,
                         ctx->sm_count, ctx->stream, &l));
        ctx->d_last_counts = ws.counts; ctx->last_H = ws.h_cap;
    } else {
        CK(cudaMemsetAsync(gout, 0, 8, ctx->stream));
    }
    ctx->launches += l;
    ctx->f2m.valid = false;   // the work buffers of a previous pslam_frame_to_map were reused
    const bool want_local = xyz_local_out || uv_out;
    CK(cudaMemcpyAsync(ctx->h_out.p, o, want_local ? out.off : o_xl, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const int nk = *(const int*)(ctx->h_out.p + o_n);
    *n_kept_out = nk;
    memcpy(kept_idx_out, ctx->h_out.p + o_k, 4 * (size_t)nk);
    if (xyz_local_out) memcpy(xyz_local_out, ctx->h_out.p + o_xl, 24 * (size_t)nk);
    if (uv_out) memcpy(uv_out, ctx->h_out.p + o_uv, 16 * (size_t)nk);
    const int* g = (const int*)(ctx->h_out.p + o_g);
    const int total = g[0];
    const int n = total < cap ? total : cap;
    result->n_matches = total;
    if (n > 0) {
        memcpy(match_query_out, g + 2, 4 * (size_t)n);
        memcpy(match_train_out, g + 2 + cap, 4 * (size_t)n);
        memcpy(match_dist_out, g + 2 + 2 * cap, 4 * (size_t)n);
    }
    if (total > cap) return fail(ctx, PSLAM_ERR_CAPACITY, "guided matching produced %d matches, capacity %d", total, cap);
    if (total == 0) return PSLAM_OK;
    unpack_ransac_result((const int*)(ctx->h_out.p + o_res), result->T, inlier_idx_out, &result->n_inliers,
                         &result->best_ratio, &result->hyp_used, &result->n_filtered);
    std::vector<int> inl_t((size_t)result->n_inliers);
    for (int i = 0; i < result->n_inliers; ++i) inl_t[i] = match_train_out[inlier_idx_out[i]];
    result->inlier_ratio = point_inlier_ratio(inl_t.data(), result->n_inliers, match_train_out, n);
    return PSLAM_OK;
}

// ---- loop-closure database ----------------------------------------------------------------------
int pslam_lc_db_reserve(pslam_ctx* ctx, int64_t max_descriptors, int max_keyframes) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (max_descriptors < 0 || max_keyframes < 0) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_db_reserve: negative size");
    CK(cudaSetDevice(ctx->device));
    if (max_descriptors > ctx->db_cap) {
        uint8_t* nd = nullptr;
        CK(cudaMalloc((void**)&nd, (size_t)max_descriptors * 32));
        if (ctx->d_db) {
            CK(cudaMemcpyAsync(nd, ctx->d_db, (size_t)ctx->db_n * 32, cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(ctx->d_db));
        }
        ctx->d_db = nd;
        ctx->db_cap = max_descriptors;
    }
    if (max_keyframes > ctx->kf_cap) {
        int64_t* no = nullptr;
        int* ns = nullptr;
        CK(cudaMalloc((void**)&no, sizeof(int64_t) * ((size_t)max_keyframes + 1)));
        CK(cudaMalloc((void**)&ns, sizeof(int) * ((size_t)max_keyframes + 1)));
        int* nk = nullptr;
        CK(cudaMalloc((void**)&nk, sizeof(int) * ((size_t)max_keyframes + 1)));
        CK(cudaMemsetAsync(nk, 0, sizeof(int) * ((size_t)max_keyframes + 1), ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->d_kf_done) CK(cudaFree(ctx->d_kf_done));
        ctx->d_kf_done = nk;
        if (ctx->d_kf_off) {
            CK(cudaMemcpyAsync(no, ctx->d_kf_off, sizeof(int64_t) * ((size_t)ctx->n_kf + 1), cudaMemcpyDeviceToDevice, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            CK(cudaFree(ctx->d_kf_off));
            CK(cudaFree(ctx->d_scores));
        }
        ctx->d_kf_off = no;
        ctx->d_scores = ns;
        ctx->kf_cap = max_keyframes;
    }
    if (ctx->h_kf_off.empty()) ctx->h_kf_off.push_back(0);
    return PSLAM_OK;
}

int pslam_lc_db_append(pslam_ctx* ctx, const uint8_t* desc, const int64_t* kf_off, int n_kf) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n_kf < 0 || (n_kf > 0 && !kf_off)) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_db_append: bad argument");
    if (n_kf == 0) return PSLAM_OK;
    const int64_t nd = kf_off[n_kf] - kf_off[0];
    if (nd < 0 || (nd > 0 && !desc)) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_db_append: bad offsets");
    for (int k = 0; k < n_kf; ++k) {
        const int64_t c = kf_off[k + 1] - kf_off[k];
        if (c < 0) return fail(ctx, PSLAM_ERR_ARG, "keyframe offsets must be non-decreasing");
        if (c > PSLAM_LC_MAX_KF_DESC) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "keyframe with %lld descriptors (max %d)", (long long)c, PSLAM_LC_MAX_KF_DESC);
        if ((int)c > ctx->max_kf_desc) ctx->max_kf_desc = (int)c;
    }
    CK(cudaSetDevice(ctx->device));
    if (ctx->h_kf_off.empty()) ctx->h_kf_off.push_back(0);
    if (ctx->db_n + nd > ctx->db_cap || ctx->n_kf + n_kf > ctx->kf_cap) {
        int64_t want_d = ctx->db_n + nd, want_k = (int64_t)ctx->n_kf + n_kf;
        if (want_d < 2 * ctx->db_cap) want_d = 2 * ctx->db_cap;
        if (want_k < 2 * (int64_t)ctx->kf_cap) want_k = 2 * (int64_t)ctx->kf_cap;
        TRY(pslam_lc_db_reserve(ctx, want_d, (int)want_k));
    }
    // descriptors go straight from the caller's buffer (chunked through the pinned arena)
    const size_t chunk = (size_t)8 << 20;
    TRY(ensure_host(ctx, ctx->h_in, nd > 0 ? (chunk < (size_t)nd * 32 ? chunk : (size_t)nd * 32) : 64));
    const uint8_t* src = desc + (size_t)kf_off[0] * 32;
    for (size_t done = 0; done < (size_t)nd * 32; done += chunk) {
        const size_t b = ((size_t)nd * 32 - done) < chunk ? ((size_t)nd * 32 - done) : chunk;
        memcpy(ctx->h_in.p, src + done, b);
        CK(cudaMemcpyAsync(ctx->d_db + (size_t)ctx->db_n * 32 + done, ctx->h_in.p, b, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // the map is kept in the re-encoded row form the sweep kernels compare in (common.cuh ham256_key_enc); once per append
    if (nd > 0) {
        int l = 0;
        CK(launch_lc_encode_rows(ctx->d_db + (size_t)ctx->db_n * 32, (long long)nd, ctx->stream, &l));
        ctx->launches += l;
    }
    const int64_t base = ctx->db_n - kf_off[0];
    for (int k = 1; k <= n_kf; ++k) ctx->h_kf_off.push_back(kf_off[k] + base);
    CK(cudaMemcpyAsync(ctx->d_kf_off + ctx->n_kf, ctx->h_kf_off.data() + ctx->n_kf, sizeof(int64_t) * ((size_t)n_kf + 1),
                       cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->db_n += nd;
    ctx->n_kf += n_kf;
    ctx->tiles_dirty = true;
    return PSLAM_OK;
}

int pslam_lc_db_clear(pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    ctx->db_n = 0;
    ctx->n_kf = 0;
    ctx->h_kf_off.assign(1, 0);
    ctx->tiles_dirty = true;
    ctx->max_kf_desc = 0;
    return PSLAM_OK;
}

int pslam_lc_db_size(const pslam_ctx* ctx, int* n_keyframes, int64_t* n_descriptors) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (n_keyframes) *n_keyframes = ctx->n_kf;
    if (n_descriptors) *n_descriptors = ctx->db_n;
    return PSLAM_OK;
}

int pslam_lc_set_work_unit(pslam_ctx* ctx, int mode) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (mode < 0 || mode > 3) return fail(ctx, PSLAM_ERR_ARG, "work unit mode must be 0 .. 3");
    ctx->lc_work_unit = mode;
    ctx->tiles_dirty = true;
    return PSLAM_OK;
}

int pslam_lc_tensor_status(pslam_ctx* ctx, int* used_tensor_cores, int* timed_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (used_tensor_cores) *used_tensor_cores = ctx->lc_last_tensor ? 1 : 0;
    if (timed_out) {
        *timed_out = 0;
        if (ctx->d_tc_status) {
            CK(cudaSetDevice(ctx->device));
            CK(cudaMemcpyAsync(timed_out, ctx->d_tc_status, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    return PSLAM_OK;
}

int pslam_lc_set_id_base(pslam_ctx* ctx, int kf_id_base) {
    if (!ctx) return PSLAM_ERR_ARG;
    ctx->kf_id_base = kf_id_base;
    return PSLAM_OK;
}

static int lc_prepare(pslam_ctx* ctx, const uint8_t* query, int nq, int k) {
    if (nq <= 0 || nq > PSLAM_LC_MAX_QUERY) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "query size %d outside 1..%d", nq, PSLAM_LC_MAX_QUERY);
    if (k <= 0 || k > PSLAM_LC_MAX_TOPK) return fail(ctx, PSLAM_ERR_ARG, "k outside 1..%d", PSLAM_LC_MAX_TOPK);
    CK(cudaSetDevice(ctx->device));
    if (!ctx->lc_configured) {
        CK(lc_sweep_configure());
        CK(lc_sweep_range_configure());
        CK(lc_sweep_tc_configure());
        {
            const char* env = getenv("PSLAM_LC_TENSOR");
            ctx->lc_tensor = !(env && env[0] == '0');
        }
        CK(cudaMalloc((void**)&ctx->d_tc_status, sizeof(int)));
        CK(cudaMemsetAsync(ctx->d_tc_status, 0, sizeof(int), ctx->stream));
        CK(cudaMalloc((void**)&ctx->d_cta_done, sizeof(unsigned int)));
        CK(cudaMemsetAsync(ctx->d_cta_done, 0, sizeof(unsigned int), ctx->stream));
        CK(cudaMalloc((void**)&ctx->d_lc_query, (size_t)PSLAM_LC_MAX_QUERY * 32));
        CK(cudaMalloc((void**)&ctx->d_lc_pairs, sizeof(int) * 2 * PSLAM_LC_MAX_TOPK * (2 + 64)));
        CK(cudaEventCreate(&ctx->ev_sweep0));
        CK(cudaEventCreate(&ctx->ev_sweep1));
        ctx->lc_configured = true;
    }
    if (!ctx->d_scores) TRY(pslam_lc_db_reserve(ctx, 1, 1));
    if (query) {
        TRY(ensure_host(ctx, ctx->h_in, (size_t)nq * 32));
        memcpy(ctx->h_in.p, query, (size_t)nq * 32);
        CK(cudaMemcpyAsync(ctx->d_lc_query, ctx->h_in.p, (size_t)nq * 32, cudaMemcpyHostToDevice, ctx->stream));
        ctx->lc_query_in_xchg = false;
    }
    ctx->lc_nq = nq;
    return PSLAM_OK;
}

// tile prefix counts of the keyframes (host mirror -> device), recomputed after appends
static int lc_refresh_tiles(pslam_ctx* ctx) {
    if (!ctx->tiles_dirty) return PSLAM_OK;
    std::vector<int> ts((size_t)ctx->n_kf + 1, 0);
    for (int k = 0; k < ctx->n_kf; ++k)
        ts[(size_t)k + 1] = ts[(size_t)k] + (int)((ctx->h_kf_off[(size_t)k + 1] - ctx->h_kf_off[(size_t)k] + 127) / 128);
    ctx->n_tiles = ts[(size_t)ctx->n_kf];
    const bool keep_f2m = ctx->f2m.valid, keep_f2f = ctx->f2f.valid;   // these buffers are not the frame arenas
    TRY(ensure_dev(ctx, ctx->d_tile_start, sizeof(int) * ts.size()));
    if (ctx->lc_work_unit == 2) {
        const size_t o_col = (lc_split_rowpart_bytes(ctx->n_tiles) + 255) & ~(size_t)255;
        TRY(ensure_dev(ctx, ctx->d_split, o_col + lc_split_colmin_bytes(ctx->n_kf)));
    }
    ctx->f2m.valid = keep_f2m; ctx->f2f.valid = keep_f2f;
    CK(cudaMemcpyAsync(ctx->d_tile_start.p, ts.data(), sizeof(int) * ts.size(), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // ts is a stack-lifetime host buffer
    ctx->tiles_dirty = false;
    return PSLAM_OK;
}

static bool lc_fused_tail(const pslam_ctx* ctx) { return ctx->lc_work_unit == 0 || ctx->lc_work_unit == 3; }
static int* lc_merged_pairs(pslam_ctx* ctx) { return ctx->d_lc_pairs + 2 * PSLAM_LC_MAX_TOPK + 2 * PSLAM_LC_MAX_TOPK * 64; }

// One sweep of the resident query over this ctx's keyframes + local top-k.  exchange: the peer exchange and merge of the
// per-rank top-k run in the kernel's tail (range form with peer access only); pushed_query: the query sits in the exchange
// buffer (pushed by the root rank) and the kernel waits for its flag.
static int lc_enqueue_local(pslam_ctx* ctx, int tau, int k, bool exchange = false, bool pushed_query = false) {
    const bool query_in_xchg = pushed_query || (ctx->lc_query_in_xchg && lc_fused_tail(ctx) && ctx->d_xchg);
    int l = 0;
    if (ctx->lc_nq > 1024 && ctx->max_kf_desc > lc_max_kf_desc_wide())
        return fail(ctx, PSLAM_ERR_UNSUPPORTED, "more than 1024 query descriptors need keyframes of at most %d descriptors (largest: %d)",
                    lc_max_kf_desc_wide(), ctx->max_kf_desc);
    TRY(lc_refresh_tiles(ctx));
    if (lc_fused_tail(ctx)) {
        // one sweep + top-k (+ peer exchange) without leaving the device.  Tensor-core form (lc_tc.cuh) when the query fits
        // its four resident quarters; otherwise the range form: contiguous 128-row tile ranges per CTA, keyframes cut by a
        // range merged by the last piece, the last CTA computes the top-k -- one launch
        const bool tensor = ctx->lc_tensor && ctx->lc_work_unit == 0 && ctx->lc_nq <= lc_tc_max_query() && ctx->max_kf_desc <= lc_max_kf_desc();
        const int grid = lc_range_grid(ctx->lc_nq, ctx->n_tiles, ctx->sm_count);
        const size_t o_col = (lc_range_rowpart_bytes(grid) + 255) & ~(size_t)255;
        const bool keep_f2m = ctx->f2m.valid, keep_f2f = ctx->f2f.valid;
        if (tensor) {
            TRY(ensure_dev(ctx, ctx->d_tc_row, lc_tc_rowbest_bytes(ctx->n_kf)));
            TRY(ensure_dev(ctx, ctx->d_tc_col, lc_tc_colbest_bytes(ctx->db_n, ctx->lc_nq)));
        } else {
            TRY(ensure_dev(ctx, ctx->d_range, o_col + lc_range_colpart_bytes(grid)));
        }
        ctx->f2m.valid = keep_f2m; ctx->f2f.valid = keep_f2f;
        LcSweepArgs a;
        a.query = reinterpret_cast<const uint4*>(query_in_xchg ? (const uint8_t*)(ctx->d_xchg + kLcXchgQueryOff) : ctx->d_lc_query);
        a.nq = ctx->lc_nq;
        a.db = reinterpret_cast<const uint4*>(ctx->d_db); a.kf_off = ctx->d_kf_off; a.tile_start = (const int*)ctx->d_tile_start.p;
        a.n_kf = ctx->n_kf; a.n_tiles = ctx->n_tiles; a.tau = tau;
        a.scores = ctx->d_scores;
        a.rowpart = (uint32_t*)ctx->d_range.p; a.colpart = (uint32_t*)(ctx->d_range.p + o_col); a.kf_done = ctx->d_kf_done;
        a.cta_done = ctx->d_cta_done;
        a.kf_id_base = ctx->kf_id_base; a.k = k;
        a.out_pairs = ctx->d_lc_pairs; a.out_merged = lc_merged_pairs(ctx);
        a.x.world = ctx->world; a.x.rank = ctx->rank; a.x.epoch = ctx->lc_epoch;
        a.x.local = exchange ? ctx->d_xchg : nullptr;
        for (int r = 0; r < kLcMaxRanks; ++r) a.x.peer[r] = (exchange && r < ctx->world) ? ctx->peer_xchg[r] : nullptr;
        a.qflag = pushed_query ? reinterpret_cast<const uint32_t*>(ctx->d_xchg) + kLcXchgQFlagOff : nullptr;
        a.qepoch = ctx->lc_qepoch;
        CK(cudaEventRecord(ctx->ev_sweep0, ctx->stream));
        if (tensor) CK(launch_lc_sweep_tc(a, ctx->db_n, (uint32_t*)ctx->d_tc_row.p, (uint32_t*)ctx->d_tc_col.p, ctx->d_tc_status, ctx->sm_count, ctx->stream, &l));
        else CK(launch_lc_sweep_range(a, grid, ctx->stream, &l));
        CK(cudaEventRecord(ctx->ev_sweep1, ctx->stream));
        ctx->lc_last_tensor = tensor;
        ctx->launches += l;
        return PSLAM_OK;
    }
    // round-1 forms, kept behind pslam_lc_set_work_unit: whole keyframes per CTA (1) / 128-row tiles + finalize (2)
    ctx->lc_last_tensor = false;
    CK(cudaEventRecord(ctx->ev_sweep0, ctx->stream));
    if (ctx->lc_work_unit == 2) {
        const size_t o_col = (lc_split_rowpart_bytes(ctx->n_tiles) + 255) & ~(size_t)255;
        CK(launch_lc_sweep_split(ctx->d_lc_query, ctx->lc_nq, ctx->d_db, ctx->d_kf_off, (const int*)ctx->d_tile_start.p,
                                 ctx->n_kf, ctx->n_tiles, tau, (uint32_t*)ctx->d_split.p, (uint32_t*)(ctx->d_split.p + o_col),
                                 ctx->d_scores, ctx->sm_count, ctx->stream, &l));
    } else {
        CK(launch_lc_sweep(ctx->d_lc_query, ctx->lc_nq, ctx->d_db, ctx->d_kf_off, ctx->n_kf, tau, ctx->d_scores,
                           ctx->sm_count, ctx->stream, &l));
    }
    CK(cudaEventRecord(ctx->ev_sweep1, ctx->stream));
    CK(launch_lc_topk(ctx->d_scores, ctx->n_kf, ctx->kf_id_base, k, ctx->d_lc_pairs, ctx->stream, &l));
    ctx->launches += l;
    return PSLAM_OK;
}

static void unpack_pairs(const int* pairs, int k, int* ids, int* scores) {
    for (int i = 0; i < k; ++i) {
        scores[i] = pairs[2 * i];
        ids[i] = pairs[2 * i + 1];
    }
}

int pslam_lc_query(pslam_ctx* ctx, const uint8_t* query, int nq, int tau, int k, int* out_kf_ids, int* out_scores,
                   int* scores_out) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!query || !out_kf_ids || !out_scores) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_query: null buffer");
    TRY(lc_prepare(ctx, query, nq, k));
    TRY(lc_enqueue_local(ctx, tau, k));
    const size_t need = sizeof(int) * (2 * (size_t)k + (scores_out ? (size_t)ctx->n_kf : 0));
    TRY(ensure_host(ctx, ctx->h_out, need));
    CK(cudaMemcpyAsync(ctx->h_out.p, ctx->d_lc_pairs, sizeof(int) * 2 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
    if (scores_out && ctx->n_kf > 0)
        CK(cudaMemcpyAsync(ctx->h_out.p + sizeof(int) * 2 * (size_t)k, ctx->d_scores, sizeof(int) * (size_t)ctx->n_kf,
                           cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unpack_pairs((const int*)ctx->h_out.p, k, out_kf_ids, out_scores);
    if (scores_out && ctx->n_kf > 0) memcpy(scores_out, ctx->h_out.p + sizeof(int) * 2 * (size_t)k, sizeof(int) * (size_t)ctx->n_kf);
    return PSLAM_OK;
}

int pslam_lc_last_sweep_ms(pslam_ctx* ctx, float* ms_out) {
    if (!ctx || !ms_out) return PSLAM_ERR_ARG;
    if (!ctx->lc_configured) return fail(ctx, PSLAM_ERR_ARG, "no sweep has run on this ctx");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev_sweep1));
    CK(cudaEventElapsedTime(ms_out, ctx->ev_sweep0, ctx->ev_sweep1));
    return PSLAM_OK;
}

int pslam_lc_query_resident(pslam_ctx* ctx, int tau, int k) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->lc_nq <= 0 || !ctx->lc_configured) return fail(ctx, PSLAM_ERR_ARG, "no resident query");
    if (k <= 0 || k > PSLAM_LC_MAX_TOPK) return fail(ctx, PSLAM_ERR_ARG, "k outside 1..%d", PSLAM_LC_MAX_TOPK);
    CK(cudaSetDevice(ctx->device));
    return lc_enqueue_local(ctx, tau, k);
}

// ---- NCCL ----------------------------------------------------------------------------------------
int pslam_comm_unique_id(uint8_t id_out[128]) {
    NcclApi* api = nccl_api();
    if (!api->ok || !id_out) return PSLAM_ERR_NCCL;
    nccl_uid uid;
    if (api->GetUniqueId(&uid) != 0) return PSLAM_ERR_NCCL;
    memcpy(id_out, uid.internal, 128);
    return PSLAM_OK;
}

// Peer exchange buffers of the sharded sweep: every rank allocates one (cudaMalloc), the CUDA IPC handles travel through
// ncclAllGather, every rank opens its peers' buffers (NVLink / NVSwitch peer access) and the ranks agree -- again through an
// all-gather -- whether ALL of them succeeded.  PSLAM_LC_P2P=0 in the environment keeps the NCCL path.
static void lc_peer_teardown(pslam_ctx* ctx) {
    for (int r = 0; r < 64; ++r) {
        if (ctx->peer_xchg[r] && r != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_xchg[r]);
        ctx->peer_xchg[r] = nullptr;
    }
    if (ctx->d_xchg) cudaFree(ctx->d_xchg);
    ctx->d_xchg = nullptr;
    ctx->p2p = false;
    ctx->lc_query_in_xchg = false;
}
static int lc_peer_setup(pslam_ctx* ctx) {
    lc_peer_teardown(ctx);
    ctx->lc_epoch = 0; ctx->lc_qepoch = 0;
    if (ctx->world <= 1) return PSLAM_OK;
    NcclApi* api = nccl_api();
    const char* env = getenv("PSLAM_LC_P2P");
    int want = !(env && env[0] == '0');
    const int W = ctx->world;
    uint8_t* d_tmp = nullptr;
    if (cudaMalloc((void**)&d_tmp, 128 * (size_t)(W + 1)) != cudaSuccess) return PSLAM_OK;
    std::vector<uint8_t> h(128 * (size_t)(W + 1), 0);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (want) {
        if (cudaMalloc((void**)&ctx->d_xchg, sizeof(int) * kLcXchgWords) != cudaSuccess) { ctx->d_xchg = nullptr; want = 0; }
        else if (cudaMemset(ctx->d_xchg, 0, sizeof(int) * kLcXchgWords) != cudaSuccess || cudaIpcGetMemHandle(&mine, ctx->d_xchg) != cudaSuccess) want = 0;
        cudaGetLastError();
    }
    // round 1: handles (64 bytes) + the rank's own verdict so far
    memcpy(h.data(), &mine, sizeof(mine));
    h[64] = (uint8_t)want;
    bool ok = cudaMemcpyAsync(d_tmp, h.data(), 128, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
              api->AllGather(d_tmp, d_tmp + 128, 128, kNcclUint8, ctx->comm, ctx->stream) == 0 &&
              cudaMemcpyAsync(h.data() + 128, d_tmp + 128, 128 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
              cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    int all = ok ? 1 : 0;
    for (int r = 0; r < W && all; ++r) all = h[128 + 128 * (size_t)r + 64] ? 1 : 0;
    int opened = all;
    if (all) {
        for (int r = 0; r < W; ++r) {
            if (r == ctx->rank) { ctx->peer_xchg[r] = ctx->d_xchg; continue; }
            cudaIpcMemHandle_t hr;
            memcpy(&hr, h.data() + 128 + 128 * (size_t)r, sizeof(hr));
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
            ctx->peer_xchg[r] = (int*)p;
        }
    }
    // round 2: did every rank open every buffer?
    if (ok) {
        h[0] = (uint8_t)opened;
        ok = cudaMemcpyAsync(d_tmp, h.data(), 128, cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
             api->AllGather(d_tmp, d_tmp + 128, 128, kNcclUint8, ctx->comm, ctx->stream) == 0 &&
             cudaMemcpyAsync(h.data() + 128, d_tmp + 128, 128 * (size_t)W, cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
             cudaStreamSynchronize(ctx->stream) == cudaSuccess;
        for (int r = 0; r < W && ok; ++r) if (!h[128 + 128 * (size_t)r]) opened = 0;
    }
    cudaFree(d_tmp);
    if (!ok || !opened) { lc_peer_teardown(ctx); return PSLAM_OK; }
    ctx->p2p = true;
    return PSLAM_OK;
}

int pslam_comm_init(pslam_ctx* ctx, const uint8_t id[128], int rank, int world) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!id || world < 1 || rank < 0 || rank >= world || world > 64) return fail(ctx, PSLAM_ERR_ARG, "pslam_comm_init: bad argument");
    NcclApi* api = nccl_api();
    if (!api->ok) return fail(ctx, PSLAM_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CK(cudaSetDevice(ctx->device));
    if (ctx->comm) { api->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    nccl_uid uid;
    memcpy(uid.internal, id, 128);
    const int r = api->CommInitRank(&ctx->comm, world, uid, rank);
    if (r != 0) return fail(ctx, PSLAM_ERR_NCCL, "ncclCommInitRank: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
    ctx->rank = rank;
    ctx->world = world;
    lc_peer_setup(ctx);          // best effort: without peer access the NCCL all-gather path is used
    return PSLAM_OK;
}

int pslam_lc_exchange_mode(const pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->world <= 1) return 0;
    return (ctx->p2p && lc_fused_tail(ctx)) ? 2 : 1;
}

int pslam_comm_destroy(pslam_ctx* ctx) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->comm && nccl_api()->ok) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        lc_peer_teardown(ctx);
        nccl_api()->CommDestroy(ctx->comm);
    }
    ctx->comm = nullptr;
    ctx->world = 1;
    ctx->rank = 0;
    return PSLAM_OK;
}

static int lc_enqueue_sharded(pslam_ctx* ctx, int root, int tau, int k) {
    NcclApi* api = nccl_api();
    if (ctx->world <= 1) return lc_enqueue_local(ctx, tau, k);
    const bool peer = ctx->p2p && lc_fused_tail(ctx);
    ++ctx->lc_epoch;
    bool pushed = false;
    if (root >= 0) {
        if (peer) {
            // the root writes the query into every rank's exchange buffer over NVLink and raises the query flags; the sweep
            // kernels wait for the flag (no collective launch on the critical path)
            ++ctx->lc_qepoch;
            pushed = true;
            ctx->lc_query_in_xchg = true;
            if (ctx->rank == root) {
                LcExchange x;
                x.world = ctx->world; x.rank = ctx->rank; x.epoch = ctx->lc_epoch; x.local = ctx->d_xchg;
                for (int r = 0; r < kLcMaxRanks; ++r) x.peer[r] = r < ctx->world ? ctx->peer_xchg[r] : nullptr;
                int l = 0;
                CK(launch_lc_push_query(ctx->d_lc_query, ctx->lc_nq, x, ctx->lc_qepoch, ctx->stream, &l));
                ctx->launches += l;
            }
        } else {
            const int r = api->Broadcast(ctx->d_lc_query, ctx->d_lc_query, (size_t)ctx->lc_nq * 32, kNcclUint8, root, ctx->comm, ctx->stream);
            if (r != 0) return fail(ctx, PSLAM_ERR_NCCL, "ncclBroadcast: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
            ctx->lc_query_in_xchg = false;
        }
    }
    TRY(lc_enqueue_local(ctx, tau, k, peer, pushed));
    if (!peer) {
        int* local = ctx->d_lc_pairs;
        int* gathered = ctx->d_lc_pairs + 2 * PSLAM_LC_MAX_TOPK;
        int* merged = lc_merged_pairs(ctx);
        const int r = api->AllGather(local, gathered, 2 * (size_t)k, kNcclInt32, ctx->comm, ctx->stream);
        if (r != 0) return fail(ctx, PSLAM_ERR_NCCL, "ncclAllGather: %s", api->GetErrorString ? api->GetErrorString(r) : "error");
        int l = 0;
        CK(launch_lc_merge_topk(gathered, ctx->world * k, k, merged, ctx->stream, &l));
        ctx->launches += l;
    }
    return PSLAM_OK;
}

int pslam_lc_query_sharded(pslam_ctx* ctx, const uint8_t* query, int nq, int root, int tau, int k, int* out_kf_ids,
                           int* out_scores) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!out_kf_ids || !out_scores) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_query_sharded: null buffer");
    if (ctx->world > 1 && !ctx->comm) return fail(ctx, PSLAM_ERR_NCCL, "communicator not initialised");
    if (ctx->world * k > 1024) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "world*k above 1024");
    const bool have_query = (root < 0) || (root == ctx->rank) || ctx->world == 1;
    if (have_query && !query) return fail(ctx, PSLAM_ERR_ARG, "query is NULL on a rank that must supply it");
    TRY(lc_prepare(ctx, have_query ? query : nullptr, nq, k));
    TRY(lc_enqueue_sharded(ctx, root, tau, k));
    const int* src = ctx->world > 1 ? ctx->d_lc_pairs + 2 * PSLAM_LC_MAX_TOPK + 2 * PSLAM_LC_MAX_TOPK * 64 : ctx->d_lc_pairs;
    TRY(ensure_host(ctx, ctx->h_out, sizeof(int) * 2 * (size_t)k));
    CK(cudaMemcpyAsync(ctx->h_out.p, src, sizeof(int) * 2 * (size_t)k, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    unpack_pairs((const int*)ctx->h_out.p, k, out_kf_ids, out_scores);
    return PSLAM_OK;
}

int pslam_lc_query_sharded_resident(pslam_ctx* ctx, int tau, int k) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->lc_nq <= 0 || !ctx->lc_configured) return fail(ctx, PSLAM_ERR_ARG, "no resident query");
    if (ctx->world > 1 && !ctx->comm) return fail(ctx, PSLAM_ERR_NCCL, "communicator not initialised");
    CK(cudaSetDevice(ctx->device));
    return lc_enqueue_sharded(ctx, -1, tau, k);
}

int pslam_lc_query_sharded_resident_bcast(pslam_ctx* ctx, int root, int tau, int k) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->lc_nq <= 0 || !ctx->lc_configured) return fail(ctx, PSLAM_ERR_ARG, "no resident query");
    if (ctx->world > 1 && !ctx->comm) return fail(ctx, PSLAM_ERR_NCCL, "communicator not initialised");
    if (root >= ctx->world) return fail(ctx, PSLAM_ERR_ARG, "root outside the communicator");
    CK(cudaSetDevice(ctx->device));
    return lc_enqueue_sharded(ctx, root, tau, k);
}

int pslam_lc_set_desc_base(pslam_ctx* ctx, int64_t desc_id_base) {
    if (!ctx) return PSLAM_ERR_ARG;
    ctx->desc_id_base = desc_id_base;
    return PSLAM_OK;
}

// layout of d_knn: [partials grid*nq*16][keys nq*16][gathered world*nq*16][idx nq*16][dist nq*8]
static int lc_knn2_run(pslam_ctx* ctx, const uint8_t* query, int nq, int root, bool sharded, int64_t* out_idx,
                       float* out_dist, bool copy_out) {
    if (nq <= 0 || nq > PSLAM_LC_MAX_QUERY) return fail(ctx, PSLAM_ERR_UNSUPPORTED, "query size %d outside 1..%d", nq, PSLAM_LC_MAX_QUERY);
    TRY(lc_prepare(ctx, query, nq, 1));
    const int world = sharded ? ctx->world : 1;
    const bool tensor = ctx->lc_tensor && ctx->lc_work_unit == 0 && nq <= lc_tc_max_query();
    const int grid = tensor ? lc_knn2_tc_parts(nq, ctx->sm_count) : lc_knn2_grid(ctx->db_n, ctx->sm_count);   // partial results per query
    Arena A;
    const size_t o_part = A.take(16 * (size_t)grid * nq), o_keys = A.take(16 * (size_t)nq);
    const size_t o_gath = A.take(16 * (size_t)world * nq), o_idx = A.take(16 * (size_t)nq), o_dist = A.take(8 * (size_t)nq);
    TRY(ensure_dev(ctx, ctx->d_knn, A.off));
    uint8_t* d = ctx->d_knn.p;
    NcclApi* api = nccl_api();
    int l = 0;
    if (world > 1 && root >= 0) {
        const int r = api->Broadcast(ctx->d_lc_query, ctx->d_lc_query, (size_t)nq * 32, kNcclUint8, root, ctx->comm, ctx->stream);
        if (r != 0) return fail(ctx, PSLAM_ERR_NCCL, "ncclBroadcast failed");
    }
    if (ctx->db_n > 0) {
        if (tensor) CK(launch_lc_knn2_tc(ctx->d_lc_query, nq, ctx->d_db, ctx->db_n, ctx->desc_id_base, d + o_part, ctx->d_tc_status, ctx->sm_count, ctx->stream, &l));
        else CK(launch_lc_knn2(ctx->d_lc_query, nq, ctx->d_db, ctx->db_n, ctx->desc_id_base, d + o_part, grid, ctx->sm_count, ctx->stream, &l));
    }
    ctx->lc_last_tensor = tensor && ctx->db_n > 0;
    CK(launch_lc_knn2_merge(d + o_part, ctx->db_n > 0 ? grid : 0, nq, (unsigned long long*)(d + o_keys),
                            world > 1 ? nullptr : (long long*)(d + o_idx), (float*)(d + o_dist), ctx->stream, &l));
    if (world > 1) {
        const int r = api->AllGather(d + o_keys, d + o_gath, 16 * (size_t)nq, kNcclUint8, ctx->comm, ctx->stream);
        if (r != 0) return fail(ctx, PSLAM_ERR_NCCL, "ncclAllGather failed");
        CK(launch_lc_knn2_merge(d + o_gath, world, nq, nullptr, (long long*)(d + o_idx), (float*)(d + o_dist), ctx->stream, &l));
    }
    ctx->launches += l;
    if (!copy_out) return PSLAM_OK;
    TRY(ensure_host(ctx, ctx->h_out, 24 * (size_t)nq));
    CK(cudaMemcpyAsync(ctx->h_out.p, d + o_idx, 16 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_out.p + 16 * (size_t)nq, d + o_dist, 8 * (size_t)nq, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    memcpy(out_idx, ctx->h_out.p, 16 * (size_t)nq);
    memcpy(out_dist, ctx->h_out.p + 16 * (size_t)nq, 8 * (size_t)nq);
    return PSLAM_OK;
}

int pslam_lc_knn2(pslam_ctx* ctx, const uint8_t* query, int nq, int64_t* out_idx, float* out_dist) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!query || !out_idx || !out_dist) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_knn2: null buffer");
    return lc_knn2_run(ctx, query, nq, -1, false, out_idx, out_dist, true);
}

int pslam_lc_knn2_sharded(pslam_ctx* ctx, const uint8_t* query, int nq, int root, int64_t* out_idx, float* out_dist) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (!out_idx || !out_dist) return fail(ctx, PSLAM_ERR_ARG, "pslam_lc_knn2_sharded: null buffer");
    if (ctx->world > 1 && !ctx->comm) return fail(ctx, PSLAM_ERR_NCCL, "communicator not initialised");
    const bool have_query = (root < 0) || (root == ctx->rank) || ctx->world == 1;
    if (have_query && !query) return fail(ctx, PSLAM_ERR_ARG, "query is NULL on a rank that must supply it");
    return lc_knn2_run(ctx, have_query ? query : nullptr, nq, root, true, out_idx, out_dist, true);
}

int pslam_lc_knn2_resident(pslam_ctx* ctx, int sharded) {
    if (!ctx) return PSLAM_ERR_ARG;
    if (ctx->lc_nq <= 0 || !ctx->lc_configured) return fail(ctx, PSLAM_ERR_ARG, "no resident query");
    if (sharded && ctx->world > 1 && !ctx->comm) return fail(ctx, PSLAM_ERR_NCCL, "communicator not initialised");
    return lc_knn2_run(ctx, nullptr, ctx->lc_nq, -1, sharded != 0, nullptr, nullptr, false);
}

}  // extern "C"
