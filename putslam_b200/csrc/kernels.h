// kernels.h -- host-callable launchers of the sm_100a kernels (internal to libpslam_b200.so).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pslam_b200.h"
#include "klt_point.cuh"

namespace pslam {

// ---- hamming.cu ------------------------------------------------------------------------------
// d_out: int[1 + 3*cap] = {n, queryIdx[cap], trainIdx[cap], distance(float)[cap]}
cudaError_t launch_bf_mutual(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, uint32_t* d_rowmin,
                             uint32_t* d_colmin, int* d_out, int cap, int sm_count, cudaStream_t st, int* launches);
int knn2_parts(int nq, int nt, int sm_count);
cudaError_t launch_knn2(const uint8_t* d_query, int nq, const uint8_t* d_train, int nt, uint2* d_partial, int* d_idx,
                        float* d_dist, int sm_count, cudaStream_t st, int* launches);
cudaError_t lc_sweep_configure();
int lc_max_kf_desc();
int lc_max_query();
int lc_max_kf_desc_wide();
cudaError_t launch_lc_sweep(const uint8_t* d_query, int nq, const uint8_t* d_db, const int64_t* d_kf_off, int n_kf,
                            int tau, int* d_scores, int sm_count, cudaStream_t st, int* launches);
// split form (tile-granular work units) for small maps; tile_start[n_kf + 1] = prefix count of 128-row tiles
size_t lc_split_rowpart_bytes(int n_tiles);
size_t lc_split_colmin_bytes(int n_kf);
cudaError_t launch_lc_sweep_split(const uint8_t* d_query, int nq, const uint8_t* d_db, const int64_t* d_kf_off,
                                  const int* d_tile_start, int n_kf, int n_tiles, int tau, uint32_t* d_rowpart,
                                  uint32_t* d_colmin, int* d_scores, int sm_count, cudaStream_t st, int* launches);
// ---- lc_sweep.cu: range-form sweep with the fused tail (local top-k, peer exchange, merge) ---------------------------
// Exchange buffer of one rank (cudaMalloc'ed, opened by the peers through CUDA IPC), in 4-byte words:
//   [kLcXchgPairsOff ...)  pairs  [2 banks][kLcMaxRanks][2 * kLcMaxTopk] int   {score, id} written by rank r into slot r
//   [kLcXchgFlagsOff ...)  flags  [2 banks][kLcMaxRanks] uint32               = epoch of the query the slot belongs to
//   [kLcXchgQFlagOff]      qflag  uint32                                       = epoch of the query in the query buffer
//   [kLcXchgQueryOff ...)  query  2048 x 32 B                                  pushed by the root rank
constexpr int kLcMaxRanks = 64;
constexpr int kLcMaxTopk = 64;
constexpr size_t kLcXchgPairsOff = 0;
constexpr size_t kLcXchgFlagsOff = kLcXchgPairsOff + 2 * (size_t)kLcMaxRanks * 2 * kLcMaxTopk;
constexpr size_t kLcXchgQFlagOff = kLcXchgFlagsOff + 2 * (size_t)kLcMaxRanks;
constexpr size_t kLcXchgQueryOff = ((kLcXchgQFlagOff + 1 + 63) / 64) * 64;           // 256-byte aligned
constexpr size_t kLcXchgWords = kLcXchgQueryOff + 2048 * 8;
struct LcExchange {
    int world = 1, rank = 0;
    uint32_t epoch = 0;              // of this query; bank = epoch & 1
    int* local = nullptr;            // this rank's exchange buffer (null: no peer exchange, NCCL path)
    int* peer[kLcMaxRanks];          // every rank's buffer as seen from this GPU (peer[rank] == local)
};
struct LcSweepArgs {
    const uint4* query; int nq;
    const uint4* db; const int64_t* kf_off; const int* tile_start; int n_kf, n_tiles, tau;
    int* scores;
    uint32_t* rowpart; uint32_t* colpart; int* kf_done;      // pieces of the keyframes cut by a CTA's tile range
    unsigned int* cta_done;                                  // grid-wide completion counter (zero between launches)
    int kf_id_base, k;
    int* out_pairs;                                          // local top-k: k x {score, id}
    int* out_merged;                                         // global top-k after the peer exchange
    LcExchange x;
    const uint32_t* qflag; uint32_t qepoch;                  // wait for the pushed query (null: the query is already here)
};
cudaError_t lc_sweep_range_configure();
int lc_range_grid(int nq, int n_tiles, int sm_count);
size_t lc_range_rowpart_bytes(int grid);
size_t lc_range_colpart_bytes(int grid);
cudaError_t launch_lc_sweep_range(const LcSweepArgs& a, int grid, cudaStream_t st, int* launches);
// tensor-core form (lc_tc.cuh): tcgen05 kind::i8 sweep into per-split partial results + finalize with the same fused tail.
// Needs nq <= lc_tc_max_query() and keyframes of at most lc_max_kf_desc() rows.  d_status: one int, non-zero after a time-out.
cudaError_t lc_sweep_tc_configure();
int lc_tc_max_query();
size_t lc_tc_rowbest_bytes(int n_kf);
size_t lc_tc_colbest_bytes(long long n_desc, int nq);
cudaError_t launch_lc_sweep_tc(const LcSweepArgs& a, long long n_desc, uint32_t* d_rowbest, uint32_t* d_colbest, int* d_status,
                               int sm_count, cudaStream_t st, int* launches);
int lc_knn2_tc_parts(int nq, int sm_count);
cudaError_t launch_lc_knn2_tc(const uint8_t* d_query, int nq, const uint8_t* d_db, long long n_desc, long long desc_id_base,
                              void* d_partial, int* d_status, int sm_count, cudaStream_t st, int* launches);
cudaError_t launch_lc_push_query(const uint8_t* d_query, int nq, const LcExchange& x, uint32_t qepoch, cudaStream_t st, int* launches);
// in-place re-encoding of n descriptor rows (32 B each) for the encoded Hamming compare of the sweep kernels
cudaError_t launch_lc_encode_rows(uint8_t* d_rows, long long n, cudaStream_t st, int* launches);
cudaError_t launch_lc_topk(const int* d_scores, int n_kf, int kf_id_base, int k, int* d_out_pairs, cudaStream_t st,
                           int* launches);
cudaError_t launch_lc_merge_topk(const int* d_gathered, int n_pairs, int k, int* d_out_pairs, cudaStream_t st,
                                 int* launches);

int lc_knn2_grid(long long n_desc, int sm_count);
// d_partial: grid x nq ulonglong2.  Keys are dist << 40 | global descriptor index.
cudaError_t launch_lc_knn2(const uint8_t* d_query, int nq, const uint8_t* d_db, long long n_desc, long long desc_id_base,
                           void* d_partial, int grid, int sm_count, cudaStream_t st, int* launches);
cudaError_t launch_lc_knn2_merge(const void* d_partial, int nparts, int nq, unsigned long long* d_keys, long long* d_idx,
                                 float* d_dist, cudaStream_t st, int* launches);

// ---- guided.cu -------------------------------------------------------------------------------
// d_out: int[2 + 3*cap] = {n_total, perfect, queryIdx[cap], trainIdx[cap], distance(float)[cap]}
cudaError_t launch_predict_levels(const double* d_map_xyz, const int* d_map_oct, const double* d_map_det, int M,
                                  const float* d_cur_xyz, const int* d_cur_oct, const double* d_cur_det, int N,
                                  const double* pow_tab, const int* lvl_tab, double log_sf, float* d_map_xyz_f,
                                  int* d_map_level, int* d_cur_level, cudaStream_t st, int* launches,
                                  const int* d_M = nullptr);   // d_M: device-resident feature count, M = capacity
size_t guided_cache_bytes(int M);
cudaError_t launch_guided_match(const float* d_map_xyz, const uint8_t* d_map_desc, const int* d_map_level, int M,
                                const float* d_cur_xyz, const uint8_t* d_cur_desc, const int* d_cur_level, int N,
                                float sq_radius_f, double accept_ratio, int mode, int* d_count, int* d_best, void* d_cache,
                                int* d_out, int cap, unsigned long long* d_cta_sums /* >= 2 x sm_count slots, zeroed once */,
                                unsigned int epoch /* != 0, different on every launch */, int sm_count, cudaStream_t st,
                                int* launches, const int* d_M = nullptr);

// ---- ransac.cu -------------------------------------------------------------------------------
// result header of the RANSAC pipeline (ints): [0] n_inliers  [1] hyp_used, or -1 = "finish on the host" (adaptive
// reference rule: hyp_used = max(last + 1, min([24], iterations(best_ratio))), evaluated with the host libm like the
// reference)  [2] n_filtered  [3] best_count  [4..19] T (col-major float)  [20..21] best_ratio (double)
// [22] winner hypothesis  [23] flags  [24] iterations(minimalInlierRatioThreshold)  [25] last improving hypothesis
// [26] hypotheses scored  [27] spare  [28..] inlier match indices
constexpr int kRansacHdrInts = 28;

struct RansacDeviceParams {
    int error_version;
    float thr_euclid_f;       // smallest float >= inlierThresholdEuclidean (float<double compare, exact)
    float sq_thr_euclid_f;    // smallest float T with sqrtf(T) >= thr_euclid_f (squared-norm form of the same test)
    double thr_euclid;
    double thr_reproj;
    double min_inlier_ratio;
    int min_matches;
    float fx, fy, cx, cy;
    uint32_t seed_lo, seed_hi;
    int num_hyp;              // 0 = adaptive (reference bound 487, shrinking), >0 fixed
    int stop_rule;            // 0 = reference RANSAC rule, 1 = USAC standard stopping (capped by the budget)
    double usac_conf;
    int iters_min_ratio;      // computeRANSACIteration(minimalInlierRatioThreshold), evaluated on the host
};
// Hypotheses scored per call.  Fixed mode: num_hyp.  Adaptive mode: the reference's loop starts with the bound
// computeRANSACIteration(0.20) = 487 (RANSAC.cpp:30) and, after the first improvement, never exceeds
// computeRANSACIteration(minimalInlierRatioThreshold) (:450-453) -- which is larger than 487 for thresholds below 0.2.
// All of them are scored in one launch (capped at 2^20), the replay then decides how many "were run".
// USAC stopping without an explicit budget keeps the 487 cap.
inline int ransac_hypothesis_budget(const RansacDeviceParams& P) {
    if (P.num_hyp > 0) return P.num_hyp;
    int h = 487;
    if (P.stop_rule != 1 && P.iters_min_ratio > h) h = P.iters_min_ratio;
    return h < (1 << 20) ? h : (1 << 20);
}
struct RansacWorkspace {
    // all device pointers; sized for m_cap matches and h_cap hypotheses
    float* pts;        // 6 * m_cap : filtered prev xyz | cur xyz as SoA px,py,pz,cx,cy,cz
    int* keep;         // m_cap     : original match index of filtered match k
    int* n_filtered;   // 1
    int* counts;       // h_cap
    float* models;     // 12 * h_cap : per-hypothesis R (row-major) and t
    int* result;       // header (see ransac.cu) + inlier list (m_cap)
    int m_cap, h_cap;
    // optional: when the selection kernel has written the result it copies out_bytes of the device output arena (match list
    // + result) into page-locked host memory itself (posted writes over PCIe), so the call needs no device->host copy
    void* out_host = nullptr; const void* out_dev = nullptr; size_t out_bytes = 0;
    // layout hint: the arena starts with a match list {total, perfect, q[cap], t[cap], d[cap]} and holds the result at
    // int offset out_res_ints -- only the used prefixes are copied (out_cap == 0: the whole arena)
    int out_cap = 0, out_res_ints = 0;
};
size_t ransac_result_ints(int m_cap);
cudaError_t launch_ransac(const float* d_prev, const float* d_cur, const int* d_mq, const int* d_mt,
                          const int* d_m /* device count, may be NULL */, int m_host, const RansacDeviceParams& P,
                          const RansacWorkspace& ws, int sm_count, cudaStream_t st, int* launches);

// ---- backproject.cu --------------------------------------------------------------------------
cudaError_t launch_backproject(const float* d_uv, int n, const uint16_t* d_depth, int W, int H, int stride,
                               const pslam_camera& cam, int undistort, double depth_scale, float* d_uv_und,
                               float* d_xyz, double* d_det_dist, double* d_cov, const pslam_cov_params* cov,
                               cudaStream_t st, int* launches);

cudaError_t launch_normal_cov(const int* d_px, int n, const uint16_t* d_depth, int W, int H, int stride,
                              const pslam_camera& cam, double depth_scale, double scale_unc, double* d_normals,
                              double* d_cov, double* d_info, cudaStream_t st, int* launches);
// diag16: the host-libm direction table of the diagonals (see gradient_cov_kernel)
cudaError_t launch_gradient_cov(const int* d_px, int n, const uint8_t* d_rgb, int rgb_row_bytes, const uint16_t* d_depth,
                                int W, int H, int stride, const pslam_camera& cam, double depth_scale, double scale_unc,
                                const int* diag16, double* d_grads, double* d_cov, double* d_info, cudaStream_t st,
                                int* launches);
cudaError_t launch_information(const double* d_uvz, int n, const pslam_cov_params& cp, double* d_cov, double* d_info,
                               cudaStream_t st, int* launches);

// ---- kabsch.cu -------------------------------------------------------------------------------
cudaError_t launch_kabsch_batch(const double* d_A, const double* d_B, const int* d_off, int batch, double* d_T,
                                cudaStream_t st, int* launches);

// ---- uncertainty.cu --------------------------------------------------------------------------
// mode 0: Euler angles (computeUncertainty), 1: quaternion vector part (computeUncertaintyG2O); see uncertainty.cu
cudaError_t launch_uncertainty_batch(const double* d_A, const double* d_B, const double* d_CA, const double* d_CB,
                                     const int* d_off, const double* d_T, int batch, int mode, double* d_U, int* d_ok,
                                     cudaStream_t st, int* launches);

// ---- mapprep.cu ------------------------------------------------------------------------------
cudaError_t launch_map_prepare(const double* d_xyz, const float* d_view_axis, int M, const double* pose_colmajor,
                               double fx, double fy, double cx, double cy, double img_w, double img_h, double max_angle,
                               double max_z, int* d_kept, double* d_xyz_local, double* d_uv, double* d_angles, int* d_n,
                               unsigned long long* d_cta_counts /* >= sm_count slots, zeroed once */,
                               unsigned int epoch /* != 0, different on every launch */, int sm_count, cudaStream_t st,
                               int* launches, const uint8_t* d_desc_in = nullptr,
                               uint8_t* d_desc_out = nullptr, const int* d_oct_in = nullptr, int* d_oct_out = nullptr,
                               const double* d_det_in = nullptr, double* d_det_out = nullptr);

// ---- orb.cu ----------------------------------------------------------------------------------
constexpr int kOrbMaxLevels = 12;
struct OrbPlan {   // geometry of the pyramid of one image size (offsets in bytes / elements from the buffer bases)
    int n;
    int w[kOrbMaxLevels], h[kOrbMaxLevels];
    int plain_off[kOrbMaxLevels], ext_off[kOrbMaxLevels], tab_off[kOrbMaxLevels];
    int pix_start[kOrbMaxLevels + 1], ext_start[kOrbMaxLevels + 1];
    size_t plain_bytes, ext_bytes, row_floats, tab_ints;
};
float orb_level_scale(int level);
size_t orb_plan(int W, int H, int nlevels, OrbPlan* P);
void orb_fill_tables(const OrbPlan& P, int* tab /* P.tab_ints */);
cudaError_t orb_upload_constants(void* d_pattern /* 1024 bytes */, cudaStream_t st);
// detection: candidates = 6 ints each {level, x, y, FAST score, Harris bits, angle bits}, raster order per level;
// d_header[0] = number found (may exceed cap; only cap are stored).  d_bgr / rgb_order as orb_gray_kernel.
cudaError_t launch_orb_detect(const uint8_t* d_bgr, int rgb_order, int W, int H, int row_bytes, const OrbPlan& P,
                              uint8_t* d_plain, uint8_t* d_score, const int* d_tab, int fast_threshold, int* d_cand, int cap,
                              int* d_header, unsigned long long* d_cta_counts, unsigned int epoch, int sm_count,
                              cudaStream_t st, int* launches);
cudaError_t launch_fast_detect(const uint8_t* d_bgr, int rgb_order, int W, int H, int row_bytes, const OrbPlan& P,
                               uint8_t* d_plain, uint8_t* d_score, int threshold, int* d_cand, int cap, int* d_header,
                               unsigned long long* d_cta_counts, unsigned int epoch, int sm_count, cudaStream_t st,
                               int* launches);
cudaError_t launch_orb_describe(const uint8_t* d_bgr, int W, int H, int row_bytes, const OrbPlan& P, uint8_t* d_plain,
                                uint8_t* d_ext, float* d_rowbuf, const int* d_tab, const void* d_pattern, const int* d_rec,
                                int n_kp, uint8_t* d_desc, cudaStream_t st, int* launches);

// ---- klt.cu ----------------------------------------------------------------------------------
// level 0 of each pyramid holds the frame; builds levels 1.. of the current frame (d_pyrJ) and, unless d_pyrI is
// nullptr, of the previous frame
cudaError_t launch_klt_pyramid(uint8_t* d_pyrI, uint8_t* d_pyrJ, const KltPlan& P, int cn, cudaStream_t st, int* launches);
cudaError_t launch_klt_track(const KltParams& P, const float* d_prev_xy, float* d_cur_xy, int n, uint8_t* d_status,
                             float* d_err, cudaStream_t st, int* launches);
// sq_thr: smallest double whose square root reaches the distance threshold; lim: smallest float >= that threshold
cudaError_t launch_klt_prune(const float* d_xy, const float* d_err, const uint8_t* d_status, int n, double err_thr,
                             double sq_thr, float lim, uint8_t* d_keep, cudaStream_t st, int* launches);
// d_mout: int[1 + 3 * cap] = {m, kept[cap], j[cap], 0[cap]} (match-list layout of the RANSAC launcher), d_cxy: cap x 2
cudaError_t launch_klt_compact(const uint8_t* d_keep, const float* d_xy, int n, int cap, int* d_mout, float* d_cxy,
                               cudaStream_t st, int* launches);

}  // namespace pslam
