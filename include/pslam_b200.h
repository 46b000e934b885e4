/*
 * pslam_b200.h -- C ABI of libpslam_b200.so: PUTSLAM's data-parallel front-end hot path
 * (depth back-projection -> Hamming matching -> RANSAC/Umeyama/Kabsch) on one B200 (sm_100a).
 *
 * The reference (LRMPUT/PUTSLAM) has no FFI layer; its seams are C++ virtuals and free functions.
 * Each entry point below names the reference interface it replaces (file:line relative to the
 * reference root); adapter/ holds the C++ classes that put these calls behind the reference's own
 * signatures, INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (PSLAM_OK) or a negative pslam_status; pslam_last_error(ctx) has text.
 *   - all pointers are HOST pointers owned by the caller unless the name says `_resident`;
 *     device memory, the CUDA stream and pinned staging live inside pslam_ctx.
 *   - one pslam_ctx per reference Matcher instance (tracking thread / loop-closure thread); calls on
 *     one ctx are not re-entrant, separate ctxs are independent.  No global mutable state.
 *   - there is NO CPU fallback: without a usable sm_100 device pslam_ctx_create fails.
 *   - descriptors are 32 bytes (ORB / LDB-256, reference src/LDB/ldb.cpp:61,657), rows contiguous.
 *   - 4x4 transforms are column-major float[16] (Eigen::Matrix4f layout); 3x4 Kabsch results are
 *     column-major double[12] (Eigen::Transform<double,3,Affine>::matrix().topRows(3) layout).
 */
#ifndef PSLAM_B200_H_
#define PSLAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PSLAM_API __attribute__((visibility("default")))
#else
#define PSLAM_API
#endif

typedef struct pslam_ctx pslam_ctx;

typedef enum {
    PSLAM_OK = 0,
    PSLAM_ERR_ARG = -1,          /* null pointer, negative size, unsupported descriptor width ... */
    PSLAM_ERR_CUDA = -2,         /* a CUDA runtime call failed; see pslam_last_error */
    PSLAM_ERR_CAPACITY = -3,     /* output truncated: more results than the caller's capacity */
    PSLAM_ERR_UNSUPPORTED = -4,  /* size beyond a documented kernel limit, or dead reference mode */
    PSLAM_ERR_NCCL = -5,
    PSLAM_ERR_NO_DEVICE = -6
} pslam_status;

#define PSLAM_DESC_BYTES 32
#define PSLAM_MAX_BF_ROWS 65535      /* pslam_match_bf_mutual / knn2: nq, nt <= 65535 (16-bit packed index) */
#define PSLAM_LC_MAX_KF_DESC 4096    /* descriptors per keyframe in the loop-closure database */
#define PSLAM_LC_MAX_QUERY 2048      /* query descriptors per loop-closure sweep; above 1024 the V1 sweep needs every
                                        keyframe to hold at most 2048 descriptors */
#define PSLAM_LC_MAX_TOPK 64

/* ---- context ------------------------------------------------------------------------------ */
PSLAM_API int pslam_ctx_create(int device, pslam_ctx** out);
PSLAM_API void pslam_ctx_destroy(pslam_ctx* ctx);
PSLAM_API const char* pslam_last_error(const pslam_ctx* ctx);
PSLAM_API int pslam_version(void);
/* cudaStream_t every kernel of this ctx is launched on (for CUDA-event timing by the caller) */
PSLAM_API void* pslam_ctx_stream(pslam_ctx* ctx);
PSLAM_API int pslam_ctx_sync(pslam_ctx* ctx);
/* number of kernels of this library launched on the ctx since creation */
PSLAM_API uint64_t pslam_kernel_launches(const pslam_ctx* ctx);
PSLAM_API int pslam_sm_count(const pslam_ctx* ctx);

/* ---- stage 1: back-projection --------------------------------------------------------------
 * Replaces RGBD::removeImageDistortion + RGBD::keypoints2Dto3D (include/putslam/RGBD/RGBD.h:38-73,
 * src/RGBD/RGBD.cpp:30-65,254-314), the detDist loop (src/Matcher/matcher.cpp:51-58) and, when
 * cov_out != NULL, DepthSensorModel::computeCov (src/Grabber/depthSensorModel.cpp:28-36). */
typedef struct {
    float fx, fy, cx, cy;     /* cameraMatrixMat (CV_32F) */
    float dist[5];            /* distortionCoeffsMat k1 k2 p1 p2 k3 (CV_32F) */
} pslam_camera;

typedef struct {
    double fx, fy, cx, cy;       /* DepthSensorModel::Config focalLength / focalAxis */
    double var_u, var_v;         /* Ruvd(0,0), Ruvd(1,1) */
    double dist_var_coefs[4];    /* c3 c2 c1 c0 of the depth variance polynomial */
} pslam_cov_params;

/* uv: n x 2 float keypoints (distorted pixels if undistort != 0). depth: H rows of row_stride uint16.
 * Outputs (nullable except xyz_out): uv_undist_out n x 2, xyz_out n x 3 float, det_dist_out n double,
 * cov_out n x 9 double row-major (needs cov != NULL). */
PSLAM_API int pslam_backproject(pslam_ctx* ctx, const float* uv, int n, const uint16_t* depth, int W, int H,
                                int row_stride, const pslam_camera* cam, int undistort, double depth_scale,
                                float* uv_undist_out, float* xyz_out, double* det_dist_out, double* cov_out,
                                const pslam_cov_params* cov);

/* Batched DepthSensorModel::informationMatrixFromImageCoordinates (src/Grabber/depthSensorModel.cpp:55-59), the
 * measurement information FeaturesMap::addFeatures / addMeasurements attach to every map measurement
 * (src/Map/featuresMap.cpp:110-114, 263-267): uvz = n x {u, v, depth} doubles (u, v truncated to unsigned like the
 * reference), info_out n x 9 row-major = inverse of the sensor covariance, cov_out (nullable) the covariance. */
PSLAM_API int pslam_information_matrices(pslam_ctx* ctx, const double* uvz, int n, const pslam_cov_params* cov,
                                         double* info_out, double* cov_out);

/* Uncertainty model 1: RGBD::computeNormals (include/putslam/RGBD/RGBD.h:91-95, src/RGBD/RGBD.cpp:101-144) followed
 * by DepthSensorModel::uncertinatyFromNormal (src/Grabber/depthSensorModel.cpp:62-76).  px = n x {u, v} integer
 * pixels; normals_out n x 3, cov_out n x 9 row-major, info_out n x 9 = cov.inverse() (what FeaturesMap attaches to
 * the measurement, src/Map/featuresMap.cpp:115-117, 268-270); all double, all nullable.  A pixel with fewer than
 * two neighbours that have depth yields NaN, as in the reference. */
PSLAM_API int pslam_normal_uncertainty(pslam_ctx* ctx, const int* px, int n, const uint16_t* depth, int W, int H,
                                       int row_stride, const pslam_camera* cam, double depth_scale,
                                       double scale_uncertainty_normal, double* normals_out, double* cov_out,
                                       double* info_out);

/* Uncertainty model 2: RGBD::computeRGBGradients (include/putslam/RGBD/RGBD.h:98-106, src/RGBD/RGBD.cpp:147-187)
 * followed by DepthSensorModel::uncertinatyFromRGBGradient (src/Grabber/depthSensorModel.cpp:79-95).  rgb = H rows of
 * rgb_row_bytes bytes, 3 bytes per pixel (CV_8UC3).  The reference reads the 3x3 colour patch through
 * at<uint16_t>, i.e. as little-endian 16-bit words at byte offsets 3(u-1)+2c of rows v-1..v+1 (:156-159); that is
 * reproduced, the Scharr sums are exact integers.  The direction offsets int(sqrt(2)*sin/cos(atan2(gy,gx)+pi/2)) are
 * evaluated in integer form; on the diagonals |gx| == |gy|, where the truncation depends on the last bit of libm,
 * they come from a table computed with the host libm when the context is created.  Pixels failing the border test
 * (:154) get the un-normalised (1,1,1) of :162.  grad_out n x 3, cov_out / info_out n x 9 row-major; double, nullable. */
PSLAM_API int pslam_gradient_uncertainty(pslam_ctx* ctx, const int* px, int n, const uint8_t* rgb, int rgb_row_bytes,
                                         const uint16_t* depth, int W, int H, int row_stride, const pslam_camera* cam,
                                         double depth_scale, double scale_uncertainty_gradient, double* grad_out,
                                         double* cov_out, double* info_out);

/* ---- descriptor production (second half) -----------------------------------
 * pslam_orb_describe replaces MatcherOpenCV::describeFeatures for the ORB descriptor (virtual, include/putslam/Matcher/
 * matcher.h:405-422; src/Matcher/matcherOpenCV.cpp:181-195 == cv::ORB::create()->compute(rgbImage, features,
 * descriptors) with OpenCV's defaults: 8 levels, scale 1.2, patch 31, edge threshold 31, WTA_K 2).
 * image: H rows of row_bytes; channels 1 (gray) or 3 (BGR order as stored -- ORB converts with COLOR_BGR2GRAY).
 * Keypoints: kp_xy n x 2 float (cv::KeyPoint::pt), kp_octave, kp_angle_deg (cv::KeyPoint::angle).  Like cv::ORB::compute,
 * keypoints whose rounded position is closer than 31 px to the image border are dropped and the rest are regrouped by
 * octave (stable): order_out (capacity n) receives the indices of the surviving keypoints in output order -- the caller
 * permutes its keypoint vector with it, as ORB does to `features` -- and desc_out (capacity n x 32) one 32-byte
 * descriptor per surviving keypoint in that order.  cos / sin of the keypoint angles are evaluated on the host
 * (double libm, rounded to float, as OpenCV does) and the sampling itself on the device.  Octaves 0 .. 11; every
 * pyramid level must stay larger than 32 x 32 pixels.
 * image == NULL: describe on the frame this context uploaded last (the preceding pslam_orb_detect or pslam_orb_describe
 * with the same W, H and channels) -- detectFeatures followed by describeFeatures on one frame then uploads it once.
 * The gray conversion is still this call's own (BGR order), exactly as when the image is passed again. */
PSLAM_API int pslam_orb_describe(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels,
                                 const float* kp_xy, const int* kp_octave, const float* kp_angle_deg, int n,
                                 int* order_out, int* n_out, uint8_t* desc_out);

/* pslam_orb_detect == cv::ORB::create(nfeatures)->detect(image), the detector call inside MatcherOpenCV::detectFeatures
 * (src/Matcher/matcherOpenCV.cpp:62-63,118-176; OpenCV defaults: 8 levels, scale 1.2, edge threshold 31, patch 31, Harris
 * score, FAST threshold 20).  image: channels 1 (gray) or 3; colour_order 0 = convert with COLOR_BGR2GRAY (what ORB does
 * to a colour input), 1 = COLOR_RGB2GRAY (what detectFeatures does first, :122).  Per pyramid level the device computes
 * the FAST-9/16 corner scores, non-maximum suppression, the 31-px border filter, and for every surviving corner the
 * Harris response and the intensity-centroid angle; OpenCV's two retainBest passes per level run on the host with the
 * same libstdc++ algorithms (nth_element, partition), so the keypoints come out in OpenCV's order.
 * Outputs (capacity cap each; n_out receives the count, PSLAM_ERR_CAPACITY if it exceeds cap): kp_xy n x 2 (pt),
 * kp_size, kp_angle (degrees), kp_response (Harris), kp_octave; class_id is -1 for all, as in OpenCV.  Levels narrower
 * than 63 pixels simply yield no keypoints (nothing is farther than 31 px from their border). */
PSLAM_API int pslam_orb_detect(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels,
                               int colour_order, int nfeatures, float* kp_xy, float* kp_size, float* kp_angle,
                               float* kp_response, int* kp_octave, int cap, int* n_out);

/* pslam_fast_detect == cv::FastFeatureDetector::create(threshold, true)->detect(image) (TYPE_9_16), the detector option
 * "FAST" of MatcherOpenCV (src/Matcher/matcherOpenCV.cpp:60-61 with OpenCV's default threshold 10): keypoints in raster
 * order, kp_xy n x 2 (integer pixel positions as float), kp_response = corner score; size 7, angle -1, octave 0 are
 * constants of the detector.  threshold 1 .. 254.  image / channels / colour_order as pslam_orb_detect.  PSLAM_ERR_CAPACITY when more than cap
 * corners are found (n_out still receives the count). */
PSLAM_API int pslam_fast_detect(pslam_ctx* ctx, const uint8_t* image, int W, int H, int row_bytes, int channels,
                                int colour_order, int threshold, float* kp_xy, float* kp_response, int cap, int* n_out);

/* ---- stage 2: Hamming matching --------------------------------------------------------------
 * pslam_match_bf_mutual replaces MatcherOpenCV::performMatching for ORB/LDB
 * (include/putslam/Matcher/matcher.h:412-413, src/Matcher/matcherOpenCV.cpp:198-206 ==
 * cv::BFMatcher(NORM_HAMMING, crossCheck=true).match(query=prev, train=cur)): mutual nearest neighbours,
 * first-argmin ties, ascending queryIdx, imgIdx 0.  Outputs sized min(nq, nt). */
PSLAM_API int pslam_match_bf_mutual(pslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                                    int desc_bytes, int* out_query_idx, int* out_train_idx, float* out_distance,
                                    int* n_out);

/* knnMatch(k = 2) extension (north_star "ratio test"): out_idx / out_dist are nq x 2, ascending
 * distance, lowest train index first on ties; -1 where nt < 2.  The caller applies d1 < r*d2. */
PSLAM_API int pslam_match_knn2(pslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                               int desc_bytes, int* out_idx, float* out_dist);

/* Guided frame-to-map matching: the loop nest of Matcher::matchXYZ (src/Matcher/matcher.cpp:662-748).
 * map_xyz: M x 3 float (MapFeature.position cast to float, :665), map_level / cur_level: predicted
 * pyramid levels computed by the caller in double exactly as :639-651 / :682-692.  radius and
 * accept_ratio are the values after the retry adjustment (:617-622).  distance_mode 0 = the reference's
 * saturating-subtract popcount (:719-721), 1 = XOR Hamming.  Matches come out in (map j, cur i) order.
 * n_out receives the full count; if it exceeds cap the first cap are written and PSLAM_ERR_CAPACITY is
 * returned.  perfect_out (nullable) = features whose best value is 0 (:729-731). */
PSLAM_API int pslam_match_guided_xyz(pslam_ctx* ctx, const float* map_xyz, const uint8_t* map_desc,
                                     const int* map_level, int M, const float* cur_xyz, const uint8_t* cur_desc,
                                     const int* cur_level, int N, int desc_bytes, double radius, double accept_ratio,
                                     int distance_mode, int* out_query_idx, int* out_train_idx, float* out_distance,
                                     int cap, int* n_out, int* perfect_out);

/* ---- stage 3: RANSAC + Umeyama, Kabsch ------------------------------------------------------
 * Mirrors RANSAC::parameters (include/putslam/TransformEst/RANSAC.h:23-31) plus the intrinsics the
 * reprojection metrics read from cameraMatrix. */
typedef struct {
    int error_version;                  /* RANSAC::ERROR_VERSION: 0 EUCLIDEAN, 1 REPROJECTION, 2 EUCLIDEAN_AND_
                                           REPROJECTION, 4 ADAPTIVE; 3 (MAHALANOBIS) is dead code in the reference
                                           (RANSAC.cpp:301-303) -> PSLAM_ERR_UNSUPPORTED */
    double inlier_threshold_euclidean;
    double inlier_threshold_reprojection;
    double minimal_inlier_ratio_threshold;
    int minimal_number_of_matches;
    int used_pairs;                     /* must be 3 */
    float fx, fy, cx, cy;
} pslam_ransac_params;

/* Replaces RANSAC::estimateTransformation (RANSAC.h:44-48, src/TransformEst/RANSAC.cpp:50-174).
 * prev: n_prev x 3, cur: n_cur x 3 float; match k pairs prev[match_query[k]] with cur[match_train[k]].
 * Hypothesis i samples 3 matches with Philox4x32-10(key = seed, counter = {block, i, 0, 0}) % m with
 * rejection (pslam_ransac_sample reproduces the draw on the host).  num_hyp = 0: the reference's
 * adaptive loop: the bound starts at computeRANSACIteration(0.20) = 487 (RANSAC.cpp:30) and after every improvement
 * becomes min(computeRANSACIteration(minimalInlierRatioThreshold), computeRANSACIteration(best)) (:450-453); all
 * hypotheses the loop can reach -- max(487, computeRANSACIteration(minimalInlierRatioThreshold)), at most 2^20 -- are
 * scored in one launch and the loop is replayed over their counts.  num_hyp > 0: exactly that many.
 * T_out: column-major 4x4, prev ~= R*cur + t.  inlier_idx_out (capacity m): indices into the match list,
 * ascending.  Failure conventions are the reference's: too few matches or best ratio below the
 * threshold -> identity and zero inliers, return value still PSLAM_OK. */
PSLAM_API int pslam_ransac_estimate(pslam_ctx* ctx, const float* prev, int n_prev, const float* cur, int n_cur,
                                    const int* match_query, const int* match_train, int m,
                                    const pslam_ransac_params* params, uint64_t seed, int num_hyp, float* T_out,
                                    int* inlier_idx_out, int* n_inliers_out, double* best_ratio_out,
                                    int* hyp_used_out);

/* Termination rule of the hypothesis loop for this ctx.  rule 0 (default): RANSAC::computeRANSACIteration /
 * saveBetterModel (src/TransformEst/RANSAC.cpp:438-461).  rule 1: the standard stopping criterion of the reference's
 * (not compiled) USAC framework, USAC<T>::updateStandardStopping (include/putslam/USAC/USAC.h:944-971), confidence
 * 0.99 in src/USAC/USAC_wrapper.cpp:66; the bound is capped by the hypothesis budget (num_hyp, or 487 when 0). */
PSLAM_API int pslam_ransac_set_stopping(pslam_ctx* ctx, int rule, double confidence);

/* Per-hypothesis inlier counts of the last pslam_ransac_estimate / frame call on this ctx
 * (-1 = degenerate model); for diagnostics and parity tests.  Copies min(cap, hypotheses) ints. */
PSLAM_API int pslam_ransac_last_counts(pslam_ctx* ctx, int* counts_out, int cap, int* n_out);

/* Host mirror of the device sampler: the 3 match indices hypothesis `hyp` draws out of m matches. */
PSLAM_API void pslam_ransac_sample(uint64_t seed, uint32_t hyp, int m, int out3[3]);

/* RANSAC::pointInlierRatio (RANSAC.h:56-66): unique trainIdx of inliers / unique trainIdx of all matches. */
PSLAM_API double pslam_point_inlier_ratio(const int* inlier_train, int n_inliers, const int* all_train, int n_all);

/* Replaces KabschEst::computeTransformation (include/putslam/TransformEst/kabschEst.h:34,
 * src/TransformEst/kabschEst.cpp:24-68) for a batch of independent point-set pairs.
 * A, B: concatenated n_i x 3 ROW-major double points; offsets[batch + 1].  T_out: batch x 12, column-major
 * 3x4 with B ~= R*A + t.  An empty pair yields identity. */
PSLAM_API int pslam_kabsch_batch(pslam_ctx* ctx, const double* A, const double* B, const int* offsets, int batch,
                                 double* T_out);

/* ---- fused per-frame pipelines (one submission, no host round trips between stages) ---------
 * Frame-to-frame VO: Matcher::match (src/Matcher/matcher.cpp:452-516) minus detection/description:
 * performMatching(prev_desc, cur_desc) -> removeImageDistortion + keypoints2Dto3D on the current frame
 * -> RANSAC(prev_xyz, cur_xyz, matches). */
typedef struct {
    int n_matches;
    int n_inliers;
    int hyp_used;
    int n_filtered;
    double best_ratio;        /* bestInlierRatio */
    double inlier_ratio;      /* RANSAC::pointInlierRatio(inliers, matches) -- the value match()/matchXYZ return */
    float T[16];              /* column-major */
} pslam_frame_result;

PSLAM_API int pslam_frame_to_frame(pslam_ctx* ctx, const uint8_t* prev_desc, const float* prev_xyz, int n_prev,
                                   const uint8_t* cur_desc, const float* cur_uv, int n_cur, const uint16_t* depth,
                                   int W, int H, int row_stride, const pslam_camera* cam, int undistort,
                                   double depth_scale, const pslam_ransac_params* params, uint64_t seed, int num_hyp,
                                   float* cur_xyz_out, float* cur_uv_undist_out, double* cur_det_dist_out,
                                   int* match_query_out, int* match_train_out, float* match_dist_out,
                                   int* inlier_idx_out, pslam_frame_result* result);

/* Frame-to-map: Matcher::matchXYZ private overload (src/Matcher/matcher.cpp:606-798) minus the
 * MapFeature marshalling: guided matching -> RANSAC(map_xyz, cur_xyz, matches).  match_cap bounds the
 * guided matches kept (PSLAM_ERR_CAPACITY if exceeded).  result->inlier_ratio = -1 when no matches (:755). */
PSLAM_API int pslam_frame_to_map(pslam_ctx* ctx, const float* map_xyz, const uint8_t* map_desc, const int* map_level,
                                 int M, const float* cur_xyz, const uint8_t* cur_desc, const int* cur_level, int N,
                                 double radius, double accept_ratio, int distance_mode,
                                 const pslam_ransac_params* params, uint64_t seed, int num_hyp, int match_cap,
                                 int* match_query_out, int* match_train_out, float* match_dist_out,
                                 int* inlier_idx_out, pslam_frame_result* result);

/* Same pipeline from the raw feature attributes: map_xyz M x 3 DOUBLE (MapFeature.position), octave and detDist of the
 * descriptor's view (ExtendedDescriptor.octave / .detDist), current keypoints' octave and detDist.  The predicted
 * pyramid levels (matcher.cpp:639-651, 682-692) and the double->float cast (:665) then run on the device, which takes
 * ~0.2 ms of pow()/log() per frame off the host.  The one case that depends on the last bit of the host's log() --
 * detDist == curDist exactly -- is answered from tables built with the host libm, so the levels equal the host-computed
 * ones (tests/test_gpu_parity.py::test_frame_to_map_device_levels). */
PSLAM_API int pslam_frame_to_map_features(pslam_ctx* ctx, const double* map_xyz, const uint8_t* map_desc,
                                          const int* map_octave, const double* map_det_dist, int M, const float* cur_xyz,
                                          const uint8_t* cur_desc, const int* cur_octave, const double* cur_det_dist, int N,
                                          double radius, double accept_ratio, int distance_mode,
                                          const pslam_ransac_params* params, uint64_t seed, int num_hyp, int match_cap,
                                          int* match_query_out, int* match_train_out, float* match_dist_out,
                                          int* inlier_idx_out, pslam_frame_result* result);

/* Loop-closure pair verification: Matcher::matchFeatureLoopClosure (src/Matcher/matcher.cpp:802-861) minus the
 * MapFeature marshalling: performMatching(desc0, desc1) -> RANSAC(xyz0, xyz1, matches) in one submission.
 * result->inlier_ratio follows the reference: 0 when either set has fewer than 10 features (:830-834), -1 when
 * there are no matches (:838-839), else pointInlierRatio.  Match / inlier buffers sized min(n0, n1). */
PSLAM_API int pslam_loop_closure_pair(pslam_ctx* ctx, const uint8_t* desc0, const float* xyz0, int n0,
                                      const uint8_t* desc1, const float* xyz1, int n1, const pslam_ransac_params* params,
                                      uint64_t seed, int num_hyp, int* match_query_out, int* match_train_out,
                                      float* match_dist_out, int* inlier_idx_out, pslam_frame_result* result);

/* Re-run the kernels of the last pslam_frame_to_map call on the inputs already resident in HBM:
 * no host<->device copies, no synchronisation (device-time measurements; matchXYZ retries). */
PSLAM_API int pslam_frame_to_map_resident(pslam_ctx* ctx);
PSLAM_API int pslam_frame_to_frame_resident(pslam_ctx* ctx);

/* ---- map-side preparation (the step in front of matchXYZ) ------------------------
 * Body of PUTSLAM::getAndFilterFeaturesFromMap (src/PUTSLAM/PUTSLAM.cpp:624-674) after getCovisibleFeatures:
 * FeaturesMap::findNearestFrame's view-angle test (src/Map/featuresMap.cpp:528-563), removal of features without a
 * good observation angle (PUTSLAM.cpp:932-950), moveMapFeaturesToLocalCordinateSystem (PUTSLAM.cpp:28-51) with
 * DepthSensorModel::inverseModel (src/Grabber/depthSensorModel.cpp:18-25), RGBD::removeFarMapFeatures (RGBD.cpp:232-252).
 * map_xyz: M x 3 double global positions; view_axis: M x 3 float, third column of the rotation of the view that
 * holds each feature's descriptor; camera_pose: column-major 4x4 double (camera -> global, Mat34::matrix()).
 * Outputs, compacted in input order (capacity M): kept_idx, xyz_local M x 3 double, uv M x 2 double ((-1,-1) when the
 * projection leaves the image or the depth box), angles.  */
typedef struct {
    double fx, fy, cx, cy;      /* DepthSensorModel focalLength / focalAxis */
    double image_w, image_h;    /* DepthSensorModel imageSize */
    double max_angle;           /* matcherParameters.maxAngleBetweenFrames */
    double max_z;               /* 5.0 in PUTSLAM.cpp:662 */
} pslam_map_prepare_params;
PSLAM_API int pslam_map_prepare(pslam_ctx* ctx, const double* map_xyz, const float* view_axis, int M,
                                const double camera_pose[16], const pslam_map_prepare_params* params, int* kept_idx,
                                double* xyz_local, double* uv, double* angles, int* n_out);

/* ---- resident feature map ------------------------------------------------
 * The map side of Matcher::matchXYZ kept in HBM in SoA form between frames, so that per frame only the camera pose
 * and the current keypoints cross PCIe.  A slot holds what PUTSLAM::getAndFilterFeaturesFromMap and matchXYZ read
 * of one MapFeature (include/putslam/Defs/putslam_defs.h:184-216): global position (double), and of the view that
 * holds its descriptor (ExtendedDescriptor, :120-151) the 32-byte descriptor, octave, detDist and optical axis.
 * Slots are the caller's feature indices.  pslam_map_write stores [first, first + count): a range that extends the
 * map needs every array, a range inside it may pass NULL for attributes that did not change (e.g. only xyz after
 * a pose-graph update, src/Map/featuresMap.cpp updateMap).  Buffers may be reused when the call returns. */
PSLAM_API int pslam_map_reserve(pslam_ctx* ctx, int max_features);
PSLAM_API int pslam_map_write(pslam_ctx* ctx, int first, int count, const double* xyz, const uint8_t* desc,
                              const int* octave, const double* det_dist, const float* view_axis);
PSLAM_API int pslam_map_truncate(pslam_ctx* ctx, int n_features);
PSLAM_API int pslam_map_size(const pslam_ctx* ctx, int* n_features);

/* One tracking frame against the resident map == pslam_map_prepare on the stored features followed by
 * pslam_frame_to_map_features on the kept ones (PUTSLAM.cpp:624-674 then Matcher::matchXYZ, matcher.cpp:606-798),
 * in one submission: filter + move to the camera frame (K8, which also gathers the kept descriptors), level
 * prediction, guided matching, RANSAC.  kept_idx_out (capacity = map size) lists the slots that passed the filters,
 * in slot order; match_query_out indexes that list, exactly as the reference's queryIdx indexes the filtered
 * mapFeatures vector.  xyz_local_out / uv_out (nullable, capacity map size x 3 / x 2) are the camera-frame positions
 * and projections moveMapFeaturesToLocalCordinateSystem writes back into the features.  Other arguments and
 * results as pslam_frame_to_map_features. */
PSLAM_API int pslam_frame_to_resident_map(pslam_ctx* ctx, const double camera_pose[16],
                                          const pslam_map_prepare_params* prep, const float* cur_xyz,
                                          const uint8_t* cur_desc, const int* cur_octave, const double* cur_det_dist,
                                          int N, double radius, double accept_ratio, int distance_mode,
                                          const pslam_ransac_params* params, uint64_t seed, int num_hyp, int match_cap,
                                          int* kept_idx_out, int* n_kept_out, double* xyz_local_out, double* uv_out,
                                          int* match_query_out, int* match_train_out, float* match_dist_out,
                                          int* inlier_idx_out, pslam_frame_result* result);

/* Replaces TransformEst::computeUncertainty (parametrization EULER: 6 x 6 covariance over x, y, z, roll, pitch, yaw;
 * include/putslam/TransformEst/transformEst.h:29-144) and computeUncertaintyG2O (QUATERNION: x, y, z, qx, qy, qz; :147-272)
 * for a batch of independent problems (callers demos/demoKabsch.cpp:655,731,1028):
 * U = H^-1 G^T Cx G H^-1 with H = d2J/dtheta2 and G = d2J/dtheta dX of J = sum |a_i - R b_i - t|^2 at the given
 * transformation, Cx = the block-diagonal covariance of the points.  Layouts as pslam_kabsch_batch:
 *   A, B        concatenated n_i x 3 ROW-major points (setA, setB), offsets[batch + 1]
 *   covA, covB  n_i x 9, one ROW-major 3 x 3 per point (setAUncertainty, setBUncertainty)
 *   T           batch x 12, column-major 3 x 4 with A ~= R*B + t   (the Mat34 passed as `transformation`)
 *   U_out       batch x 36, column-major 6 x 6 (Mat66)
 *   ok_out      nullable, batch: 0 where the set is empty or the Hessian is singular (U = 0 there; Eigen's inverse()
 *               would return non-finite values) */
enum { PSLAM_UNCERTAINTY_EULER = 0, PSLAM_UNCERTAINTY_QUATERNION = 1 };
PSLAM_API int pslam_transform_uncertainty_batch(pslam_ctx* ctx, const double* A, const double* B, const double* covA,
                                                const double* covB, const int* offsets, const double* T, int batch,
                                                int parametrization, double* U_out, int* ok_out);

/* ---- KLT tracking: the performTracking seam ------------------------------
 * pslam_klt_track == cv::calcOpticalFlowPyrLK(prevImg, img, prevPts, nextPts, status, err, Size(win, win), max_level,
 * TermCriteria(criteria_type, max_iter, eps), flags, min_eig_threshold) as MatcherOpenCV::performTracking calls it
 * (src/Matcher/matcherOpenCV.cpp:209-241, on the colour frames: src/Matcher/matcher.cpp:151-158): pyramids of both
 * frames (cv::pyrDown), then every point from the coarsest level to the base -- Scharr gradients, 2^14 fixed-point
 * bilinear windows, the 2 x 2 normal equations, Newton steps -- with OpenCV's arithmetic, so cur_xy / status / err equal
 * OpenCV's bit for bit (tests/golden/klt_cv2.npz).
 *   prev_image  NULL: the previous frame is the cur_image of the preceding pslam_klt_* call on this context (same size
 *               and channels, at least as many pyramid levels) -- its pyramid is still resident, so a tracked
 *               sequence uploads every frame once.
 *   channels    1 or 3 (the reference tracks on the 3-channel frames); rows of row_bytes >= channels * W bytes.
 *   cur_xy      n x 2, in/out: read as the initial guess when flags has PSLAM_KLT_USE_INITIAL_FLOW.
 *   win         odd or even, 3 .. 21 (reference: 7); max_level 0 .. 7 (reference: 3; like OpenCV the pyramid ends
 *               where the next level would not exceed the window).
 *   criteria_type  bit 0: max_iter counts (else 30), bit 1: eps counts (else 0.01) -- cv::TermCriteria::COUNT / EPS;
 *               max_iter is clamped to [0, 100] and eps to [0, 10] as OpenCV does.
 *   err         mean absolute window difference in grey levels, or the minimum eigenvalue with
 *               PSLAM_KLT_GET_MIN_EIGENVALS; status 1 = tracked. */
enum { PSLAM_KLT_USE_INITIAL_FLOW = 4, PSLAM_KLT_GET_MIN_EIGENVALS = 8 };   /* cv::OPTFLOW_* values */
PSLAM_API int pslam_klt_track(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H,
                              int row_bytes, int channels, const float* prev_xy, float* cur_xy, int n, int win,
                              int max_level, int criteria_type, int max_iter, double eps, int flags,
                              double min_eig_threshold, uint8_t* status, float* err);

/* pslam_klt_perform_tracking == the whole of MatcherOpenCV::performTracking (matcherOpenCV.cpp:209-300) in one
 * submission: pslam_klt_track, then status cleared where err > error_threshold (:247-251), then of every pair of tracked
 * positions (all features, whatever their status) closer than min_distance the one with the larger err is dropped, the
 * second on ties (:254-266; N^2 / 2 pair tests on the device).  kept_idx_out (capacity n) lists the surviving feature
 * indices in order: survivor j is DMatch(kept_idx_out[j], j, 0) and row j of the compacted features / keyPoints /
 * detDists (:269-290).  cur_xy, status (as returned by the tracker, before the threshold) and err cover all n. */
PSLAM_API int pslam_klt_perform_tracking(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H,
                                         int row_bytes, int channels, const float* prev_xy, float* cur_xy, int n, int win,
                                         int max_level, int criteria_type, int max_iter, double eps, int flags,
                                         double min_eig_threshold, double error_threshold, double min_distance,
                                         uint8_t* status, float* err, int* kept_idx_out, int* n_kept_out);

/* One tracking frame of the VO_TRACKING mode == the data-parallel part of Matcher::trackKLT (src/Matcher/matcher.cpp:
 * 151-207) in one submission: performTracking (as pslam_klt_perform_tracking), RGBD::removeImageDistortion and
 * RGBD::keypoints2Dto3D on the survivors (as pslam_backproject; undistort = 0 skips the first), then
 * RANSAC::estimateTransformation(prevFeatures3D, features3D, matches, inliers) with matches = DMatch(kept[j], j, 0)
 * (as pslam_ransac_estimate).  The survivors' list, positions and 3-D points stay in HBM between the steps.
 *   prev_xyz            n x 3, the previous frame's 3-D features (prevFeatures3D), row i belongs to prev_xy row i
 *   kept_*_out          capacity n rows; row j belongs to survivor j = feature kept_idx_out[j]: undistorted position
 *                       (nullable), 3-D point, detection distance |p| (nullable)
 *   inlier_idx_out      capacity n: indices j into the survivor list
 *   result              n_matches = survivors, T = estimated transformation (column-major, identity when RANSAC rejects),
 *                       inlier_ratio = RANSAC::pointInlierRatio(inliers, matches) (0 without survivors), as trackKLT returns
 * Other arguments as pslam_klt_perform_tracking / pslam_frame_to_frame. */
PSLAM_API int pslam_klt_frame(pslam_ctx* ctx, const uint8_t* prev_image, const uint8_t* cur_image, int W, int H, int row_bytes,
                              int channels, const float* prev_xy, const float* prev_xyz, float* cur_xy, int n, int win,
                              int max_level, int criteria_type, int max_iter, double eps, int flags,
                              double min_eig_threshold, double error_threshold, double min_distance, const uint16_t* depth,
                              int depth_row_stride, const pslam_camera* cam, int undistort, double depth_scale,
                              const pslam_ransac_params* params, uint64_t seed, int num_hyp, uint8_t* status, float* err,
                              int* kept_idx_out, int* n_kept_out, float* kept_uv_undist_out, float* kept_xyz_out,
                              double* kept_det_dist_out, int* inlier_idx_out, pslam_frame_result* result);

/* ---- loop-closure sweep: query frame vs every keyframe of the map --------------------------
 * Generalises Matcher::matchFeatureLoopClosure's performMatching step (src/Matcher/matcher.cpp:802-861,
 * :835) from one FABMAP-proposed pair to all keyframes: score(k) = number of mutual-NN matches between
 * the query and keyframe k with distance <= tau; result = top-k keyframes, score descending, keyframe id
 * ascending on ties.  The keyframe descriptors stay resident in HBM. */
PSLAM_API int pslam_lc_db_reserve(pslam_ctx* ctx, int64_t max_descriptors, int max_keyframes);
/* append keyframes: desc = sum(counts) x 32 bytes, kf_off[n_kf + 1] descriptor offsets relative to desc */
PSLAM_API int pslam_lc_db_append(pslam_ctx* ctx, const uint8_t* desc, const int64_t* kf_off, int n_kf);
PSLAM_API int pslam_lc_db_clear(pslam_ctx* ctx);
PSLAM_API int pslam_lc_db_size(const pslam_ctx* ctx, int* n_keyframes, int64_t* n_descriptors);
/* global id of local keyframe 0 (rank r of a sharded map owns ids [base, base + n_keyframes)) */
PSLAM_API int pslam_lc_set_id_base(pslam_ctx* ctx, int kf_id_base);

/* Form of the V1 sweep: 0 = automatic -- the tensor-core form (tcgen05 kind::i8 on +-1-expanded rows, exact) for up to
 * 1024 query descriptors, the popcount range form beyond; 3 = popcount range form always; 1 = whole keyframes per CTA,
 * 2 = 128-row tiles + finalize pass (the round-1 forms).  Results are identical; this is a performance knob.
 * PSLAM_LC_TENSOR=0 in the environment makes 0 behave like 3. */
PSLAM_API int pslam_lc_set_work_unit(pslam_ctx* ctx, int mode);
/* After a sweep: *used_tensor_cores = 1 when the last V1 sweep ran in the tensor-core form; *timed_out != 0 when one of its
 * internal waits ever gave up (a bug, never expected: results of that sweep are then undefined).  Synchronises the stream. */
PSLAM_API int pslam_lc_tensor_status(pslam_ctx* ctx, int* used_tensor_cores, int* timed_out);

/* Single-GPU query.  out_* sized k (<= PSLAM_LC_MAX_TOPK); unused slots -1.  scores_out (nullable):
 * n_keyframes per-keyframe scores. */
PSLAM_API int pslam_lc_query(pslam_ctx* ctx, const uint8_t* query, int nq, int tau, int k, int* out_kf_ids,
                             int* out_scores, int* scores_out);
/* kernels only, on the query already resident from the last pslam_lc_query / _sharded call */
PSLAM_API int pslam_lc_query_resident(pslam_ctx* ctx, int tau, int k);
/* device time (CUDA events on the ctx stream) of the sweep kernel of the most recent query on this ctx */
PSLAM_API int pslam_lc_last_sweep_ms(pslam_ctx* ctx, float* ms_out);

/* Caller-pinned input buffers.  Every entry point copies its host inputs through a page-locked staging arena (one memcpy +
 * one cudaMemcpyAsync).  A caller whose buffers live across calls (a map side kept on the host, descriptor matrices reused
 * from frame to frame) can page-lock them once: inputs that lie inside a registered range are copied to the device straight
 * from where they are, without the staging memcpy (pslam_frame_to_map / pslam_frame_to_map_features use this for every
 * array).  The range must stay valid and unchanged in size until pslam_host_unregister (or pslam_ctx_destroy). */
/* Diagnostics: host-side phase times (microseconds since the call began) of the last pslam_frame_to_map* call on this ctx:
 * [1] inputs packed  [2] host->device copies enqueued  [3] kernels enqueued  [4] device->host copy enqueued
 * [5] stream synchronised  [6] results unpacked. */
PSLAM_API int pslam_debug_host_stamps(const pslam_ctx* ctx, double out8[8]);
PSLAM_API int pslam_host_register(pslam_ctx* ctx, const void* ptr, size_t bytes);
PSLAM_API int pslam_host_unregister(pslam_ctx* ctx, const void* ptr);

/* Multi-GPU (one process per GPU, keyframes sharded by rank).  The library brings up its own NCCL
 * communicator from a caller-distributed ncclUniqueId (128 bytes): rank 0 calls pslam_comm_unique_id,
 * ships the bytes to the other ranks over whatever the host application already has, every rank calls
 * pslam_comm_init.  pslam_lc_query_sharded: the query descriptors go from `root` to every rank (root < 0: every rank already
 * passes the same query), local sweep + local top-k, exchange of k {score, id} pairs per rank, merge -> the same global
 * top-k on every rank.  With peer access between the GPUs the exchange runs inside the sweep kernel over NVLink
 * (pslam_lc_exchange_mode); otherwise ncclBroadcast / ncclAllGather + a merge kernel. */
PSLAM_API int pslam_comm_unique_id(uint8_t id_out[128]);
PSLAM_API int pslam_comm_init(pslam_ctx* ctx, const uint8_t id[128], int rank, int world);
PSLAM_API int pslam_comm_destroy(pslam_ctx* ctx);
/* How a sharded sweep exchanges its per-rank results: 0 = single rank, 1 = NCCL (ncclBroadcast of the query, ncclAllGather of
 * the top-k, merge kernel), 2 = peer memory over NVLink (pslam_comm_init opened every rank's exchange buffer through CUDA IPC):
 * the root rank stores the query into every peer's buffer, every rank stores its top-k into every peer's buffer from the tail
 * of the sweep kernel and merges there.  PSLAM_LC_P2P=0 in the environment selects 1. */
PSLAM_API int pslam_lc_exchange_mode(const pslam_ctx* ctx);
PSLAM_API int pslam_lc_query_sharded(pslam_ctx* ctx, const uint8_t* query, int nq, int root, int tau, int k,
                                     int* out_kf_ids, int* out_scores);
PSLAM_API int pslam_lc_query_sharded_resident(pslam_ctx* ctx, int tau, int k);
/* As pslam_lc_query_sharded_resident, but the query that is resident on `root` (from its last pslam_lc_query_sharded call)
 * is first broadcast to the other ranks: the whole exchange of a sharded query -- broadcast, sweep, top-k, gather, merge --
 * enqueued on the ctx stream without any host copy (root < 0: no broadcast). */
PSLAM_API int pslam_lc_query_sharded_resident_bcast(pslam_ctx* ctx, int root, int tau, int k);

/* Per-query-descriptor 2-NN against the whole resident database (SURVEY 8e variant V2; the oracle is
 * cv::BFMatcher::knnMatch(query, whole_db, 2)): out_idx nq x 2 GLOBAL descriptor indices (int64, -1 = none;
 * global = desc_id_base + position in this ctx's database), out_dist nq x 2 float, ascending distance, lowest
 * index first on ties.  _sharded: ncclBroadcast(query) / local sweep / ncclAllGather of nq x 2 keys / merge.
 * Like the V1 sweep it runs on the tensor cores (tcgen05 kind::i8, exact) for up to 1024 query descriptors and on the
 * integer pipes beyond or when pslam_lc_set_work_unit / PSLAM_LC_TENSOR say so; pslam_lc_tensor_status reports which. */
PSLAM_API int pslam_lc_set_desc_base(pslam_ctx* ctx, int64_t desc_id_base);
PSLAM_API int pslam_lc_knn2(pslam_ctx* ctx, const uint8_t* query, int nq, int64_t* out_idx, float* out_dist);
PSLAM_API int pslam_lc_knn2_sharded(pslam_ctx* ctx, const uint8_t* query, int nq, int root, int64_t* out_idx,
                                    float* out_dist);
PSLAM_API int pslam_lc_knn2_resident(pslam_ctx* ctx, int sharded);

#ifdef __cplusplus
}
#endif
#endif /* PSLAM_B200_H_ */
