// klt_emul.cpp -- TEST INFRASTRUCTURE ONLY (built by tests/conftest into tests/_build/, never linked into the product).
// Runs the SOURCE of the device routine putslam_b200/csrc/klt_point.cuh on the CPU, its 32 lanes as a loop, so that the
// kernel's arithmetic and warp choreography can be checked against the oracle / cv2 golden vectors on a machine
// without a GPU.  The GPU parity tests (tests/test_gpu_klt.py) are the parity tests proper.
//   g++ -O2 -ffp-contract=off -fPIC -shared -o tests/_build/libklt_emul.so tests/klt_emul.cpp
#include <stdint.h>
#include <string.h>

#include <vector>

#include "../putslam_b200/csrc/klt_point.cuh"

using namespace pslam;

extern "C" int klt_emul_pyrdown(const uint8_t* src, int w, int h, int cn, uint8_t* dst) {
    const int ow = (w + 1) / 2, oh = (h + 1) / 2;
    for (int oy = 0; oy < oh; ++oy)
        for (int ox = 0; ox < ow; ++ox)
            for (int c = 0; c < cn; ++c) dst[((size_t)oy * ow + ox) * cn + c] = klt_pyrdown_px(src, w, h, cn, ox, oy, c);
    return 0;
}

// flags: bit 0 = OPTFLOW_USE_INITIAL_FLOW, bit 1 = OPTFLOW_LK_GET_MIN_EIGENVALS; specialised != 0 selects the
// compile-time window instantiations exactly like launch_klt_track
extern "C" int klt_emul_track(const uint8_t* I0, const uint8_t* J0, int W, int H, int cn, const float* prev_xy, float* cur_xy,
                              int n, int win, int max_level, int max_iter, double eps, int flags, double min_eig_thr,
                              uint8_t* status, float* err, int specialised, int* levels_used) {
    if (win < 3 || win > kKltMaxWin || (cn != 1 && cn != 3)) return -1;
    KltPlan plan;
    klt_plan(W, H, cn, win, max_level, &plan);
    std::vector<uint8_t> pyrI(plan.bytes), pyrJ(plan.bytes);
    memcpy(pyrI.data(), I0, (size_t)W * H * cn);
    memcpy(pyrJ.data(), J0, (size_t)W * H * cn);
    for (int l = 1; l < plan.n_levels; ++l) {
        klt_emul_pyrdown(pyrI.data() + plan.off[l - 1], plan.w[l - 1], plan.h[l - 1], cn, pyrI.data() + plan.off[l]);
        klt_emul_pyrdown(pyrJ.data() + plan.off[l - 1], plan.w[l - 1], plan.h[l - 1], cn, pyrJ.data() + plan.off[l]);
    }
    KltParams P;
    memset(&P, 0, sizeof(P));
    P.n_levels = plan.n_levels; P.win = win; P.cn = cn;
    klt_criteria(1, max_iter, 1, eps, &P.max_iter, &P.eps_sq);
    P.min_eig_thr = min_eig_thr;
    P.use_initial_flow = flags & 1; P.min_eig_err = (flags >> 1) & 1;
    for (int l = 0; l < plan.n_levels; ++l) {
        P.lv[l].I = pyrI.data() + plan.off[l]; P.lv[l].J = pyrJ.data() + plan.off[l];
        P.lv[l].w = plan.w[l]; P.lv[l].h = plan.h[l];
    }
    if (levels_used) *levels_used = plan.n_levels;
    std::vector<uint8_t> smem(klt_work_bytes(win, cn) + 16);
    const KltWork Wk = klt_carve(smem.data(), win, cn);
    for (int i = 0; i < n; ++i) {
        float nx = 0.f, ny = 0.f, e = 0.f;
        uint8_t st = 0;
        if (P.use_initial_flow) { nx = cur_xy[2 * i]; ny = cur_xy[2 * i + 1]; }
        if (specialised && win == 7 && cn == 3) klt_track_point<7, 3>(P, Wk, prev_xy[2 * i], prev_xy[2 * i + 1], nx, ny, st, e);
        else if (specialised && win == 7 && cn == 1) klt_track_point<7, 1>(P, Wk, prev_xy[2 * i], prev_xy[2 * i + 1], nx, ny, st, e);
        else klt_track_point<0, 0>(P, Wk, prev_xy[2 * i], prev_xy[2 * i + 1], nx, ny, st, e);
        cur_xy[2 * i] = nx; cur_xy[2 * i + 1] = ny; status[i] = st; err[i] = e;
    }
    return 0;
}

// klt_prune_kernel's decision per feature: the 32 lanes' shares of the pair tests (klt_prune_lane), OR-ed like the warp vote
extern "C" int klt_emul_prune(const float* xy, const float* err, const uint8_t* status, int n, double err_thr, double sq_thr,
                              float lim, uint8_t* keep) {
    for (int i = 0; i < n; ++i) {
        bool removed = false;
        for (int lane = 0; lane < 32; ++lane) removed = klt_prune_lane(i, lane, n, xy, err, sq_thr, lim) || removed;
        keep[i] = (uint8_t)(status[i] != 0 && !((double)err[i] > err_thr) && !removed);
    }
    return 0;
}

// shared-memory bytes one warp needs (launch_klt_track sizes its CTAs with this)
extern "C" int klt_emul_work_bytes(int win, int cn) { return (int)klt_work_bytes(win, cn); }
