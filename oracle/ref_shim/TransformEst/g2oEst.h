// Stand-in for the reference's include/putslam/TransformEst/g2oEst.h, which includes the whole g2o pose-graph stack
// (PoseGraph/graph_g2o.h).  RANSAC.cpp only needs the factory's declaration: its G2O branch (RANSAC.cpp:227-232) is
// never selected (every caller leaves usedType = UMEYAMA).  TEST INFRASTRUCTURE, used only by oracle/Makefile's ref target.
#ifndef PSLAM_REF_SHIM_G2OEST_H
#define PSLAM_REF_SHIM_G2OEST_H
#include "TransformEst/transformEst.h"
namespace putslam {
TransformEst* createG2OEstimator(void);     // defined in ref_frontend_wrap.cpp: aborts (not reachable on the path)
}
#endif
