// Stand-in for the reference's include/putslam/Defs/opencv.h (which pulls in all of OpenCV): forwards to the small
// cv:: subset of opencv2/shim_cv.h.  Used ONLY to compile reference sources into oracle/_ref/ (oracle/Makefile, target
// ref).  TEST INFRASTRUCTURE: nothing in the product includes this file.
#ifndef PSLAM_REF_SHIM_OPENCV_H
#define PSLAM_REF_SHIM_OPENCV_H
#include <opencv2/shim_cv.h>
#endif
