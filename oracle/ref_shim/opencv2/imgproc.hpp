// forwards to the stand-in (TEST INFRASTRUCTURE, see opencv2/shim_cv.h)
#include <opencv2/shim_cv.h>
