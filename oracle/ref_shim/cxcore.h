/* see cv.h in this directory: shim of the legacy OpenCV C API for compiling the reference's VarFlow.cpp (test infrastructure) */
#include "cv.h"
