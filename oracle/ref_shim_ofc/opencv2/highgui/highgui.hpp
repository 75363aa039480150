// opencv2/highgui/highgui.hpp -- SHIM: nothing of highgui is used by optical_flow_calculator.cpp
#include <opencv2/core/core.hpp>
