/* opencv/highgui.h -- SHIM: nothing of highgui is used */
#include <opencv/cv.h>
