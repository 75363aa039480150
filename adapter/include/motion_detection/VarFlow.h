/* VarFlow.h -- drop-in replacement of common/include/motion_detection/VarFlow.h:31-84 (public interface). */
#ifndef VARFLOW_H
#define VARFLOW_H

#include <opencv2/core/core.hpp>

struct md_ctx;

class VarFlow{

    public:

        /* common/src/VarFlow.cpp:27-162 */
        VarFlow(int width_in, int height_in, int max_level_in, int start_level_in, int n1_in, int n2_in,
        float rho_in, float alpha_in, float sigma_in);
        ~VarFlow();
        /* common/src/VarFlow.cpp:600-697; imgA/imgB 8-bit 1-channel, imgU/imgV IPL_DEPTH_32F; returns 1 on success, 0 otherwise */
        int CalcFlow(IplImage* imgA, IplImage* imgB, IplImage* imgU, IplImage* imgV, bool saved_data);

        /* the GPU the next VarFlow objects are created on (the reference's constructor has no such argument) */
        static void setDefaultDevice(int device) { default_device_ = device; }

    private:
        static int default_device_;
        VarFlow(const VarFlow &);
        VarFlow &operator=(const VarFlow &);
        md_ctx *ctx_;
        int width, height;
        int initialized;
};

#endif
