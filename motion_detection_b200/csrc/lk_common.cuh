// lk_common.cuh -- arithmetic shared by the LK kernels (k_lk.cu generic, k_lk_tma.cu production).
//
// Fixed-point scheme of OpenCV's LKTrackerInvoker (what cv::calcOpticalFlowPyrLK at
// common/src/optical_flow_calculator.cpp:71,172 runs): W_BITS = 14 bilinear weights, window samples descaled to
// 5 extra bits, derivative samples to 0 extra bits, sums scaled by 2^-20, f32 2x2 solve.
// The scalar f32 expressions use explicit _rn intrinsics so that no FMA contraction changes their rounding.
#pragma once
#include <float.h>

#include "md_internal.h"

#define W_BITS 14

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 2^-level exactly (what 1.f / (float)(1 << level) evaluates to), without the division routine
__device__ __forceinline__ float lk_level_scale(int level) { return __int_as_float((127 - level) << 23); }

__device__ __forceinline__ void lk_weights(float a, float b, int &w00, int &w01, int &w10, int &w11)
{
    float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), (float)(1 << W_BITS)));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), (float)(1 << W_BITS)));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), (float)(1 << W_BITS)));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// The scalar tail of one LK iteration: 2x2 solve, position update, both stopping rules.
// Returns true when the iteration loop must stop.
struct LkIterState { float npx, npy, pdx, pdy; };
__device__ __forceinline__ bool lk_update(float A11, float A12, float A22, float D, float fb1, float fb2, float half, int j,
                                          double eps2, LkIterState &s, float2 &nxt)
{
    const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), D);
    const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), D);
    s.npx = __fadd_rn(s.npx, ddx); s.npy = __fadd_rn(s.npy, ddy);
    nxt = make_float2(__fadd_rn(s.npx, half), __fadd_rn(s.npy, half));
    if (__dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)) <= eps2) return true;
    if (j > 0 && (double)fabsf(__fadd_rn(ddx, s.pdx)) < 0.01 && (double)fabsf(__fadd_rn(ddy, s.pdy)) < 0.01) {
        nxt.x = __fsub_rn(nxt.x, __fmul_rn(ddx, 0.5f));
        nxt.y = __fsub_rn(nxt.y, __fmul_rn(ddy, 0.5f));
        return true;
    }
    s.pdx = ddx; s.pdy = ddy;
    return false;
}

__device__ __forceinline__ bool lk_min_eig_ok(float A11, float A12, float A22, int win, float thr, float &Dinv)
{
    const float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float disc = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
    const float min_eig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(disc)), (float)(2 * win * win));
    if (min_eig < thr || D < FLT_EPSILON) return false;
    Dinv = __fdiv_rn(1.f, D);
    return true;
}

cudaError_t launch_lk_tma(const LkParams &p, const void *maps, int pairs, cudaStream_t s);
