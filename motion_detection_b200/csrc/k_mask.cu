// k_mask.cu -- K4: fused warpPerspective + absdiff + threshold + erode + dilate (one pass over HBM: prev, cur, mask).
//
// Replaces, in one kernel,
//   cv::warpPerspective(gray1, compensated, H, size)   common/src/optical_flow_calculator.cpp:124
//   cv::absdiff(compensated, gray2, comp)              common/src/optical_flow_calculator.cpp:125
//   cv::threshold(comp, comp, 190, 255, BINARY)        common/src/optical_flow_calculator.cpp:127
//   cv::erode / cv::dilate (3x3, 1 iteration)          common/src/background_subtractor.cpp:31-32
//
// Bit-exact integer model of OpenCV's fixed-point bilinear remap: per destination pixel the f64
// projective coordinates are formed exactly like WarpPerspectiveInvoker does (row base at the 64-column block origin
// plus per-column increment, no FMA contraction), scaled by 32, rounded half-to-even, split into integer position and
// 5-bit fraction; out = ((32-ax)(32-ay) p00 + ax(32-ay) p01 + (32-ax)ay p10 + ax ay p11 + 512) >> 10, taps outside the
// image read 0.  Threshold is strict '>'.  Erode ignores out-of-image pixels (identity 255), dilate likewise (identity 0).
//
// v1 layout: CTA = 256 threads, output tile 128 x 16, thresholded bits for the (tile + 2) halo kept in shared memory.
#include "md_internal.h"

#define TW 128
#define TH 16
#define RW (TW + 4)
#define RH (TH + 4)

__device__ __forceinline__ int warp_sample(const uint8_t *__restrict__ src, int pitch, int w, int h, const double *M,
                                           int x, int y, int bw0)
{
    const int bx = x - x % bw0, x1 = x - bx;
    const double dbx = (double)bx, dy = (double)y, dx1 = (double)x1;
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dbx), __dmul_rn(M[1], dy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dbx), __dmul_rn(M[4], dy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dbx), __dmul_rn(M[7], dy)), M[8]);
    double W = __dadd_rn(W0, __dmul_rn(M[6], dx1));
    W = W != 0.0 ? __ddiv_rn(32.0, W) : 0.0;
    double fX = __dmul_rn(__dadd_rn(X0, __dmul_rn(M[0], dx1)), W);
    double fY = __dmul_rn(__dadd_rn(Y0, __dmul_rn(M[3], dx1)), W);
    fX = fmax(-2147483648.0, fmin(2147483647.0, fX));
    fY = fmax(-2147483648.0, fmin(2147483647.0, fY));
    const int X = __double2int_rn(fX), Y = __double2int_rn(fY);
    int sx = X >> 5, sy = Y >> 5;
    sx = max(-32768, min(32767, sx));
    sy = max(-32768, min(32767, sy));
    const int ax = X & 31, ay = Y & 31;
    int p00 = 0, p01 = 0, p10 = 0, p11 = 0;
    const bool x0in = (unsigned)sx < (unsigned)w, x1in = (unsigned)(sx + 1) < (unsigned)w;
    if ((unsigned)sy < (unsigned)h) {
        const uint8_t *r = src + (size_t)sy * pitch;
        if (x0in) p00 = __ldg(r + sx);
        if (x1in) p01 = __ldg(r + sx + 1);
    }
    if ((unsigned)(sy + 1) < (unsigned)h) {
        const uint8_t *r = src + (size_t)(sy + 1) * pitch;
        if (x0in) p10 = __ldg(r + sx);
        if (x1in) p11 = __ldg(r + sx + 1);
    }
    const int v = (32 - ax) * (32 - ay) * p00 + ax * (32 - ay) * p01 + (32 - ax) * ay * p10 + ax * ay * p11;
    return (v + 512) >> 10;
}

__global__ void __launch_bounds__(256) k_mask(const MaskParams p)
{
    __shared__ uint8_t T[RH][RW + 4];
    __shared__ uint8_t E[RH][RW + 4];
    __shared__ double sM[9];
    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    uint8_t *out = p.mask + (size_t)b * p.mask_stride;
    const bool valid = p.valid ? p.valid[b] != 0 : true;
    if (!valid) {
        // no egomotion (fewer than the minimal number of vectors): empty mask
        for (int i = threadIdx.x; i < TW * TH; i += 256) {
            int x = tx0 + i % TW, y = ty0 + i / TW;
            if (x < p.w && y < p.h) out[(size_t)y * p.mask_pitch + x] = 0;
        }
        return;
    }
    if (threadIdx.x < 9) sM[threadIdx.x] = p.Hinv[b * 9 + threadIdx.x];
    __syncthreads();
    const uint8_t *prev = p.prev + (long long)(p.nslots ? (p.prev_slot0 + b) % p.nslots : b) * p.stride;
    const uint8_t *cur = p.cur + (long long)(p.nslots ? (p.cur_slot0 + b) % p.nslots : b) * p.stride;
    const int bh0 = p.h < 16 ? p.h : 16;
    const int bw0 = (1024 / bh0 < p.w) ? 1024 / bh0 : p.w;
    for (int i = threadIdx.x; i < RW * RH; i += 256) {
        const int rx = i % RW, ry = i / RW;
        const int x = tx0 + rx - 2, y = ty0 + ry - 2;
        uint8_t t = 255;   // erode identity outside the image
        if (x >= 0 && y >= 0 && x < p.w && y < p.h) {
            int wv = warp_sample(prev, p.pitch, p.w, p.h, sM, x, y, bw0);
            int d = wv - (int)__ldg(cur + (size_t)y * p.pitch + x);
            d = d < 0 ? -d : d;
            t = d > p.thresh ? 255 : 0;
        }
        T[ry][rx] = t;
    }
    __syncthreads();
    int local = 0;
    if (p.morph) {
        for (int i = threadIdx.x; i < (TW + 2) * (TH + 2); i += 256) {
            const int rx = i % (TW + 2) + 1, ry = i / (TW + 2) + 1;
            const int x = tx0 + rx - 2, y = ty0 + ry - 2;
            uint8_t e = 0;     // dilate identity outside the image
            if (x >= 0 && y >= 0 && x < p.w && y < p.h) {
                e = 255;
#pragma unroll
                for (int j = -1; j <= 1; j++)
#pragma unroll
                    for (int k = -1; k <= 1; k++) e = min(e, T[ry + j][rx + k]);
            }
            E[ry][rx] = e;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < TW * TH; i += 256) {
            const int rx = i % TW + 2, ry = i / TW + 2;
            const int x = tx0 + rx - 2, y = ty0 + ry - 2;
            if (x < p.w && y < p.h) {
                uint8_t m = 0;
#pragma unroll
                for (int j = -1; j <= 1; j++)
#pragma unroll
                    for (int k = -1; k <= 1; k++) m = max(m, E[ry + j][rx + k]);
                out[(size_t)y * p.mask_pitch + x] = m;
                local += m != 0;
            }
        }
    } else {
        for (int i = threadIdx.x; i < TW * TH; i += 256) {
            const int rx = i % TW + 2, ry = i / TW + 2;
            const int x = tx0 + rx - 2, y = ty0 + ry - 2;
            if (x < p.w && y < p.h) {
                out[(size_t)y * p.mask_pitch + x] = T[ry][rx];
                local += T[ry][rx] != 0;
            }
        }
    }
    if (p.stat_mask) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((threadIdx.x & 31) == 0 && local) atomicAdd(p.stat_mask, (unsigned long long)local);
    }
}

cudaError_t launch_mask(const MaskParams &p, int pairs, cudaStream_t s)
{
    dim3 grid((p.w + TW - 1) / TW, (p.h + TH - 1) / TH, pairs);
    k_mask<<<grid, 256, 0, s>>>(p);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
