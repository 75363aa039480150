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
// Layout: CTA = 256 threads, output tile 128 x 32.  A thread owns runs of 4 consecutive pixels: the thresholded
// bytes (0x00 / 0xFF) of a run are one 32-bit word in shared memory, so the 3x3 erode / dilate are AND / OR of nine
// funnel-shifted words (four pixels per instruction) and the mask leaves as coalesced 32-bit stores.
// The per-pixel projective divide is the FP64 cost centre: 32/den is taken from a Newton-refined reciprocal
// (error < 1e-15 relative), and only when the scaled coordinate lands within 1e-7 of a rounding boundary -- where the
// last-bit difference to the correctly rounded quotient could change rint() -- is the IEEE division redone, so the
// result stays bit-identical to the reference at a third of the FP64 work.
#include "md_internal.h"

#define TW 128
#define TH 32
#define GW (TW / 4 + 2)      // 34 groups of 4 columns: [tx0 - 4, tx0 + 132)
#define TR (TH + 4)          // 36 rows of thresholded words
#define ER (TH + 2)          // 34 rows of eroded words
#define SP (GW + 1)          // shared-memory pitch in words

__device__ __forceinline__ int bilinear_fetch(const uint8_t *__restrict__ src, int pitch, int w, int h, int X, int Y)
{
    int sx = X >> 5, sy = Y >> 5;
    const int ax = X & 31, ay = Y & 31;
    int p00 = 0, p01 = 0, p10 = 0, p11 = 0;
    if ((unsigned)sx < (unsigned)(w - 1) && (unsigned)sy < (unsigned)(h - 1)) {
        // common case: the 2x2 footprint is inside the image -> four unconditional loads from one base
        const uint8_t *r = src + (sy * pitch + sx);
        p00 = __ldg(r); p01 = __ldg(r + 1); p10 = __ldg(r + pitch); p11 = __ldg(r + pitch + 1);
    } else {
        sx = max(-32768, min(32767, sx));      // saturate_cast<short> of the remap tables
        sy = max(-32768, min(32767, sy));
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
    }
    // (32-ax)(32-ay), ax(32-ay), (32-ax)ay, ax*ay from one product
    const int w11 = ax * ay, w01 = (ax << 5) - w11, w10 = (ay << 5) - w11, w00 = 1024 - (ax << 5) - w10;
    const int v = w00 * p00 + w01 * p01 + w10 * p10 + w11 * p11;
    return (v + 512) >> 10;
}

// rint(num * (32 / den)) exactly as the reference computes it (W = den ? 32/den : 0; fX = num * W; saturating rint)
__device__ __forceinline__ void project(double nx, double ny, double den, int &X, int &Y)
{
    if (den == 0.0) { X = 0; Y = 0; return; }
    // Newton-refined reciprocal: two steps from the f32 seed give < 1e-15 relative error
    double r = (double)__frcp_rn((float)den);
    r = fma(r, fma(-den, r, 1.0), r);
    r = fma(r, fma(-den, r, 1.0), r);
    const double W = 32.0 * r;
    const double fX = nx * W, fY = ny * W;
    X = __double2int_rn(fX);
    Y = __double2int_rn(fY);
    // distance to the nearest rounding boundary (k + 0.5); tiny -> the last bits of W matter -> exact path
    const double dx = 0.5 - fabs(fX - (double)X), dy = 0.5 - fabs(fY - (double)Y);
    const double tol = 1e-7;
    // (|fX| < 1e6 bounds the absolute error of the fast path by 1e6 * 7e-16 << tol; it also catches inf / NaN)
    if (dx < tol || dy < tol || !(fabs(fX) < 1e6) || !(fabs(fY) < 1e6)) {
        const double We = __ddiv_rn(32.0, den);
        X = __double2int_rn(__dmul_rn(nx, We));     // cvt.rni.s32.f64 saturates like the reference's clamp
        Y = __double2int_rn(__dmul_rn(ny, We));
    }
}

template <bool ALIGNED>
__global__ void __launch_bounds__(256) k_mask(const MaskParams p)
{
    __shared__ uint32_t T[TR][SP];
    __shared__ uint32_t E[ER][SP];
    __shared__ double sM[9];
    const int b = blockIdx.z;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    uint8_t *out = p.mask + (size_t)b * p.mask_stride;
    const bool valid = p.valid ? p.valid[b] != 0 : true;
    const int w = p.w, h = p.h;
    if (!valid) {
        // no egomotion (fewer than the minimal number of vectors): empty mask
        for (int i = threadIdx.x; i < TW * TH; i += 256) {
            int x = tx0 + i % TW, y = ty0 + i / TW;
            if (x < w && y < h) out[(size_t)y * p.mask_pitch + x] = 0;
        }
        return;
    }
    if (threadIdx.x < 9) sM[threadIdx.x] = p.Hinv[b * 9 + threadIdx.x];
    __syncthreads();
    const uint8_t *prev = p.prev + (long long)(p.nslots ? (p.prev_slot0 + b) % p.nslots : b) * p.stride;
    const uint8_t *cur = p.cur + (long long)(p.nslots ? (p.cur_slot0 + b) % p.nslots : b) * p.stride;
    const int bh0 = h < 16 ? h : 16;
    const int bw0 = (1024 / bh0 < w) ? 1024 / bh0 : w;
    const double M0 = sM[0], M1 = sM[1], M2 = sM[2], M3 = sM[3], M4 = sM[4], M5 = sM[5], M6 = sM[6], M7 = sM[7], M8 = sM[8];
    const bool uniform_block = (bw0 & 3) == 0;     // a run of 4 aligned columns never straddles a 64-column block

    // ---- phase 1: warp + absdiff + threshold, 4 pixels per thread-iteration ---------------------------------------
    for (int g = threadIdx.x; g < GW * TR; g += 256) {
        const int ry = g / GW, gx = g - ry * GW;
        const int x = tx0 - 4 + 4 * gx, y = ty0 - 2 + ry;
        uint32_t word = 0xffffffffu;            // erode identity outside the image
        if (y >= 0 && y < h && x + 3 >= 0 && x < w) {
            uint32_t c4;
            if (ALIGNED && x >= 0 && x + 3 < w) c4 = __ldg(reinterpret_cast<const uint32_t *>(cur + (size_t)y * p.pitch + x));
            else {
                c4 = 0;
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j >= 0 && x + j < w) c4 |= (uint32_t)__ldg(cur + (size_t)y * p.pitch + x + j) << (8 * j);
            }
            const double dy = (double)y;
            double X0 = 0, Y0 = 0, W0 = 0;
            int bx = 0;
            if (uniform_block) {
                const int xc = x < 0 ? 0 : x;
                bx = bw0 == 64 ? (xc & ~63) : xc - xc % bw0;
                const double dbx = (double)bx;
                X0 = __dadd_rn(__dadd_rn(__dmul_rn(M0, dbx), __dmul_rn(M1, dy)), M2);
                Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M3, dbx), __dmul_rn(M4, dy)), M5);
                W0 = __dadd_rn(__dadd_rn(__dmul_rn(M6, dbx), __dmul_rn(M7, dy)), M8);
            }
            word = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int xj = x + j;
                uint32_t t = 0xffu;
                if (xj >= 0 && xj < w) {
                    if (!uniform_block) {
                        bx = xj - xj % bw0;
                        const double dbx = (double)bx;
                        X0 = __dadd_rn(__dadd_rn(__dmul_rn(M0, dbx), __dmul_rn(M1, dy)), M2);
                        Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M3, dbx), __dmul_rn(M4, dy)), M5);
                        W0 = __dadd_rn(__dadd_rn(__dmul_rn(M6, dbx), __dmul_rn(M7, dy)), M8);
                    }
                    const double dx1 = (double)(xj - bx);
                    const double den = __dadd_rn(W0, __dmul_rn(M6, dx1));
                    const double nx = __dadd_rn(X0, __dmul_rn(M0, dx1)), ny = __dadd_rn(Y0, __dmul_rn(M3, dx1));
                    int X, Y;
                    project(nx, ny, den, X, Y);
                    const int wv = bilinear_fetch(prev, p.pitch, w, h, X, Y);
                    int d = wv - (int)((c4 >> (8 * j)) & 0xffu);
                    d = d < 0 ? -d : d;
                    t = d > p.thresh ? 0xffu : 0u;
                }
                word |= t << (8 * j);
            }
        }
        T[ry][gx] = word;
    }
    __syncthreads();

    int local = 0;
    if (p.morph) {
        // ---- phase 2: erode = AND of the nine neighbours, four pixels per word ----------------------------------
        for (int g = threadIdx.x; g < GW * ER; g += 256) {
            const int ry = g / GW, gx = g - ry * GW;
            const int x = tx0 - 4 + 4 * gx, y = ty0 - 1 + ry;
            uint32_t e = 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const uint32_t c = T[ry + j][gx];
                const uint32_t l = gx > 0 ? T[ry + j][gx - 1] : 0xffffffffu;
                const uint32_t r = gx < GW - 1 ? T[ry + j][gx + 1] : 0xffffffffu;
                e &= c & __funnelshift_l(l, c, 8) & __funnelshift_r(c, r, 8);
            }
            // dilate identity (0) for pixels outside the image
            uint32_t m = 0;
            if (y >= 0 && y < h) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j >= 0 && x + j < w) m |= 0xffu << (8 * j);
            }
            E[ry][gx] = e & m;
        }
        __syncthreads();
        // ---- phase 3: dilate = OR of the nine neighbours, coalesced 32-bit stores -------------------------------
        for (int g = threadIdx.x; g < (TW / 4) * TH; g += 256) {
            const int ry = g / (TW / 4), go = g - ry * (TW / 4);
            const int gx = go + 1;
            const int x = tx0 + 4 * go, y = ty0 + ry;
            if (x >= w || y >= h) continue;
            uint32_t o = 0;
#pragma unroll
            for (int j = 0; j < 3; j++) {
                const uint32_t c = E[ry + j][gx], l = E[ry + j][gx - 1], r = E[ry + j][gx + 1];
                o |= c | __funnelshift_l(l, c, 8) | __funnelshift_r(c, r, 8);
            }
            uint8_t *dst = out + (size_t)y * p.mask_pitch + x;
            if (ALIGNED && x + 3 < w) {
                *reinterpret_cast<uint32_t *>(dst) = o;
                local += __popc(o & 0x01010101u);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j < w) { dst[j] = (uint8_t)(o >> (8 * j)); local += (o >> (8 * j)) & 1; }
            }
        }
    } else {
        for (int g = threadIdx.x; g < (TW / 4) * TH; g += 256) {
            const int ry = g / (TW / 4), go = g - ry * (TW / 4);
            const int x = tx0 + 4 * go, y = ty0 + ry;
            if (x >= w || y >= h) continue;
            const uint32_t o = T[ry + 2][go + 1];
            uint8_t *dst = out + (size_t)y * p.mask_pitch + x;
            if (ALIGNED && x + 3 < w) {
                *reinterpret_cast<uint32_t *>(dst) = o;
                local += __popc(o & 0x01010101u);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j < w) { dst[j] = (uint8_t)(o >> (8 * j)); local += (o >> (8 * j)) & 1; }
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
    const bool aligned = (((uintptr_t)p.cur | (uintptr_t)p.mask | (uintptr_t)p.pitch | (uintptr_t)p.mask_pitch |
                           (uintptr_t)p.stride | (uintptr_t)p.mask_stride) & 3) == 0;
    if (aligned) k_mask<true><<<grid, 256, 0, s>>>(p);
    else k_mask<false><<<grid, 256, 0, s>>>(p);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
