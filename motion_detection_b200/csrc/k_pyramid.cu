// k_pyramid.cu -- K0 (gray) + K1 (pyramid + Scharr planes) for sm_100a.
//
// Replaces cv::cvtColor(CV_BGR2GRAY) (common/src/optical_flow_calculator.cpp:50-51) and
// cv::buildOpticalFlowPyramid(gray, pyr, Size(40,40), 5, true) (cpp:67,170):
//   level l+1 = pyrDown(level l): separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8, size (w+1)/2;
//   per level Scharr derivative planes (int16 x2), image planes padded REFLECT_101 by the window size,
//   derivative planes padded with zeros.
// Integer arithmetic throughout: results are bit-exact with the CPU oracle.
#include "md_internal.h"

__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__device__ __forceinline__ uint32_t gray_px(const uint8_t *s)
{
    return (s[0] * 3735u + s[1] * 19235u + s[2] * 9798u + (1u << 14)) >> 15;
}

// ---- K0 standalone: 8UC3 -> 8UC1 ------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gray(const uint8_t *__restrict__ src, int spitch, int w, int h,
                                              uint8_t *__restrict__ dst, int dpitch)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= w) return;
    dst[(size_t)y * dpitch + x] = (uint8_t)gray_px(src + (size_t)y * spitch + 3 * x);
}

cudaError_t launch_gray(const uint8_t *src3, int src_pitch, int w, int h, uint8_t *dst, int dst_pitch, cudaStream_t s)
{
    dim3 grid((w + 255) / 256, h);
    k_gray<<<grid, 256, 0, s>>>(src3, src_pitch, w, h, dst, dst_pitch);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

// ---- level 0: plain frame (gray or 8UC3) -> padded plane with REFLECT_101 frame --------------------------------
// One thread writes 4 consecutive padded pixels as one 32-bit store (padded rows are 128-byte aligned).
template <int CH>
__global__ void __launch_bounds__(256) k_level0(const uint8_t *__restrict__ frames, int fpitch, long long fstride,
                                                uint8_t *__restrict__ img, size_t slot_bytes, int slot0, int nslots,
                                                LevelGeom L, int padx, int pady)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;      // group of 4 padded columns
    int py = blockIdx.y;
    int f = blockIdx.z;
    if (q * 4 >= L.pitch) return;
    const uint8_t *src = frames + (size_t)f * fstride;
    uint8_t *plane = img + (size_t)((slot0 + f) % nslots) * slot_bytes + L.img_off;
    int sy = reflect101(py - pady, L.h);
    const uint8_t *row = src + (size_t)sy * fpitch;
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int px = q * 4 + i - padx;
        uint32_t g = 0;
        if (px >= -padx && px < L.w + padx) {
            int sx = reflect101(px, L.w);
            g = CH == 1 ? row[sx] : gray_px(row + 3 * sx);
        }
        v |= g << (8 * i);
    }
    *reinterpret_cast<uint32_t *>(plane + (size_t)py * L.pitch + q * 4) = v;
}

// ---- level l -> l+1: pyrDown straight into the padded plane (frame pixels computed at reflected coordinates) -----
__global__ void __launch_bounds__(256) k_pyrdown(uint8_t *__restrict__ img, size_t slot_bytes, int slot0, int nslots,
                                                 LevelGeom S, LevelGeom D, int padx, int pady)
{
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    int py = blockIdx.y;
    int f = blockIdx.z;
    if (q * 4 >= D.pitch) return;
    uint8_t *base = img + (size_t)((slot0 + f) % nslots) * slot_bytes;
    const uint8_t *sp = base + S.img_off + (size_t)pady * S.pitch + padx;   // source interior origin
    uint8_t *dp = base + D.img_off;
    int oy = reflect101(py - pady, D.h);
    int px0 = q * 4 - padx;
    uint32_t v = 0;
    if (px0 >= 0 && px0 + 3 < D.w) {
        // fast path: 4 interior outputs share an 11-column source strip
        int acc[4] = {0, 0, 0, 0};
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const uint8_t *r = sp + (ptrdiff_t)(2 * oy + j - 2) * S.pitch + 2 * px0 - 2;
            int c[11];
#pragma unroll
            for (int i = 0; i < 11; i++) c[i] = r[i];
            const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
#pragma unroll
            for (int o = 0; o < 4; o++)
                acc[o] += kj * (c[2 * o] + 4 * c[2 * o + 1] + 6 * c[2 * o + 2] + 4 * c[2 * o + 3] + c[2 * o + 4]);
        }
#pragma unroll
        for (int o = 0; o < 4; o++) v |= (uint32_t)((acc[o] + 128) >> 8) << (8 * o);
    } else {
#pragma unroll
        for (int o = 0; o < 4; o++) {
            int px = px0 + o;
            if (px < -padx || px >= D.w + padx) continue;
            int ox = reflect101(px, D.w);
            int acc = 0;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const uint8_t *r = sp + (ptrdiff_t)(2 * oy + j - 2) * S.pitch + 2 * ox - 2;
                const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
                acc += kj * (r[0] + 4 * r[1] + 6 * r[2] + 4 * r[3] + r[4]);
            }
            v |= (uint32_t)((acc + 128) >> 8) << (8 * o);
        }
    }
    *reinterpret_cast<uint32_t *>(dp + (size_t)py * D.pitch + q * 4) = v;
}

// ---- Scharr planes (calcScharrDeriv): Ix = t0[x+1]-t0[x-1], t0 = 3(up+down)+10 mid; Iy = 3(t1[x-1]+t1[x+1])+10 t1[x] ---
__global__ void __launch_bounds__(256) k_scharr(const uint8_t *__restrict__ img, short2 *__restrict__ der,
                                                size_t slot_bytes, size_t slot_der, int slot0, int nslots, LevelGeom L,
                                                int padx, int pady)
{
    // one thread = 4 consecutive pixels: three aligned 32-bit loads per row (x-4 .. x+7), one 16-byte store
    const int x = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int f = blockIdx.z;
    if (x >= L.w) return;
    const int slot = (slot0 + f) % nslots;
    const uint8_t *p = img + (size_t)slot * slot_bytes + L.img_off + (size_t)(pady + y) * L.pitch + padx + x;
    int t0[6], t1[6];      // column sums for x-1 .. x+4
    {
        const uint32_t *u = reinterpret_cast<const uint32_t *>(p - L.pitch), *m = reinterpret_cast<const uint32_t *>(p),
                       *d = reinterpret_cast<const uint32_t *>(p + L.pitch);
        const uint32_t ul = __ldg(u - 1), uc = __ldg(u), ur = __ldg(u + 1);
        const uint32_t ml = __ldg(m - 1), mc = __ldg(m), mr = __ldg(m + 1);
        const uint32_t dl = __ldg(d - 1), dc = __ldg(d), dr = __ldg(d + 1);
#pragma unroll
        for (int i = 0; i < 6; i++) {
            // byte i-1 relative to x: i = 0 -> last byte of the left word, 1..4 -> centre word, 5 -> first byte of the right word
            const uint32_t uw = i == 0 ? ul >> 24 : (i == 5 ? ur & 0xffu : (uc >> (8 * (i - 1))) & 0xffu);
            const uint32_t mw = i == 0 ? ml >> 24 : (i == 5 ? mr & 0xffu : (mc >> (8 * (i - 1))) & 0xffu);
            const uint32_t dw = i == 0 ? dl >> 24 : (i == 5 ? dr & 0xffu : (dc >> (8 * (i - 1))) & 0xffu);
            t0[i] = (int)(uw + dw) * 3 + (int)mw * 10;
            t1[i] = (int)dw - (int)uw;
        }
    }
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int ix = t0[i + 2] - t0[i];
        const int iy = (t1[i] + t1[i + 2]) * 3 + t1[i + 1] * 10;
        o[i] = ((uint32_t)ix & 0xffffu) | ((uint32_t)iy << 16);
    }
    short2 *dst = der + (size_t)slot * slot_der + L.der_off + (size_t)(pady + y) * L.pitch + padx + x;
    if (x + 3 < L.w) *reinterpret_cast<uint4 *>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
    else
        for (int i = 0; i < 4 && x + i < L.w; i++) reinterpret_cast<uint32_t *>(dst)[i] = o[i];
}

// The three parts of the build, separately launchable (md_process_batch spreads the levels over several streams):
// level-0 image (gray conversion + padding), pyrDown for levels l0..l1 (needs level l0-1), Scharr planes for levels l0..l1.
cudaError_t launch_pyramid_level0(const PyrGeom &g, uint8_t *img, int slot0, int nframes, const uint8_t *frames, int channels,
                                  int fpitch, long long fstride, cudaStream_t s)
{
    const LevelGeom &L = g.lv[0];
    dim3 grid((L.pitch / 4 + 255) / 256, L.rows, nframes);
    if (channels == 1)
        k_level0<1><<<grid, 256, 0, s>>>(frames, fpitch, fstride, img, g.slot_img_bytes, slot0, g.nslots, L, g.padx, g.pady);
    else
        k_level0<3><<<grid, 256, 0, s>>>(frames, fpitch, fstride, img, g.slot_img_bytes, slot0, g.nslots, L, g.padx, g.pady);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

cudaError_t launch_pyramid_down(const PyrGeom &g, uint8_t *img, int slot0, int nframes, int l0, int l1, cudaStream_t s)
{
    for (int l = l0 < 1 ? 1 : l0; l <= l1 && l < g.nlev; l++) {
        const LevelGeom &D = g.lv[l];
        dim3 grid((D.pitch / 4 + 255) / 256, D.rows, nframes);
        k_pyrdown<<<grid, 256, 0, s>>>(img, g.slot_img_bytes, slot0, g.nslots, g.lv[l - 1], D, g.padx, g.pady);
        MD_COUNT_LAUNCH(1);
    }
    return cudaGetLastError();
}

cudaError_t launch_pyramid_scharr(const PyrGeom &g, uint8_t *img, short2 *der, int slot0, int nframes, int l0, int l1, cudaStream_t s)
{
    for (int l = l0; l <= l1 && l < g.nlev; l++) {
        const LevelGeom &L = g.lv[l];
        dim3 grid(((L.w + 3) / 4 + 127) / 128, L.h, nframes);
        k_scharr<<<grid, 128, 0, s>>>(img, der, g.slot_img_bytes, g.slot_der_elems, slot0, g.nslots, L, g.padx, g.pady);
        MD_COUNT_LAUNCH(1);
    }
    return cudaGetLastError();
}

cudaError_t launch_pyramid(const PyrGeom &g, uint8_t *img, short2 *der, int slot0, int nframes, const uint8_t *frames,
                           int channels, int fpitch, long long fstride, cudaStream_t s)
{
    cudaError_t e = launch_pyramid_level0(g, img, slot0, nframes, frames, channels, fpitch, fstride, s);
    if (e == cudaSuccess) e = launch_pyramid_down(g, img, slot0, nframes, 1, g.nlev - 1, s);
    if (e == cudaSuccess) e = launch_pyramid_scharr(g, img, der, slot0, nframes, 0, g.nlev - 1, s);
    return e;
}
