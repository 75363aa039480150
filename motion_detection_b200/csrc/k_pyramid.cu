// k_pyramid.cu -- K0 (gray) + K1 (pyramid + Scharr planes) for sm_100a.
//
// Replaces cv::cvtColor(CV_BGR2GRAY) (common/src/optical_flow_calculator.cpp:50-51) and
// cv::buildOpticalFlowPyramid(gray, pyr, Size(40,40), 5, true) (cpp:67,170):
//   level l+1 = pyrDown(level l): separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8, size (w+1)/2
//   (shared-memory tiled, 16-byte loads);
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
// One thread writes 16 consecutive padded pixels as one 16-byte store (padded rows are 128-byte aligned, padx is a multiple of
// 16).  Groups that lie inside the image read their source with aligned 16-byte loads when the frame's base and pitch allow it
// (ALIGNED, decided by the host); groups that touch the REFLECT_101 frame take the per-pixel path.
template <int CH>
__device__ __forceinline__ uint32_t level0_px4(const uint8_t *row, int px, int w, int padx)
{
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t g = 0;
        if (px + i >= -padx && px + i < w + padx) {
            const int sx = reflect101(px + i, w);
            g = CH == 1 ? row[sx] : gray_px(row + 3 * sx);
        }
        v |= g << (8 * i);
    }
    return v;
}
__device__ __forceinline__ uint32_t gray_word4(uint32_t a, uint32_t b, uint32_t c)      // 12 source bytes -> 4 gray pixels
{
    const uint32_t g0 = ((a & 0xffu) * 3735u + ((a >> 8) & 0xffu) * 19235u + ((a >> 16) & 0xffu) * 9798u + (1u << 14)) >> 15;
    const uint32_t g1 = ((a >> 24) * 3735u + (b & 0xffu) * 19235u + ((b >> 8) & 0xffu) * 9798u + (1u << 14)) >> 15;
    const uint32_t g2 = (((b >> 16) & 0xffu) * 3735u + (b >> 24) * 19235u + (c & 0xffu) * 9798u + (1u << 14)) >> 15;
    const uint32_t g3 = (((c >> 8) & 0xffu) * 3735u + ((c >> 16) & 0xffu) * 19235u + (c >> 24) * 9798u + (1u << 14)) >> 15;
    return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}
template <int CH, bool ALIGNED>
__global__ void __launch_bounds__(128) k_level0(const uint8_t *__restrict__ frames, int fpitch, long long fstride,
                                                uint8_t *__restrict__ img, size_t slot_bytes, int slot0, int nslots,
                                                LevelGeom L, int padx, int pady)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;      // group of 16 padded columns
    const int py = blockIdx.y;
    const int f = blockIdx.z;
    if (q * 16 >= L.pitch) return;
    const uint8_t *src = frames + (size_t)f * fstride;
    uint8_t *plane = img + (size_t)((slot0 + f) % nslots) * slot_bytes + L.img_off;
    const int sy = reflect101(py - pady, L.h);
    const uint8_t *row = src + (size_t)sy * fpitch;
    const int px = q * 16 - padx;
    uint4 v;
    if (ALIGNED && px >= 0 && px + 16 <= L.w) {
        if (CH == 1) v = __ldg(reinterpret_cast<const uint4 *>(row + px));
        else {
            const uint4 a = __ldg(reinterpret_cast<const uint4 *>(row + 3 * px)), b = __ldg(reinterpret_cast<const uint4 *>(row + 3 * px + 16)),
                        c = __ldg(reinterpret_cast<const uint4 *>(row + 3 * px + 32));
            v = make_uint4(gray_word4(a.x, a.y, a.z), gray_word4(a.w, b.x, b.y), gray_word4(b.z, b.w, c.x), gray_word4(c.y, c.z, c.w));
        }
    } else {
        v = make_uint4(level0_px4<CH>(row, px, L.w, padx), level0_px4<CH>(row, px + 4, L.w, padx), level0_px4<CH>(row, px + 8, L.w, padx),
                       level0_px4<CH>(row, px + 12, L.w, padx));
    }
    *reinterpret_cast<uint4 *>(plane + (size_t)py * L.pitch + q * 16) = v;
}

// ---- level l -> l+1: pyrDown straight into the padded plane (frame pixels computed at reflected coordinates) -----
// One pixel of the padded destination plane at padded coordinates (px, py): the general path (REFLECT_101 frame, ragged edges).
__device__ __forceinline__ uint32_t pyrdown_px(const uint8_t *sp, int spitch, int ox, int oy)
{
    int acc = 0;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const uint8_t *r = sp + (ptrdiff_t)(2 * oy + j - 2) * spitch + 2 * ox - 2;
        const int kj = (j == 0 || j == 4) ? 1 : ((j == 2) ? 6 : 4);
        acc += kj * (r[0] + 4 * r[1] + 6 * r[2] + 4 * r[3] + r[4]);
    }
    return (uint32_t)((acc + 128) >> 8);
}

// Shared-memory tiled: a CTA owns a PD_TW x PD_TH tile of the padded destination plane.  Interior tiles (every output inside the
// level) stage the (2 PD_TW + 3) x (2 PD_TH + 3) source patch with aligned 16-byte loads, run the horizontal [1 4 6 4 1] pass
// once per source row into 16-bit sums (4 outputs per thread from three 32-bit shared loads), then the vertical pass from 8-byte
// shared loads with one coalesced 32-bit store per 4 outputs: 4.9 source bytes and ~26 instructions per output instead of 13.75
// byte loads and ~50.  Tiles that touch the REFLECT_101 frame take the per-pixel path (the source's own frame already holds the
// reflected taps; the destination's frame pixels are computed at reflected coordinates, no second border pass).
#define PD_TW 128
#define PD_TH 16
#define PD_SW (2 * PD_TW + 32)         // staged source row: 16-byte aligned window around 2 ox0 - 2 .. 2 ox0 + 2 PD_TW + 1
#define PD_SH (2 * PD_TH + 3)
__global__ void __launch_bounds__(256) k_pyrdown(uint8_t *__restrict__ img, size_t slot_bytes, int slot0, int nslots,
                                                 LevelGeom S, LevelGeom D, int padx, int pady)
{
    __shared__ __align__(16) uint8_t sS[PD_SH][PD_SW];
    __shared__ __align__(8) uint16_t sH[PD_SH][PD_TW];
    const int px0 = blockIdx.x * PD_TW, py0 = blockIdx.y * PD_TH;
    const int f = blockIdx.z;
    uint8_t *base = img + (size_t)((slot0 + f) % nslots) * slot_bytes;
    const uint8_t *sp = base + S.img_off + (size_t)pady * S.pitch + padx;   // source interior origin
    uint8_t *dp = base + D.img_off;
    const int ox0 = px0 - padx, oy0 = py0 - pady;
    const bool interior = ox0 >= 0 && ox0 + PD_TW <= D.w && oy0 >= 0 && oy0 + PD_TH <= D.h && ((padx | S.pitch) & 15) == 0;
    if (interior) {
        // ---- stage: source rows 2 oy0 - 2 .. 2 oy0 + 2 PD_TH, columns from the 16-byte boundary at or below 2 ox0 - 2
        const int c0 = (2 * ox0 - 2) & ~15;            // padx is a multiple of 16: the boundary is one in memory as well
        const int skew = 2 * ox0 - 2 - c0;             // 14 (ox0 is a multiple of 128 minus padx = a multiple of 64)
        const uint8_t *src = sp + (ptrdiff_t)(2 * oy0 - 2) * S.pitch + c0;
        for (int i = threadIdx.x; i < PD_SH * (PD_SW / 16); i += 256) {
            const int r = i / (PD_SW / 16), c = i - r * (PD_SW / 16);
            *reinterpret_cast<uint4 *>(&sS[r][16 * c]) = __ldg(reinterpret_cast<const uint4 *>(src + (ptrdiff_t)r * S.pitch + 16 * c));
        }
        __syncthreads();
        // ---- horizontal pass: 4 outputs (source bytes skew + 8 q .. + 10) per thread-iteration
        for (int i = threadIdx.x; i < PD_SH * (PD_TW / 4); i += 256) {
            const int r = i / (PD_TW / 4), q = i - r * (PD_TW / 4);
            const int o = skew + 8 * q;                // even; the three words around it cover 11 bytes whatever o & 3 is
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(&sS[r][o & ~3]);
            const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2], w3 = wp[3];
            const int sh = (o & 3) * 8;
            const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh), c = __funnelshift_r(w2, w3, sh);
            int p[11];
#pragma unroll
            for (int k = 0; k < 4; k++) { p[k] = (a >> (8 * k)) & 0xff; p[4 + k] = (b >> (8 * k)) & 0xff; }
            p[8] = c & 0xff; p[9] = (c >> 8) & 0xff; p[10] = (c >> 16) & 0xff;
            uint32_t h[4];
#pragma unroll
            for (int k = 0; k < 4; k++) h[k] = (uint32_t)(p[2 * k] + 4 * p[2 * k + 1] + 6 * p[2 * k + 2] + 4 * p[2 * k + 3] + p[2 * k + 4]);
            *reinterpret_cast<uint2 *>(&sH[r][4 * q]) = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
        }
        __syncthreads();
        // ---- vertical pass + store
        for (int i = threadIdx.x; i < PD_TH * (PD_TW / 4); i += 256) {
            const int r = i / (PD_TW / 4), q = i - r * (PD_TW / 4);
            uint32_t lo = 0, hi = 0;                   // two packed 16-bit sums each (max 255 * 256 < 65536)
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const uint2 v = *reinterpret_cast<const uint2 *>(&sH[2 * r + j][4 * q]);
                const uint32_t kj = (j == 0 || j == 4) ? 1u : ((j == 2) ? 6u : 4u);
                lo += kj * v.x; hi += kj * v.y;
            }
            // (sum + 128) >> 8 per half: the sums stay below 65536, so the halves do not carry into each other
            lo += 0x00800080u; hi += 0x00800080u;
            const uint32_t v = ((lo >> 8) & 0xffu) | ((lo >> 16) & 0xff00u) | (((hi >> 8) & 0xffu) << 16) | ((hi >> 24) << 24);
            *reinterpret_cast<uint32_t *>(dp + (size_t)(py0 + r) * D.pitch + px0 + 4 * q) = v;
        }
        return;
    }
    // ---- tiles on the frame / ragged edges: per-pixel path
    for (int i = threadIdx.x; i < PD_TH * (PD_TW / 4); i += 256) {
        const int r = i / (PD_TW / 4), q = i - r * (PD_TW / 4);
        const int py = py0 + r;
        if (py >= D.rows || px0 + 4 * q >= D.pitch) continue;
        const int oy = reflect101(py - pady, D.h);
        uint32_t v = 0;
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const int px = ox0 + 4 * q + o;
            if (px < -padx || px >= D.w + padx) continue;
            v |= pyrdown_px(sp, S.pitch, reflect101(px, D.w), oy) << (8 * o);
        }
        *reinterpret_cast<uint32_t *>(dp + (size_t)py * D.pitch + px0 + 4 * q) = v;
    }
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
    dim3 grid((L.pitch / 16 + 127) / 128, L.rows, nframes);
    // 16-byte source loads need a 16-byte aligned base, pitch and frame stride (and padx a multiple of 16: column groups of the
    // padded plane then start on multiples of 16 source pixels)
    const bool al = (((uintptr_t)frames | (uintptr_t)fpitch | (uintptr_t)fstride | (uintptr_t)g.padx) & 15) == 0;
#define MD_L0(CH, AL) k_level0<CH, AL><<<grid, 128, 0, s>>>(frames, fpitch, fstride, img, g.slot_img_bytes, slot0, g.nslots, L, g.padx, g.pady)
    if (channels == 1) { if (al) MD_L0(1, true); else MD_L0(1, false); }
    else { if (al) MD_L0(3, true); else MD_L0(3, false); }
#undef MD_L0
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

cudaError_t launch_pyramid_down(const PyrGeom &g, uint8_t *img, int slot0, int nframes, int l0, int l1, cudaStream_t s)
{
    for (int l = l0 < 1 ? 1 : l0; l <= l1 && l < g.nlev; l++) {
        const LevelGeom &D = g.lv[l];
        dim3 grid((D.pitch + PD_TW - 1) / PD_TW, (D.rows + PD_TH - 1) / PD_TH, nframes);
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
