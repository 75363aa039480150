// k_draw.cu -- OpticalFlowVisualizer::showOpticalFlowVectors on the device.
//
// Replaces the drawing loop of common/src/optical_flow_visualizer.cpp:23-71 (call sites ros/src/motion_detection_node.cpp:83,101):
// the original image with one anti-aliased arrow (shaft + two head strokes, cv::line(..., 1, CV_AA)) per flow vector whose
// components pass the reference's test (:37).  With it the flow field never has to travel to the host just to be drawn: the
// vectors come straight from the batch outputs on the device (next_pts / status / keep), or from a caller's Vec4d list.
//
// cv::line's anti-aliased rasteriser (OpenCV imgproc/src/drawing.cpp, LineAA) walks a line one major-axis pixel at a time and
// blends three minor-axis pixels per step; where arrows overlap, the result depends on the drawing ORDER (the blend is not
// commutative).  The device version keeps that order exactly and still runs pixel-parallel:
//   k_draw_setup : one thread per vector (in the reference's visiting order: rows outer, columns inner) forms the three integer
//                  segments (float / double arithmetic of the reference, cvRound), clips them like cv::clipLine, and stores per
//                  segment the closed form of LineAA's walk: first major pixel, step count, 16.16 minor start + per-step
//                  increment (integer adds: step k is start + k * increment exactly), end-point correction table, pixel box.
//   k_draw_tiles : a CTA owns a 32 x 16 pixel tile; it scans the arrows' pixel boxes 256 at a time, compacts the ones that touch
//                  the tile IN ORDER (ballot + prefix), and every thread then applies those segments to its own two pixels in that
//                  order: for a pixel a segment either misses or touches it exactly once, with coverage read off the closed form.
// Bit-identical to the oracle's sequential loop (tests/test_gpu_draw.py), which is pinned against cv2.line.
#include <math.h>

#include "md_internal.h"

#define XY_SHIFT 16
#define XY_ONE (1 << XY_SHIFT)

__constant__ uint8_t c_slope_corr[32] = {181, 181, 181, 182, 182, 183, 184, 185, 187, 188, 190, 192, 194, 196, 198, 201,
                                         203, 206, 209, 211, 214, 218, 221, 224, 227, 231, 235, 238, 242, 246, 250, 254};
__constant__ uint8_t c_aa_filter[64] = {168, 177, 185, 194, 202, 210, 218, 224, 231, 236, 241, 246, 249, 252, 254, 254,
                                        254, 254, 252, 249, 246, 241, 236, 231, 224, 218, 210, 202, 194, 185, 177, 168,
                                        158, 149, 140, 131, 122, 114, 105, 97,  89,  82,  75,  68,  62,  56,  50,  45,
                                        40,  36,  32,  28,  25,  22,  19,  16,  14,  12,  11,  9,   8,   7,   5,   5};

struct alignas(16) DrawSeg {
    long long m1, step;          // minor coordinate (16.16) at step 0, increment per major pixel
    int a0, ecount0;             // first major pixel, number of steps - 1
    int bx0, by0, bx1, by1;      // pixels the segment can touch (inclusive, clipped to the image); bx0 > bx1 = nothing
    unsigned short ep[9];
    unsigned short xmajor;
};

// cv::clipLine(Size2l, Point2l&, Point2l&)
__device__ bool draw_clip(long long w, long long h, long long &x1, long long &y1, long long &x2, long long &y2)
{
    const long long right = w - 1, bottom = h - 1;
    if (w <= 0 || h <= 0) return false;
    int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
    int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
        long long a;
        if (c1 & 12) {
            a = c1 < 8 ? 0 : bottom;
            x1 += (long long)__ddiv_rn(__dmul_rn((double)(a - y1), (double)(x2 - x1)), (double)(y2 - y1));
            y1 = a;
            c1 = (x1 < 0) + (x1 > right) * 2;
        }
        if (c2 & 12) {
            a = c2 < 8 ? 0 : bottom;
            x2 += (long long)__ddiv_rn(__dmul_rn((double)(a - y2), (double)(x2 - x1)), (double)(y2 - y1));
            y2 = a;
            c2 = (x2 < 0) + (x2 > right) * 2;
        }
        if ((c1 & c2) == 0 && (c1 | c2) != 0) {
            if (c1) {
                a = c1 == 1 ? 0 : right;
                y1 += (long long)__ddiv_rn(__dmul_rn((double)(a - x1), (double)(y2 - y1)), (double)(x2 - x1));
                x1 = a;
                c1 = 0;
            }
            if (c2) {
                a = c2 == 1 ? 0 : right;
                y2 += (long long)__ddiv_rn(__dmul_rn((double)(a - x2), (double)(y2 - y1)), (double)(x2 - x1));
                x2 = a;
                c2 = 0;
            }
        }
    }
    return (c1 | c2) == 0;
}

// LineAA's set-up for integer end points; false = nothing to draw
__device__ bool draw_seg_setup(DrawSeg &s, int px1, int py1, int px2, int py2, int w, int h)
{
    long long x1 = (long long)px1 << XY_SHIFT, y1 = (long long)py1 << XY_SHIFT, x2 = (long long)px2 << XY_SHIFT, y2 = (long long)py2 << XY_SHIFT;
    s.bx0 = 1; s.bx1 = 0; s.by0 = 1; s.by1 = 0;
    if (!draw_clip((long long)w << XY_SHIFT, (long long)h << XY_SHIFT, x1, y1, x2, y2)) return false;
    long long dx = x2 - x1, dy = y2 - y1;
    long long j = dx < 0 ? -1 : 0, ax = (dx ^ j) - j;
    long long i = dy < 0 ? -1 : 0, ay = (dy ^ i) - i;
    long long stepv;
    int ecount, slope;
    const bool xmajor = ax > ay;
    if (xmajor) {
        dy = (dy ^ j) - j;
        if (j) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        stepv = (dy * XY_ONE) / (ax | 1);
        x2 += XY_ONE;
        ecount = (int)((x2 >> XY_SHIFT) - (x1 >> XY_SHIFT));
        j = -(x1 & (XY_ONE - 1));
        y1 += ((stepv * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((stepv >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (stepv < 0 ? 0x3f : 0);
        i = (x1 >> (XY_SHIFT - 7)) & 0x78;
        j = (x2 >> (XY_SHIFT - 7)) & 0x78;
        s.a0 = (int)(x1 >> XY_SHIFT); s.m1 = y1;
    } else {
        dx = (dx ^ i) - i;
        if (i) { long long t = x1; x1 = x2; x2 = t; t = y1; y1 = y2; y2 = t; }
        stepv = (dx * XY_ONE) / (ay | 1);
        y2 += XY_ONE;
        ecount = (int)((y2 >> XY_SHIFT) - (y1 >> XY_SHIFT));
        j = -(y1 & (XY_ONE - 1));
        x1 += ((stepv * j) >> XY_SHIFT) + (XY_ONE >> 1);
        slope = (int)((stepv >> (XY_SHIFT - 5)) & 0x3f);
        slope ^= (stepv < 0 ? 0x3f : 0);
        i = (y1 >> (XY_SHIFT - 7)) & 0x78;
        j = (y2 >> (XY_SHIFT - 7)) & 0x78;
        s.a0 = (int)(y1 >> XY_SHIFT); s.m1 = x1;
    }
    slope = (slope & 0x20) ? 0x100 : c_slope_corr[slope];
    {
        const int t0 = slope << 7, t1 = ((0x78 - (int)i) | 4) * slope, t2 = ((int)j | 4) * slope;
        s.ep[0] = 0;
        s.ep[8] = (unsigned short)slope;
        s.ep[1] = s.ep[3] = (unsigned short)(((((int)(j - i) & 0x78) | 4) * slope >> 8) & 0x1ff);
        s.ep[2] = (unsigned short)((t1 >> 8) & 0x1ff);
        s.ep[4] = (unsigned short)(((((int)(j - i) + 0x80) | 4) * slope >> 8) & 0x1ff);
        s.ep[5] = (unsigned short)(((t1 + t0) >> 8) & 0x1ff);
        s.ep[6] = (unsigned short)((t2 >> 8) & 0x1ff);
        s.ep[7] = (unsigned short)(((t2 + t0) >> 8) & 0x1ff);
    }
    s.step = stepv; s.ecount0 = ecount; s.xmajor = xmajor ? 1 : 0;
    if (ecount < 0) return false;
    // pixels the walk can touch: major a0 .. a0 + ecount0, minor (m >> 16) - 1 .. + 1 at the two ends (the walk is monotonic)
    const long long mA = s.m1, mB = s.m1 + (long long)ecount * stepv;
    const int lo = (int)((mA < mB ? mA : mB) >> XY_SHIFT) - 1, hi = (int)((mA < mB ? mB : mA) >> XY_SHIFT) + 1;
    int bx0 = xmajor ? s.a0 : lo, bx1 = xmajor ? s.a0 + ecount : hi, by0 = xmajor ? lo : s.a0, by1 = xmajor ? hi : s.a0 + ecount;
    bx0 = max(bx0, 0); by0 = max(by0, 0); bx1 = min(bx1, w - 1); by1 = min(by1, h - 1);
    s.bx0 = bx0; s.bx1 = bx1; s.by0 = by0; s.by1 = by1;
    return bx0 <= bx1 && by0 <= by1;
}

struct DrawParams {
    int w, h, n;
    int ps;
    double min_vec;
    const double *vec4;          // [n][4] in drawing order, or nullptr: derive from the grid arrays below
    const float2 *next;          // [P] one pair
    const uint8_t *status, *keep;
    int gx, gy;
    DrawSeg *segs;               // [n][3]
    int4 *abox;                  // [n] pixel box of the whole arrow (x0, y0, x1, y1), x0 > x1 = not drawn
    int *drawn;
};

__global__ void __launch_bounds__(128) k_draw_setup(const DrawParams p)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= p.n) return;
    double e0, e1, e2, e3;
    if (p.vec4) { e0 = p.vec4[4 * q]; e1 = p.vec4[4 * q + 1]; e2 = p.vec4[4 * q + 2]; e3 = p.vec4[4 * q + 3]; }
    else {
        // the flow-field element of grid point (column gxi, row gyi), visited row by row (optical_flow_calculator.cpp:78-117)
        const int gyi = q / p.gx, gxi = q - gyi * p.gx, k = gxi * p.gy + gyi;
        const float sx = (float)(gxi * p.ps), sy = (float)(gyi * p.ps);
        e0 = -1.0; e1 = -1.0; e2 = 0.0; e3 = 0.0;
        if (p.status[k]) {
            e0 = (double)sx; e1 = (double)sy;
            if (p.keep[k]) { const float2 nx = p.next[k]; e2 = (double)__fsub_rn(nx.x, sx); e3 = (double)__fsub_rn(nx.y, sy); }
        }
    }
    DrawSeg *S = p.segs + (size_t)3 * q;
    int4 box = make_int4(1, 1, 0, 0);
    const double lim = (double)(p.ps * 5);
    if ((fabs(e2) > p.min_vec || fabs(e3) > p.min_vec) && fabs(e2) < lim && fabs(e3) < lim) {
        const float sx = (float)e0, sy = (float)e1;
        const float ex = (float)__dadd_rn((double)sx, e2), ey = (float)__dadd_rn((double)sy, e3);
        const double back = (double)atan2f(__fsub_rn(sy, ey), __fsub_rn(sx, ex));
        const double a1 = __dadd_rn(back, M_PI / 4.0), a2 = __dsub_rn(back, M_PI / 4.0);
        const float h1x = (float)__dadd_rn((double)ex, __dmul_rn(3.0, cos(a1))), h1y = (float)__dadd_rn((double)ey, __dmul_rn(3.0, sin(a1)));
        const float h2x = (float)__dadd_rn((double)ex, __dmul_rn(3.0, cos(a2))), h2y = (float)__dadd_rn((double)ey, __dmul_rn(3.0, sin(a2)));
        const int isx = __float2int_rn(sx), isy = __float2int_rn(sy), iex = __float2int_rn(ex), iey = __float2int_rn(ey);
        const int seg[3][4] = {{isx, isy, iex, iey}, {iex, iey, __float2int_rn(h1x), __float2int_rn(h1y)},
                               {iex, iey, __float2int_rn(h2x), __float2int_rn(h2y)}};
        int x0 = 1 << 30, y0 = 1 << 30, x1 = -(1 << 30), y1 = -(1 << 30);
#pragma unroll
        for (int l = 0; l < 3; l++) {
            if (draw_seg_setup(S[l], seg[l][0], seg[l][1], seg[l][2], seg[l][3], p.w, p.h)) {
                x0 = min(x0, S[l].bx0); y0 = min(y0, S[l].by0); x1 = max(x1, S[l].bx1); y1 = max(y1, S[l].by1);
            }
        }
        if (x0 <= x1) box = make_int4(x0, y0, x1, y1);
        if (p.drawn) atomicAdd(p.drawn, 1);
    } else {
#pragma unroll
        for (int l = 0; l < 3; l++) { S[l].bx0 = 1; S[l].bx1 = 0; S[l].by0 = 1; S[l].by1 = 0; }
    }
    p.abox[q] = box;
}

#define DT_W 32
#define DT_H 16

template <int NCH>
__device__ __forceinline__ void draw_apply(const DrawSeg &L, int x, int y, const uint8_t *colour, int (&c)[NCH])
{
    if (x < L.bx0 || x > L.bx1 || y < L.by0 || y > L.by1) return;
    const int maj = L.xmajor ? x : y, mino = L.xmajor ? y : x;
    const int k = maj - L.a0;
    if (k < 0 || k > L.ecount0) return;
    const long long m = L.m1 + (long long)k * L.step;
    const int r = mino - ((int)(m >> XY_SHIFT) - 1);
    if (r < 0 || r > 2) return;
    const int dist = (int)((m >> (XY_SHIFT - 5)) & 31);
    const int idx = r == 0 ? dist + 32 : (r == 1 ? dist : 63 - dist);
    const int scount = k, ecount = L.ecount0 - k;
    const int epc = L.ep[((((scount >= 2) + 1) & (scount | 2)) * 3) + (((ecount >= 2) + 1) & (ecount | 2))];
    const int a = (epc * (int)c_aa_filter[idx] >> 8) & 0xff;
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        int v = c[ch];
        v += (((int)colour[ch] - v) * a + 127) >> 8;
        v += (((int)colour[ch] - v) * a + 127) >> 8;
        c[ch] = v;
    }
}

template <int NCH>
__global__ void __launch_bounds__(256) k_draw_tiles(const uint8_t *__restrict__ src, int spitch, uint8_t *__restrict__ dst, int dpitch, int w, int h,
                                                    const DrawSeg *__restrict__ segs, const int4 *__restrict__ abox, int n, uchar4 colour4)
{
    __shared__ int s_list[256];
    __shared__ int s_wcount[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx0 = blockIdx.x * DT_W, ty0 = blockIdx.y * DT_H;
    const int tx1 = min(tx0 + DT_W, w) - 1, ty1 = min(ty0 + DT_H, h) - 1;
    const uint8_t colour[4] = {colour4.x, colour4.y, colour4.z, colour4.w};
    const int x = tx0 + lane, ya = ty0 + warp, yb = ty0 + warp + 8;
    const bool ina = x < w && ya < h, inb = x < w && yb < h;
    int ca[NCH], cb[NCH];
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        ca[ch] = ina ? src[(size_t)ya * spitch + (size_t)x * NCH + ch] : 0;
        cb[ch] = inb ? src[(size_t)yb * spitch + (size_t)x * NCH + ch] : 0;
    }
    for (int base = 0; base < n; base += 256) {
        // ---- the arrows of this round that touch the tile, in drawing order
        const int k = base + tid;
        bool hit = false;
        if (k < n) {
            const int4 b = __ldg(abox + k);
            hit = b.x <= b.z && b.x <= tx1 && b.z >= tx0 && b.y <= ty1 && b.w >= ty0;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (lane == 0) s_wcount[warp] = __popc(bal);
        __syncthreads();
        int off = 0, total = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const int c = s_wcount[i]; off += i < warp ? c : 0; total += c; }
        if (hit) s_list[off + __popc(bal & ((1u << lane) - 1))] = k;
        __syncthreads();
        // ---- applied to this thread's two pixels, in that order
        for (int i = 0; i < total; i++) {
            const DrawSeg *S = segs + (size_t)3 * s_list[i];
#pragma unroll 1
            for (int l = 0; l < 3; l++) {
                const DrawSeg L = S[l];
                if (L.bx0 > L.bx1) continue;
                if (ina) draw_apply<NCH>(L, x, ya, colour, ca);
                if (inb) draw_apply<NCH>(L, x, yb, colour, cb);
            }
        }
        __syncthreads();            // s_list / s_wcount are rewritten by the next round
    }
#pragma unroll
    for (int ch = 0; ch < NCH; ch++) {
        if (ina) dst[(size_t)ya * dpitch + (size_t)x * NCH + ch] = (uint8_t)ca[ch];
        if (inb) dst[(size_t)yb * dpitch + (size_t)x * NCH + ch] = (uint8_t)cb[ch];
    }
}

cudaError_t launch_draw_flow(const uint8_t *src, int channels, int spitch, uint8_t *dst, int dpitch, int w, int h, int n, int ps, double min_vec,
                             const double *vec4, const float2 *next, const uint8_t *status, const uint8_t *keep, int gx, int gy,
                             const uint8_t *colour, void *segs, void *abox, int *drawn, cudaStream_t s)
{
    DrawParams p;
    p.w = w; p.h = h; p.n = n; p.ps = ps; p.min_vec = min_vec; p.vec4 = vec4; p.next = next; p.status = status; p.keep = keep;
    p.gx = gx; p.gy = gy; p.segs = (DrawSeg *)segs; p.abox = (int4 *)abox; p.drawn = drawn;
    cudaError_t e = cudaMemsetAsync(drawn, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    if (n > 0) k_draw_setup<<<(n + 127) / 128, 128, 0, s>>>(p);
    const dim3 grid((w + DT_W - 1) / DT_W, (h + DT_H - 1) / DT_H);
    const uchar4 c4 = make_uchar4(colour[0], channels > 1 ? colour[1] : 0, channels > 2 ? colour[2] : 0, 0);
    if (channels == 1) k_draw_tiles<1><<<grid, 256, 0, s>>>(src, spitch, dst, dpitch, w, h, p.segs, p.abox, n, c4);
    else k_draw_tiles<3><<<grid, 256, 0, s>>>(src, spitch, dst, dpitch, w, h, p.segs, p.abox, n, c4);
    MD_COUNT_LAUNCH(2);
    return cudaGetLastError();
}

size_t draw_workspace_bytes(int n) { return (size_t)n * (3 * sizeof(DrawSeg) + sizeof(int4)) + 256; }
size_t draw_segs_bytes(int n) { return ((size_t)n * 3 * sizeof(DrawSeg) + 255) / 256 * 256; }
