// k_lk.cu -- K2: pyramidal Lucas-Kanade, one warp per tracked point, all levels in one launch.
//
// Replaces cv::calcOpticalFlowPyrLK(pyr, gray2, pts1, pts2, status, err, Size(40,40), 5,
// TermCriteria(COUNT|EPS, 10, 0.03), 0, 0.001) as called at common/src/optical_flow_calculator.cpp:71,172
// (algorithm: OpenCV LKTrackerInvoker; OpenCV is not vendored by the reference).
//
// Arithmetic follows the reference's fixed-point scheme exactly (W_BITS = 14 bilinear weights, window samples
// descaled to 5 extra bits, derivative samples to 0 extra bits).  The structure-tensor sums (A11, A12, A22) and the
// mismatch sums (b1, b2) are accumulated EXACTLY -- int32 per lane (<= 64 taps per lane keeps them below 2^31),
// then a warp-shuffle butterfly in f64 -- and rounded once to f32, where OpenCV accumulates in f32 in
// SIMD-lane order; the two agree to ~1e-7 relative (flow differences ~1e-5 px, tests/test_gpu_parity.py).
// The scalar f32 expressions use explicit _rn intrinsics so that no FMA contraction changes their rounding.
//
// Two kernels: k_lk_tiled (production, window = 40) and k_lk (generic window size, simple).
#include <float.h>

#include "md_internal.h"

#define W_BITS 14

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void lk_weights(float a, float b, int &w00, int &w01, int &w10, int &w11)
{
    float na = __fsub_rn(1.f, a), nb = __fsub_rn(1.f, b);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(na, nb), (float)(1 << W_BITS)));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, nb), (float)(1 << W_BITS)));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(na, b), (float)(1 << W_BITS)));
    w11 = (1 << W_BITS) - w00 - w01 - w10;
}

// The scalar tail of one LK iteration, shared by both kernels: 2x2 solve, position update, both stopping rules.
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

// =================================================================================================================
// k_lk: generic window size; lanes stride over the taps, window samples in shared memory, source reads from L1/L2.
// =================================================================================================================
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_lk(const LkParams p)
{
    extern __shared__ int16_t lk_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * WARPS + warp;
    const int b = blockIdx.y;
    if (k >= p.P) return;
    const int win = p.win, W2 = win * win;
    int16_t *Iw = lk_smem + (size_t)warp * 3 * W2;
    int16_t *Ixw = Iw + W2, *Iyw = Ixw + W2;

    float2 pt;
    if (p.pts_in) pt = p.pts_in[(size_t)b * p.P + k];
    else pt = make_float2((float)(p.ps * (k / p.gy)), (float)(p.ps * (k % p.gy)));

    const int slotI = (p.prev_slot0 + b) % p.g.nslots, slotJ = (p.next_slot0 + b) % p.g.nslots;
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nxt = make_float2(0.f, 0.f);
    int st = 1;

    for (int level = p.g.nlev - 1; level >= 0; level--) {
        const LevelGeom L = p.g.lv[level];
        const float scale = 1.f / (float)(1 << level);
        float ppx = pt.x * scale, ppy = pt.y * scale;
        LkIterState s;
        if (level == p.g.nlev - 1) { s.npx = ppx; s.npy = ppy; }
        else { s.npx = nxt.x * 2.f; s.npy = nxt.y * 2.f; }
        nxt = make_float2(s.npx, s.npy);

        ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) {
            if (level == 0) st = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);

        const uint8_t *I0 = p.img + (size_t)slotI * p.g.slot_img_bytes + L.img_off + (size_t)(p.g.pady + ipy) * L.pitch + p.g.padx + ipx;
        const short2 *D0 = p.der + (size_t)slotI * p.g.slot_der_elems + L.der_off + (size_t)(p.g.pady + ipy) * L.pitch + p.g.padx + ipx;
        int a11 = 0, a12 = 0, a22 = 0;
        for (int t = lane; t < W2; t += 32) {
            const int y = t / win, x = t - y * win;
            const uint8_t *sp = I0 + (size_t)y * L.pitch + x;
            const short2 *d = D0 + (size_t)y * L.pitch + x;
            const int ival = (sp[0] * w00 + sp[1] * w01 + sp[L.pitch] * w10 + sp[L.pitch + 1] * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
            const short2 d00 = __ldg(d), d01 = __ldg(d + 1), d10 = __ldg(d + L.pitch), d11 = __ldg(d + L.pitch + 1);
            const int ix = (d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11 + (1 << (W_BITS - 1))) >> W_BITS;
            const int iy = (d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11 + (1 << (W_BITS - 1))) >> W_BITS;
            Iw[t] = (int16_t)ival; Ixw[t] = (int16_t)ix; Iyw[t] = (int16_t)iy;
            a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
        }
        const float A11 = (float)warp_sum((double)a11) * FLT_SCALE;
        const float A12 = (float)warp_sum((double)a12) * FLT_SCALE;
        const float A22 = (float)warp_sum((double)a22) * FLT_SCALE;
        float D;
        if (!lk_min_eig_ok(A11, A12, A22, win, p.min_eig, D)) {
            if (level == 0) st = 0;
            continue;
        }
        s.npx = __fsub_rn(s.npx, half); s.npy = __fsub_rn(s.npy, half);
        s.pdx = 0.f; s.pdy = 0.f;
        const uint8_t *Jbase = p.img + (size_t)slotJ * p.g.slot_img_bytes + L.img_off + (size_t)p.g.pady * L.pitch + p.g.padx;
        for (int j = 0; j < p.max_iters; j++) {
            const int inx = __float2int_rd(s.npx), iny = __float2int_rd(s.npy);
            if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            lk_weights(__fsub_rn(s.npx, (float)inx), __fsub_rn(s.npy, (float)iny), w00, w01, w10, w11);
            const uint8_t *J0 = Jbase + (ptrdiff_t)iny * L.pitch + inx;
            int b1 = 0, b2 = 0;
            for (int t = lane; t < W2; t += 32) {
                const int y = t / win, x = t - y * win;
                const uint8_t *sp = J0 + (size_t)y * L.pitch + x;
                const int diff = ((sp[0] * w00 + sp[1] * w01 + sp[L.pitch] * w10 + sp[L.pitch + 1] * w11 + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5)) - Iw[t];
                b1 += diff * Ixw[t];
                b2 += diff * Iyw[t];
            }
            const float fb1 = (float)warp_sum((double)b1) * FLT_SCALE;
            const float fb2 = (float)warp_sum((double)b2) * FLT_SCALE;
            if (lk_update(A11, A12, A22, D, fb1, fb2, half, j, p.eps2, s, nxt)) break;
        }
    }
    if (lane == 0) {
        p.next[(size_t)b * p.P + k] = nxt;
        p.status[(size_t)b * p.P + k] = (uint8_t)st;
    }
}

// =================================================================================================================
// k_lk_tiled: the production kernel for the reference's 40x40 window (any WIN that is a multiple of 8).
//
//  * one warp per tracked point, all pyramid levels in one launch (a point's level-l result feeds only itself);
//  * lane (lx, ly) of the 4 x 8 lane grid owns a TW x TH = 10 x 5 block of the window's taps, so the 11 x 6 source
//    samples it needs are shared between its taps, and the derivative window samples Ix, Iy stay in REGISTERS for
//    the whole iteration loop (packed s16 pairs, 50 registers);
//  * the mismatch sums are split algebraically: sum (J - I) Ix = sum J Ix - sum I Ix.  The second term is constant
//    per level (computed with the window), so the iteration loop never touches I;
//  * the bilinear samples use dp2a (two 14-bit weights x two u8 pixels per instruction) on row words kept packed
//    in registers (an aligned and a 1-byte-shifted copy), so no byte is ever extracted;
//  * the previous-frame patch (u8) + its Scharr planes (short2), and the next-frame patch with a +-3 px drift
//    margin, are staged once per level in shared memory with cp.async; the J tile is restaged only when the point
//    drifts out of the margin;
//  * structure-tensor / mismatch sums: exact int32 per lane, f64 warp-shuffle butterfly, one rounding to f32.
// Arithmetic is identical to k_lk; tap order differs only inside exact integer sums.
// =================================================================================================================
template <int WIN>
struct LkTile {
    static constexpr int LXN = 4, LYN = 8;
    static constexpr int TW = WIN / LXN, TH = WIN / LYN, NP = TW / 2;
    static constexpr int MARGIN = 3;
    static constexpr int T8P = (WIN + 2 * MARGIN + 4 + 3) / 4 + 1;   // u8 tile pitch in 32-bit words (14 for WIN = 40)
    static constexpr int T8R = WIN + 1 + 2 * MARGIN;                 // rows of the J tile (47)
    static constexpr int TDP = WIN + 4;                              // derivative tile pitch in words (44)
    static constexpr int TDR = WIN + 1;
    static constexpr int WORDS = T8P * T8R + TDP * TDR;              // per warp
    static_assert(WIN % 8 == 0 && TW % 2 == 0 && TW <= 12, "window must be a multiple of 8, at most 48");
};

__device__ __forceinline__ void cp_async4(uint32_t *smem_dst, const void *gmem_src)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// signed 16-bit pair x unsigned byte pair dot products: a.lo*b.b0 + a.hi*b.b1 + c (lo) / a.lo*b.b2 + a.hi*b.b3 + c (hi)
__device__ __forceinline__ int dp2a_lo(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int a, uint32_t b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// One staged u8 row -> aligned words a[0..2] (bytes c0 .. c0+11) and the same shifted by one byte, s[0..2]
struct RowWords { uint32_t a[3], s[3]; };
__device__ __forceinline__ RowWords load_row(const uint32_t *row, int wb, int sh)
{
    const uint32_t w0 = row[wb], w1 = row[wb + 1], w2 = row[wb + 2], w3 = row[wb + 3];
    RowWords r;
    r.a[0] = __funnelshift_r(w0, w1, sh);
    r.a[1] = __funnelshift_r(w1, w2, sh);
    r.a[2] = __funnelshift_r(w2, w3, sh);
    r.s[0] = __funnelshift_r(r.a[0], r.a[1], 8);
    r.s[1] = __funnelshift_r(r.a[1], r.a[2], 8);
    r.s[2] = r.a[2] >> 8;
    return r;
}
// sum over the pixel pair (i, i+1) of a row: wpair.lo * p[i] + wpair.hi * p[i+1] + c   (i compile-time)
template <int I>
__device__ __forceinline__ int row_pair(const RowWords &r, int wpair, int c)
{
    if constexpr ((I & 1) == 0) {
        if constexpr ((I & 2) == 0) return dp2a_lo(wpair, r.a[I >> 2], c);
        else return dp2a_hi(wpair, r.a[I >> 2], c);
    } else {
        if constexpr (((I - 1) & 2) == 0) return dp2a_lo(wpair, r.s[(I - 1) >> 2], c);
        else return dp2a_hi(wpair, r.s[(I - 1) >> 2], c);
    }
}

template <int WIN, int I, int TW>
struct TapLoop {
    // build: I sample, bilinear derivative samples, A sums, the constant sum I*Ix / I*Iy
    template <int NP>
    static __device__ __forceinline__ void build(const RowWords &r0, const RowWords &r1, const uint32_t *d0, const uint32_t *d1,
                                                 int wtop, int wbot, int w00, int w01, int w10, int w11, int (&Xpk)[NP],
                                                 int (&Ypk)[NP], int &a11, int &a12, int &a22, int &c1, int &c2, int &xprev, int &yprev)
    {
        const int iv = row_pair<I>(r1, wbot, row_pair<I>(r0, wtop, 1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
        const int e00 = (int)d0[I], e01 = (int)d0[I + 1], e10 = (int)d1[I], e11 = (int)d1[I + 1];
        const int xv = ((int)(short)e00 * w00 + (int)(short)e01 * w01 + (int)(short)e10 * w10 + (int)(short)e11 * w11 + (1 << (W_BITS - 1))) >> W_BITS;
        const int yv = ((e00 >> 16) * w00 + (e01 >> 16) * w01 + (e10 >> 16) * w10 + (e11 >> 16) * w11 + (1 << (W_BITS - 1))) >> W_BITS;
        a11 += xv * xv; a12 += xv * yv; a22 += yv * yv;
        c1 += iv * xv; c2 += iv * yv;
        if constexpr (I & 1) {
            Xpk[I >> 1] = (int)__byte_perm((uint32_t)xprev, (uint32_t)xv, 0x5410);
            Ypk[I >> 1] = (int)__byte_perm((uint32_t)yprev, (uint32_t)yv, 0x5410);
        } else { xprev = xv; yprev = yv; }
        if constexpr (I + 1 < TW)
            TapLoop<WIN, I + 1, TW>::template build<NP>(r0, r1, d0, d1, wtop, wbot, w00, w01, w10, w11, Xpk, Ypk, a11, a12, a22, c1, c2, xprev, yprev);
    }
    // iteration: q = bilinear J sample (5 extra bits); b1 += q * Ix; b2 += q * Iy
    template <int NP>
    static __device__ __forceinline__ void iter(const RowWords &r0, const RowWords &r1, int wtop, int wbot, const int (&Xpk)[NP],
                                                const int (&Ypk)[NP], int &b1, int &b2)
    {
        const int q = row_pair<I>(r1, wbot, row_pair<I>(r0, wtop, 1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
        const int xp = Xpk[I >> 1], yp = Ypk[I >> 1];
        b1 += q * ((I & 1) ? (xp >> 16) : (int)(short)xp);
        b2 += q * ((I & 1) ? (yp >> 16) : (int)(short)yp);
        if constexpr (I + 1 < TW) TapLoop<WIN, I + 1, TW>::template iter<NP>(r0, r1, wtop, wbot, Xpk, Ypk, b1, b2);
    }
};

template <int WIN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 2) k_lk_tiled(const LkParams p)
{
    using T = LkTile<WIN>;
    constexpr int TW = T::TW, TH = T::TH, NP = T::NP;
    extern __shared__ uint32_t lk_tiles[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * WARPS + warp;
    const int b = blockIdx.y;
    if (k >= p.P) return;
    uint32_t *t8 = lk_tiles + (size_t)warp * T::WORDS;     // u8 tile (I during the build, J during the iterations)
    uint32_t *td = t8 + T::T8P * T::T8R;                   // derivative tile (build only)
    const int lx = lane & 3, ly = lane >> 2;

    float2 pt;
    if (p.pts_in) pt = p.pts_in[(size_t)b * p.P + k];
    else pt = make_float2((float)(p.ps * (k / p.gy)), (float)(p.ps * (k % p.gy)));
    const int slotI = (p.prev_slot0 + b) % p.g.nslots, slotJ = (p.next_slot0 + b) % p.g.nslots;
    const float half = (WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nxt = make_float2(0.f, 0.f);
    int st = 1;

    for (int level = p.g.nlev - 1; level >= 0; level--) {
        const LevelGeom L = p.g.lv[level];
        const float scale = 1.f / (float)(1 << level);
        float ppx = pt.x * scale, ppy = pt.y * scale;
        LkIterState s;
        if (level == p.g.nlev - 1) { s.npx = ppx; s.npy = ppy; }
        else { s.npx = nxt.x * 2.f; s.npy = nxt.y * 2.f; }
        nxt = make_float2(s.npx, s.npy);
        ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -WIN || ipx >= L.w || ipy < -WIN || ipy >= L.h) {
            if (level == 0) st = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);

        // ---- stage the I patch and its derivative planes (cp.async: no register round trip) -----------------------
        __syncwarp();
        {
            const int gx0 = ipx & ~3;
            const uint8_t *Ig = p.img + (size_t)slotI * p.g.slot_img_bytes + L.img_off + (size_t)(p.g.pady + ipy) * L.pitch + (p.g.padx + gx0);
            constexpr int NW8 = (WIN + 4 + 3) / 4 + 1;     // 12 words cover ox + WIN + 1 bytes plus the word load_row over-reads
            for (int idx = lane; idx < (WIN + 1) * NW8; idx += 32) {
                const int r = idx / NW8, c = idx - r * NW8;
                cp_async4(t8 + r * T::T8P + c, Ig + (size_t)r * L.pitch + 4 * c);
            }
            const short2 *Dg = p.der + (size_t)slotI * p.g.slot_der_elems + L.der_off + (size_t)(p.g.pady + ipy) * L.pitch + (p.g.padx + ipx);
            for (int idx = lane; idx < (WIN + 1) * (WIN + 1); idx += 32) {
                const int r = idx / (WIN + 1), c = idx - r * (WIN + 1);
                cp_async4(td + r * T::TDP + c, Dg + (size_t)r * L.pitch + c);
            }
            cp_async_wait_all();
        }
        __syncwarp();

        // ---- window samples into registers, structure tensor, constant part of the mismatch ---------------------
        int Xpk[TH][NP], Ypk[TH][NP];
        int a11 = 0, a12 = 0, a22 = 0, c1 = 0, c2 = 0;
        {
            const int c0 = (ipx & 3) + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
            const int wtop = (int)__byte_perm((uint32_t)w00, (uint32_t)w01, 0x5410);
            const int wbot = (int)__byte_perm((uint32_t)w10, (uint32_t)w11, 0x5410);
            RowWords r0 = load_row(t8 + (TH * ly) * T::T8P, wb, sh);
#pragma unroll
            for (int r = 0; r < TH; r++) {
                const RowWords r1 = load_row(t8 + (TH * ly + r + 1) * T::T8P, wb, sh);
                const uint32_t *d0 = td + (TH * ly + r) * T::TDP + TW * lx, *d1 = d0 + T::TDP;
                int xprev = 0, yprev = 0;
                TapLoop<WIN, 0, TW>::template build<NP>(r0, r1, d0, d1, wtop, wbot, w00, w01, w10, w11, Xpk[r], Ypk[r], a11, a12, a22,
                                                        c1, c2, xprev, yprev);
                r0 = r1;
            }
        }
        const float A11 = (float)warp_sum((double)a11) * FLT_SCALE;
        const float A12 = (float)warp_sum((double)a12) * FLT_SCALE;
        const float A22 = (float)warp_sum((double)a22) * FLT_SCALE;
        float D;
        if (!lk_min_eig_ok(A11, A12, A22, WIN, p.min_eig, D)) {
            if (level == 0) st = 0;
            continue;
        }
        s.npx = __fsub_rn(s.npx, half); s.npy = __fsub_rn(s.npy, half);
        s.pdx = 0.f; s.pdy = 0.f;
        const uint8_t *Jbase = p.img + (size_t)slotJ * p.g.slot_img_bytes + L.img_off + (size_t)p.g.pady * L.pitch + p.g.padx;
        int tx0 = 0, ty0 = 0;
        bool staged = false;
        for (int j = 0; j < p.max_iters; j++) {
            const int inx = __float2int_rd(s.npx), iny = __float2int_rd(s.npy);
            if (inx < -WIN || inx >= L.w || iny < -WIN || iny >= L.h) {
                if (level == 0) st = 0;
                break;
            }
            lk_weights(__fsub_rn(s.npx, (float)inx), __fsub_rn(s.npy, (float)iny), w00, w01, w10, w11);
            // ---- (re)stage the J tile when the window leaves it (warp-uniform) --------------------------------------
            if (!staged || inx < tx0 || inx - tx0 > 2 * T::MARGIN + 3 || iny < ty0 || iny - ty0 > 2 * T::MARGIN) {
                tx0 = (inx - T::MARGIN) & ~3;
                ty0 = iny - T::MARGIN;
                __syncwarp();
                const uint8_t *Jg = Jbase + (ptrdiff_t)ty0 * L.pitch + tx0;
                for (int idx = lane; idx < T::T8R * T::T8P; idx += 32) {
                    const int r = idx / T::T8P, c = idx - r * T::T8P;
                    cp_async4(t8 + idx, Jg + (ptrdiff_t)r * L.pitch + 4 * c);
                }
                cp_async_wait_all();
                __syncwarp();
                staged = true;
            }
            const int c0 = (inx - tx0) + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
            const uint32_t *rowp = t8 + ((iny - ty0) + TH * ly) * T::T8P;
            const int wtop = (int)__byte_perm((uint32_t)w00, (uint32_t)w01, 0x5410);
            const int wbot = (int)__byte_perm((uint32_t)w10, (uint32_t)w11, 0x5410);
            int b1 = -c1, b2 = -c2;
            RowWords r0 = load_row(rowp, wb, sh);
#pragma unroll
            for (int r = 0; r < TH; r++) {
                const RowWords r1 = load_row(rowp + (r + 1) * T::T8P, wb, sh);
                TapLoop<WIN, 0, TW>::template iter<NP>(r0, r1, wtop, wbot, Xpk[r], Ypk[r], b1, b2);
                r0 = r1;
            }
            const float fb1 = (float)warp_sum((double)b1) * FLT_SCALE;
            const float fb2 = (float)warp_sum((double)b2) * FLT_SCALE;
            if (lk_update(A11, A12, A22, D, fb1, fb2, half, j, p.eps2, s, nxt)) break;
        }
    }
    if (lane == 0) {
        p.next[(size_t)b * p.P + k] = nxt;
        p.status[(size_t)b * p.P + k] = (uint8_t)st;
    }
}

cudaError_t launch_lk(const LkParams &p, int pairs, cudaStream_t s)
{
    constexpr int WARPS = 8;
    dim3 grid((p.P + WARPS - 1) / WARPS, pairs);
    if (p.win == 40) {
        size_t smem = (size_t)WARPS * LkTile<40>::WORDS * sizeof(uint32_t);
        cudaError_t e = cudaFuncSetAttribute(k_lk_tiled<40, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_lk_tiled<40, WARPS><<<grid, WARPS * 32, smem, s>>>(p);
    } else {
        size_t smem = (size_t)WARPS * 3 * p.win * p.win * sizeof(int16_t);
        cudaError_t e = cudaFuncSetAttribute(k_lk<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_lk<WARPS><<<grid, WARPS * 32, smem, s>>>(p);
    }
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
