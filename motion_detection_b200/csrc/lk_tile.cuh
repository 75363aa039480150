// lk_tile.cuh -- register-tile helpers shared by the 40x40-window LK kernels (k_lk_tma.cu: arbitrary points,
// k_lk_phase.cu: grid points with precomputed sub-pixel phase planes).  Lane (lx, ly) of a 4 x 8 lane grid owns a
// 10 x 5 block of the window's taps; rows of u8 pixels are kept packed in registers and sampled with dp2a.
#pragma once
#include "lk_common.cuh"

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
// the same from a 32-bit shared-window byte address (word wb of the row at `row`); volatile: ordered after the mbarrier wait
template <int OFF>
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
    return v;
}
template <int ROW_OFF>
__device__ __forceinline__ RowWords load_row_a(uint32_t addr, int sh)
{
    const uint32_t w0 = lds_u32<ROW_OFF>(addr), w1 = lds_u32<ROW_OFF + 4>(addr), w2 = lds_u32<ROW_OFF + 8>(addr), w3 = lds_u32<ROW_OFF + 12>(addr);
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

template <int I, int TW>
struct TapLoop {
    // build: I sample, bilinear derivative samples, A sums, the constant sums I*Ix / I*Iy
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
            TapLoop<I + 1, TW>::template build<NP>(r0, r1, d0, d1, wtop, wbot, w00, w01, w10, w11, Xpk, Ypk, a11, a12, a22, c1, c2, xprev, yprev);
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
        if constexpr (I + 1 < TW) TapLoop<I + 1, TW>::template iter<NP>(r0, r1, wtop, wbot, Xpk, Ypk, b1, b2);
    }
    // Same sums, two taps at a time, without unpacking the s16 pairs and without shifting the samples: q = v >> 9 with v < 2^22, so
    // byte 2 of v is q >> 7 and byte 1 of v is 2 * (q & 127) + (bit 8 of v).  One PRMT gathers (v0.b1, v1.b1, v0.b2, v1.b2), one AND
    // clears the two stray bits, and dp2a.lo / dp2a.hi against the packed (Ix_i, Ix_i+1) give lo = sum 2 (q & 127) Ix (even) and
    // hi = sum (q >> 7) Ix:  sum q * Ix = 128 * hi + lo / 2, exactly.
    template <int NP>
    static __device__ __forceinline__ void iter2(const RowWords &r0, const RowWords &r1, int wtop, int wbot, const int (&Xpk)[NP],
                                                 const int (&Ypk)[NP], int &b1lo, int &b1hi, int &b2lo, int &b2hi)
    {
        static_assert((I & 1) == 0, "pairs start at even taps");
        const int v0 = row_pair<I>(r1, wbot, row_pair<I>(r0, wtop, 1 << (W_BITS - 5 - 1)));
        const int v1 = row_pair<I + 1>(r1, wbot, row_pair<I + 1>(r0, wtop, 1 << (W_BITS - 5 - 1)));
        const uint32_t qb = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x6251) & 0xfffffefeu;
        const int xp = Xpk[I >> 1], yp = Ypk[I >> 1];
        b1lo = dp2a_lo(xp, qb, b1lo); b1hi = dp2a_hi(xp, qb, b1hi);
        b2lo = dp2a_lo(yp, qb, b2lo); b2hi = dp2a_hi(yp, qb, b2hi);
        if constexpr (I + 2 < TW) TapLoop<I + 2, TW>::template iter2<NP>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
    }
};

// rolled row loop of the window build: the finished row is moved into its (compile-time indexed) register row
template <int R, int TH, int NP>
struct RowStore {
    static __device__ __forceinline__ void put(int r, int (&Xpk)[TH][NP], int (&Ypk)[TH][NP], const int (&Xr)[NP], const int (&Yr)[NP])
    {
        if (r == R) {
#pragma unroll
            for (int i = 0; i < NP; i++) { Xpk[R][i] = Xr[i]; Ypk[R][i] = Yr[i]; }
        } else if constexpr (R + 1 < TH) RowStore<R + 1, TH, NP>::put(r, Xpk, Ypk, Xr, Yr);
    }
};

// sum q * Ix from the two accumulators of TapLoop::iter2
__device__ __forceinline__ int iter2_total(int lo, int hi) { return hi * 128 + (lo >> 1); }

// exact warp sum of per-lane int32 partial sums (|v| < 2^31): two REDUX adds on the 16-bit halves, one rounding to f32
__device__ __forceinline__ long long warp_sum_exact_i64(int v)
{
    const int hi = __reduce_add_sync(0xffffffffu, v >> 16);
    const unsigned lo = __reduce_add_sync(0xffffffffu, (unsigned)v & 0xffffu);
    return (long long)hi * 65536 + (long long)lo;
}
__device__ __forceinline__ float warp_sum_exact_f32(int v) { return (float)warp_sum_exact_i64(v); }

