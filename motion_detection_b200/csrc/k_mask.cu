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
// Layout: CTA = 256 threads, output tile 128 x 32 (+ the 2-pixel halo erode-then-dilate needs).  A thread owns runs of 4
// consecutive pixels: the thresholded bytes (0x00 / 0xFF) of a run are one 32-bit word in shared memory, so the 3x3
// erode / dilate are AND / OR of nine funnel-shifted words and the mask leaves as coalesced 32-bit stores.
//
// TMA-staged tiles (the common case).  The `cur` tile and the SOURCE BOUNDING BOX of the tile in `prev` under H^-1 are
// fetched by two cp.async.bulk.tensor boxes of tensor maps laid over the image INTERIOR: every tap outside the image is
// zero-filled by the TMA unit, which is exactly BORDER_CONSTANT 0, so the sampling loop has no bounds logic and no global
// loads.  A projective map with a denominator of one sign over the tile sends the tile to the convex hull of its corner
// images, so the box is sized from the four corners (plus slack for rounding and the bilinear footprint); tiles whose box
// does not fit (strong zoom / rotation, horizon inside the tile, misaligned caller buffers) take the per-pixel gather path.
//
// Coordinates.  The reference divides per pixel in f64 and rounds to 1/32 px; the result must be the same integer.  Fast
// path: 32 * 2^k / den from a third-order expansion around the exact reciprocal at the tile centre (relative error e^4,
// e <= 1e-3 checked per tile), numerators by FMA, and rint(v * 2^k) read off the low mantissa bits after adding 1.5 * 2^52
// -- eight DFMA and no conversion per pixel.  The k extra bits tell how close v is to a rounding boundary: only when it is
// within 2^-k of one (where the < 1e-3 unit error of the fast path could flip the rint) the pixel is recomputed with the
// reference's exact operation sequence, so the mask stays bit-identical.
#include "md_internal.h"
#include "tma.h"

#define TW 128
#define TH 32
#define GW (TW / 4 + 2)      // 34 groups of 4 columns: [tx0 - 4, tx0 + 132)
#define TR (TH + 4)          // 36 rows of thresholded words
#define ER (TH + 2)          // 34 rows of eroded words
#define SP (GW + 1)          // shared-memory pitch in words

// TMA boxes (bytes x rows); the inner extent and the start column are multiples of 16 bytes
#define CBW 160              // cur: columns [tx0 - 16, tx0 + 144)
#define CBH TR
#define PBW MD_MASK_PREV_BOX_W
#define PBH MD_MASK_PREV_BOX_H
#define NCONS 256            // consumer threads (8 warps); warp 8 is the producer
#define NSTAGE 2             // stages of the box ring
#define PREV_BYTES 8192      // staging buffer of the prev box (PBW * PBH = 7744 bytes used): a power of two, indices are wrapped
#define PREV_STAGE ((PREV_BYTES + 2 * PBW + 127) / 128 * 128)   // + the footprint of a wrapped index; a TMA destination is 128-byte aligned

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

// The reference's numerators and denominator of pixel (xj, y): row base at the origin of its 64-column block (bw0 columns in
// general) plus the per-column increment, every product and sum rounded separately (WarpPerspectiveInvoker).
__device__ __forceinline__ void ref_terms(const double *M, int bw0, int xj, int y, double &nx, double &ny, double &den)
{
    const int bx = bw0 == 64 ? (xj & ~63) : xj - xj % bw0;
    const double dbx = (double)bx, dy = (double)y, dx1 = (double)(xj - bx);
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(M[0], dbx), __dmul_rn(M[1], dy)), M[2]);
    const double Y0 = __dadd_rn(__dadd_rn(__dmul_rn(M[3], dbx), __dmul_rn(M[4], dy)), M[5]);
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(M[6], dbx), __dmul_rn(M[7], dy)), M[8]);
    den = __dadd_rn(W0, __dmul_rn(M[6], dx1));
    nx = __dadd_rn(X0, __dmul_rn(M[0], dx1));
    ny = __dadd_rn(Y0, __dmul_rn(M[3], dx1));
}
// the exact operation sequence for one pixel (rare: rounding-boundary pixels of the fast path)
__device__ __noinline__ void exact_xy(const double *M, int bw0, int xj, int y, int &X, int &Y)
{
    double nx, ny, den;
    ref_terms(M, bw0, xj, y, nx, ny, den);
    if (den == 0.0) { X = 0; Y = 0; return; }
    const double We = __ddiv_rn(32.0, den);
    X = __double2int_rn(__dmul_rn(nx, We));
    Y = __double2int_rn(__dmul_rn(ny, We));
}

// a fast-path pixel that sits on a rounding boundary (or, never expected, outside the staged box): the reference's exact
// coordinates and a bounds-checked fetch through L2
__device__ __noinline__ int slow_pixel(const uint8_t *__restrict__ prev, int pitch, int w, int h, const double *M, int bw0, int xj, int y)
{
    int X, Y;
    exact_xy(M, bw0, xj, y, X, Y);
    return bilinear_fetch(prev, pitch, w, h, X, Y);
}

// What the producer warp hands to the consumers with every staged tile
struct MaskTile {
    double M[9];                 // H^-1 of the tile's pair
    double rc, RS, de6;          // reciprocal of the denominator at the tile centre, 32 * 2^KB * rc, -M6 * rc
    int bxs, bys;                // image coordinates of the first column / row of the prev box
    int mode;                    // 0 = gather path, 1 = TMA-staged fast path, 2 = pair without egomotion (empty mask)
    int tx0, ty0, b;             // tile origin, pair
    int zp;                      // frame index of `prev` (ring slot or pair)
    int pad;
};

// Fixed point of the fast path: v = 32 * coordinate * 2^KB is rounded to an integer n by the FP adder itself (sum with
// 1.5 * 2^52: the 52-bit fraction field of the result is n + 2^51).  Folding 2^(KB-1) + 1 into that constant makes
//   bits [0, KB) of the low word  <= 2   <=>  v within one unit of a rounding boundary of rint(v / 2^KB)  (-> exact path),
//   bits [KB, KB + 5)                  =  the 5-bit fraction ax of the reference's 1/32-pixel grid,
//   bits [KB + 5, 52)                  =  the integer sample column, offset by 2^31 (one funnel shift across hi:lo).
#define KB 15
#define MASK_MAGIC (6755399441055744.0 + (double)(1 << (KB - 1)) + 1.0)

__device__ __forceinline__ int lds_u8(uint32_t addr)
{
    int v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

template <bool INTERIOR>
__device__ __forceinline__ void mask_phase1_fast(const MaskParams &p, const MaskTile &sT, uint32_t aPrev, uint32_t aCur,
                                                 uint32_t (*T)[SP], const uint8_t *prev, int tx0, int ty0, int bw0, int tid)
{
    const double *sM = sT.M;
    const int w = p.w, h = p.h;
    const uint32_t thr4 = (uint32_t)min(max(p.thresh, 0), 255) * 0x01010101u;
    const double M0 = sM[0], M1 = sM[1], M2 = sM[2], M3 = sM[3], M4 = sM[4], M5 = sM[5], M6 = sM[6], M7 = sM[7], M8 = sM[8];
    const double rc = sT.rc, RS = sT.RS, de6 = sT.de6;
    const uint32_t bxo = (uint32_t)sT.bxs ^ 0x80000000u, byo = (uint32_t)sT.bys ^ 0x80000000u;
    for (int g = tid; g < GW * TR; g += NCONS) {
        const int ry = g / GW, gx = g - ry * GW;
        const int x = tx0 - 4 + 4 * gx, y = ty0 - 2 + ry;
        uint32_t word = 0xffffffffu;            // erode identity outside the image
        if (INTERIOR || (y >= 0 && y < h && x + 3 >= 0 && x < w)) {
            const uint32_t c4 = lds_u32(aCur + ry * CBW + 12 + 4 * gx);
            const double xd = (double)x, yd = (double)y;
            const double Xr = fma(M0, xd, fma(M1, yd, M2)), Yr = fma(M3, xd, fma(M4, yd, M5));
            const double e0 = fma(-fma(M6, xd, fma(M7, yd, M8)), rc, 1.0);
            uint32_t gmin = 0xffffffffu;        // smallest distance (in 2^-KB units, offset by one) of a coordinate to a rounding boundary
            bool inbox = true;
            int wv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double nx = j ? fma(M0, (double)j, Xr) : Xr, ny = j ? fma(M3, (double)j, Yr) : Yr;
                const double e = j ? fma(de6, (double)j, e0) : e0;
                const double e3 = fma(fma(e, e, e), e, e);             // e + e^2 + e^3
                const double rs = fma(RS, e3, RS);
                const double vX = fma(nx, rs, MASK_MAGIC), vY = fma(ny, rs, MASK_MAGIC);
                const uint32_t loX = (uint32_t)__double2loint(vX), loY = (uint32_t)__double2loint(vY);
                const uint32_t lx = __funnelshift_r(loX, (uint32_t)__double2hiint(vX), KB + 5) - bxo;
                const uint32_t ly = __funnelshift_r(loY, (uint32_t)__double2hiint(vY), KB + 5) - byo;
                const int ax = (loX >> KB) & 31, ay = (loY >> KB) & 31;
                gmin = min(gmin, min(loX & ((1u << KB) - 1), loY & ((1u << KB) - 1)));
                // the ends of the run prove that every footprint is inside the staged box (the run maps to a segment: its inner
                // pixels lie between the ends)
                if (j == 0 || j == 3) inbox = inbox && lx < (uint32_t)(PBW - 1) && ly < (uint32_t)(PBH - 1);
                // the fetch itself is unconditional: the index is wrapped into the (8 KB) staging buffer, so a run that is redone
                // below reads harmless bytes
                const uint32_t r = aPrev + ((ly * PBW + lx) & (PREV_BYTES - 1));
                const int p00 = lds_u8(r), p01 = lds_u8(r + 1), p10 = lds_u8(r + PBW), p11 = lds_u8(r + PBW + 1);
                const int h0 = (p00 << 5) + ax * (p01 - p00), h1 = (p10 << 5) + ax * (p11 - p10);
                wv[j] = ((h0 << 5) + ay * (h1 - h0) + 512) >> 10;
            }
            if (gmin <= 2u || !inbox) {
                // rare: a coordinate of the run sits on a rounding boundary (or, never expected, a footprint leaves the box):
                // redo the run with the reference's exact operation sequence and bounds-checked fetches through L2
                for (int j = 0; j < 4; j++)
                    if (INTERIOR || (x + j >= 0 && x + j < w)) wv[j] = slow_pixel(prev, p.pitch, w, h, sM, bw0, x + j, y);
            }
            // |warp - cur| > thresh on four pixels at once
            const uint32_t w4 = __byte_perm(__byte_perm((uint32_t)wv[0], (uint32_t)wv[1], 0x0040), __byte_perm((uint32_t)wv[2], (uint32_t)wv[3], 0x0040), 0x5410);
            word = p.thresh < 0 ? 0xffffffffu : __vcmpgtu4(__vabsdiffu4(w4, c4), thr4);      // d >= 0 > a negative threshold
            if (!INTERIOR) {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j < 0 || x + j >= w) word |= 0xffu << (8 * j);
            }
        }
        T[ry][gx] = word;
    }
}

struct alignas(64) MaskKernelMaps { CUtensorMap prev, cur; };


__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(NCONS) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// The producer's per-tile work: where the tile's source box lies and whether the tile can run from TMA-staged boxes.
__device__ __forceinline__ void mask_tile_setup(const MaskParams &p, MaskTile &t, int tile, int ntx, int nty)
{
    const int w = p.w, h = p.h;
    const int b = tile / (ntx * nty), r = tile - b * ntx * nty, tyi = r / ntx, txi = r - tyi * ntx;
    const int tx0 = txi * TW, ty0 = tyi * TH;
    t.tx0 = tx0; t.ty0 = ty0; t.b = b;
    t.zp = p.nslots ? (p.prev_slot0 + b) % p.nslots : b;
    t.rc = 0; t.RS = 0; t.de6 = 0; t.bxs = 0; t.bys = 0; t.pad = 0;
    if (p.valid && p.valid[b] == 0) { t.mode = 2; return; }
    double M[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { M[i] = p.Hinv[b * 9 + i]; t.M[i] = M[i]; }
    int fast = p.use_tma;
    const int x0 = max(tx0 - 4, 0), x1 = min(tx0 + TW + 3, w - 1), y0 = max(ty0 - 2, 0), y1 = min(ty0 + TH + 1, h - 1);
    double lo_x = 1e300, hi_x = -1e300, lo_y = 1e300, hi_y = -1e300, dmin = 1e300, dmax = -1e300;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const double cx = (double)((c & 1) ? x1 : x0), cy = (double)((c & 2) ? y1 : y0);
        const double den = fma(M[6], cx, fma(M[7], cy, M[8]));
        const double rr = 1.0 / den;
        const double sx = fma(M[0], cx, fma(M[1], cy, M[2])) * rr, sy = fma(M[3], cx, fma(M[4], cy, M[5])) * rr;
        lo_x = fmin(lo_x, sx); hi_x = fmax(hi_x, sx); lo_y = fmin(lo_y, sy); hi_y = fmax(hi_y, sy);
        dmin = fmin(dmin, den); dmax = fmax(dmax, den);
    }
    // one sign, away from zero, finite; |32 * coordinate * 2^KB| stays far below the 2^51 the fixed point holds
    if (!(dmin * dmax > 0.0) || !(fmin(fabs(dmin), fabs(dmax)) > 1e-9) || !(lo_x > -1e6 && hi_x < 1e6 && lo_y > -1e6 && hi_y < 1e6)) fast = 0;
    if (fast) {
        const int bx0 = (int)floor(lo_x) - 1, bx1 = (int)floor(hi_x) + 2, by0 = (int)floor(lo_y) - 1, by1 = (int)floor(hi_y) + 2;
        t.bxs = bx0 & ~15;                    // 16-byte aligned box start (two's complement: rounds towards -inf)
        t.bys = by0;
        if (bx1 - t.bxs + 1 > PBW || by1 - t.bys + 1 > PBH) fast = 0;
        // a box that misses the image altogether is left to the gather path
        if (t.bxs >= w || t.bxs + PBW <= 0 || t.bys >= h || t.bys + PBH <= 0) fast = 0;
        const double xc = 0.5 * (x0 + x1), yc = 0.5 * (y0 + y1);
        const double denc = fma(M[6], xc, fma(M[7], yc, M[8]));
        t.rc = 1.0 / denc;
        t.RS = t.rc * (double)(32 << KB);
        t.de6 = -M[6] * t.rc;
        // |1 - den * rc| over the tile: the third-order expansion is good to e^4 <= 1e-12
        if (!((fabs(t.de6) * (0.5 * (x1 - x0) + 1.0) + fabs(M[7] * t.rc) * (0.5 * (y1 - y0) + 1.0)) < 1e-3)) fast = 0;
    }
    t.mode = fast ? 1 : 0;
}

// Persistent, warp-specialised: warp 8 (one lane) walks the CTA's tiles ahead of the others -- source box, expansion
// constants, two TMA box loads per tile into a two-stage ring guarded by full / empty mbarriers -- while warps 0-7 sample,
// threshold, erode, dilate and store the tile that has landed.
template <bool ALIGNED>
__global__ void __launch_bounds__(NCONS + 32, 3) k_mask(const MaskParams p, const __grid_constant__ MaskKernelMaps maps, int ntx, int nty,
                                                        int ntiles)
{
    __shared__ __align__(128) uint8_t sPrev[NSTAGE][PREV_STAGE];
    __shared__ __align__(128) uint8_t sCur[NSTAGE][CBW * CBH];
    static_assert((CBW * CBH) % 128 == 0 && PREV_STAGE % 128 == 0, "TMA destinations are 128-byte aligned");
    __shared__ uint32_t T[TR][SP];
    __shared__ uint32_t E[ER][SP];
    __shared__ MaskTile sT[NSTAGE];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];
    const int w = p.w, h = p.h;
    const int bh0 = h < 16 ? h : 16;
    const int bw0 = (1024 / bh0 < w) ? 1024 / bh0 : w;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS / 32); }
        mbar_fence_init();
    }
    __syncthreads();

    if (threadIdx.x >= NCONS) {
        // ---- producer -------------------------------------------------------------------------------------------------------
        if (threadIdx.x == NCONS) {
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
                const int s = it % NSTAGE;
                if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
                MaskTile &t = sT[s];
                mask_tile_setup(p, t, tile, ntx, nty);
                if (t.mode == 1) {
                    const int zc = p.nslots ? (p.cur_slot0 + t.b) % p.nslots : t.b;
                    mbar_expect_tx(&full[s], PBW * PBH + CBW * CBH);
                    tma_load_3d(sPrev[s], &maps.prev, t.bxs, t.bys, t.zp, &full[s]);
                    tma_load_3d(sCur[s], &maps.cur, t.tx0 - 16, t.ty0 - 2, zc, &full[s]);
                } else mbar_arrive(&full[s]);
            }
        }
        return;
    }

    // ---- consumers ----------------------------------------------------------------------------------------------------------
    const int tid = threadIdx.x;
    int local = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int s = it % NSTAGE;
        mbar_wait(&full[s], (it / NSTAGE) & 1);
        const MaskTile &t = sT[s];
        const int tx0 = t.tx0, ty0 = t.ty0, b = t.b, mode = t.mode;
        uint8_t *out = p.mask + (size_t)b * p.mask_stride;
        if (mode == 2) {
            // no egomotion (fewer than the minimal number of vectors): empty mask
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&empty[s]);
            for (int i = tid; i < TW * TH; i += NCONS) {
                int x = tx0 + i % TW, y = ty0 + i / TW;
                if (x < w && y < h) out[(size_t)y * p.mask_pitch + x] = 0;
            }
            continue;
        }
        const uint8_t *prev = p.prev + (long long)t.zp * p.stride;
        const bool interior = tx0 - 4 >= 0 && tx0 + TW + 4 <= w && ty0 - 2 >= 0 && ty0 + TH + 2 <= h;
        if (mode == 1) {
            // ---- phase 1 (fast): warp + absdiff + threshold from the staged boxes, 4 pixels per thread-iteration -----------
            if (interior) mask_phase1_fast<true>(p, t, smem_u32(sPrev[s]), smem_u32(sCur[s]), T, prev, tx0, ty0, bw0, tid);
            else mask_phase1_fast<false>(p, t, smem_u32(sPrev[s]), smem_u32(sCur[s]), T, prev, tx0, ty0, bw0, tid);
        } else {
            // ---- phase 1 (gather): per-pixel loads through L2, reference operation sequence with the Newton shortcut ---------
            const double *sM = t.M;
            const uint8_t *cur = p.cur + (long long)(p.nslots ? (p.cur_slot0 + b) % p.nslots : b) * p.stride;
            const bool uniform_block = (bw0 & 3) == 0;     // a run of 4 aligned columns never straddles a 64-column block
            for (int g = tid; g < GW * TR; g += NCONS) {
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
                        X0 = __dadd_rn(__dadd_rn(__dmul_rn(sM[0], dbx), __dmul_rn(sM[1], dy)), sM[2]);
                        Y0 = __dadd_rn(__dadd_rn(__dmul_rn(sM[3], dbx), __dmul_rn(sM[4], dy)), sM[5]);
                        W0 = __dadd_rn(__dadd_rn(__dmul_rn(sM[6], dbx), __dmul_rn(sM[7], dy)), sM[8]);
                    }
                    word = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int xj = x + j;
                        uint32_t tt = 0xffu;
                        if (xj >= 0 && xj < w) {
                            double nx, ny, den;
                            if (uniform_block) {
                                const double dx1 = (double)(xj - bx);
                                den = __dadd_rn(W0, __dmul_rn(sM[6], dx1));
                                nx = __dadd_rn(X0, __dmul_rn(sM[0], dx1)); ny = __dadd_rn(Y0, __dmul_rn(sM[3], dx1));
                            } else ref_terms(sM, bw0, xj, y, nx, ny, den);
                            int X, Y;
                            project(nx, ny, den, X, Y);
                            const int wv = bilinear_fetch(prev, p.pitch, w, h, X, Y);
                            int d = wv - (int)((c4 >> (8 * j)) & 0xffu);
                            d = d < 0 ? -d : d;
                            tt = d > p.thresh ? 0xffu : 0u;
                        }
                        word |= tt << (8 * j);
                    }
                }
                T[ry][gx] = word;
            }
        }
        consumer_sync();
        // the staged boxes and the tile record are consumed: hand the stage back to the producer
        // (the tile's scalars were copied to registers above; phases 2 and 3 only touch T / E)
        if ((tid & 31) == 0) mbar_arrive(&empty[s]);

        if (p.morph) {
            // ---- phase 2: erode = AND of the nine neighbours, four pixels per word ----------------------------------
            for (int g = tid; g < GW * ER; g += NCONS) {
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
                uint32_t m = interior ? 0xffffffffu : 0u;
                if (!interior && y >= 0 && y < h) {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (x + j >= 0 && x + j < w) m |= 0xffu << (8 * j);
                }
                E[ry][gx] = e & m;
            }
            consumer_sync();
            // ---- phase 3: dilate = OR of the nine neighbours, coalesced 32-bit stores -------------------------------
            for (int g = tid; g < (TW / 4) * TH; g += NCONS) {
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
            for (int g = tid; g < (TW / 4) * TH; g += NCONS) {
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
        // T is rewritten by the next tile's phase 1: every consumer must be done reading it (phase 2 / the raw copy)
        consumer_sync();
    }
    if (p.stat_mask) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if ((tid & 31) == 0 && local) atomicAdd(p.stat_mask, (unsigned long long)local);
    }
}

// Tensor maps over the image interior (x, y, frame): out-of-image box elements are zero-filled = BORDER_CONSTANT 0.
bool mask_encode_maps(MaskTmaMaps *m, const uint8_t *prev, const uint8_t *cur, int w, int h, int pitch, long long stride, int nframes)
{
    m->valid = 0;
    if (((uintptr_t)prev | (uintptr_t)cur | (uintptr_t)pitch | (uintptr_t)stride) & 15) return false;
    if (nframes < 1) nframes = 1;
    const uint64_t fs = stride > 0 ? (uint64_t)stride : (uint64_t)pitch * h;
    if (fs & 15) return false;
    if (!tma_encode_3d(&m->prev, 1, (void *)prev, w, h, nframes, pitch, fs, PBW, PBH)) return false;
    if (!tma_encode_3d(&m->cur, 1, (void *)cur, w, h, nframes, pitch, fs, CBW, CBH)) return false;
    m->valid = 1;
    return true;
}

cudaError_t launch_mask(const MaskParams &p0, int pairs, const MaskTmaMaps *maps, cudaStream_t s)
{
    MaskParams p = p0;
    const bool aligned = (((uintptr_t)p.cur | (uintptr_t)p.mask | (uintptr_t)p.pitch | (uintptr_t)p.mask_pitch |
                           (uintptr_t)p.stride | (uintptr_t)p.mask_stride) & 3) == 0;
    p.use_tma = maps && maps->valid && aligned ? 1 : 0;
    MaskKernelMaps km;
    if (p.use_tma) { km.prev = maps->prev; km.cur = maps->cur; }
    else memset(&km, 0, sizeof km);
    const int ntx = (p.w + TW - 1) / TW, nty = (p.h + TH - 1) / TH;
    const long long nt = (long long)ntx * nty * pairs;
    if (nt > 0x7fffffff) return cudaErrorInvalidValue;
    // persistent CTAs: 3 per SM (register bound), every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
    const int grid = (int)(nt < (long long)sms * 3 ? nt : (long long)sms * 3);
    if (aligned) k_mask<true><<<grid, NCONS + 32, 0, s>>>(p, km, ntx, nty, (int)nt);
    else k_mask<false><<<grid, NCONS + 32, 0, s>>>(p, km, ntx, nty, (int)nt);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
