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
// Layout: persistent CTAs of 8 consumer warps + 1 producer warp; output tile 120 x 36, computed region 128 x 40 (the 2-pixel
// halo erode-then-dilate needs, rounded to whole 4-pixel runs).  A WARP owns a row of the computed region, a lane a run of 4
// consecutive pixels; the thresholded pixels leave the sampling loop as predicates and __ballot_sync turns a row into four
// 32-bit words (word j, bit l = pixel 4 l + j), so the 3x3 erode / dilate of a whole 128-pixel row are a few dozen AND / OR /
// shift instructions on ONE thread (bit-parallel morphology: 1/30 of the byte-SIMD version's instructions), and the mask
// leaves as coalesced 32-bit stores expanded from the bit planes.
//
// TMA-staged tiles (the common case).  The `cur` tile and the SOURCE BOUNDING BOX of the tile in `prev` under H^-1 are
// fetched by two cp.async.bulk.tensor boxes of tensor maps laid over the image INTERIOR: every tap outside the image is
// zero-filled by the TMA unit, which is exactly BORDER_CONSTANT 0, so the sampling loop has no bounds logic and no global
// loads.  A projective map with a denominator of one sign over the tile sends the tile to the convex hull of its corner
// images, so the box is sized from the four corners (plus slack for rounding and the bilinear footprint); tiles whose box
// does not fit (strong zoom / rotation, horizon inside the tile, misaligned caller buffers) take the per-pixel gather path.
//
// Coordinates.  The reference divides per pixel in f64 and rounds to 1/32 px; the result must be the same integer.  Fast
// path: 32 * 2^KB / den from a third-order expansion around the exact reciprocal at the tile centre (relative error e^4,
// e <= 1e-3 checked per tile), numerators by FMA from per-row terms the producer warp tabulates, and rint(v * 2^KB) of the
// BOX-RELATIVE coordinate read off the low mantissa word after adding 1.5 * 2^52 (-32 * 2^KB * box origin folded in): one
// shift gives the sample column, one shift + mask the 5-bit fraction, the low 16 bits tell how close v is to a rounding
// boundary.  Only when it is within 2^-(KB-2) of one (where the < 1e-3 unit error of the fast path could flip the rint) the
// run is recomputed with the reference's exact operation sequence, so the mask stays bit-identical.
#include "md_internal.h"
#include "tma.h"

#define TW 120               // output tile
#define TH 36
#define CW (TW + 8)          // computed region: columns [tx0 - 4, tx0 + 124) = 32 runs of 4 pixels = one warp per row
#define TR (TH + 4)          // rows [ty0 - 2, ty0 + TH + 2) = 5 rows per consumer warp

// TMA boxes (bytes x rows); the inner extent and the start column are multiples of 16 bytes
#define CBW 144              // cur: columns [(tx0 - 4) & ~15, + 144)
#define CBH TR
#define PBW MD_MASK_PREV_BOX_W
#define PBH MD_MASK_PREV_BOX_H
#define NCONS 256            // consumer threads (8 warps); warp 8 is the producer
#define NSTAGE 2             // stages of the box ring
#ifndef MASK_CTAS_PER_SM
#define MASK_CTAS_PER_SM 3     // persistent CTAs per SM (register bound: 72 registers x 288 threads)
#endif
#define PREV_STAGE (PBW * PBH)
#define PREV_AMAX (PBW * PBH - PBW - 2)      // largest index whose 2 x 2 footprint stays inside the staged box
static_assert(PBW == 256, "the sample address is (row byte, column byte) of the fixed-point coordinates: the box pitch is 256");

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
    double RS, de6;              // 32 * 2^KB * rc (rc = reciprocal of the denominator at the tile centre), -M6 * rc
    double magx, magy;           // MASK_MAGIC - 32 * 2^KB * (box origin): the fixed point holds BOX-RELATIVE coordinates
    double rowX[TR], rowY[TR], rowE[TR];   // per row y: M1 y + M2, M4 y + M5, 1 - (M7 y + M8) rc
    int mode;                    // 0 = gather path, 1 = TMA-staged fast path (3 = its first-order variant), 2 = pair without egomotion (empty mask)
    int tx0, ty0, b;             // tile origin, pair
    int zp;                      // frame index of `prev` (ring slot or pair)
    int cxo;                     // byte offset of column tx0 - 4 inside the cur box
    uint32_t colin[4];           // plane j, bit l: column tx0 - 4 + 4 l + j lies inside the image
    int pad[2];
};

// Fixed point of the fast path: v = 32 * (coordinate - box origin) * 2^KB is rounded to an integer n by the FP adder itself
// (sum with 1.5 * 2^52: the 52-bit fraction field of the result is n + 2^51).  Folding 2^(KB-1) + 2 into that constant makes
//   bits [0, KB) of the low word  <= 4   <=>  v within two units of a rounding boundary of rint(v / 2^KB)  (-> exact path;
//                                              only the low 16 of the 19 bits are looked at: a superset, 8e-5 of the coordinates),
//   bits [KB, KB + 5)                  =  the 5-bit fraction ax of the reference's 1/32-pixel grid,
//   bits [KB + 5, 32) = byte 3         =  the sample column / row inside the box (box pitch 256: PRMT of the two bytes = the address).
// (the fast path's own error is < 1e-2 unit: series truncation e^4 <= 1e-12 relative of v < 2^32, FMA roundings 2^-53 relative of
// the absolute coordinate)
#define KB 19
#define MASK_MAGIC (6755399441055744.0 + (double)(1 << (KB - 1)) + 2.0)
#define MASK_GUARD 4u

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
__device__ __forceinline__ double lds_f64(uint32_t addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
// A fast-path pixel that sits on a rounding boundary: the reference's exact coordinates; the four taps from the staged box when
// the footprint lies inside it (always, short of a pixel outside the image), else bounds-checked through L2.
__device__ __noinline__ int redo_pixel(const uint8_t *__restrict__ prev, int pitch, int w, int h, const MaskTile &t, int bw0, int xj, int y,
                                       uint32_t aPrev)
{
    int X, Y;
    exact_xy(t.M, bw0, xj, y, X, Y);
    const int lx = (X >> 5) - t.pad[0], ly = (Y >> 5) - t.pad[1];
    if ((unsigned)lx < (unsigned)(PBW - 1) && (unsigned)ly < (unsigned)(PBH - 1)) {
        const uint32_t r = aPrev + ly * PBW + lx;
        const int ax = X & 31, ay = Y & 31;
        const int p00 = lds_u8(r), p01 = lds_u8(r + 1), p10 = lds_u8(r + PBW), p11 = lds_u8(r + PBW + 1);
        const int h0 = (p00 << 5) + ax * (p01 - p00), h1 = (p10 << 5) + ax * (p11 - p10);
        return ((h0 << 5) + ay * (h1 - h0) + 512) >> 10;
    }
    return bilinear_fetch(prev, pitch, w, h, X, Y);
}

// prmt.b32 with the full selector semantics (bit 3 of a selector nibble replicates the sign bit of the chosen byte; __byte_perm
// only looks at the low three bits)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// per-halfword unsigned minimum of three packed pairs (DPX: VIMNMX3.U16x2)
__device__ __forceinline__ uint32_t min3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }

// The thresholded pixel as a predicate without unpacking the warped value: with v = the bilinear sum before its `+ 512 >> 10`,
// |((v + 512) >> 10) - c| <= thr  <=>  0 <= v + 512 - (c - thr) * 1024 < (2 thr + 1) * 1024, so "motion" is ONE unsigned compare of
// v + kc against lim, kc = 512 - (c - thr) * 1024 folded into the interpolation.  thr < 0 (everything is motion): lim = 0.
struct ThreshConst { int kt; uint32_t lim; };      // kc = kt - c * 1024
__device__ __forceinline__ ThreshConst thresh_const(int thresh)
{
    ThreshConst t;
    const int thr = min(max(thresh, 0), 255);
    t.kt = 512 + thr * 1024;
    t.lim = thresh < 0 ? 0u : (uint32_t)(2 * thr + 1) * 1024u;
    return t;
}

// one row of three-neighbour AND (erode) / OR (dilate) along x on the interleaved planes: pixel 4 l + j has its left
// neighbour in plane j - 1 (plane 3 of lane l - 1 for j = 0) and its right neighbour in plane j + 1 (plane 0 of lane l + 1)
template <bool AND>
__device__ __forceinline__ uint4 morph_row_x(const uint4 v)
{
    uint4 r;
    if (AND) {
        r.x = v.x & (v.w << 1 | 1u) & v.y; r.y = v.y & v.x & v.z; r.z = v.z & v.y & v.w; r.w = v.w & v.z & (v.x >> 1 | 0x80000000u);
    } else {
        r.x = v.x | (v.w << 1) | v.y; r.y = v.y | v.x | v.z; r.z = v.z | v.y | v.w; r.w = v.w | v.z | (v.x >> 1);
    }
    return r;
}

// ---- phase 1 (fast): one row of the computed region per warp-iteration, a run of 4 pixels per lane -----------------------------
// ORDER = 3: rs = RS (1 + e + e^2 + e^3) per pixel.  ORDER = 1 (|e| < 2e-6 over the tile, i.e. a nearly affine map -- the usual
// frame-to-frame egomotion): rs = RS (1 + e), which is linear along the run: one FMA per pixel (e^2 < 4e-12 relative).
template <int ORDER>
__device__ __forceinline__ void mask_phase1_fast(const MaskParams &p, const MaskTile &sT, uint32_t aT, uint32_t aPrev, uint32_t aCur,
                                                 uint4 *Wp, const uint8_t *prev, int bw0, int warp, int lane)
{
    const int w = p.w, h = p.h;
    const ThreshConst tc = thresh_const(p.thresh);
    const int tx0 = sT.tx0, ty0 = sT.ty0;
    const int x = tx0 - 4 + 4 * lane;
    const double xd = (double)x;
    const double M0 = sT.M[0], M3 = sT.M[3], de6 = sT.de6, RS = sT.RS, magx = sT.magx, magy = sT.magy;
    const double drs = RS * de6;                     // ORDER 1: d rs / d x
    uint32_t aRow = aT + (uint32_t)offsetof(MaskTile, rowX) + 8 * warp;
    uint32_t aC = aCur + sT.cxo + 4 * lane + warp * CBW;
    const uint32_t ci0 = sT.colin[0], ci1 = sT.colin[1], ci2 = sT.colin[2], ci3 = sT.colin[3];
    int y = ty0 - 2 + warp;
#pragma unroll 1
    for (int ry = warp; ry < TR; ry += NCONS / 32, y += NCONS / 32, aRow += 8 * (NCONS / 32), aC += (NCONS / 32) * CBW) {
        const uint32_t c4 = lds_u32(aC);
        const double Xr = fma(M0, xd, lds_f64(aRow)), Yr = fma(M3, xd, lds_f64(aRow + 8 * TR));
        const double e0 = fma(de6, xd, lds_f64(aRow + 16 * TR));
        const double rs0 = fma(RS, e0, RS);
        uint32_t dist[4], a16[4];
        int vp[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const double nx = j ? fma(M0, (double)j, Xr) : Xr, ny = j ? fma(M3, (double)j, Yr) : Yr;
            double rs;
            if (ORDER == 1) rs = j ? fma(drs, (double)j, rs0) : rs0;
            else {
                const double e = j ? fma(de6, (double)j, e0) : e0;
                rs = fma(RS, fma(fma(e, e, e), e, e), RS);            // RS (1 + e + e^2 + e^3)
            }
            const double vX = fma(nx, rs, magx), vY = fma(ny, rs, magy);
            const uint32_t loX = (uint32_t)__double2loint(vX), loY = (uint32_t)__double2loint(vY);
            const int ax = (loX >> KB) & 31, ay = (loY >> KB) & 31;
            dist[j] = __byte_perm(loX, loY, 0x5410);               // the two 16-bit boundary distances side by side
            // byte 3 of the low words = column / row inside the box.  Every footprint of the tile lies inside the staged box: the
            // producer sized it from the images of the region's corners (a projective map with a denominator of one sign maps the
            // rectangle into the convex hull of its corner images) and sends tiles that do not fit down the gather path.  The
            // index is clamped all the same, so the loads stay inside the staging buffer whatever the coordinates are.
            a16[j] = __byte_perm(loX, loY, 0x7373);
            const uint32_t r = aPrev + min3_u16x2(a16[j], (uint32_t)PREV_AMAX, (uint32_t)PREV_AMAX);   // low half clamped, high half -> 0
            const uint32_t p00 = (uint32_t)lds_u8(r), p01 = (uint32_t)lds_u8(r + 1), p10 = (uint32_t)lds_u8(r + PBW), p11 = (uint32_t)lds_u8(r + PBW + 1);
            // both rows at once: (h0, h1) = (32 - ax) (p00, p10) + ax (p01, p11) as 16-bit halves, then (32 - ay) h0 + ay h1 by dp2a
            const uint32_t hh = (uint32_t)(32 - ax) * __byte_perm(p00, p10, 0x5410) + (uint32_t)ax * __byte_perm(p01, p11, 0x5410);
            const int kc = tc.kt - (int)__byte_perm(c4, 0, 0x4440 + j) * 1024;
            int v;
            asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(v) : "r"(hh), "r"(32 + 255 * ay), "r"(kc));
            vp[j] = v;
        }
        uint32_t gmin = min3_u16x2(min3_u16x2(dist[0], dist[1], dist[2]), dist[3], 0xffffffffu);
        gmin = min(gmin & 0xffffu, gmin >> 16);     // smallest distance (in 2^-KB units, offset by two) of a coordinate to a rounding boundary
        if (gmin <= MASK_GUARD) {
            // rare (1.5e-4 of the pixels): a coordinate sits within the guard band of a rounding boundary -> that pixel is redone
            // with the reference's exact operation sequence (its taps come from the staged box when they lie inside it)
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (min(dist[j] & 0xffffu, dist[j] >> 16) <= MASK_GUARD)
                    vp[j] = (redo_pixel(prev, p.pitch, w, h, sT, bw0, x + j, y, aPrev) - (int)((c4 >> (8 * j)) & 0xffu)) * 1024 + tc.kt;
        }
        // the row as four bit planes; pixels outside the image are the erode identity (1)
        uint32_t w0 = __ballot_sync(0xffffffffu, (uint32_t)vp[0] >= tc.lim), w1 = __ballot_sync(0xffffffffu, (uint32_t)vp[1] >= tc.lim);
        uint32_t w2 = __ballot_sync(0xffffffffu, (uint32_t)vp[2] >= tc.lim), w3 = __ballot_sync(0xffffffffu, (uint32_t)vp[3] >= tc.lim);
        if (y < 0 || y >= h) w0 = w1 = w2 = w3 = 0xffffffffu;
        if (lane == 0) {
            // the erode's pass along x happens here, on the row's four words (phase 2 only combines rows)
            const uint4 row = make_uint4(w0 | ~ci0, w1 | ~ci1, w2 | ~ci2, w3 | ~ci3);
            Wp[ry] = p.morph ? morph_row_x<true>(row) : row;
        }
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
// Executed by the whole producer warp (the row tables are filled lane-parallel); returns the same record to every lane.
__device__ __forceinline__ void mask_tile_setup(const MaskParams &p, MaskTile &t, int tile, int ntx, int nty, int lane)
{
    const int w = p.w;
    const int b = tile / (ntx * nty), r = tile - b * ntx * nty, tyi = r / ntx, txi = r - tyi * ntx;
    const int tx0 = txi * TW, ty0 = tyi * TH;
    double M[9];
    int mode = 0;
    const bool novalid = p.valid && p.valid[b] == 0;
    if (!novalid) {
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = p.Hinv[b * 9 + i];
    } else {
#pragma unroll
        for (int i = 0; i < 9; i++) M[i] = 0.0;
    }
    int fast = p.use_tma && !novalid;
    // corners of the computed region (not clamped to the image: pixels outside it are computed and discarded)
    const int x0 = tx0 - 4, x1 = tx0 + CW - 5, y0 = ty0 - 2, y1 = ty0 + TH + 1;
    // the four corners on lanes 0-3 (one division each instead of four in a row), min / max by shuffles
    double lo_x, hi_x, lo_y, hi_y, dmin, dmax;
    {
        const int c = lane & 3;
        const double cx = (double)((c & 1) ? x1 : x0), cy = (double)((c & 2) ? y1 : y0);
        const double den = fma(M[6], cx, fma(M[7], cy, M[8]));
        const double rr = 1.0 / den;
        const double sx = fma(M[0], cx, fma(M[1], cy, M[2])) * rr, sy = fma(M[3], cx, fma(M[4], cy, M[5])) * rr;
        lo_x = hi_x = sx; lo_y = hi_y = sy; dmin = dmax = den;
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            lo_x = fmin(lo_x, __shfl_xor_sync(0xffffffffu, lo_x, o)); hi_x = fmax(hi_x, __shfl_xor_sync(0xffffffffu, hi_x, o));
            lo_y = fmin(lo_y, __shfl_xor_sync(0xffffffffu, lo_y, o)); hi_y = fmax(hi_y, __shfl_xor_sync(0xffffffffu, hi_y, o));
            dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o)); dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
        }
    }
    // one sign, away from zero, finite; |32 * coordinate * 2^KB| stays far below the 2^51 the fixed point holds
    if (!(dmin * dmax > 0.0) || !(fmin(fabs(dmin), fabs(dmax)) > 1e-9) || !(lo_x > -1e6 && hi_x < 1e6 && lo_y > -1e6 && hi_y < 1e6)) fast = 0;
    int bxs = 0, bys = 0;
    double rc = 0.0, de6 = 0.0;
    if (fast) {
        const int bx0 = (int)floor(lo_x) - 1, bx1 = (int)floor(hi_x) + 2, by0 = (int)floor(lo_y) - 1, by1 = (int)floor(hi_y) + 2;
        bxs = bx0 & ~15;                      // 16-byte aligned box start (two's complement: rounds towards -inf)
        bys = by0;
        if (bx1 - bxs + 1 > PBW || by1 - bys + 1 > PBH) fast = 0;
        // a box that misses the image altogether is left to the gather path
        if (bxs >= w || bxs + PBW <= 0 || bys >= p.h || bys + PBH <= 0) fast = 0;
        const double xc = 0.5 * (x0 + x1), yc = 0.5 * (y0 + y1);
        const double denc = fma(M[6], xc, fma(M[7], yc, M[8]));
        rc = 1.0 / denc;
        de6 = -M[6] * rc;
        // |1 - den * rc| over the tile: the third-order expansion is good to e^4 <= 1e-12, the first-order one to e^2 <= 4e-12
        const double emax = fabs(de6) * (0.5 * (x1 - x0) + 1.0) + fabs(M[7] * rc) * (0.5 * (y1 - y0) + 1.0);
        if (!(emax < 1e-3)) fast = 0;
        else if (fast && emax < 2e-6) fast = 3;
    }
    mode = novalid ? 2 : fast;
    const double RS = rc * (double)(32u << KB) ;
    // the per-row terms of the numerators and of e = 1 - den * rc
    for (int ry = lane; ry < TR; ry += 32) {
        const double yd = (double)(ty0 - 2 + ry);
        t.rowX[ry] = fma(M[1], yd, M[2]);
        t.rowY[ry] = fma(M[4], yd, M[5]);
        t.rowE[ry] = fma(-fma(M[7], yd, M[8]), rc, 1.0);
    }
    if (lane < 4) {
        // plane `lane`, bit l: column x0 + 4 l + lane inside [0, w)  <=>  l in [lmin, lmax]
        const int xs = x0 + lane;
        const int lmin = xs >= 0 ? 0 : (-xs + 3) >> 2;
        const int lmax = min(31, (w - 1 - xs) >> 2);              // arithmetic shift: negative when the plane starts beyond the image
        uint32_t m = 0;
        if (lmax >= lmin) m = (0xffffffffu >> (31 - lmax)) & (0xffffffffu << lmin);
        t.colin[lane] = m;
    }
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < 9; i++) t.M[i] = M[i];
        t.tx0 = tx0; t.ty0 = ty0; t.b = b;
        t.zp = p.nslots ? (p.prev_slot0 + b) % p.nslots : b;
        t.RS = RS; t.de6 = de6;
        t.magx = MASK_MAGIC - (double)bxs * (double)(32u << KB);
        t.magy = MASK_MAGIC - (double)bys * (double)(32u << KB);
        t.cxo = x0 - (x0 & ~15);
        t.mode = mode;
        t.pad[0] = bxs; t.pad[1] = bys;
    }
    __syncwarp();
}

// Persistent, warp-specialised: warp 8 walks the CTA's tiles ahead of the others -- source box, expansion constants, row
// tables, two TMA box loads per tile into a two-stage ring guarded by full / empty mbarriers -- while warps 0-7 sample and
// threshold the tile that has landed (phase 1), erode + dilate it on bit planes (phase 2) and expand / store it (phase 3).
template <bool ALIGNED>
__global__ void __launch_bounds__(NCONS + 32, MASK_CTAS_PER_SM) k_mask(const MaskParams p, const __grid_constant__ MaskKernelMaps maps, int ntx, int nty,
                                                        int ntiles)
{
    __shared__ __align__(128) uint8_t sPrev[NSTAGE][PREV_STAGE];
    __shared__ __align__(128) uint8_t sCur[NSTAGE][CBW * CBH];
    static_assert((CBW * CBH) % 128 == 0 && PREV_STAGE % 128 == 0, "TMA destinations are 128-byte aligned");
    __shared__ __align__(16) uint4 Wp[2][TR];    // thresholded rows, four interleaved bit planes each (double buffered over tiles)
    __shared__ __align__(16) uint4 Dp[2][TR];    // after erode + dilate (rows 2 .. TR - 3 are the output rows)
    __shared__ __align__(16) MaskTile sT[NSTAGE];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE];
    const int w = p.w, h = p.h;
    const int bh0 = h < 16 ? h : 16;
    const int bw0 = (1024 / bh0 < w) ? 1024 / bh0 : w;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS / 32); }
        mbar_fence_init();
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;

    if (threadIdx.x >= NCONS) {
        // ---- producer warp ------------------------------------------------------------------------------------------------------
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const int s = it % NSTAGE;
            if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
            MaskTile &t = sT[s];
            mask_tile_setup(p, t, tile, ntx, nty, lane);
            if (lane == 0) {
                if (t.mode == 1 || t.mode == 3) {
                    const int zc = p.nslots ? (p.cur_slot0 + t.b) % p.nslots : t.b;
                    mbar_expect_tx(&full[s], PBW * PBH + CBW * CBH);
                    tma_load_3d(sPrev[s], &maps.prev, t.pad[0], t.pad[1], t.zp, &full[s]);
                    tma_load_3d(sCur[s], &maps.cur, (t.tx0 - 4) & ~15, t.ty0 - 2, zc, &full[s]);
                } else mbar_arrive(&full[s]);
            }
            __syncwarp();
        }
        return;
    }

    // ---- consumers ----------------------------------------------------------------------------------------------------------
    // One CTA barrier per tile: phase 1 of tile i fills Wp[i & 1]; after the barrier the first TH threads erode + dilate it into
    // Dp[i & 1] while everybody expands and stores tile i - 1 from Dp[(i - 1) & 1] (finished before the barrier).
    const int tid = threadIdx.x, warp = tid >> 5;
    int local = 0;
    int it = 0;
    int ptx0 = 0, pty0 = 0, pb = -1;                  // the tile whose bit planes wait in Dp for phase 3
    // phase 3: bits -> bytes, coalesced 32-bit stores.  Lane l owns the 4-pixel run at bit l of the four planes: rotating a plane so
    // that the bit becomes the sign bit of byte 0 lets PRMT's sign replication write 0x00 / 0xff.
    const int rot = (lane - 7) & 31;
    auto phase3 = [&](const uint4 *D, int tx0, int ty0, int b) {
        const int x = tx0 - 4 + 4 * lane;
        const int nrow = min(TH, h - ty0);
        if (p.packed) {
            // 1 bit per pixel: a lane's run is a nibble, an odd lane and its right neighbour make the byte at x / 8 (the tile starts at
            // a multiple of 120 pixels = 15 bytes)
            uint8_t *dst = p.mask + (size_t)b * p.mask_stride + (size_t)(ty0 + warp) * p.mask_pitch + (x >> 3);
            const size_t step = (size_t)(NCONS / 32) * p.mask_pitch;
            for (int ry = warp; ry < nrow; ry += NCONS / 32, dst += step) {
                uint4 d = D[ry + 2];
                if (p.morph) {
                    const uint4 u = D[ry + 1], v = D[ry + 3];
                    d.x |= u.x | v.x; d.y |= u.y | v.y; d.z |= u.z | v.z; d.w |= u.w | v.w;
                }
                uint32_t nib = ((d.x >> lane) & 1u) | (((d.y >> lane) & 1u) << 1) | (((d.z >> lane) & 1u) << 2) | (((d.w >> lane) & 1u) << 3);
                if (lane < 1 || lane > TW / 4) nib = 0;
                const uint32_t right = __shfl_down_sync(0xffffffffu, nib, 1);
                local += 8 * __popc(nib);
                if ((lane & 1) && lane <= TW / 4 && x < w) *dst = (uint8_t)(nib | (right << 4));
            }
            return;
        }
        if (lane < 1 || lane > TW / 4) return;
        if (x >= w) return;
        uint8_t *dst = p.mask + (size_t)b * p.mask_stride + (size_t)(ty0 + warp) * p.mask_pitch + x;
        const size_t step = (size_t)(NCONS / 32) * p.mask_pitch;
        for (int ry = warp; ry < nrow; ry += NCONS / 32, dst += step) {
            uint4 d = D[ry + 2];
            if (p.morph) {
                // the dilate's pass along y
                const uint4 u = D[ry + 1], v = D[ry + 3];
                d.x |= u.x | v.x; d.y |= u.y | v.y; d.z |= u.z | v.z; d.w |= u.w | v.w;
            }
            const uint32_t r0 = __funnelshift_r(d.x, d.x, rot), r1 = __funnelshift_r(d.y, d.y, rot);
            const uint32_t r2 = __funnelshift_r(d.z, d.z, rot), r3 = __funnelshift_r(d.w, d.w, rot);
            const uint32_t o = prmt(prmt(r0, r1, 0x00c8u), prmt(r2, r3, 0x00c8u), 0x5410u);
            if (ALIGNED && x + 3 < w) { *reinterpret_cast<uint32_t *>(dst) = o; local += __popc(o); }
            else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x + j < w) { dst[j] = (uint8_t)(o >> (8 * j)); local += __popc((o >> (8 * j)) & 0xffu); }
            }
        }
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        const int s = it % NSTAGE;
        mbar_wait(&full[s], (it / NSTAGE) & 1);
        const MaskTile &t = sT[s];
        const int tx0 = t.tx0, ty0 = t.ty0, b = t.b, mode = t.mode;
        uint4 *W = Wp[it & 1], *D = Dp[it & 1];
        const uint32_t ci0 = t.colin[0], ci1 = t.colin[1], ci2 = t.colin[2], ci3 = t.colin[3];
        if (mode == 2) {
            // no egomotion (fewer than the minimal number of vectors): empty mask (all-zero planes, stored like any other tile)
            if (tid < TR) W[tid] = make_uint4(0u, 0u, 0u, 0u);
        } else if (mode == 1 || mode == 3) {
            const uint8_t *prev = p.prev + (long long)t.zp * p.stride;
            if (mode == 3) mask_phase1_fast<1>(p, t, smem_u32(&t), smem_u32(sPrev[s]), smem_u32(sCur[s]), W, prev, bw0, warp, lane);
            else mask_phase1_fast<3>(p, t, smem_u32(&t), smem_u32(sPrev[s]), smem_u32(sCur[s]), W, prev, bw0, warp, lane);
        } else {
            // ---- phase 1 (gather): per-pixel loads through L2, reference operation sequence with the Newton shortcut ---------
            const uint8_t *prev = p.prev + (long long)t.zp * p.stride;
            const double *sM = t.M;
            const ThreshConst tc = thresh_const(p.thresh);
            const uint8_t *cur = p.cur + (long long)(p.nslots ? (p.cur_slot0 + b) % p.nslots : b) * p.stride;
            const bool uniform_block = (bw0 & 3) == 0;     // a run of 4 aligned columns never straddles a 64-column block
            const int x = tx0 - 4 + 4 * lane;
            for (int ry = warp; ry < TR; ry += NCONS / 32) {
                const int y = ty0 - 2 + ry;
                bool mot[4] = {true, true, true, true};
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
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int xj = x + j;
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
                            mot[j] = (uint32_t)((wv - (int)((c4 >> (8 * j)) & 0xffu)) * 1024 + tc.kt) >= tc.lim;
                        }
                    }
                }
                const uint32_t w0 = __ballot_sync(0xffffffffu, mot[0]), w1 = __ballot_sync(0xffffffffu, mot[1]);
                const uint32_t w2 = __ballot_sync(0xffffffffu, mot[2]), w3 = __ballot_sync(0xffffffffu, mot[3]);
                if (lane == 0) {
                    const uint4 row = make_uint4(w0 | ~ci0, w1 | ~ci1, w2 | ~ci2, w3 | ~ci3);
                    W[ry] = p.morph ? morph_row_x<true>(row) : row;
                }
            }
        }
        consumer_sync();
        // the staged boxes and the tile record are consumed: hand the stage back to the producer
        // (the tile's scalars were copied to registers above; phases 2 and 3 only touch the bit planes)
        if (lane == 0) mbar_arrive(&empty[s]);

        // ---- phase 2: the erode's pass along y, then the dilate's pass along x; one thread per row (rows 1 .. TR - 2), on the two
        // warps that own one row less of phase 3 ---------------------------------------------------------------------------------
        if (tid >= 4 * 32 && tid < 4 * 32 + TR - 2) {
            const int r = tid - 4 * 32 + 1;                     // row of the computed region
            uint4 o;
            if (p.morph && mode != 2) {
                const uint4 a0 = W[r - 1], a1 = W[r], a2 = W[r + 1];
                // eroded row r; outside the image it is the dilate identity (0)
                const int y = ty0 - 2 + r;
                const bool in = y >= 0 && y < h;
                uint4 e;
                e.x = in ? (a0.x & a1.x & a2.x & ci0) : 0u; e.y = in ? (a0.y & a1.y & a2.y & ci1) : 0u;
                e.z = in ? (a0.z & a1.z & a2.z & ci2) : 0u; e.w = in ? (a0.w & a1.w & a2.w & ci3) : 0u;
                o = morph_row_x<false>(e);
            } else o = W[r];
            // only columns inside the image count and are stored
            o.x &= ci0; o.y &= ci1; o.z &= ci2; o.w &= ci3;
            D[r] = o;
        }
        // ---- phase 3 of the PREVIOUS tile (its planes were finished before the barrier above) -------------------------------------
        if (pb >= 0) phase3(Dp[(it - 1) & 1], ptx0, pty0, pb);
        ptx0 = tx0; pty0 = ty0; pb = b;
    }
    consumer_sync();
    if (pb >= 0) phase3(Dp[(it - 1) & 1], ptx0, pty0, pb);
    if (p.stat_mask) {
        for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
        if (lane == 0 && local) atomicAdd(p.stat_mask, (unsigned long long)(local >> 3));     // eight set bits per 0xff byte
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
    // persistent CTAs: MASK_CTAS_PER_SM per SM (register bound), every CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms < 1) sms = 148; }
    const int grid = (int)(nt < (long long)sms * MASK_CTAS_PER_SM ? nt : (long long)sms * MASK_CTAS_PER_SM);
    if (aligned) k_mask<true><<<grid, NCONS + 32, 0, s>>>(p, km, ntx, nty, (int)nt);
    else k_mask<false><<<grid, NCONS + 32, 0, s>>>(p, km, ntx, nty, (int)nt);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
