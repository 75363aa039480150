// k_lk_phase.cu -- K2 for GRID points (pts_in == NULL, window 40): pyramidal Lucas-Kanade on precomputed sub-pixel
// phase planes.
//
// Replaces cv::calcOpticalFlowPyrLK(...) as called at common/src/optical_flow_calculator.cpp:71 (grid built at :56-64).
// Arithmetic is identical to k_lk_tma.cu / k_lk.cu (OpenCV's LKTrackerInvoker fixed-point scheme, lk_common.cuh); only
// WHERE the window samples are computed differs:
//
//  * The tracked grid sits at integer multiples of pixel_step, so at level l a point's window origin
//    (x / 2^l - 19.5) has one of (2^l / gcd(ps, 2^l))^2 fractional parts.  LKTrackerInvoker's bilinear weights depend
//    on that fraction only, so the window samples I (5 extra bits), Ix, Iy of ALL points of a class are the same
//    function of the integer origin.  k_phase_planes evaluates them once per pixel and class (2 N pixel evaluations
//    per frame at pixel_step 10 instead of 80 N per-point tap evaluations); the per-point window build of k_lk_tma
//    (48 % of its instructions) becomes one TMA box load of the Ix, Iy planes.
//  * The five integer window sums A11, A12, A22, sum I*Ix, sum I*Iy of a point are box sums over its class planes; points
//    of one grid row and class share their 40 window rows, so k_window_sums forms the column sums once per (row, class)
//    and every point adds up 40 of them: 28 M pixel visits per 1080p pair instead of 166 M per-point taps, exact int64.
//    Where the lattice step of a level divides the window (every level of pixel_step 5 / 10 / 20 / 40) k_window_sums_ring does both
//    jobs in one pass down the level: the samples are evaluated on the way, Ix / Iy stored, the I plane never materialised, no row
//    visited twice (running prefix + ring of lattice-boundary snapshots in shared memory).
//  * The iteration loop (J tile staged by TMA, dp2a bilinear samples, dp2a mismatch accumulation on packed tap pairs,
//    REDUX reductions) is the one of k_lk_tma.cu.
#include <stdlib.h>

#include "lk_tile.cuh"
#include "tma.h"

// ---- phase planes ----------------------------------------------------------------------------------------------------
struct PhParams {
    PyrGeom g;
    PhaseGeom pg;
    const uint8_t *img;
    const short2 *der;
    int16_t *ph;
    LkLevelRec *wsum;
    int prev_slot0, pair0;
    int P, ps, gx, gy;
    float half, min_eig;
};

// One thread = 8 consecutive plane pixels x 4 rows of one (class, pair): aligned 8-byte / 32-byte loads of the padded
// pyramid planes, 16-byte stores.  Horizontal partial sums of a source row serve as the bottom half of one output row and
// the top half of the next.
__global__ void __launch_bounds__(256) k_phase_planes(const PhParams q, int level)
{
    const PhaseLevel PL = q.pg.lv[level];
    const LevelGeom L = q.g.lv[level];
    const int nxt = PL.pitch >> 3;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int xt = t % nxt, yt = t / nxt;
    const int Y0 = yt * 4;
    if (Y0 >= PL.h) return;
    const int cls = blockIdx.y, b = blockIdx.z;
    const int cx = cls % PL.ncx, cy = cls / PL.ncx;
    int w00, w01, w10, w11;
    {
        // the fraction of a class representative's window origin, computed exactly like the LK kernels compute a point's
        const float scale = lk_level_scale(level);
        const float ppx = __fsub_rn((float)(cx << PL.shift) * scale, q.half), ppy = __fsub_rn((float)(cy << PL.shift) * scale, q.half);
        lk_weights(__fsub_rn(ppx, floorf(ppx)), __fsub_rn(ppy, floorf(ppy)), w00, w01, w10, w11);
    }
    const int slot = (q.prev_slot0 + b) % q.g.nslots;
    const int X0 = xt * 8;
    // plane pixel (X, Y) <- source pixel (X - margin, Y - margin) of the padded level plane
    const size_t src0 = (size_t)(q.g.pady - MD_PH_MARGIN + Y0) * L.pitch + (q.g.padx - MD_PH_MARGIN + X0);
    const uint8_t *ip = q.img + (size_t)slot * q.g.slot_img_bytes + L.img_off + src0;
    const short2 *dp = q.der + (size_t)slot * q.g.slot_der_elems + L.der_off + src0;
    const size_t plane = (size_t)PL.pitch * PL.h;
    int16_t *out = q.ph + (size_t)(q.pair0 + b) * q.pg.pair_elems + PL.off + (size_t)cls * 3 * plane + (size_t)Y0 * PL.pitch + X0;

    int ti[8], tx[8], ty[8];      // top halves carried from the previous source row
#pragma unroll
    for (int r = 0; r < 5; r++) {
        if (Y0 + r > PL.h) break;             // source row r feeds output rows r-1 and r
        const uint2 iw = *reinterpret_cast<const uint2 *>(ip + (size_t)r * L.pitch);
        const uint32_t i8 = ip[(size_t)r * L.pitch + 8];
        const uint4 da = *reinterpret_cast<const uint4 *>(dp + (size_t)r * L.pitch);
        const uint4 db = *reinterpret_cast<const uint4 *>(dp + (size_t)r * L.pitch + 4);
        const uint32_t d8 = *reinterpret_cast<const uint32_t *>(dp + (size_t)r * L.pitch + 8);
        int pv[9], dx[9], dy[9];
        const uint32_t dd[9] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w, d8};
#pragma unroll
        for (int i = 0; i < 9; i++) {
            pv[i] = i < 4 ? (int)((iw.x >> (8 * i)) & 0xffu) : (i < 8 ? (int)((iw.y >> (8 * (i - 4))) & 0xffu) : (int)i8);
            dx[i] = (int)(short)(dd[i] & 0xffffu);
            dy[i] = (int)dd[i] >> 16;
        }
        uint32_t oi[4], ox[4], oy[4];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int bi = w10 * pv[i] + w11 * pv[i + 1], bx = w10 * dx[i] + w11 * dx[i + 1], by = w10 * dy[i] + w11 * dy[i + 1];
            if (r > 0) {
                const int iv = (ti[i] + bi + (1 << (W_BITS - 5 - 1))) >> (W_BITS - 5);
                const int xv = (tx[i] + bx + (1 << (W_BITS - 1))) >> W_BITS;
                const int yv = (ty[i] + by + (1 << (W_BITS - 1))) >> W_BITS;
                if (i & 1) {
                    oi[i >> 1] |= (uint32_t)iv << 16; ox[i >> 1] |= (uint32_t)xv << 16; oy[i >> 1] |= (uint32_t)yv << 16;
                } else {
                    oi[i >> 1] = (uint32_t)iv & 0xffffu; ox[i >> 1] = (uint32_t)xv & 0xffffu; oy[i >> 1] = (uint32_t)yv & 0xffffu;
                }
            }
            ti[i] = w00 * pv[i] + w01 * pv[i + 1]; tx[i] = w00 * dx[i] + w01 * dx[i + 1]; ty[i] = w00 * dy[i] + w01 * dy[i + 1];
        }
        if (r > 0) {
            int16_t *o = out + (size_t)(r - 1) * PL.pitch;
            *reinterpret_cast<uint4 *>(o) = make_uint4(oi[0], oi[1], oi[2], oi[3]);
            *reinterpret_cast<uint4 *>(o + plane) = make_uint4(ox[0], ox[1], ox[2], ox[3]);
            *reinterpret_cast<uint4 *>(o + 2 * plane) = make_uint4(oy[0], oy[1], oy[2], oy[3]);
        }
    }
}

// ---- per-point window sums -----------------------------------------------------------------------------------------------
// The window origins of one phase class form a lattice with step s = ps / gcd(ps, 2^level) in both directions, and the
// windows of neighbouring lattice rows share 40 - s of their 40 rows.  One CTA owns a block of the lattice (<= WS_SEG origin
// rows x the origin columns that fit 512 plane columns) of one (class, pair); a thread keeps the five column sums
// (x^2, x*y, y^2, i*x, i*y over the window's 40 rows; int32: 40 * 8160 * 4080 fits) of two adjacent plane columns and
// SLIDES them from one origin row to the next (subtract s rows, add s rows); after every origin row the points of that row
// add up 40 consecutive column sums from shared memory (int64).  Same integers as the per-point loops of k_lk_tma.
#define WS_COLS 512
#define WS_SEG 16
#define WS_GRP 4
__device__ __forceinline__ void ws_accumulate(int (&acc)[2][5], uint32_t wi, uint32_t wx, uint32_t wy, int sign)
{
    const int i0 = (int)(wi & 0xffffu), i1 = (int)(wi >> 16);
    const int x0 = (int)(short)wx * sign, x1 = ((int)wx >> 16) * sign, y0 = (int)(short)wy, y1 = (int)wy >> 16;
    const int xs0 = (int)(short)wx, xs1 = (int)wx >> 16;
    acc[0][0] += x0 * xs0; acc[0][1] += x0 * y0; acc[0][2] += y0 * y0 * sign; acc[0][3] += i0 * x0; acc[0][4] += i0 * y0 * sign;
    acc[1][0] += x1 * xs1; acc[1][1] += x1 * y1; acc[1][2] += y1 * y1 * sign; acc[1][3] += i1 * x1; acc[1][4] += i1 * y1 * sign;
}

__global__ void __launch_bounds__(256) k_window_sums(const PhParams q, int level, int cp, int nchunk_x)
{
    __shared__ int ws_col[WS_GRP][5][WS_COLS + 8];
    const PhaseLevel PL = q.pg.lv[level];
    const int ncx = PL.ncx, cls = blockIdx.y, cx = cls % ncx, cy = cls / ncx, b = blockIdx.z;
    const int chunk = blockIdx.x % nchunk_x, seg = blockIdx.x / nchunk_x;
    const int msk = (1 << level) - 1;
    const int step = q.ps >> PL.shift;                          // lattice step in plane pixels: ps / gcd(ps, 2^level)
    const float scale = lk_level_scale(level);
    // first grid index of each axis that falls into this class; the class members are every ncx-th index from there
    int kx0 = -1, ky0 = -1;
    for (int t = 0; t < ncx; t++) {
        if (kx0 < 0 && t < q.gx && (((q.ps * t) & msk) >> PL.shift) == cx) kx0 = t;
        if (ky0 < 0 && t < q.gy && (((q.ps * t) & msk) >> PL.shift) == cy) ky0 = t;
    }
    if (kx0 < 0 || ky0 < 0) return;
    const int nx = (q.gx - kx0 + ncx - 1) / ncx, ny = (q.gy - ky0 + ncx - 1) / ncx;
    const int i0 = chunk * cp, i1 = min(nx, i0 + cp);           // lattice columns of this CTA
    const int j0 = seg * WS_SEG, j1 = min(ny, j0 + WS_SEG);     // lattice rows of this CTA
    if (i0 >= i1 || j0 >= j1) return;
    const int ox0 = __float2int_rd(__fsub_rn((float)(q.ps * (kx0 + i0 * ncx)) * scale, q.half)) + MD_PH_MARGIN;
    const int oy0 = __float2int_rd(__fsub_rn((float)(q.ps * (ky0 + j0 * ncx)) * scale, q.half)) + MD_PH_MARGIN;
    const int c_lo = ox0 & ~1;                                   // word aligned first column
    const int ncols = ox0 + (i1 - i0 - 1) * step + 40 - c_lo;    // <= WS_COLS by the choice of cp
    const size_t plane = (size_t)PL.pitch * PL.h;
    const int16_t *base = q.ph + (size_t)(q.pair0 + b) * q.pg.pair_elems + PL.off + (size_t)cls * 3 * plane + c_lo;
    const int wp = PL.pitch >> 1;
    const uint32_t *pI = reinterpret_cast<const uint32_t *>(base) + threadIdx.x, *pX = pI + (plane >> 1), *pY = pX + (plane >> 1);
    const bool col_on = 2 * (int)threadIdx.x < ncols;
    int acc[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    if (col_on) {
#pragma unroll 4
        for (int r = 0; r < 40; r++) {
            const size_t o = (size_t)(oy0 + r) * wp;
            ws_accumulate(acc, __ldg(pI + o), __ldg(pX + o), __ldg(pY + o), 1);
        }
    }
    // WS_GRP origin rows per barrier pair: the sliding column sums of a group are parked in shared memory, then all the points
    // of the group's rows add up their 40 columns together (the horizontal phase of a single row keeps only 240 threads busy)
    for (int jg = j0; jg < j1; jg += WS_GRP) {
        const int gn = min(WS_GRP, j1 - jg);
        __syncthreads();                                         // the previous group's readers are done
        for (int g = 0; g < gn; g++) {
            const int j = jg + g;
            if (j > j0 && col_on) {
                const int top = oy0 + (j - 1 - j0) * step;      // previous origin row
#pragma unroll 5
                for (int r = 0; r < step; r++) {
                    const size_t om = (size_t)(top + r) * wp, op = (size_t)(top + 40 + r) * wp;
                    const uint32_t mi = __ldg(pI + om), mx = __ldg(pX + om), my = __ldg(pY + om);
                    const uint32_t ai = __ldg(pI + op), ax = __ldg(pX + op), ay = __ldg(pY + op);
                    ws_accumulate(acc, mi, mx, my, -1);
                    ws_accumulate(acc, ai, ax, ay, 1);
                }
            }
            if (col_on) {
#pragma unroll
                for (int t = 0; t < 5; t++) { ws_col[g][t][2 * threadIdx.x] = acc[0][t]; ws_col[g][t][2 * threadIdx.x + 1] = acc[1][t]; }
            }
        }
        __syncthreads();
        // one thread per point: its five window sums (40 column sums each, int64), then the level scalars k_lk_phase needs --
        // the f32 matrix entries, the minimum-eigenvalue / determinant test of calcOpticalFlowPyrLK and 1 / det, evaluated
        // once here instead of by every lane of the point's warp
        const int per_row = i1 - i0;
        for (int e = threadIdx.x; e < gn * per_row; e += blockDim.x) {
            const int g = e / per_row, pi = e - g * per_row;
            const int *v = &ws_col[g][0][ox0 + pi * step - c_lo];
            long long sum[5] = {0, 0, 0, 0, 0};
#pragma unroll 4
            for (int c = 0; c < 40; c++) {
#pragma unroll
                for (int t = 0; t < 5; t++) sum[t] += v[t * (WS_COLS + 8) + c];
            }
            const float FLT_SCALE = 1.f / (1 << 20);
            LkLevelRec rec;
            rec.A11 = (float)sum[0] * FLT_SCALE; rec.A12 = (float)sum[1] * FLT_SCALE; rec.A22 = (float)sum[2] * FLT_SCALE;
            float D;
            rec.Dinv = lk_min_eig_ok(rec.A11, rec.A12, rec.A22, 40, q.min_eig, D) ? D : 0.f;
            rec.C1 = sum[3]; rec.C2 = sum[4];
            const int kx = kx0 + (i0 + pi) * ncx, ky = ky0 + (jg + g) * ncx;
            LkLevelRec *o = q.wsum + (((size_t)(q.pair0 + b) * q.g.nlev + level) * q.P) + (size_t)kx * q.gy + ky;
            *reinterpret_cast<float4 *>(o) = make_float4(rec.A11, rec.A12, rec.A22, rec.Dinv);
            *reinterpret_cast<longlong2 *>(&o->C1) = make_longlong2(rec.C1, rec.C2);
        }
    }
}

// ---- single-pass window sums (and, FUSED, the Ix / Iy phase planes themselves) ------------------------------------------------
// For lattice steps that divide the window (40 % step == 0, step >= 5: every level of pixel_step 5, 10, 20, 40) each plane row is
// visited ONCE: a thread walks down two adjacent plane columns with a running prefix of the five products (uint32, wrap-around: a
// difference of two prefixes 40 rows apart is the exact window column sum, which fits 31 bits), parks the prefix of every lattice row
// boundary in a thread-private ring of R = 40 / step entries in shared memory, and at boundary k emits prefix(k) - prefix(k - R) =
// the column sums of the window of lattice row k - R.  No row is subtracted again (k_window_sums reads every row 1.7 - 2.4 times).
// The horizontal phase adds the column sums of a lattice row in two levels -- sums of `step` adjacent columns (int64), then R of
// those per point -- instead of 40 columns per point.
// FUSED: the window samples are not loaded from the phase planes but evaluated on the way down from the padded pyramid level
// (image + Scharr planes, exactly k_phase_planes' expressions; the I sample by dp2a on the packed pixel bytes, the rounding constants
// riding in the carried top halves); the Ix / Iy samples are stored for k_lk_phase, the I samples (a third of the plane bytes, read
// by nobody else) are never materialised, and k_phase_planes is not launched for the level.  Only plane pixels inside some grid
// point's window are written: k_lk_phase uses nothing else of a plane.  The rows of the next lattice step are requested before the
// boundary work (ring, horizontal phase, barriers) of the current one.
// NT threads = 2 NT plane columns per CTA; the launcher picks NT per level (wide levels: 256, i.e. less overlap between
// neighbouring CTAs, whose windows share 40 - step columns; narrow class planes: 128 / 96 / 64).
#define WR_GRP 2
#define WR_NG(NT) (2 * (NT) / 5 + 3)
#define WR_BATCH 5
__device__ __forceinline__ void wr_accumulate(unsigned (&acc)[2][5], int i0, int i1, int x0, int x1, int y0, int y1)
{
    acc[0][0] += (unsigned)(x0 * x0); acc[0][1] += (unsigned)(x0 * y0); acc[0][2] += (unsigned)(y0 * y0);
    acc[0][3] += (unsigned)(i0 * x0); acc[0][4] += (unsigned)(i0 * y0);
    acc[1][0] += (unsigned)(x1 * x1); acc[1][1] += (unsigned)(x1 * y1); acc[1][2] += (unsigned)(y1 * y1);
    acc[1][3] += (unsigned)(i1 * x1); acc[1][4] += (unsigned)(i1 * y1);
}
static inline size_t wr_smem_bytes(int NT, int R)
{
    return ((size_t)R * 5 * 2 * NT + (size_t)WR_GRP * 5 * (2 * NT + 8)) * sizeof(int) + (size_t)WR_GRP * 5 * WR_NG(NT) * sizeof(long long);
}

template <bool FUSED, int NT, int MINB = 512 / NT>
__global__ void __launch_bounds__(NT, MINB) k_window_sums_ring(const PhParams q, int level, int cp, int nchunk_x, int seg)
{
    constexpr int COLS = 2 * NT, CPITCH = COLS + 8, NGP = WR_NG(NT);
    extern __shared__ __align__(16) int wr_smem[];
    const PhaseLevel PL = q.pg.lv[level];
    const int ncx = PL.ncx, cls = blockIdx.y, cx = cls % ncx, cy = cls / ncx, b = blockIdx.z;
    const int chunk = blockIdx.x % nchunk_x, sg = blockIdx.x / nchunk_x;
    const int msk = (1 << level) - 1;
    const int step = q.ps >> PL.shift;
    const int R = 40 / step;
    const float scale = lk_level_scale(level);
    int kx0 = -1, ky0 = -1;
    for (int t = 0; t < ncx; t++) {
        if (kx0 < 0 && t < q.gx && (((q.ps * t) & msk) >> PL.shift) == cx) kx0 = t;
        if (ky0 < 0 && t < q.gy && (((q.ps * t) & msk) >> PL.shift) == cy) ky0 = t;
    }
    if (kx0 < 0 || ky0 < 0) return;
    const int nx = (q.gx - kx0 + ncx - 1) / ncx, ny = (q.gy - ky0 + ncx - 1) / ncx;
    const int i0 = chunk * cp, i1 = min(nx, i0 + cp);
    const int j0 = sg * seg, j1 = min(ny, j0 + seg);
    if (i0 >= i1 || j0 >= j1) return;
    const int per_row = i1 - i0, nlat = j1 - j0;
    const int ox0 = __float2int_rd(__fsub_rn((float)(q.ps * (kx0 + i0 * ncx)) * scale, q.half)) + MD_PH_MARGIN;
    const int oy0 = __float2int_rd(__fsub_rn((float)(q.ps * (ky0 + j0 * ncx)) * scale, q.half)) + MD_PH_MARGIN;
    const int c_lo = ox0 & ~1, off = ox0 - c_lo;
    const int ncols = off + (per_row - 1) * step + 40;             // <= COLS by the choice of cp
    const int tid = threadIdx.x;
    const bool col_on = 2 * tid < ncols;
    const size_t plane = (size_t)PL.pitch * PL.h;
    int16_t *pbase = q.ph + (size_t)(q.pair0 + b) * q.pg.pair_elems + PL.off + (size_t)cls * 3 * plane + (size_t)oy0 * PL.pitch + c_lo + 2 * tid;
    const int wp = PL.pitch >> 1;

    int *ring = wr_smem;                                             // [R][5][COLS], thread private columns
    int *col = ring + R * 5 * COLS;                                  // [WR_GRP][5][CPITCH] window column sums of a group of lattice rows
    long long *grp = reinterpret_cast<long long *>(col + WR_GRP * 5 * CPITCH);            // [WR_GRP][5][NGP] sums of `step` columns

    // row sources
    const uint32_t *pI = reinterpret_cast<const uint32_t *>(pbase), *pX = pI + (plane >> 1), *pY = pX + (plane >> 1);   // !FUSED
    uint32_t *oX = reinterpret_cast<uint32_t *>(pbase) + (plane >> 1), *oY = oX + (plane >> 1);                      // FUSED
    const uint8_t *ip = nullptr;
    const short2 *dp = nullptr;
    int spitch = 0;
    int w00 = 0, w01 = 0, w10 = 0, w11 = 0, wtop = 0, wbot = 0;
    // top halves of the current plane row (FUSED), the rounding constants of the final shifts included
    constexpr int RND_I = 1 << (W_BITS - 5 - 1), RND_D = 1 << (W_BITS - 1);
    int ti[2] = {0, 0}, tx[2] = {0, 0}, ty[2] = {0, 0};
    if (FUSED) {
        const LevelGeom L = q.g.lv[level];
        const float ppx = __fsub_rn((float)(cx << PL.shift) * scale, q.half), ppy = __fsub_rn((float)(cy << PL.shift) * scale, q.half);
        lk_weights(__fsub_rn(ppx, floorf(ppx)), __fsub_rn(ppy, floorf(ppy)), w00, w01, w10, w11);
        wtop = w00 | (w01 << 16); wbot = w10 | (w11 << 16);       // 14-bit weights as the 16-bit halves dp2a multiplies with two pixel bytes
        const int slot = (q.prev_slot0 + b) % q.g.nslots;
        spitch = L.pitch;
        const size_t src0 = (size_t)(q.g.pady - MD_PH_MARGIN + oy0) * L.pitch + (q.g.padx - MD_PH_MARGIN + c_lo + 2 * tid);
        ip = q.img + (size_t)slot * q.g.slot_img_bytes + L.img_off + src0;
        dp = q.der + (size_t)slot * q.g.slot_der_elems + L.der_off + src0;
        if (col_on) {
            const uint32_t pw = *reinterpret_cast<const uint16_t *>(ip), p2 = ip[2];
            const uint2 dw = *reinterpret_cast<const uint2 *>(dp);
            const uint32_t d2 = *reinterpret_cast<const uint32_t *>(dp + 2);
            const uint32_t pa = pw | (p2 << 16);
            const int x0 = (int)(short)(dw.x & 0xffffu), x1 = (int)(short)(dw.y & 0xffffu), x2 = (int)(short)(d2 & 0xffffu);
            const int y0 = (int)dw.x >> 16, y1 = (int)dw.y >> 16, y2 = (int)d2 >> 16;
            ti[0] = dp2a_lo(wtop, pa, RND_I); ti[1] = dp2a_lo(wtop, pa >> 8, RND_I);
            tx[0] = w00 * x0 + (w01 * x1 + RND_D); tx[1] = w00 * x1 + (w01 * x2 + RND_D);
            ty[0] = w00 * y0 + (w01 * y1 + RND_D); ty[1] = w00 * y1 + (w01 * y2 + RND_D);
        }
        ip += spitch; dp += spitch;                                  // plane row Y takes its bottom half from source row Y + 1
    }

    // one batch of rows in registers: FUSED pw / p2 = pixels (X, X + 1) / (X + 2), dw / d2 = their Scharr pairs; else wi, wx, wy words
    uint32_t ba[WR_BATCH], bb[WR_BATCH], bc[WR_BATCH];
    uint2 bd[WR_BATCH];
    auto load_batch = [&](int nr) {
#pragma unroll
        for (int u = 0; u < WR_BATCH; u++) {
            if (u < nr) {
                if (FUSED) {
                    ba[u] = __ldg(reinterpret_cast<const uint16_t *>(ip + (size_t)u * spitch)); bb[u] = __ldg(ip + (size_t)u * spitch + 2);
                    bd[u] = __ldg(reinterpret_cast<const uint2 *>(dp + (size_t)u * spitch));
                    bc[u] = __ldg(reinterpret_cast<const uint32_t *>(dp + (size_t)u * spitch + 2));
                } else {
                    ba[u] = __ldg(pI + u * wp); bb[u] = __ldg(pX + u * wp); bc[u] = __ldg(pY + u * wp);
                }
            }
        }
        if (FUSED) { ip += (size_t)nr * spitch; dp += (size_t)nr * spitch; }
        else { pI += nr * wp; pX += nr * wp; pY += nr * wp; }
    };
    unsigned acc[2][5] = {{0, 0, 0, 0, 0}, {0, 0, 0, 0, 0}};
    auto compute_batch = [&](int nr) {
#pragma unroll
        for (int u = 0; u < WR_BATCH; u++) {
            if (u < nr) {
                if (FUSED) {
                    const uint32_t pa = ba[u] | (bb[u] << 16), pb = pa >> 8;      // bytes (a0, a1, a2) and (a1, a2)
                    const int x0 = (int)(short)(bd[u].x & 0xffffu), x1 = (int)(short)(bd[u].y & 0xffffu), x2 = (int)(short)(bc[u] & 0xffffu);
                    const int y0 = (int)bd[u].x >> 16, y1 = (int)bd[u].y >> 16, y2 = (int)bc[u] >> 16;
                    const int iv0 = dp2a_lo(wbot, pa, ti[0]) >> (W_BITS - 5), iv1 = dp2a_lo(wbot, pb, ti[1]) >> (W_BITS - 5);
                    const int xv0 = (tx[0] + w10 * x0 + w11 * x1) >> W_BITS, xv1 = (tx[1] + w10 * x1 + w11 * x2) >> W_BITS;
                    const int yv0 = (ty[0] + w10 * y0 + w11 * y1) >> W_BITS, yv1 = (ty[1] + w10 * y1 + w11 * y2) >> W_BITS;
                    ti[0] = dp2a_lo(wtop, pa, RND_I); ti[1] = dp2a_lo(wtop, pb, RND_I);
                    tx[0] = w00 * x0 + (w01 * x1 + RND_D); tx[1] = w00 * x1 + (w01 * x2 + RND_D);
                    ty[0] = w00 * y0 + (w01 * y1 + RND_D); ty[1] = w00 * y1 + (w01 * y2 + RND_D);
                    oX[u * wp] = ((uint32_t)xv0 & 0xffffu) | ((uint32_t)xv1 << 16);
                    oY[u * wp] = ((uint32_t)yv0 & 0xffffu) | ((uint32_t)yv1 << 16);
                    wr_accumulate(acc, iv0, iv1, xv0, xv1, yv0, yv1);
                } else {
                    wr_accumulate(acc, (int)(ba[u] & 0xffffu), (int)(ba[u] >> 16), (int)(short)bb[u], (int)bb[u] >> 16, (int)(short)bc[u],
                                  (int)bc[u] >> 16);
                }
            }
        }
        if (FUSED) { oX += nr * wp; oY += nr * wp; }
    };

    const int K = nlat + R;                                          // lattice row boundaries met on the way down
    const int nr0 = min(WR_BATCH, step);
    if (col_on) load_batch(nr0);
    int slot_r = 0, g = 0, jg = j0;
    for (int k = 0; k < K; k++) {
        if (col_on) {
            int *rg = ring + slot_r * 5 * COLS + 2 * tid;
            if (k >= R) {
                int *cg = col + g * 5 * CPITCH + 2 * tid;
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    const int2 old = *reinterpret_cast<const int2 *>(rg + t * COLS);
                    *reinterpret_cast<int2 *>(cg + t * CPITCH) = make_int2((int)(acc[0][t] - (unsigned)old.x), (int)(acc[1][t] - (unsigned)old.y));
                }
            }
#pragma unroll
            for (int t = 0; t < 5; t++) *reinterpret_cast<int2 *>(rg + t * COLS) = make_int2((int)acc[0][t], (int)acc[1][t]);
        }
        if (++slot_r == R) slot_r = 0;
        if (k >= R && (++g == WR_GRP || k == K - 1)) {
            // the points of g lattice rows (first: jg): sums of `step` adjacent column sums, then R of those per point and the level
            // scalars k_lk_phase needs (f32 matrix entries, minimum-eigenvalue / determinant test, 1 / det), once per point
            const int NG = per_row - 1 + R;
            __syncthreads();
            // a warp takes a (lattice row, quantity) line of column sums, a lane a group of `step` columns
            for (int row = tid >> 5; row < g * 5; row += NT / 32) {
                for (int gi = tid & 31; gi < NG; gi += 32) {
                    const int *v = col + row * CPITCH + off + gi * step;
                    long long s;
                    if (step == 5) s = ((long long)v[0] + v[1]) + ((long long)v[2] + v[3]) + v[4];
                    else if (step == 10)
                        s = (((long long)v[0] + v[1]) + ((long long)v[2] + v[3])) + (((long long)v[4] + v[5]) + ((long long)v[6] + v[7])) +
                            ((long long)v[8] + v[9]);
                    else {
                        s = 0;
                        for (int c0 = 0; c0 < step; c0 += 8) {
                            int t8[8];
#pragma unroll
                            for (int u = 0; u < 8; u++) t8[u] = c0 + u < step ? v[c0 + u] : 0;
#pragma unroll
                            for (int u = 0; u < 8; u++) s += t8[u];
                        }
                    }
                    grp[row * NGP + gi] = s;
                }
            }
            __syncthreads();
            for (int e = tid; e < g * per_row; e += NT) {
                static_assert(WR_GRP == 2, "a group holds at most two lattice rows");
                const int gg = e >= per_row ? 1 : 0, pi = e - gg * per_row;
                long long sum[5];
#pragma unroll
                for (int t = 0; t < 5; t++) {
                    const long long *v = grp + (gg * 5 + t) * NGP + pi;
                    long long part[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) part[c] = c < R ? v[c] : 0;
                    sum[t] = ((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7]));
                }
                const float FLT_SCALE = 1.f / (1 << 20);
                const float A11 = (float)sum[0] * FLT_SCALE, A12 = (float)sum[1] * FLT_SCALE, A22 = (float)sum[2] * FLT_SCALE;
                float D;
                const float Dinv = lk_min_eig_ok(A11, A12, A22, 40, q.min_eig, D) ? D : 0.f;
                const int kx = kx0 + (i0 + pi) * ncx, ky = ky0 + (jg + gg) * ncx;
                LkLevelRec *o = q.wsum + (((size_t)(q.pair0 + b) * q.g.nlev + level) * q.P) + (size_t)kx * q.gy + ky;
                *reinterpret_cast<float4 *>(o) = make_float4(A11, A12, A22, Dinv);
                *reinterpret_cast<longlong2 *>(&o->C1) = make_longlong2(sum[3], sum[4]);
            }
            jg += g; g = 0;
        }
        if (k == K - 1 || !col_on) continue;
        // the `step` rows down to the next boundary, WR_BATCH at a time; the batch in registers was requested before the boundary work
        for (int r0 = 0; r0 < step; r0 += WR_BATCH) {
            compute_batch(min(WR_BATCH, step - r0));
            const int left = step - r0 - WR_BATCH;
            if (left > 0) load_batch(min(WR_BATCH, left));
            else if (k + 1 < K - 1) load_batch(nr0);
        }
    }
}

// ---- LK on phase planes ------------------------------------------------------------------------------------------------
struct PhTile {
    static constexpr int WIN = 40, LXN = 4, LYN = 8;
    static constexpr int TW = WIN / LXN, TH = WIN / LYN, NP = TW / 2;
    static constexpr int PW = MD_PH_BOX_W / 2;        // phase tile pitch in words (24)
    static constexpr int PLANE_WORDS = PW * WIN;      // 960
    static constexpr int JP = MD_LK_J_BOX_W / 4;      // J tile pitch in words (20)
    static constexpr int JROWS = WIN + 1 + 2 * MD_LK_J_MARGIN_Y;
    static constexpr int P_BYTES = 2 * PLANE_WORDS * 4, J_BYTES = MD_LK_J_BOX_W * JROWS;      // Ix, Iy planes
    static constexpr int P_OFF = 0;
    static constexpr int J_OFF = (P_BYTES + 127) / 128 * 128;
    static constexpr int BAR_OFF = J_OFF + (J_BYTES + 127) / 128 * 128;
    static constexpr int WARP_BYTES = BAR_OFF + 128;
    static constexpr int JX_MAX = MD_LK_J_BOX_W - (WIN + 1) - 3;
};

// register fill of a lane's 10 x 5 block of the Ix / Iy window from the staged phase tile (rows unrolled by recursion: the row
// offsets are immediates of the shared-memory loads)
template <int R, bool ODD>
__device__ __forceinline__ void lkp_fill_rows(uint32_t aX, int (&Xpk)[PhTile::TH][PhTile::NP], int (&Ypk)[PhTile::TH][PhTile::NP])
{
    using T = PhTile;
    constexpr int NP = T::NP, NW = ODD ? NP + 1 : NP, YO = T::PLANE_WORDS * 4;
    static_assert(NP == 5, "five tap pairs per row and plane");
    uint32_t wx[NP + 1], wy[NP + 1];
    wx[0] = lds_u32<(R * T::PW + 0) * 4>(aX); wx[1] = lds_u32<(R * T::PW + 1) * 4>(aX); wx[2] = lds_u32<(R * T::PW + 2) * 4>(aX);
    wx[3] = lds_u32<(R * T::PW + 3) * 4>(aX); wx[4] = lds_u32<(R * T::PW + 4) * 4>(aX);
    wy[0] = lds_u32<YO + (R * T::PW + 0) * 4>(aX); wy[1] = lds_u32<YO + (R * T::PW + 1) * 4>(aX); wy[2] = lds_u32<YO + (R * T::PW + 2) * 4>(aX);
    wy[3] = lds_u32<YO + (R * T::PW + 3) * 4>(aX); wy[4] = lds_u32<YO + (R * T::PW + 4) * 4>(aX);
    if constexpr (NW > NP) { wx[5] = lds_u32<(R * T::PW + 5) * 4>(aX); wy[5] = lds_u32<YO + (R * T::PW + 5) * 4>(aX); }
#pragma unroll
    for (int i = 0; i < NP; i++) {
        if constexpr (ODD) {       // the window starts on the odd element of a word: pairs straddle two words
            Xpk[R][i] = (int)__byte_perm(wx[i], wx[i + 1], 0x5432);
            Ypk[R][i] = (int)__byte_perm(wy[i], wy[i + 1], 0x5432);
        } else { Xpk[R][i] = (int)wx[i]; Ypk[R][i] = (int)wy[i]; }
    }
    if constexpr (R + 1 < T::TH) lkp_fill_rows<R + 1, ODD>(aX, Xpk, Ypk);
}
// the tap rows of one LK iteration
template <int R>
__device__ __forceinline__ void lkp_iter_rows(uint32_t rowa, int sh, const RowWords &r0, int wtop, int wbot, const int (&Xpk)[PhTile::TH][PhTile::NP],
                                              const int (&Ypk)[PhTile::TH][PhTile::NP], int &b1lo, int &b1hi, int &b2lo, int &b2hi)
{
    using T = PhTile;
    const RowWords r1 = load_row_a<(R + 1) * T::JP * 4>(rowa, sh);
    TapLoop<0, T::TW>::template iter2<T::NP>(r0, r1, wtop, wbot, Xpk[R], Ypk[R], b1lo, b1hi, b2lo, b2hi);
    if constexpr (R + 1 < T::TH) lkp_iter_rows<R + 1>(rowa, sh, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
}

// exactly one lane of the (converged) warp
__device__ __forceinline__ bool lkp_elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// One warp per CTA; a CTA tracks NPTS consecutive grid points (one grid column segment: neighbouring windows, shared L2 lines) one
// after the other, so the barrier setup, the kernel parameters and the first window request are paid once per NPTS points and
// the top-level window of point k + 1 is in flight while point k iterates on level 0.  One-warp CTAs keep the scheduling dynamic:
// a CTA's shared memory goes back to the SM as soon as its own points are done.
template <int NPTS>
__global__ void __launch_bounds__(32, 18) k_lk_phase(const LkParams p, const __grid_constant__ LkPhaseMaps maps)
{
    using T = PhTile;
    constexpr int WIN = T::WIN, TW = T::TW, TH = T::TH, NP = T::NP;
    extern __shared__ __align__(1024) uint8_t lkp_sm_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.y;
    const int k0 = blockIdx.x * NPTS, k1 = min(p.P, k0 + NPTS);
    uint64_t *barP = reinterpret_cast<uint64_t *>(lkp_sm_raw + T::BAR_OFF), *barJ = barP + 1;
    const int lx = lane & 3, ly = lane >> 2;
    if (lane == 0) {
        mbar_init(barP, 1);
        mbar_init(barJ, 1);
        mbar_fence_init();
    }
    __syncwarp();
    // shared-window addresses of the tiles and barriers, computed once and kept opaque (the compiler otherwise re-derives the window
    // base from SR_CgaCtaId at every use)
    uint32_t sm0 = smem_u32(lkp_sm_raw);
    asm volatile("" : "+r"(sm0));
    const uint32_t aP = sm0 + T::P_OFF, aJ = sm0 + T::J_OFF, aBarP = sm0 + T::BAR_OFF, aBarJ = aBarP + 8;
    uint32_t phP = 0, phJ = 0;
    const int top = p.g.nlev - 1;
    const int slotJ = (p.next_slot0 + b) % p.g.nslots;
    const float half = (WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    unsigned n_iters = 0, n_levels = 0;         // measurement: work actually done by this CTA's points

    // request the window (Ix, Iy planes of the point's phase class) of grid point kk at `level`; nothing is requested (and nothing
    // will be waited for) when the window lies outside the level
    auto issue_P = [&](int gxi, int gyi, int level) {
        const float scale = lk_level_scale(level);
        const float ppx = __fsub_rn((float)gxi * scale, half), ppy = __fsub_rn((float)gyi * scale, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -WIN || ipx >= p.g.lv[level].w || ipy < -WIN || ipy >= p.g.lv[level].h) return;
        const int msk = (1 << level) - 1, sh = p.pg.lv[level].shift;
        const int cls = ((gyi & msk) >> sh) * p.pg.lv[level].ncx + ((gxi & msk) >> sh);
        if (lkp_elect_one()) {
            mbar_expect_tx_a(aBarP, T::P_BYTES);
            tma_load_5d_a(aP, &maps.ph[level], (ipx + MD_PH_MARGIN) & ~7, ipy + MD_PH_MARGIN, 1, cls, p.ph_pair0 + b, aBarP);
        }
    };
    // the window after (k, level) in processing order: the next finer level of k, or the top level of k + 1
    // the grid point (cpp:56-64: x outer, y inner) and its successor in the CTA, stepped without divisions
    int kx = k0 / p.gy, ky = k0 - kx * p.gy;
    int gxi = p.ps * kx, gyi = p.ps * ky, gxn, gyn;
    // Which levels of a point passed the minimum-eigenvalue test (Dinv != 0 in its level records, k_window_sums): lane l looks at
    // level l, one ballot.  Tracking of a point starts at its HIGHEST usable level: the failed levels above it (half of the
    // coarsest-level windows of the bench sequence fail) cost neither tiles nor a register fill; a failed level further down is
    // handled like before (found after its fill).  The mask of point k + 1 is fetched while point k is tracked.
    const LkLevelRec *rec_top = p.wsum + ((size_t)(p.ph_pair0 + b) * p.g.nlev + top) * p.P;      // level `top`, point 0
    auto load_D = [&](int kk) { return lane <= top ? __ldg(&(rec_top - (size_t)(top - lane) * p.P + kk)->Dinv) : 0.f; };
    int first_cur = -1, first_next = -1;        // highest usable level of the current / the next point (-1: none)
    // the window after (k, level) in processing order: the next finer level of k, or the first usable level of k + 1
    auto issue_next_P = [&](int k, int level) {
        if (level > 0) issue_P(gxi, gyi, level - 1);
        else if (k + 1 < k1 && first_next >= 0) issue_P(gxn, gyn, first_next);
    };
    if (k0 < k1) {
        first_next = 31 - __clz(__ballot_sync(0xffffffffu, load_D(k0) != 0.f));
        if (first_next >= 0) issue_P(gxi, gyi, first_next);
    }

    for (int k = k0; k < k1; k++, gxi = gxn, gyi = gyn) {
        if (++ky == p.gy) { ky = 0; gxn = gxi + p.ps; gyn = 0; }
        else { gxn = gxi; gyn = gyi + p.ps; }
        const float2 pt = make_float2((float)gxi, (float)gyi);
        float2 nxt = make_float2(0.f, 0.f);
        int st = 1;
        const LkLevelRec *rec = rec_top + k;      // the level scalars (k_window_sums), walked down with the levels
        first_cur = first_next;
        first_next = -1;
        const float d_next = k + 1 < k1 ? load_D(k + 1) : 0.f;      // lands long before level 0 looks at it
        if (first_cur < 0) {                    // no usable level at all: nothing will be requested for this point, keep the chain going
            first_next = 31 - __clz(__ballot_sync(0xffffffffu, d_next != 0.f));
            issue_next_P(k, 0);
        }

        for (int level = top; level >= 0; level--, rec -= p.P) {
            const int Lw = p.g.lv[level].w, Lh = p.g.lv[level].h;
            const float scale = lk_level_scale(level);
            float ppx = pt.x * scale, ppy = pt.y * scale;
            LkIterState s;
            if (level == top) { s.npx = ppx; s.npy = ppy; }
            else { s.npx = nxt.x * 2.f; s.npy = nxt.y * 2.f; }
            nxt = make_float2(s.npx, s.npy);
            if (level > first_cur) {            // failed level above the first usable one (or no usable level): stepped over
                if (level == 0) st = 0;
                continue;
            }
            if (level == 0) first_next = 31 - __clz(__ballot_sync(0xffffffffu, d_next != 0.f));
            ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
            const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
            if (ipx < -WIN || ipx >= Lw || ipy < -WIN || ipy >= Lh) {
                if (level == 0) st = 0;
                issue_next_P(k, level);
                continue;
            }
            // the level scalars land while the window is copied into registers
            const float4 A = __ldg(reinterpret_cast<const float4 *>(rec));
            const longlong2 C = __ldg(reinterpret_cast<const longlong2 *>(&rec->C1));
            // ---- request the next-frame tile around the initial guess; it lands while the window is copied ------------------
            s.npx = __fsub_rn(s.npx, half); s.npy = __fsub_rn(s.npy, half);
            int tx0 = 0, ty0 = 0;
            bool j_ok = false;
            {
                const int inx = __float2int_rd(s.npx), iny = __float2int_rd(s.npy);
                if (!(inx < -WIN || inx >= Lw || iny < -WIN || iny >= Lh)) {
                    tx0 = ((inx - MD_LK_J_MARGIN_X + p.g.padx) & ~15) - p.g.padx; ty0 = iny - MD_LK_J_MARGIN_Y;
                    j_ok = true;
                    __syncwarp();      // every lane is done with the J tile of the previous level
                    if (lkp_elect_one()) {
                        mbar_expect_tx_a(aBarJ, T::J_BYTES);
                        tma_load_3d_a(aJ, &maps.imgJ[level], tx0 + p.g.padx, ty0 + p.g.pady, slotJ, aBarJ);
                    }
                }
            }

            // ---- derivative window from the phase planes into registers ----------------------------------------------------
            mbar_wait_a(aBarP, phP); phP ^= 1;
            int Xpk[TH][NP], Ypk[TH][NP];
            {
                const int e0 = ((ipx + MD_PH_MARGIN) & 7) + TW * lx;       // first tap column of this lane inside the box
                const uint32_t aX = aP + ((TH * ly) * T::PW + (e0 >> 1)) * 4;
                if ((ipx + MD_PH_MARGIN) & 1) lkp_fill_rows<0, true>(aX, Xpk, Ypk);       // warp-uniform: 10 lx is even
                else lkp_fill_rows<0, false>(aX, Xpk, Ypk);
            }
            __syncwarp();                       // the phase tile is consumed: prefetch the next window
            issue_next_P(k, level);
            if (j_ok) { mbar_wait_a(aBarJ, phJ); phJ ^= 1; }

            const float A11 = A.x, A12 = A.y, A22 = A.z, D = A.w;
            const long long C1 = C.x, C2 = C.y;
            if (D == 0.f) {                     // minimum-eigenvalue / determinant test failed (k_window_sums)
                if (level == 0) st = 0;
                continue;
            }
            s.pdx = 0.f; s.pdy = 0.f;
            n_levels++;
            for (int j = 0; j < p.max_iters; j++) {
                const int inx = __float2int_rd(s.npx), iny = __float2int_rd(s.npy);
                if (inx < -WIN || inx >= Lw || iny < -WIN || iny >= Lh) {
                    if (level == 0) st = 0;
                    break;
                }
                int w00, w01, w10, w11;
                lk_weights(__fsub_rn(s.npx, (float)inx), __fsub_rn(s.npy, (float)iny), w00, w01, w10, w11);
                n_iters++;
                // ---- restage the J tile when the window drifts out of it (warp-uniform, rare) -------------------------------
                if (inx < tx0 || inx - tx0 > T::JX_MAX || iny < ty0 || iny - ty0 > 2 * MD_LK_J_MARGIN_Y) {
                    tx0 = ((inx - MD_LK_J_MARGIN_X + p.g.padx) & ~15) - p.g.padx; ty0 = iny - MD_LK_J_MARGIN_Y;
                    __syncwarp();
                    if (lkp_elect_one()) {
                        mbar_expect_tx_a(aBarJ, T::J_BYTES);
                        tma_load_3d_a(aJ, &maps.imgJ[level], tx0 + p.g.padx, ty0 + p.g.pady, slotJ, aBarJ);
                    }
                    mbar_wait_a(aBarJ, phJ); phJ ^= 1;
                }
                const int c0 = (inx - tx0) + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
                const uint32_t rowa = aJ + (((iny - ty0) + TH * ly) * T::JP + wb) * 4;
                const int wtop = (int)__byte_perm((uint32_t)w00, (uint32_t)w01, 0x5410);
                const int wbot = (int)__byte_perm((uint32_t)w10, (uint32_t)w11, 0x5410);
                int b1lo = 0, b1hi = 0, b2lo = 0, b2hi = 0;
                lkp_iter_rows<0>(rowa, sh, load_row_a<0>(rowa, sh), wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
                const int b1 = iter2_total(b1lo, b1hi), b2 = iter2_total(b2lo, b2hi);
                // sum (J - I) Ix = sum J Ix - sum I Ix, exactly, in 64-bit integers; one rounding to f32
                const float fb1 = (float)(warp_sum_exact_i64(b1) - C1) * FLT_SCALE;
                const float fb2 = (float)(warp_sum_exact_i64(b2) - C2) * FLT_SCALE;
                if (lk_update(A11, A12, A22, D, fb1, fb2, half, j, p.eps2, s, nxt)) break;
            }
        }
        if (lane == 0) {
            p.next[(size_t)b * p.P + k] = nxt;
            p.status[(size_t)b * p.P + k] = (uint8_t)st;
        }
    }
    // 64 striped counter pairs: one hot address would serialise the atomics of every finishing warp
    if (lane == 0 && p.stat_iters) {
        unsigned long long *c = p.stat_iters + 2 * (blockIdx.x & 63);
        atomicAdd(c, (unsigned long long)n_iters); atomicAdd(c + 1, (unsigned long long)n_levels);
    }
}

// the phase planes of the PREVIOUS frames of `pairs` pairs (k_lk_phase reads them); may run on another stream than LK
// planes + window sums of pyramid levels l0..l1 (the levels are independent of each other)
cudaError_t launch_lk_planes_levels(const LkParams &p, int pairs, int l0, int l1, cudaStream_t s)
{
    PhParams q;
    q.g = p.g; q.pg = p.pg; q.img = p.img; q.der = p.der; q.ph = p.ph; q.prev_slot0 = p.prev_slot0; q.pair0 = p.ph_pair0;
    q.wsum = p.wsum; q.P = p.P; q.ps = p.ps; q.gy = p.gy; q.gx = p.P / p.gy;
    q.half = (p.win - 1) * 0.5f; q.min_eig = p.min_eig;
    for (int l = l0; l <= l1 && l < p.g.nlev; l++) {
        const PhaseLevel &PL = p.pg.lv[l];
        const int step = p.ps >> PL.shift;
        if (step > WS_COLS - 42) return cudaErrorInvalidConfiguration;          // one point per chunk must fit (pixel_step <= 470)
        const int nx = (q.gx + PL.ncx - 1) / PL.ncx, ny = (q.gy + PL.ncx - 1) / PL.ncx;   // upper bounds per class
        // MD_WS_MODE: 2 (default) = planes and sums in one pass (k_window_sums_ring<true>), 1 = k_phase_planes + single-pass sums,
        // 0 = k_phase_planes + sliding sums; lattice steps the ring kernel does not take (40 % step != 0 or step < 5) always get 0
        static const int ws_mode = [] { const char *e = getenv("MD_WS_MODE"); return e ? atoi(e) : 2; }();
        // plane rows per CTA segment: neighbouring segments recompute the 40 - step rows their windows share, shorter segments fill the
        // GPU better.  Measured (frames/s): 16 pairs per launch 80 / 120 / 160 / 240 / 320 rows: 5 724 / 5 754 / 5 780 / 5 740 / 5 718;
        // 32 pairs per launch 160 / 240 / 320: 5 966 / 5 969 / 5 993
        static const int ws_rows_env = [] { const char *e = getenv("MD_WS_ROWS"); return e && atoi(e) > 0 ? atoi(e) : 0; }();
        const int ws_rows = ws_rows_env > 0 ? ws_rows_env : (pairs >= 24 ? 320 : 160);
        // CTAs per SM the 256-thread build is compiled for: 2 (default, 128 registers) or 3 (80 registers: 134 instead of ~100 instructions per
        // row visit, measured 5 839 against 5 867 frames/s)
        static const int ws_minb = [] { const char *e = getenv("MD_WS_MINB"); return e ? atoi(e) : 2; }();
        static const int ws_nt = [] { const char *e = getenv("MD_WS_NT"); const int v = e ? atoi(e) : 0; return v == 256 || v == 128 || v == 96 || v == 64 ? v : 0; }();
        const bool ring = ws_mode > 0 && step >= 5 && step <= 40 && 40 % step == 0;
        if (!ring || ws_mode == 1) {
            const int threads = (PL.pitch / 8) * ((PL.h + 3) / 4);
            dim3 grid((threads + 255) / 256, PL.ncx * PL.ncx, pairs);
            k_phase_planes<<<grid, 256, 0, s>>>(q, l);
            MD_COUNT_LAUNCH(1);
        }
        if (ring) {
            const int R = 40 / step, seg = max(2, ws_rows / step);
            // threads per CTA (2 columns each): the fewest thread slots over the chunks of a lattice row, the wider CTA on a tie
            int NT = 128, cpr = 0, ncx_r = 0;
            {
                const int cand[4] = {256, 128, 96, 64};
                long best = -1;
                for (int c = 0; c < 4; c++) {
                    if (ws_mode == 1 && cand[c] != 128) continue;
                    const int cpmax = (2 * cand[c] - 41) / step + 1;
                    if (cpmax < 1) continue;
                    const int nch = (nx + cpmax - 1) / cpmax, cpe = (nx + nch - 1) / nch;
                    const long cost = (long)nch * cand[c];
                    if (best < 0 || cost < best) { best = cost; NT = cand[c]; cpr = cpe; ncx_r = nch; }
                }
                if (ws_nt > 0 && ws_mode != 1) { NT = ws_nt; const int cpmax = (2 * NT - 41) / step + 1; ncx_r = (nx + cpmax - 1) / cpmax; cpr = (nx + ncx_r - 1) / ncx_r; }
            }
            const int nseg_r = (ny + seg - 1) / seg;
            const size_t smem = wr_smem_bytes(NT, R);
            const dim3 grid(ncx_r * nseg_r, PL.ncx * PL.ncx, pairs);
            // (the attribute is per device: set at every launch, like launch_lk_phase does, so that a process with contexts on several
            // GPUs never launches with the 48 KB default)
            auto go = [&](auto kern, int nt) {
                cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) return e;
                kern<<<grid, nt, smem, s>>>(q, l, cpr, ncx_r, seg);
                return cudaSuccess;
            };
            cudaError_t e;
            if (ws_mode == 1) e = go(k_window_sums_ring<false, 128>, 128);
            else if (NT == 256 && ws_minb == 3) e = go(k_window_sums_ring<true, 256, 3>, 256);
            else if (NT == 256) e = go(k_window_sums_ring<true, 256>, 256);
            else if (NT == 128) e = go(k_window_sums_ring<true, 128>, 128);
            else if (NT == 96) e = go(k_window_sums_ring<true, 96>, 96);
            else e = go(k_window_sums_ring<true, 64>, 64);
            if (e != cudaSuccess) return e;
            MD_COUNT_LAUNCH(1);
            continue;
        }
        const int cp = (WS_COLS - 42) / step + 1;                                // lattice columns per CTA
        const int nchunk_x = (nx + cp - 1) / cp, nseg = (ny + WS_SEG - 1) / WS_SEG;
        k_window_sums<<<dim3(nchunk_x * nseg, PL.ncx * PL.ncx, pairs), 256, 0, s>>>(q, l, cp, nchunk_x);
        MD_COUNT_LAUNCH(1);
    }
    return cudaGetLastError();
}

cudaError_t launch_lk_planes(const LkParams &p, int pairs, cudaStream_t s) { return launch_lk_planes_levels(p, pairs, 0, p.g.nlev - 1, s); }

cudaError_t launch_lk_phase(const LkParams &p, const LkPhaseMaps *maps, int pairs, cudaStream_t s)
{
    if (!p.ph_ready) {
        cudaError_t e0 = launch_lk_planes(p, pairs, s);
        if (e0 != cudaSuccess) return e0;
    }
    static const int npts_env = [] { const char *e = getenv("MD_LK_NPTS"); return e ? atoi(e) : 0; }();      // read once (thread-safe init)
    auto go = [&](auto kern, int NPTS) {
        const size_t smem = (size_t)PhTile::WARP_BYTES + 128;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dim3 grid((p.P + NPTS - 1) / NPTS, pairs);
        kern<<<grid, 32, smem, s>>>(p, *maps);
        MD_COUNT_LAUNCH(1);
        return cudaGetLastError();
    };
    // points per one-warp CTA (MD_LK_NPTS = 1 | 2 | 8 | 16 selects the other builds: tuning aid).  Measured on the default bench,
    // value / e2e frames/s: 1: 5 397 / 4 696, 4: 5 480 / 4 773, 8: 5 437 / 4 725, 16: 5 384 (longer CTAs leave a longer tail per launch,
    // and the host-buffer pipeline launches six LK kernels per batch)
    if (npts_env == 1) return go(k_lk_phase<1>, 1);
    if (npts_env == 2) return go(k_lk_phase<2>, 2);
    if (npts_env == 8) return go(k_lk_phase<8>, 8);
    if (npts_env == 16) return go(k_lk_phase<16>, 16);
    return go(k_lk_phase<4>, 4);
}
