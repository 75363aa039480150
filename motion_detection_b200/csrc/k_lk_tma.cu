// k_lk_tma.cu -- K2 production kernel: pyramidal Lucas-Kanade for the reference's 40x40 window on sm_100a.
//
// Replaces cv::calcOpticalFlowPyrLK(...) as called at common/src/optical_flow_calculator.cpp:71,172 (see lk_common.cuh).
//
//  * one warp per tracked point, all pyramid levels in one launch (a point's level-l result feeds only itself);
//  * every patch is staged by TMA (cp.async.bulk.tensor.3d on per-level tensor maps over the padded pyramid planes,
//    dims = x, y, slot) into per-warp shared-memory tiles, completion on per-warp mbarriers -- no index arithmetic,
//    no register round trip.  The next-frame (J) tile is requested at the start of a level and lands while the
//    window is being built; the previous-frame (I + Scharr) tiles of the NEXT level are requested as soon as the
//    window of the current level sits in registers and land during the iteration loop;
//  * lane (lx, ly) of the 4 x 8 lane grid owns a 10 x 5 block of the window's taps; the derivative window samples
//    Ix, Iy stay in registers for the whole iteration loop;
//  * the mismatch sums are split algebraically: sum (J - I) Ix = sum J Ix - sum I Ix; the second term is constant per
//    level, so the iteration loop never touches I;
//  * bilinear samples use dp2a (two 14-bit weights x two u8 pixels per instruction) on row words kept packed in
//    registers (an aligned and a one-byte-shifted copy): no byte is ever extracted;
//  * structure-tensor / mismatch sums: exact int32 per lane, f64 warp-shuffle butterfly, one rounding to f32.
// Arithmetic is identical to k_lk (k_lk.cu); tap order differs only inside exact integer sums.
#include <stdlib.h>

#include "lk_tile.cuh"
#include "tma.h"

template <int WIN>
struct LkTile {
    static constexpr int LXN = 4, LYN = 8;
    static constexpr int TW = WIN / LXN, TH = WIN / LYN, NP = TW / 2;
    static constexpr int IP = MD_LK_I_BOX_W / 4;      // I tile pitch in words (12)
    static constexpr int JP = MD_LK_J_BOX_W / 4;      // J tile pitch in words (20)
    static constexpr int DP = MD_LK_D_BOX_W;          // derivative tile pitch in words (44)
    static constexpr int IROWS = WIN + 1, JROWS = WIN + 1 + 2 * MD_LK_J_MARGIN_Y;
    static constexpr int I_BYTES = MD_LK_I_BOX_W * IROWS, J_BYTES = MD_LK_J_BOX_W * JROWS, D_BYTES = DP * 4 * IROWS;
    static constexpr int I_OFF = 0;
    static constexpr int J_OFF = (I_BYTES + 127) / 128 * 128;
    static constexpr int D_OFF = J_OFF + (J_BYTES + 127) / 128 * 128;
    static constexpr int BAR_OFF = D_OFF + (D_BYTES + 127) / 128 * 128;
    static constexpr int WARP_BYTES = BAR_OFF + 128;
    static constexpr int JX_MAX = MD_LK_J_BOX_W - (WIN + 1) - 3;     // window + the word load_row over-reads must fit
    static_assert(WIN == 40, "tile geometry in md_internal.h is laid out for the 40x40 window");
};

template <int WIN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 16 / WARPS) k_lk_tma(const LkParams p, const __grid_constant__ LkTmaMaps maps)
{
    using T = LkTile<WIN>;
    constexpr int TW = T::TW, TH = T::TH, NP = T::NP;
    extern __shared__ __align__(1024) uint8_t lk_sm_raw[];     // dynamic smem starts 1 KB aligned; keep shared-space pointers
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * WARPS + warp;
    const int b = blockIdx.y;
    if (k >= p.P) return;
    uint8_t *wbase = lk_sm_raw + (size_t)warp * T::WARP_BYTES;
    uint32_t *tI = reinterpret_cast<uint32_t *>(wbase + T::I_OFF);
    uint32_t *tJ = reinterpret_cast<uint32_t *>(wbase + T::J_OFF);
    uint32_t *tD = reinterpret_cast<uint32_t *>(wbase + T::D_OFF);
    uint64_t *barI = reinterpret_cast<uint64_t *>(wbase + T::BAR_OFF), *barJ = barI + 1;
    const int lx = lane & 3, ly = lane >> 2;
    if (lane == 0) {
        mbar_init(barI, 1);
        mbar_init(barJ, 1);
        mbar_fence_init();
    }
    __syncwarp();
    uint32_t phI = 0, phJ = 0;

    float2 pt;
    if (p.pts_in) pt = p.pts_in[(size_t)b * p.P + k];
    else pt = make_float2((float)(p.ps * (k / p.gy)), (float)(p.ps * (k % p.gy)));
    const int slotI = (p.prev_slot0 + b) % p.g.nslots, slotJ = (p.next_slot0 + b) % p.g.nslots;
    const float half = (WIN - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    float2 nxt = make_float2(0.f, 0.f);
    int st = 1;
    int n_iters = 0, n_levels = 0;              // measurement: work actually done for this point

    // request the previous-frame patch + Scharr planes of `level` (no-op when the window is out of range there)
    auto issue_I = [&](int level) {
        const float scale = lk_level_scale(level);
        const float ppx = __fsub_rn(pt.x * scale, half), ppy = __fsub_rn(pt.y * scale, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -WIN || ipx >= p.g.lv[level].w || ipy < -WIN || ipy >= p.g.lv[level].h) return;
        if (lane == 0) {
            // TMA box origins must be 16-byte aligned in x: the sub-offset is applied when the rows are read
            const int gx = ipx + p.g.padx;
            mbar_expect_tx(barI, T::I_BYTES + T::D_BYTES);
            tma_load_3d(tI, &maps.imgI[level], gx & ~15, ipy + p.g.pady, slotI, barI);
            tma_load_3d(tD, &maps.der[level], gx & ~3, ipy + p.g.pady, slotI, barI);
        }
    };
    issue_I(p.g.nlev - 1);

    for (int level = p.g.nlev - 1; level >= 0; level--) {
        const int Lw = p.g.lv[level].w, Lh = p.g.lv[level].h;
        const float scale = lk_level_scale(level);
        float ppx = pt.x * scale, ppy = pt.y * scale;
        LkIterState s;
        if (level == p.g.nlev - 1) { s.npx = ppx; s.npy = ppy; }
        else { s.npx = nxt.x * 2.f; s.npy = nxt.y * 2.f; }
        nxt = make_float2(s.npx, s.npy);
        ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
        const int ipx = __float2int_rd(ppx), ipy = __float2int_rd(ppy);
        if (ipx < -WIN || ipx >= Lw || ipy < -WIN || ipy >= Lh) {
            if (level == 0) st = 0;
            else issue_I(level - 1);
            continue;
        }
        // ---- request the next-frame tile around the initial guess; it lands while the window is built -------------
        s.npx = __fsub_rn(s.npx, half); s.npy = __fsub_rn(s.npy, half);
        int tx0 = 0, ty0 = 0;
        bool j_ok = false;
        {
            const int inx = __float2int_rd(s.npx), iny = __float2int_rd(s.npy);
            if (!(inx < -WIN || inx >= Lw || iny < -WIN || iny >= Lh)) {
                tx0 = ((inx - MD_LK_J_MARGIN_X + p.g.padx) & ~15) - p.g.padx; ty0 = iny - MD_LK_J_MARGIN_Y;
                j_ok = true;
                __syncwarp();      // every lane is done with the J tile of the previous level
                if (lane == 0) {
                    mbar_expect_tx(barJ, T::J_BYTES);
                    tma_load_3d(tJ, &maps.imgJ[level], tx0 + p.g.padx, ty0 + p.g.pady, slotJ, barJ);
                }
            }
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);

        // ---- window samples into registers, structure tensor, constant part of the mismatch ---------------------
        mbar_wait(barI, phI); phI ^= 1;
        int Xpk[TH][NP], Ypk[TH][NP];
        int a11 = 0, a12 = 0, a22 = 0, c1 = 0, c2 = 0;
        {
            const int gx = ipx + p.g.padx;
            const int c0 = (gx & 15) + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
            const int wtop = (int)__byte_perm((uint32_t)w00, (uint32_t)w01, 0x5410);
            const int wbot = (int)__byte_perm((uint32_t)w10, (uint32_t)w11, 0x5410);
            RowWords r0 = load_row(tI + (TH * ly) * T::IP, wb, sh);
            // the row loop stays ROLLED (the unrolled build was 26 KB of SASS and thrashed the instruction cache)
#pragma unroll 1
            for (int r = 0; r < TH; r++) {
                const RowWords r1 = load_row(tI + (TH * ly + r + 1) * T::IP, wb, sh);
                const uint32_t *d0 = tD + (TH * ly + r) * T::DP + (gx & 3) + TW * lx, *d1 = d0 + T::DP;
                int xprev = 0, yprev = 0;
                int Xr[NP], Yr[NP];
                TapLoop<0, TW>::template build<NP>(r0, r1, d0, d1, wtop, wbot, w00, w01, w10, w11, Xr, Yr, a11, a12, a22, c1, c2,
                                                   xprev, yprev);
                RowStore<0, TH, NP>::put(r, Xpk, Ypk, Xr, Yr);
                r0 = r1;
            }
        }
        __syncwarp();                       // the I / Scharr tiles are consumed: prefetch the next level's
        if (level > 0) issue_I(level - 1);
        if (j_ok) { mbar_wait(barJ, phJ); phJ ^= 1; }

        const float A11 = warp_sum_exact_f32(a11) * FLT_SCALE;
        const float A12 = warp_sum_exact_f32(a12) * FLT_SCALE;
        const float A22 = warp_sum_exact_f32(a22) * FLT_SCALE;
        float D;
        if (!lk_min_eig_ok(A11, A12, A22, WIN, p.min_eig, D)) {
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
            lk_weights(__fsub_rn(s.npx, (float)inx), __fsub_rn(s.npy, (float)iny), w00, w01, w10, w11);
            n_iters++;
            // ---- restage the J tile when the window drifts out of it (warp-uniform, rare) -----------------------------
            if (inx < tx0 || inx - tx0 > T::JX_MAX || iny < ty0 || iny - ty0 > 2 * MD_LK_J_MARGIN_Y) {
                tx0 = ((inx - MD_LK_J_MARGIN_X + p.g.padx) & ~15) - p.g.padx; ty0 = iny - MD_LK_J_MARGIN_Y;
                __syncwarp();
                if (lane == 0) {
                    mbar_expect_tx(barJ, T::J_BYTES);
                    tma_load_3d(tJ, &maps.imgJ[level], tx0 + p.g.padx, ty0 + p.g.pady, slotJ, barJ);
                }
                mbar_wait(barJ, phJ); phJ ^= 1;
            }
            const int c0 = (inx - tx0) + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
            const uint32_t *rowp = tJ + ((iny - ty0) + TH * ly) * T::JP;
            const int wtop = (int)__byte_perm((uint32_t)w00, (uint32_t)w01, 0x5410);
            const int wbot = (int)__byte_perm((uint32_t)w10, (uint32_t)w11, 0x5410);
            int b1lo = 0, b1hi = 0, b2lo = 0, b2hi = 0;
            RowWords r0 = load_row(rowp, wb, sh);
#pragma unroll
            for (int r = 0; r < TH; r++) {
                const RowWords r1 = load_row(rowp + (r + 1) * T::JP, wb, sh);
                TapLoop<0, TW>::template iter2<NP>(r0, r1, wtop, wbot, Xpk[r], Ypk[r], b1lo, b1hi, b2lo, b2hi);
                r0 = r1;
            }
            const int b1 = iter2_total(b1lo, b1hi) - c1, b2 = iter2_total(b2lo, b2hi) - c2;      // sum (J - I) Ix of this lane's taps
            const float fb1 = warp_sum_exact_f32(b1) * FLT_SCALE;
            const float fb2 = warp_sum_exact_f32(b2) * FLT_SCALE;
            if (lk_update(A11, A12, A22, D, fb1, fb2, half, j, p.eps2, s, nxt)) break;
        }
    }
    if (lane == 0) {
        p.next[(size_t)b * p.P + k] = nxt;
        p.status[(size_t)b * p.P + k] = (uint8_t)st;
        // 64 striped counter pairs: one hot address would serialise the atomics of every finishing warp
        if (p.stat_iters) { unsigned long long *c = p.stat_iters + 2 * (k & 63); atomicAdd(c, (unsigned long long)n_iters); atomicAdd(c + 1, (unsigned long long)n_levels); }
    }
}

cudaError_t launch_lk_tma(const LkParams &p, const void *maps, int pairs, cudaStream_t s)
{
    static int warps_env = -1;
    if (warps_env < 0) { const char *e = getenv("MD_LKT_WARPS"); warps_env = e ? atoi(e) : 0; }
    auto go = [&](auto kern, int WARPS) {
        const size_t smem = (size_t)WARPS * LkTile<40>::WARP_BYTES + 128;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        dim3 grid((p.P + WARPS - 1) / WARPS, pairs);
        kern<<<grid, WARPS * 32, smem, s>>>(p, *reinterpret_cast<const LkTmaMaps *>(maps));
        MD_COUNT_LAUNCH(1);
        return cudaGetLastError();
    };
    // one warp per CTA, like k_lk_phase (live path: 552 callbacks/s with 8 warps per CTA, 578 with 1); MD_LKT_WARPS=8: tuning aid
    if (warps_env == 8) return go(k_lk_tma<40, 8>, 8);
    return go(k_lk_tma<40, 1>, 1);
}
