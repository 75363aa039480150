// k_lk.cu -- K2: pyramidal Lucas-Kanade, generic window size (the reference's 40x40 window runs k_lk_tma.cu).
//
// Replaces cv::calcOpticalFlowPyrLK(pyr, gray2, pts1, pts2, status, err, Size(win,win), 5,
// TermCriteria(COUNT|EPS, 10, 0.03), 0, 0.001) as called at common/src/optical_flow_calculator.cpp:71,172
// (algorithm: OpenCV LKTrackerInvoker; OpenCV is not vendored by the reference).
//
// One warp per tracked point, all pyramid levels in one launch; lanes stride over the window taps, window samples
// live in shared memory, source reads come from L1/L2.  The structure-tensor sums (A11, A12, A22) and the mismatch
// sums (b1, b2) are accumulated EXACTLY -- int32 per lane (<= 64 taps per lane keeps them below 2^31), then a
// warp-shuffle butterfly in f64 -- and rounded once to f32, where OpenCV accumulates in f32 in SIMD-lane order;
// the two agree to ~1e-7 relative (flow differences ~1e-5 px, tests/test_gpu_parity.py).
#include "lk_common.cuh"

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
        const float scale = lk_level_scale(level);
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

cudaError_t launch_lk(const LkParams &p, const LkTmaMaps *maps, const LkPhaseMaps *pmaps, int pairs, cudaStream_t s)
{
    if (!p.pts_in && pmaps && pmaps->valid && p.ph && p.win == 40) return launch_lk_phase(p, pmaps, pairs, s);
    if (maps && maps->valid && p.win == 40) return launch_lk_tma(p, maps, pairs, s);
    constexpr int WARPS = 8;
    dim3 grid((p.P + WARPS - 1) / WARPS, pairs);
    size_t smem = (size_t)WARPS * 3 * p.win * p.win * sizeof(int16_t);
    cudaError_t e = cudaFuncSetAttribute(k_lk<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_lk<WARPS><<<grid, WARPS * 32, smem, s>>>(p);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
