// pipe_probe2.cu -- cost of the LK iteration tap body on sm_100a (the loop of k_lk_phase: 6 staged rows -> 50 taps per lane) in the
// variants considered.  Prints clocks per iteration per warp and SM sub-partition, and the implied issue rate.
// build + run on the GPU box: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I motion_detection_b200/csrc -I include \
//    -o /tmp/pp2 tools/pipe_probe2.cu && /tmp/pp2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "lk_tile.cuh"

constexpr int TW = 10, TH = 5, NP = 5, JP = 20;

// variants of TapLoop::iter2 (two taps per step)
template <int MODEX, int I>
__device__ __forceinline__ void pair_step(const RowWords &r0, const RowWords &r1, int wtop, int wbot, const int (&Xpk)[NP], const int (&Ypk)[NP],
                                          int &b1lo, int &b1hi, int &b2lo, int &b2hi)
{
    constexpr int MODE = MODEX >= 10 ? 0 : MODEX;
    const int v0 = row_pair<I>(r1, wbot, row_pair<I>(r0, wtop, 1 << (W_BITS - 5 - 1)));
    const int v1 = row_pair<I + 1>(r1, wbot, row_pair<I + 1>(r0, wtop, 1 << (W_BITS - 5 - 1)));
    const int xp = Xpk[I >> 1], yp = Ypk[I >> 1];
    if (MODE == 4) { b1lo += v0; b2lo ^= v1; b1hi += xp; b2hi += yp; return; }      // bilinear only
    uint32_t qb;
    if (MODE == 0) qb = __byte_perm((uint32_t)(v0 >> 1), (uint32_t)(v1 >> 1), 0x6251);
    else if (MODE == 1) qb = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x6251) & 0xfffffefeu;
    else qb = __byte_perm((uint32_t)v0, (uint32_t)v1, 0x6251);
    if (MODE == 6) {
        const int q0 = v0 >> 9, q1 = v1 >> 9;
        b1lo += q0 * (int)(short)xp; b1hi += q1 * (xp >> 16); b2lo += q0 * (int)(short)yp; b2hi += q1 * (yp >> 16);
        return;
    }
    b1lo = dp2a_lo(xp, qb, b1lo); b1hi = dp2a_hi(xp, qb, b1hi);
    b2lo = dp2a_lo(yp, qb, b2lo); b2hi = dp2a_hi(yp, qb, b2hi);
}
template <int MODE>
__device__ __forceinline__ void row_step(const RowWords &r0, const RowWords &r1, int wtop, int wbot, const int (&Xpk)[NP], const int (&Ypk)[NP],
                                         int &b1lo, int &b1hi, int &b2lo, int &b2hi)
{
    pair_step<MODE, 0>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
    pair_step<MODE, 2>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
    pair_step<MODE, 4>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
    pair_step<MODE, 6>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
    pair_step<MODE, 8>(r0, r1, wtop, wbot, Xpk, Ypk, b1lo, b1hi, b2lo, b2hi);
}

template <int MODE>
__global__ void __launch_bounds__(32, 18) probe(uint32_t *out, long long *clk, uint32_t seed, int iters)
{
    __shared__ uint32_t tJ[JP * 47];
    const int lane = threadIdx.x & 31, lx = lane & 3, ly = lane >> 2;
    for (int i = lane; i < JP * 47; i += 32) tJ[i] = seed * i + (i << 7);
    int Xpk[TH][NP], Ypk[TH][NP];
#pragma unroll
    for (int r = 0; r < TH; r++)
#pragma unroll
        for (int i = 0; i < NP; i++) { Xpk[r][i] = (int)out[(r * NP + i) * 64 + lane]; Ypk[r][i] = (int)out[(r * NP + i) * 64 + 32 + lane]; }
    int wtop = seed | 0x00010001, wbot = (seed >> 3) | 0x00020002;
    int b1lo = 0, b1hi = 0, b2lo = 0, b2hi = 0;
    int inx = seed & 15, iny = 1;
    uint32_t ex[4] = {seed, seed * 3, seed * 5, seed * 7};
    __syncwarp();
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        const int c0 = inx + TW * lx, wb = c0 >> 2, sh = (c0 & 3) * 8;
        const uint32_t *rowp = tJ + (iny + TH * ly) * JP;
        RowWords r0 = load_row(rowp, wb, sh);
#pragma unroll
        for (int r = 0; r < TH; r++) {
            const RowWords r1 = load_row(rowp + (r + 1) * JP, wb, sh);
            row_step<MODE>(r0, r1, wtop, wbot, Xpk[r], Ypk[r], b1lo, b1hi, b2lo, b2hi);
            r0 = r1;
        }
        if (MODE == 10) {
#pragma unroll
            for (int e = 0; e < 100; e++) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ex[e & 3]) : "r"(wtop), "r"(wbot));
        }
        if (MODE == 11) {
#pragma unroll
            for (int e = 0; e < 50; e++) ex[e & 3] ^= tJ[(iny + e) * 3 + lane];
        }
        if (MODE == 12) {
#pragma unroll
            for (int e = 0; e < 50; e++) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(ex[e & 3]) : "r"(wtop), "r"(wbot));
        }
        if (MODE == 13) {
#pragma unroll
            for (int e = 0; e < 100; e++) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(ex[e & 3]) : "r"(wtop), "r"(wbot));
        }
        inx = (inx + (b1lo & 1) + 1) & 15; iny = (b2hi & 1) + 1;
        wtop += b1hi & 1; wbot ^= b2lo & 2;
    }
    const long long t1 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (lane == 0 && blockIdx.x == 0) clk[148 * 32] = (long long)(g1 - g0);
    out[blockIdx.x * 32 + lane] = b1lo + b1hi + b2lo + b2hi + ex[0] + ex[1] + ex[2] + ex[3];
    if (lane == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int instr, uint32_t *out, long long *clk, int ctas_per_sm)
{
    const int iters = 4000, blocks = 148 * ctas_per_sm;
    probe<MODE><<<blocks, 32>>>(out, clk, 12345u, iters);
    probe<MODE><<<blocks, 32>>>(out, clk, 12345u, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    static long long h[148 * 32];
    cudaMemcpy(h, clk, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; i++) avg += (double)h[i];
    avg /= blocks;
    const double its = (double)iters * ctas_per_sm / 4.0;                  // warp iterations per SM sub-partition
    long long ns; cudaMemcpy(&ns, clk + 148 * 32, 8, cudaMemcpyDeviceToHost);
    printf("[%.0f MHz] ", 1e3 * (double)h[0] / (double)ns);
    printf("%-44s warps/SM %2d  %7.1f clk / iteration / SMSP  (~%d instr -> IPC %.3f)\n", name, ctas_per_sm, avg / its, instr, instr * its / avg);
}

int main()
{
    uint32_t *out; long long *clk;
    cudaMalloc(&out, 148 * 32 * 32 * 4); cudaMalloc(&clk, (148 * 32 + 1) * 8);
    for (int wps : {8, 18}) {
        run<10>("current + 100 LOP3", 60 + 375, out, clk, wps);
        run<13>("current + 100 SHF", 60 + 375, out, clk, wps);
        run<11>("current + 50 LDS + 50 LOP3", 60 + 375, out, clk, wps);
        run<12>("current + 50 IMAD", 60 + 325, out, clk, wps);
        run<0>("current: 4 IDP + 2 SHF + PRMT + 4 IDP", 60 + 275, out, clk, wps);
        run<1>("4 IDP + PRMT + AND + 4 IDP", 60 + 250, out, clk, wps);
        run<2>("4 IDP + PRMT + 4 IDP (bound, wrong result)", 60 + 225, out, clk, wps);
        run<4>("bilinear only: 4 IDP + 4 ALU", 60 + 200, out, clk, wps);
        run<6>("4 IDP + 2 SHF + unpack + 4 IMAD", 60 + 300, out, clk, wps);
    }
    return 0;
}
