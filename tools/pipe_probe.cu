// pipe_probe.cu -- issue rates of the integer instructions the LK iteration loop is made of (sm_100a): IDP.2A, IDP.4A, IMAD, SHF,
// PRMT, LOP3 alone and in the loop's mixes.  Prints warp-instructions per clock per SM sub-partition.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe tools/pipe_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define REPS 64

template <int MODE>
__global__ void __launch_bounds__(512) probe(uint32_t *out, long long *clk, uint32_t seed, int iters)
{
    uint32_t a[CHAINS], b[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed ^ (i * 77 + threadIdx.x); }
    const uint32_t w = seed | 0x00010001u;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REPS; r++) {
#pragma unroll
            for (int i = 0; i < CHAINS; i++) {
                if (MODE == 0) asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                if (MODE == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(w));
                if (MODE == 2) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                                 asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[(i + 1) % CHAINS]), "r"(w)); }
                if (MODE == 3) asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                if (MODE == 4) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                                 asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(w), "r"(a[(i + 1) % CHAINS])); }
                if (MODE == 5) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                if (MODE == 6) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(w));
                if (MODE == 7) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(w));
                if (MODE == 8) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));        // the loop's mix:
                                 asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(w), "r"(a[i]));        // 4 IDP : 1 SHF : 0.5 PRMT
                                 asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                                 asm volatile("dp2a.hi.s32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(w), "r"(a[i]));
                                 asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(w)); }
                if (MODE == 9) { asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                                 asm volatile("add.s32 %0, %0, %1;" : "+r"(b[i]) : "r"(a[(i + 1) % CHAINS])); }
                if (MODE == 10) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i]));
                                  asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[(i + 1) % CHAINS]), "r"(w)); }
                if (MODE == 11) { float f = __uint_as_float(a[i]), g = __uint_as_float(b[i]);
                                  asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f) : "f"(g), "f"(g)); a[i] = __float_as_uint(f);
                                  asm volatile("dp2a.lo.s32.u32 %0, %1, %2, %0;" : "+r"(b[i]) : "r"(w), "r"(a[(i + 1) % CHAINS])); }
                if (MODE == 12) { asm volatile("mad.lo.s32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(w), "r"(b[i])); asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[(i + 1) % CHAINS]), "r"(w)); asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(a[(i + 2) % CHAINS]), "r"(w)); }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += a[i] ^ b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int per_step, uint32_t *out, long long *clk, int warps_per_sm)
{
    const int iters = 64, blocks = 148, threads = warps_per_sm * 32;
    probe<MODE><<<blocks, threads>>>(out, clk, 12345u, iters);
    probe<MODE><<<blocks, threads>>>(out, clk, 12345u, iters);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; i++) avg += (double)h[i];
    avg /= 148;
    const double instr = (double)iters * REPS * CHAINS * per_step * warps_per_sm;   // warp-instructions per SM
    printf("%-34s warps/SM %2d  %.3f warp-instr / clk / SMSP\n", name, warps_per_sm, instr / avg / 4.0);
}

int main()
{
    uint32_t *out; long long *clk;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
    for (int wps : {4, 8, 16}) {
        run<0>("IDP.2A", 1, out, clk, wps);
        run<5>("IDP.4A", 1, out, clk, wps);
        run<3>("IMAD", 1, out, clk, wps);
        run<1>("SHF", 1, out, clk, wps);
        run<6>("PRMT", 1, out, clk, wps);
        run<7>("LOP3", 1, out, clk, wps);
        run<2>("IDP.2A + SHF", 2, out, clk, wps);
        run<4>("IDP.2A + IMAD", 2, out, clk, wps);
        run<9>("IDP.2A + IADD", 2, out, clk, wps);
        run<10>("IMAD + SHF", 2, out, clk, wps);
        run<11>("FFMA + IDP.2A", 2, out, clk, wps);
        run<8>("4 IDP.2A + SHF", 5, out, clk, wps);
        run<12>("IMAD + LOP3 + SHF", 3, out, clk, wps);
    }
    return 0;
}
