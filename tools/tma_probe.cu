// tma_probe.cu -- stand-alone probe of the TMA tile load used by k_lk_tma (debug aid, not part of the library).
// usage: tma_probe <variant>   0: grid_constant static index  1: grid_constant dynamic index  2: map in global memory
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../motion_detection_b200/csrc/tma.h"

struct Maps { CUtensorMap m[4]; };

__global__ void k_probe(const __grid_constant__ Maps maps, const CUtensorMap *gmap, int variant, int idx, int x, int y, int z,
                        int box_bytes, uint8_t *out, const uint8_t *gsrc)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t *tile = reinterpret_cast<uint8_t *>(((uintptr_t)sm + 127) & ~(uintptr_t)127);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + 8192);
    const int lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    const CUtensorMap *mp = variant == 0 ? &maps.m[1] : (variant == 1 ? &maps.m[idx] : gmap);
    if (variant == 3) {          // barrier only
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
    } else if (variant == 4) {   // 1-D bulk copy (no tensor map)
        if (lane == 0) {
            mbar_expect_tx(bar, 1024);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(tile)), "l"(gsrc), "r"(1024), "r"(smem_u32(bar)) : "memory");
        }
    } else if (variant == 5) {   // expect_tx then plain complete via arrive (no copy): just test expect_tx
        if (lane == 0) mbar_expect_tx(bar, 0);
    } else if (lane == 0) {
        mbar_expect_tx(bar, box_bytes);
        tma_load_3d(tile, mp, x, y, z, bar);
    }
    mbar_wait(bar, 0);
    for (int i = lane; i < box_bytes; i += 32) out[i] = tile[i];
}

int main(int argc, char **argv)
{
    int variant = argc > 1 ? atoi(argv[1]) : 0;
    const int pitch = 256, rows = 128, slots = 3, bw = 48, bh = 41;
    std::vector<uint8_t> h((size_t)pitch * rows * slots);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8) * 13);
    uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 8192);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    Maps *maps = new Maps();
    for (int i = 0; i < 4; i++)
        if (!tma_encode_3d(&maps->m[i], 1, d, pitch, rows, slots, pitch, (uint64_t)pitch * rows, bw, bh)) { printf("encode failed\n"); return 2; }
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof(CUtensorMap)); cudaMemcpy(gmap, &maps->m[1], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    int x = 13, y = 7, z = 2;
    k_probe<<<1, 32, 16384>>>(*maps, gmap, variant, 1, x, y, z, bw * bh, out, d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("variant %d: %s\n", variant, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o(bw * bh);
    cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    if (variant >= 3) return 0;
    int bad = 0;
    for (int r = 0; r < bh; r++) for (int c = 0; c < bw; c++)
        if (o[r * bw + c] != h[((size_t)z * rows + y + r) * pitch + x + c]) bad++;
    printf("variant %d: mismatches %d\n", variant, bad);
    return bad != 0;
}
