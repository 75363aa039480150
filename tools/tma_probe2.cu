// tma_probe2.cu -- probe tensor-map parameter space for the TMA tile load (debug aid).
// usage: tma_probe2 rank elem_bytes box0 box1 l2promo
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../motion_detection_b200/csrc/tma.h"

__global__ void k_probe(const __grid_constant__ CUtensorMap map, int rank, int x, int y, int z, int box_bytes, uint8_t *out)
{
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t *tile = reinterpret_cast<uint8_t *>(((uintptr_t)sm + 127) & ~(uintptr_t)127);
    uint64_t *bar = reinterpret_cast<uint64_t *>(tile + 32768);
    const int lane = threadIdx.x & 31;
    if (lane == 0) { mbar_init(bar, 1); mbar_fence_init(); }
    __syncwarp();
    if (lane == 0) {
        mbar_expect_tx(bar, box_bytes);
        if (rank == 3) tma_load_3d(tile, &map, x, y, z, bar);
        else asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                          ::"r"(smem_u32(tile)), "l"(&map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = lane; i < box_bytes; i += 32) out[i] = tile[i];
}

int main(int argc, char **argv)
{
    int rank = atoi(argv[1]), eb = atoi(argv[2]), b0 = atoi(argv[3]), b1 = atoi(argv[4]), promo = atoi(argv[5]);
    const int pitch_bytes = 1024, rows = 128, slots = 3;
    std::vector<uint8_t> h((size_t)pitch_bytes * rows * slots);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)(i * 7 + (i >> 8) * 13);
    uint8_t *d, *out; cudaMalloc(&d, h.size()); cudaMalloc(&out, 65536);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    encode_fn fn = (encode_fn)p;
    alignas(64) CUtensorMap map;
    cuuint64_t gdim[3] = {(cuuint64_t)(pitch_bytes / eb), rows, slots};
    cuuint64_t gstr[2] = {pitch_bytes, (cuuint64_t)pitch_bytes * rows};
    cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(&map, eb == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT32, rank, d, gdim, gstr, box, es,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d q=%d  ", (int)r, (int)q);
    const unsigned long long *w = reinterpret_cast<const unsigned long long *>(&map);
    for (int i = 0; i < 8; i++) printf("%016llx ", w[i]);
    printf("\n");
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    int x = argc > 6 ? atoi(argv[6]) : 13, y = 7, z = rank == 3 ? 2 : 0;
    int bytes = b0 * b1 * eb;
    k_probe<<<1, 32, 40000>>>(map, rank, x, y, z, bytes, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("rank %d eb %d box %dx%d promo %d: %s\n", rank, eb, b0, b1, promo, cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o(bytes);
    cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < b1; r2++) for (int c = 0; c < b0 * eb; c++)
        if (o[r2 * b0 * eb + c] != h[((size_t)z * rows + y + r2) * pitch_bytes + x * eb + c]) bad++;
    printf("  mismatches %d\n", bad);
    return bad != 0;
}
