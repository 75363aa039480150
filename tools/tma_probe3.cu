// tma_probe3.cu -- CUDA programming-guide style TMA load through libcu++ (debug aid).
#include <cstdio>
#include <vector>
#include <cuda/barrier>
#include "../motion_detection_b200/csrc/tma.h"
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
constexpr int SW = 64, SH = 32;

__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, uint8_t *out)
{
    __shared__ alignas(128) uint8_t smem_buffer[SH][SW];
#pragma nv_diag_suppress static_var_with_dynamic_init
    __shared__ barrier bar;
    if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
    __syncthreads();
    barrier::arrival_token token;
    if (threadIdx.x == 0) {
        cde::cp_async_bulk_tensor_2d_global_to_shared(&smem_buffer, &tensor_map, x, y, bar);
        token = cuda::device::barrier_arrive_tx(bar, 1, sizeof(smem_buffer));
    } else token = bar.arrive();
    bar.wait(std::move(token));
    for (int i = threadIdx.x; i < SW * SH; i += blockDim.x) out[i] = (&smem_buffer[0][0])[i];
}

int main()
{
    const int pitch = 1024, rows = 128;
    std::vector<uint8_t> h((size_t)pitch * rows);
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
    cuuint64_t gdim[2] = {pitch, rows};
    cuuint64_t gstr[1] = {pitch};
    cuuint32_t box[2] = {SW, SH}, es[2] = {1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d\n", (int)r);
    kernel<<<1, 128>>>(map, 16, 8, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("guide-style: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint8_t> o(SW * SH);
    cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < SH; r2++) for (int c = 0; c < SW; c++) if (o[r2 * SW + c] != h[(size_t)(8 + r2) * pitch + 16 + c]) bad++;
    printf("mismatches %d\n", bad);
    return 0;
}
