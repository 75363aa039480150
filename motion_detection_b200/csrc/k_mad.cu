// k_mad.cu -- OutlierDetector::findOutliers / createMask (common/src/outlier_detector.cpp:37-186) on the device:
// the median / MAD outlier test on the angles and magnitudes of the flow vectors.
//
//   createAngleMatrix :75-85       angle = atan2(dy, dx)                      (f64)
//   createMagnitudeMatrix :87-97   magnitude = sqrt(dy*dy + dx*dx)            (f64; exact products of f32-born values)
//   createMask :122-186            vals = values of the vectors taking part (all grid vectors, or the non-zero ones);
//                                  median (getMedian :99-121: mean of the two middle values for an even count);
//                                  diff = |val - median|; MAD = median(diff);
//                                  0.6745 * diff / MAD > 3.5  ->  mask = 1       (angle pass, then magnitude pass: OR)
//
// The medians are exact order statistics: an 8-pass MSB-first radix select over order-preserving 64-bit keys, all four
// selections of a stage (2 quantities x lower / upper middle rank) in the same launches (blockIdx.y), bucket choice by the
// last block to finish a pass (ticket counter) -- no sort, no host round trip.  HBM-bound by construction (8 reads of the
// keys per stage) and tiny: n is the number of grid vectors.
#include <float.h>

#include "md_internal.h"

struct MadSel {                 // one running selection
    unsigned long long prefix;  // key bits decided so far (high bytes)
    unsigned int k;             // remaining rank inside the prefix bucket
    unsigned int pad;
};

struct MadParams {
    const double *dxdy;         // [n][2]
    int n, include_zeros;
    double *vals;               // [2][n]   angle, magnitude; later |val - median|
    uint8_t *sel;               // [n] vector takes part
    unsigned int *count;        // [1] number of participating vectors
    unsigned int *hist;         // [4][256]
    unsigned int *ticket;       // [4]
    MadSel *state;              // [4]  (q * 2 + {lower, upper})
    double *stats;              // [4] median angle, MAD angle, median magnitude, MAD magnitude
    uint8_t *outlier;           // [n]
};

__device__ __forceinline__ unsigned long long mad_key(double v)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double mad_unkey(unsigned long long k)
{
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) k_mad_values(const MadParams p)
{
    unsigned int c = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        const double dx = p.dxdy[2 * i], dy = p.dxdy[2 * i + 1];
        p.vals[i] = atan2(dy, dx);                                                    // :82
        p.vals[p.n + i] = __dsqrt_rn(__dadd_rn(__dmul_rn(dy, dy), __dmul_rn(dx, dx)));   // :94
        const bool s = p.include_zeros || fabs(dx) > 0.0 || fabs(dy) > 0.0;           // :132-139
        p.sel[i] = s ? 1 : 0;
        p.outlier[i] = 0;
        c += s ? 1u : 0u;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(p.count, c);
}

// ranks of the two middle elements (getMedian :110-119) for both quantities
__global__ void k_mad_init_select(const MadParams p)
{
    const int t = threadIdx.x;
    if (t < 4) {
        const unsigned int m = *p.count;
        p.state[t].prefix = 0;
        p.state[t].k = m == 0 ? 0 : ((t & 1) ? m / 2 : (m - 1) / 2);
        p.ticket[t] = 0;
    }
    for (int i = t; i < 4 * 256; i += blockDim.x) p.hist[i] = 0;
}

// one radix pass (byte `pass`, 7 = most significant) of the four selections
__global__ void __launch_bounds__(256) k_mad_select_pass(const MadParams p, int pass)
{
    __shared__ unsigned int s_hist[256];
    __shared__ bool s_last;
    const int sid = blockIdx.y, q = sid >> 1;
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = p.state[sid].prefix;
    const int shift = 8 * pass;
    const double *v = p.vals + (size_t)q * p.n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        if (!p.sel[i]) continue;
        const unsigned long long key = mad_key(v[i]);
        if (pass == 7 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&s_hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (s_hist[threadIdx.x]) atomicAdd(&p.hist[sid * 256 + threadIdx.x], s_hist[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&p.ticket[sid], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    // the last block of this selection: pick the bucket holding rank k, refine the prefix, clear for the next pass
    s_hist[threadIdx.x] = atomicExch(&p.hist[sid * 256 + threadIdx.x], 0u);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int k = p.state[sid].k, acc = 0;
        int b = 0;
        for (; b < 255; b++) {
            if (acc + s_hist[b] > k) break;
            acc += s_hist[b];
        }
        p.state[sid].k = k - acc;
        p.state[sid].prefix = (pass == 7 ? 0ull : prefix) | ((unsigned long long)b << shift);
        p.ticket[sid] = 0;
    }
}

// stage 0: medians -> stats[0], stats[2], vals <- |val - median|;  stage 1: MADs -> stats[1], stats[3]
__global__ void __launch_bounds__(256) k_mad_finish(const MadParams p, int stage)
{
    const unsigned int m = *p.count;
    double med[2];
    for (int q = 0; q < 2; q++) {
        const double lo = mad_unkey(p.state[2 * q].prefix), hi = mad_unkey(p.state[2 * q + 1].prefix);
        med[q] = (m & 1) ? lo : __ddiv_rn(__dadd_rn(lo, hi), 2.0);              // :110-119
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && m) { p.stats[stage] = med[0]; p.stats[2 + stage] = med[1]; }
    if (m == 0) return;                                                          // vals.empty(): nothing is flagged (:142-145)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        if (!p.sel[i]) continue;
        if (stage == 0) {
            p.vals[i] = fabs(__dsub_rn(p.vals[i], med[0]));                      // :150-154
            p.vals[p.n + i] = fabs(__dsub_rn(p.vals[p.n + i], med[1]));
        } else {
            const double za = __ddiv_rn(__dmul_rn(0.6745, p.vals[i]), med[0]);   // :168 / :177
            const double zm = __ddiv_rn(__dmul_rn(0.6745, p.vals[p.n + i]), med[1]);
            p.outlier[i] = (fabs(za) > 3.5 || fabs(zm) > 3.5) ? 1 : 0;
        }
    }
}

struct MadWs {
    double *dxdy = nullptr, *vals = nullptr, *stats = nullptr;
    uint8_t *sel = nullptr, *outlier = nullptr;
    unsigned int *ints = nullptr;     // count, ticket[4], hist[1024], then MadSel[4]
    int cap = 0;
};

void mad_free_workspace(void *w)
{
    MadWs *ws = (MadWs *)w;
    if (!ws) return;
    void *ptrs[] = {ws->dxdy, ws->vals, ws->stats, ws->sel, ws->outlier, ws->ints};
    for (void *q : ptrs) if (q) cudaFree(q);
    delete ws;
}

#define MCK(call)                                                                                           \
    do {                                                                                                    \
        cudaError_t e_ = (call);                                                                            \
        if (e_ != cudaSuccess) {                                                                            \
            ctx->err = std::string("md_find_outliers: ") + #call + " -> " + cudaGetErrorString(e_);        \
            return MD_ERR_CUDA;                                                                             \
        }                                                                                                   \
    } while (0)

extern "C" int md_find_outliers(md_ctx *ctx, const double *flow_dxdy, int32_t n, int32_t include_zeros, uint8_t *outlier,
                                double *stats4, int mem)
{
    MD_NVTX("md_find_outliers");
    if (!ctx) return MD_ERR_INVALID;
    if (!flow_dxdy || !outlier || n < 1 || n > (1 << 26)) { ctx->err = "md_find_outliers: bad arguments"; return MD_ERR_INVALID; }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MD_ERR_CUDA;
    cudaStream_t s = ctx->stream;
    if (!ctx->mad_ws) ctx->mad_ws = new MadWs();
    MadWs *ws = (MadWs *)ctx->mad_ws;
    if (ws->cap < n) {
        MCK(cudaStreamSynchronize(s));
        void *ptrs[] = {ws->dxdy, ws->vals, ws->stats, ws->sel, ws->outlier, ws->ints};
        for (void *q : ptrs) if (q) cudaFree(q);
        ws->dxdy = ws->vals = ws->stats = nullptr; ws->sel = ws->outlier = nullptr; ws->ints = nullptr; ws->cap = 0;
        MCK(cudaMalloc((void **)&ws->dxdy, sizeof(double) * 2 * n));
        MCK(cudaMalloc((void **)&ws->vals, sizeof(double) * 2 * n));
        MCK(cudaMalloc((void **)&ws->stats, sizeof(double) * 4));
        MCK(cudaMalloc((void **)&ws->sel, n));
        MCK(cudaMalloc((void **)&ws->outlier, n));
        MCK(cudaMalloc((void **)&ws->ints, sizeof(unsigned int) * (8 + 1024) + sizeof(MadSel) * 4));
        ws->cap = n;
    }
    const bool host = mem == MD_MEM_HOST;
    MadParams p;
    if (host) MCK(cudaMemcpyAsync(ws->dxdy, flow_dxdy, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, s));
    p.dxdy = host ? ws->dxdy : flow_dxdy;
    p.n = n; p.include_zeros = include_zeros ? 1 : 0;
    p.vals = ws->vals; p.sel = ws->sel;
    p.count = ws->ints; p.ticket = ws->ints + 4; p.hist = ws->ints + 8;
    p.state = reinterpret_cast<MadSel *>(ws->ints + 8 + 1024);
    p.stats = ws->stats; p.outlier = host ? ws->outlier : outlier;
    MCK(cudaMemsetAsync(ws->ints, 0, sizeof(unsigned int) * 8, s));
    MCK(cudaMemsetAsync(ws->stats, 0, sizeof(double) * 4, s));
    int nb = (n + 255) / 256;
    if (nb > 148 * 4) nb = 148 * 4;
    k_mad_values<<<nb, 256, 0, s>>>(p);
    for (int stage = 0; stage < 2; stage++) {
        k_mad_init_select<<<1, 256, 0, s>>>(p);
        for (int pass = 7; pass >= 0; pass--) k_mad_select_pass<<<dim3(nb, 4), 256, 0, s>>>(p, pass);
        k_mad_finish<<<nb, 256, 0, s>>>(p, stage);
    }
    MD_COUNT_LAUNCH(1 + 2 * 10);
    MCK(cudaGetLastError());
    const cudaMemcpyKind outk = host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (host) MCK(cudaMemcpyAsync(outlier, ws->outlier, n, outk, s));
    if (stats4) MCK(cudaMemcpyAsync(stats4, ws->stats, sizeof(double) * 4, outk, s));
    if (host) MCK(cudaStreamSynchronize(s));
    return MD_OK;
}
