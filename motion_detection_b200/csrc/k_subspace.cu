// k_subspace.cu -- OutlierDetector::fitSubspace (common/src/outlier_detector.cpp:236-331) on the device.
//
//   fillMatrix :188-200      data (2F x T), rows x0,y0,x1,y1,...        -> read straight from traj[T][F][2]
//   meanSubtract :202-221    x rows -= mean(x0), y rows = mean(y0) - y  -> k_sub_mean (f64 fixed-order reduction)
//   fillSubset :223-234      d = 4*num_motions columns, rand() % T      -> k_sub_hyp (glibc TYPE_3 restated)
//   JacobiSVD + Pnd :266-282 Pnd = I - sum_{k<d} u_k u_k^T              -> k_sub_hyp (one-sided Jacobi, f64, one thread
//                                                                          per hypothesis, oracle operation order)
//   residual / inliers :284-299   |d_i^T Pnd d_i| < (2F-d) sigma^2     -> k_sub_score (all hypotheses x all trajectories,
//                                                                          Pnd in shared memory, ballot + popc counts)
//   best-of + threshold :300-324  first max wins; chi-square p99 table  -> k_sub_final
// Every f64 expression uses _rn intrinsics in the oracle's order, so projectors, residuals, inlier counts and labels
// are bit-identical to oracle/md_oracle_subspace.c for the same column draws.
#include <float.h>

#include "md_internal.h"

#define SUB_MAXN 32

struct SubParams {
    const float *traj;       // [T][F][2]
    int T, F, n, d, iters;
    double sigma;
    uint32_t seed;
    const int *forced_cols;  // [iters][d] or nullptr
    double *mean;            // [2]  (x0 mean, y0 mean)
    double *P;               // [iters][n*n]
    int *cols;               // [iters][d]
    int *counts;             // [iters]
    float *residual;         // [T]
    uint8_t *outlier;        // [T]
    int *best_cols;          // [d]
    int *num_inliers;        // [1]
};

__device__ __forceinline__ double sub_datum(const SubParams &p, int r, int i, float xm, float ym)
{
    // row r of column i after meanSubtract, as the f32 value the reference stores, widened to f64
    const float v = p.traj[((size_t)i * p.F + (r >> 1)) * 2 + (r & 1)];
    return (double)((r & 1) ? __fsub_rn(ym, v) : __fsub_rn(v, xm));
}

__global__ void __launch_bounds__(1024) k_sub_mean(const SubParams p)
{
    __shared__ double sx[1024], sy[1024];
    double ax = 0, ay = 0;
    for (int i = threadIdx.x; i < p.T; i += 1024) {
        ax += (double)p.traj[(size_t)i * p.F * 2];
        ay += (double)p.traj[(size_t)i * p.F * 2 + 1];
    }
    sx[threadIdx.x] = ax; sy[threadIdx.x] = ay;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sx[threadIdx.x] += sx[threadIdx.x + o]; sy[threadIdx.x] += sy[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { p.mean[0] = sx[0] / p.T; p.mean[1] = sy[0] / p.T; }
    for (int i = threadIdx.x; i < p.iters; i += 1024) p.counts[i] = 0;
}

struct SubRand { int32_t r[34]; int f, b; };
__device__ void sub_srand(SubRand &st, uint32_t seed)
{
    if (seed == 0) seed = 1;
    st.r[0] = (int32_t)seed;
    for (int i = 1; i < 31; i++) {
        long long hi = st.r[i - 1] / 127773, lo = st.r[i - 1] % 127773;
        long long word = 16807 * lo - 2836 * hi;
        if (word < 0) word += 2147483647;
        st.r[i] = (int32_t)word;
    }
    st.f = 3; st.b = 0;
}
__device__ int sub_rand(SubRand &st)
{
    uint32_t *r = reinterpret_cast<uint32_t *>(st.r);
    r[st.f] += r[st.b];
    uint32_t out = r[st.f] >> 1;
    st.f = st.f + 1 == 31 ? 0 : st.f + 1;
    st.b = st.b + 1 == 31 ? 0 : st.b + 1;
    return (int)out;
}

// ordered sum x_0 + x_1 + ... + x_{n-1} of one value per lane (the oracle's serial accumulation order), on every lane
__device__ __forceinline__ double sub_ordered_sum(double x, int n)
{
    double a = 0;
    for (int i = 0; i < n; i++) a = __dadd_rn(a, __shfl_sync(0xffffffffu, x, i));
    return a;
}

// One BLOCK per hypothesis, one warp per column pair.  The sampled columns live in shared memory (A[k * n + i], lane i
// owns row i).  The one-sided (Hestenes) Jacobi runs the oracle's ROUND-ROBIN sweep order: a round is d/2 disjoint column
// pairs whose rotations commute exactly, so the warps of the block rotate them concurrently and the result is
// bit-identical to the oracle's sequential loop (oracle/md_oracle_subspace.c:hestenes).  Inside a pair the three dot
// products are accumulated in the oracle's serial order (products in parallel, ordered sum through shuffles), the
// rotation angle is computed redundantly on every lane, the rotation itself is element-parallel.
__global__ void __launch_bounds__(512) k_sub_hyp(const SubParams p)
{
    extern __shared__ double s_hyp[];            // [d * n] doubles, [32] norms, then [d] ints
    const int n = p.n, d = p.d, nwarps = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *A = s_hyp, *s_norm = s_hyp + (size_t)n * d;
    int *s_cols = reinterpret_cast<int *>(s_norm + 32);
    const int it = blockIdx.x;
    if (threadIdx.x == 0) {
        // fillSubset (:223-234): the draws of hypothesis `it` are draws it*d .. it*d+d-1 of the stream
        if (p.forced_cols) {
            for (int k = 0; k < d; k++) s_cols[k] = p.forced_cols[it * d + k];
        } else {
            SubRand st;
            sub_srand(st, p.seed);
            for (int i = 0; i < 310 + it * d; i++) (void)sub_rand(st);
            for (int k = 0; k < d; k++) s_cols[k] = sub_rand(st) % p.T;
        }
    }
    __syncthreads();
    const float xm = (float)p.mean[0], ym = (float)p.mean[1];
    for (int e = threadIdx.x; e < n * d; e += blockDim.x) {
        const int k = e / n, i = e - k * n;
        int c = s_cols[k];
        c = c < 0 ? 0 : (c >= p.T ? p.T - 1 : c);
        if (i == 0) p.cols[it * d + k] = c;
        A[e] = sub_datum(p, i, c, xm, ym);
    }
    __syncthreads();
    const bool row = lane < n;
    const int de = d + (d & 1), m = de - 1;
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int r = 0; r < m; r++) {
            for (int i = warp; i < de / 2; i += nwarps) {
                int pc = i == 0 ? r : (r + i) % m, q = i == 0 ? m : (r - i + m) % m;
                if (pc > q) { const int t = pc; pc = q; q = t; }
                if (q >= d || pc == q) continue;
                const double vp = row ? A[pc * n + lane] : 0.0, vq = row ? A[q * n + lane] : 0.0;
                const double a = sub_ordered_sum(__dmul_rn(vp, vp), n), b = sub_ordered_sum(__dmul_rn(vq, vq), n),
                             g = sub_ordered_sum(__dmul_rn(vp, vq), n);
                if (g == 0 || fabs(g) <= __dmul_rn(1e-15, __dsqrt_rn(__dmul_rn(a, b)))) continue;          // warp-uniform
                rotated = 1;
                const double zeta = __ddiv_rn(__dsub_rn(b, a), __dmul_rn(2.0, g));
                const double t = __ddiv_rn(zeta >= 0 ? 1.0 : -1.0,
                                           __dadd_rn(fabs(zeta), __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(zeta, zeta)))));
                const double c = __ddiv_rn(1.0, __dsqrt_rn(__dadd_rn(1.0, __dmul_rn(t, t)))), sn = __dmul_rn(c, t);
                if (row) {
                    A[pc * n + lane] = __dsub_rn(__dmul_rn(c, vp), __dmul_rn(sn, vq));
                    A[q * n + lane] = __dadd_rn(__dmul_rn(sn, vp), __dmul_rn(c, vq));
                }
            }
            __syncthreads();                     // the next round pairs the columns differently
        }
        if (!__syncthreads_or(rotated)) break;
    }
    for (int k = warp; k < d; k += nwarps) {
        const double v = row ? A[k * n + lane] : 0.0;
        const double nk = __dsqrt_rn(sub_ordered_sum(__dmul_rn(v, v), n));
        if (lane == 0) s_norm[k] = nk;
    }
    __syncthreads();
    double smax = 0;
    for (int k = 0; k < d; k++) if (s_norm[k] > smax) smax = s_norm[k];
    double *P = p.P + (size_t)it * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, jj = e - i * n;
        double val = i == jj ? 1.0 : 0.0;
        for (int k = 0; k < d; k++) {
            const double nk = s_norm[k];
            if (!(nk > __dmul_rn(1e-12, smax)) || nk == 0) continue;
            const double inv = __ddiv_rn(1.0, nk);
            val = __dsub_rn(val, __dmul_rn(__dmul_rn(A[k * n + i], inv), __dmul_rn(A[k * n + jj], inv)));
        }
        P[e] = val;
    }
}

__device__ __forceinline__ double sub_residual(const double *P, const double *dcol, int n)
{
    double acc = 0;
#pragma unroll
    for (int r = 0; r < n; r++) {
        double pr = 0;
#pragma unroll
        for (int c = 0; c < n; c++) pr = __dadd_rn(pr, __dmul_rn(P[r * n + c], dcol[c]));
        acc = __dadd_rn(acc, __dmul_rn(dcol[r], pr));
    }
    return fabs(acc);
}

// all hypotheses against all trajectories: blockIdx.y = group of `hb` hypotheses (projectors in shared memory), two
// hypotheses per inner step (two independent f64 chains per thread)
template <int N>                        // N = 2F known at compile time (the datum column stays in registers), 0 = any
__global__ void __launch_bounds__(256) k_sub_score(const SubParams p, int hb)
{
    extern __shared__ double s_P[];      // [hb][n*n]
    const int n = N ? N : p.n, nn = n * n;
    const float xm = (float)p.mean[0], ym = (float)p.mean[1];
    const double inl_thr = __dmul_rn(__dmul_rn((double)(n - p.d), p.sigma), p.sigma);
    const int lane = threadIdx.x & 31;
    const int Tr = (p.T + 31) & ~31;
    const int h0 = blockIdx.y * hb;
    const int nh = min(hb, p.iters - h0);
    if (nh <= 0) return;
    for (int i = threadIdx.x; i < nh * nn; i += blockDim.x) s_P[i] = p.P[(size_t)h0 * nn + i];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < Tr; i += gridDim.x * blockDim.x) {
        double dcol[N ? N : SUB_MAXN];
        const bool act = i < p.T;
        if (act) {
#pragma unroll
            for (int r = 0; r < n; r++) dcol[r] = sub_datum(p, r, i, xm, ym);
        }
        int j = 0;
        for (; j + 1 < nh; j += 2) {
            const double r0 = act ? sub_residual(s_P + j * nn, dcol, n) : 0.0, r1 = act ? sub_residual(s_P + (j + 1) * nn, dcol, n) : 0.0;
            const unsigned b0 = __ballot_sync(0xffffffffu, act && r0 < inl_thr), b1 = __ballot_sync(0xffffffffu, act && r1 < inl_thr);
            if (lane == 0 && b0) atomicAdd(&p.counts[h0 + j], __popc(b0));
            if (lane == 0 && b1) atomicAdd(&p.counts[h0 + j + 1], __popc(b1));
        }
        if (j < nh) {
            const bool inl = act && sub_residual(s_P + j * nn, dcol, n) < inl_thr;
            const unsigned bal = __ballot_sync(0xffffffffu, inl);
            if (lane == 0 && bal) atomicAdd(&p.counts[h0 + j], __popc(bal));
        }
    }
}

// chi_square_table p99, outlier_detector.cpp:19-30
__constant__ double c_chi2_p99[10] = {0.0, 0.020, 0.115, 0.297, 0.554, 0.872, 1.239, 1.646, 2.088, 2.558};

__global__ void __launch_bounds__(256) k_sub_final(const SubParams p)
{
    extern __shared__ double s_P[];
    const int n = p.n, nn = n * n;
    int best = -1, best_cnt = 0;
    for (int j = 0; j < p.iters; j++)
        if (p.counts[j] > best_cnt) { best_cnt = p.counts[j]; best = j; }      // :300 first best wins
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *p.num_inliers = best_cnt;
        for (int k = 0; k < p.d; k++) p.best_cols[k] = best >= 0 ? p.cols[best * p.d + k] : -1;
    }
    if (best < 0) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.T; i += gridDim.x * blockDim.x) { p.residual[i] = 0.f; p.outlier[i] = 0; }
        return;
    }
    for (int i = threadIdx.x; i < nn; i += blockDim.x) s_P[i] = p.P[(size_t)best * nn + i];
    __syncthreads();
    const float xm = (float)p.mean[0], ym = (float)p.mean[1];
    double thr = 0.2;                                                           // :312-317
    const int k = n - p.d;
    if (k < 11 && k > 0) thr = __dmul_rn(__dmul_rn(p.sigma, p.sigma), c_chi2_p99[k >= 10 ? 9 : k]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.T; i += gridDim.x * blockDim.x) {
        double dcol[SUB_MAXN];
        for (int r = 0; r < n; r++) dcol[r] = sub_datum(p, r, i, xm, ym);
        const float res = (float)sub_residual(s_P, dcol, n);
        p.residual[i] = res;
        p.outlier[i] = (double)res > thr ? 1 : 0;                               // :318-324 (residual is stored as f32)
    }
}

// ---- persistent workspace (grown on demand; freed with the context) ---------------------------------------------------
struct SubWs {
    double *mean = nullptr, *P = nullptr;
    int *cols = nullptr, *counts = nullptr, *forced = nullptr, *best = nullptr, *ninl = nullptr;
    float *traj = nullptr, *res = nullptr;
    uint8_t *out = nullptr;
    size_t capP = 0, capCols = 0, capIters = 0, capForced = 0, capTraj = 0, capT = 0, capOut = 0;
};

void sub_free_workspace(void *w)
{
    SubWs *ws = (SubWs *)w;
    if (!ws) return;
    void *ptrs[] = {ws->mean, ws->P, ws->cols, ws->counts, ws->forced, ws->best, ws->ninl, ws->traj, ws->res, ws->out};
    for (void *q : ptrs) if (q) cudaFree(q);
    delete ws;
}

template <class T>
static bool sub_grow(T **ptr, size_t *cap, size_t need, cudaStream_t s)
{
    if (*cap >= need) return true;
    cudaStreamSynchronize(s);
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *cap = 0;
    if (cudaMalloc((void **)ptr, need * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return false; }
    *cap = need;
    return true;
}

// Enqueues the whole fit on the context's stream; every pointer is device memory except forced_host.
int sub_enqueue(md_ctx *ctx, const float *d_traj, int T, int F, int num_motions, double sigma, uint32_t seed,
                const int32_t *forced_host, int iters, float *d_res, uint8_t *d_out, int *d_best, int *d_ninl)
{
    const int n = 2 * F, d = 4 * num_motions;
    if (!d_traj || !d_res || !d_best || !d_out || T < 1 || F < 1 || n > SUB_MAXN || d < 1 || d > SUB_MAXN || d > n || iters < 1 ||
        iters > 4096) {
        ctx->err = "md_fit_subspace: bad arguments (need 2F <= 32, 4*num_motions <= 2F)";
        return MD_ERR_INVALID;
    }
    cudaStream_t s = ctx->stream;
    if (!ctx->sub_ws) ctx->sub_ws = new SubWs();
    SubWs *ws = (SubWs *)ctx->sub_ws;
    size_t two = ws->mean ? 2 : 0, one = ws->ninl ? 1 : 0;
    bool ok = sub_grow(&ws->mean, &two, 2, s) && sub_grow(&ws->P, &ws->capP, (size_t)iters * n * n, s) &&
              sub_grow(&ws->cols, &ws->capCols, (size_t)iters * d, s) && sub_grow(&ws->counts, &ws->capIters, (size_t)iters, s) &&
              sub_grow(&ws->ninl, &one, 1, s);
    if (ok && forced_host) ok = sub_grow(&ws->forced, &ws->capForced, (size_t)iters * d, s);
    if (!ok) { ctx->err = "md_fit_subspace: out of device memory"; return MD_ERR_NOMEM; }
    cudaError_t e = cudaSuccess;
    if (forced_host) e = cudaMemcpyAsync(ws->forced, forced_host, sizeof(int) * iters * d, cudaMemcpyHostToDevice, s);
    SubParams p;
    p.traj = d_traj;
    p.T = T; p.F = F; p.n = n; p.d = d; p.iters = iters; p.sigma = sigma; p.seed = seed;
    p.forced_cols = forced_host ? ws->forced : nullptr; p.mean = ws->mean; p.P = ws->P; p.cols = ws->cols; p.counts = ws->counts;
    p.residual = d_res; p.outlier = d_out; p.best_cols = d_best; p.num_inliers = d_ninl ? d_ninl : ws->ninl;
    k_sub_mean<<<1, 1024, 0, s>>>(p);
    const size_t hyp_smem = sizeof(double) * ((size_t)n * d + 32) + sizeof(int) * d;
    const int hyp_warps = (d + 1) / 2 < 1 ? 1 : (d + 1) / 2;                              // one warp per column pair of a round
    k_sub_hyp<<<iters, 32 * hyp_warps, hyp_smem, s>>>(p);
    int hb = 8;                                    // hypotheses per block (projectors: 8 n^2 doubles <= 64 KB)
    if (hb > iters) hb = iters;
    const size_t psm = sizeof(double) * (size_t)hb * n * n;
    int nb = (T + 255) / 256;
    if (nb > 592) nb = 592;
    const dim3 sgrid(nb, (iters + hb - 1) / hb);
    auto score = [&](auto kern) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm);
        kern<<<sgrid, 256, psm, s>>>(p, hb);
    };
    switch (n) {                                   // F = 3, 5, 7, 9 <-> num_motions 1..4 of the launch files
        case 6: score(k_sub_score<6>); break;
        case 10: score(k_sub_score<10>); break;
        case 14: score(k_sub_score<14>); break;
        case 18: score(k_sub_score<18>); break;
        default: score(k_sub_score<0>); break;
    }
    k_sub_final<<<nb, 256, sizeof(double) * n * n, s>>>(p);
    MD_COUNT_LAUNCH(4);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { ctx->err = std::string("md_fit_subspace: ") + cudaGetErrorString(e); return MD_ERR_CUDA; }
    return MD_OK;
}

#define SCK(call)                                                                                               \
    do {                                                                                                        \
        cudaError_t e_ = (call);                                                                                \
        if (e_ != cudaSuccess) {                                                                                \
            ctx->err = std::string("md_fit_subspace: ") + #call + " -> " + cudaGetErrorString(e_);             \
            return MD_ERR_CUDA;                                                                                 \
        }                                                                                                       \
    } while (0)

extern "C" int md_fit_subspace(md_ctx *ctx, const float *traj, int32_t T, int32_t F, int32_t num_motions, double sigma,
                               uint32_t seed, const int32_t *forced_cols, int32_t iters, float *residual, int32_t *best_cols,
                               uint8_t *outlier, int32_t *num_inliers, int mem)
{
    MD_NVTX("md_fit_subspace");
    if (!ctx) return MD_ERR_INVALID;
    const int n = 2 * F, d = 4 * num_motions;
    if (!traj || !residual || !best_cols || !outlier || T < 1 || F < 1 || n > SUB_MAXN || d < 1 || d > SUB_MAXN || d > n ||
        iters < 1 || iters > 4096) {
        ctx->err = "md_fit_subspace: bad arguments (need 2F <= 32, 4*num_motions <= 2F)";
        return MD_ERR_INVALID;
    }
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MD_ERR_CUDA;
    cudaStream_t s = ctx->stream;
    if (!ctx->sub_ws) ctx->sub_ws = new SubWs();
    SubWs *ws = (SubWs *)ctx->sub_ws;
    size_t capBest = ws->best ? SUB_MAXN : 0;
    if (!sub_grow(&ws->best, &capBest, SUB_MAXN, s)) { ctx->err = "md_fit_subspace: out of device memory"; return MD_ERR_NOMEM; }
    if (mem == MD_MEM_DEVICE) {
        // stream ordered; best_cols / num_inliers are device pointers too
        return sub_enqueue(ctx, traj, T, F, num_motions, sigma, seed, forced_cols, iters, residual, outlier, best_cols, num_inliers);
    }
    const size_t tb = (size_t)2 * T * F;
    if (!sub_grow(&ws->traj, &ws->capTraj, tb, s) || !sub_grow(&ws->res, &ws->capT, (size_t)T, s) || !sub_grow(&ws->out, &ws->capOut, (size_t)T, s)) {
        ctx->err = "md_fit_subspace: out of device memory";
        return MD_ERR_NOMEM;
    }
    SCK(cudaMemcpyAsync(ws->traj, traj, tb * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = sub_enqueue(ctx, ws->traj, T, F, num_motions, sigma, seed, forced_cols, iters, ws->res, ws->out, ws->best, nullptr);
    if (rc != MD_OK) return rc;
    SCK(cudaMemcpyAsync(residual, ws->res, sizeof(float) * T, cudaMemcpyDeviceToHost, s));
    SCK(cudaMemcpyAsync(outlier, ws->out, T, cudaMemcpyDeviceToHost, s));
    SCK(cudaMemcpyAsync(best_cols, ws->best, sizeof(int) * d, cudaMemcpyDeviceToHost, s));
    if (num_inliers) SCK(cudaMemcpyAsync(num_inliers, ws->ninl, sizeof(int), cudaMemcpyDeviceToHost, s));
    SCK(cudaStreamSynchronize(s));
    return MD_OK;
}
