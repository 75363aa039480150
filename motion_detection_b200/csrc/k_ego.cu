// k_ego.cu -- K3: vector filter + egomotion fit, entirely on the device (no host round trip inside a batch).
//
//  k_keep_count / k_scan / k_compact : the status / min_vector_size filter of calculateOpticalFlow
//        (common/src/optical_flow_calculator.cpp:78-117) as an ORDER-PRESERVING stream compaction: src/dst keep
//        the grid order of the reference's push_back loop, which the sampling below depends on.
//  k_hypotheses : minimal-sample hypotheses.  Indices are drawn as rand() % M with glibc's TYPE_3 generator
//        restated on the device (sampling pattern of fillSubset, common/src/outlier_detector.cpp:223-234; srand at :17);
//        each hypothesis is the 8x8 LU solve of cv::getPerspectiveTransform (cpp:120) or an exact 3-point affine.
//  k_score      : batched scoring of all hypotheses against all kept vectors; inlier counts through
//        __ballot_sync + __popc and integer atomics (deterministic).
//  k_accum / k_solve : first-best-wins selection (outlier_detector.cpp:300), least-squares refit over the winner's
//        inliers: block-reduced normal-equation sums in f64 with a fixed-order second stage (no float atomics),
//        8x8 / 3x3 solve, de-normalisation, 3x3 inverse for the warp.
// Hypothesis generation and the inlier test use _rn intrinsics in the oracle's operation order, so given identical
// flow vectors the hypotheses, inlier sets and counts are bit-identical to the CPU oracle.
#include <float.h>

#include "md_internal.h"

#define SCAN_ITEMS 2048   // items per block in keep/compact (256 threads x 8)
#define NSUM 24

// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 src_point(const EgoParams &p, int b, int k)
{
    if (p.pts_in) return p.pts_in[(size_t)b * p.P + k];
    return make_float2((float)(p.ps * (k / p.gy)), (float)(p.ps * (k % p.gy)));
}

__global__ void __launch_bounds__(256) k_keep_count(const EgoParams p)
{
    const int b = blockIdx.y, blk = blockIdx.x;
    const int base = blk * SCAN_ITEMS + threadIdx.x * 8;
    int cnt = 0, trk = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int k = base + i;
        if (k >= p.P) break;
        size_t gi = (size_t)b * p.P + k;
        int kp;
        int st = p.status[gi];
        if (p.keep_given) kp = p.keep[gi] != 0;
        else {
            kp = 0;
            if (st) {
                float2 s = src_point(p, b, k), d = p.next[gi];
                float xd = __fsub_rn(d.x, s.x), yd = __fsub_rn(d.y, s.y);
                kp = ((double)fabsf(xd) > p.min_vec) || ((double)fabsf(yd) > p.min_vec);
            }
            p.keep[gi] = (uint8_t)kp;
        }
        cnt += kp;
        trk += st != 0;
    }
    __shared__ int s_cnt, s_trk;
    if (threadIdx.x == 0) { s_cnt = 0; s_trk = 0; }
    __syncthreads();
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xffffffffu, cnt, o); trk += __shfl_xor_sync(0xffffffffu, trk, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_cnt, cnt); atomicAdd(&s_trk, trk); }
    __syncthreads();
    if (threadIdx.x == 0) {
        p.blockcnt[b * p.nblk_scan + blk] = s_cnt;
        if (p.stat_tracked && s_trk) atomicAdd(p.stat_tracked, (unsigned long long)s_trk);
    }
}

// exclusive scan of the per-block counts of one pair (sequential carry over chunks of 1024)
__global__ void __launch_bounds__(1024) k_scan(const EgoParams p)
{
    const int b = blockIdx.x;
    __shared__ int s[1024];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    for (int i = threadIdx.x; i < p.iters; i += blockDim.x) p.counts[b * p.iters + i] = 0;
    __syncthreads();
    for (int c0 = 0; c0 < p.nblk_scan; c0 += 1024) {
        int i = c0 + threadIdx.x;
        int v = i < p.nblk_scan ? p.blockcnt[b * p.nblk_scan + i] : 0;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            int t = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
            __syncthreads();
            s[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < p.nblk_scan) p.blockcnt[b * p.nblk_scan + i] = carry + s[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += s[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.M[b] = carry;
}

__global__ void __launch_bounds__(256) k_compact(const EgoParams p)
{
    const int b = blockIdx.y, blk = blockIdx.x;
    const int base = blk * SCAN_ITEMS + threadIdx.x * 8;
    int flags = 0, cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int k = base + i;
        if (k < p.P && p.keep[(size_t)b * p.P + k]) { flags |= 1 << i; cnt++; }
    }
    // block exclusive scan of cnt
    __shared__ int wsum[8];
    int incl = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; w++) woff += wsum[w];
    int off = p.blockcnt[b * p.nblk_scan + blk] + woff + incl - cnt;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (flags & (1 << i)) p.kept_idx[(size_t)b * p.P + off++] = base + i;
}

// ---------------------------------------------------------------------------------------------------------------
// glibc srand()/rand(): TYPE_3, r[i] = r[i-3] + r[i-31], output >> 1, first 310 outputs discarded
struct GlibcRand { int32_t r[34]; int f, b; };
__device__ void glibc_srand(GlibcRand &st, uint32_t seed)
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
__device__ int glibc_rand(GlibcRand &st)
{
    uint32_t *r = reinterpret_cast<uint32_t *>(st.r);
    r[st.f] += r[st.b];
    uint32_t out = r[st.f] >> 1;
    st.f = st.f + 1 == 31 ? 0 : st.f + 1;
    st.b = st.b + 1 == 31 ? 0 : st.b + 1;
    return (int)out;
}

// The same generator with its 31-word state in REGISTERS: 31 consecutive rand() calls from the initial position (f = 3, b = 0) are
// r[(3 + t) % 31] += r[t] for t = 0 .. 30 and bring f back to 3, so a round unrolls into 31 adds with static indices (the ring in a
// dynamically indexed local array costs a dependent local-memory round trip per call: ~40 us for the 510 calls of a pair, which
// was the whole duration of k_hypotheses).  out != nullptr: the round's 31 outputs (r >> 1) are stored.
__device__ __forceinline__ void glibc_rand_round31(uint32_t (&r)[31], int *out)
{
#pragma unroll
    for (int t = 0; t < 31; t++) {
        r[(3 + t) % 31] += r[t];
        if (out) out[t] = (int)(r[(3 + t) % 31] >> 1);
    }
}

// Gaussian elimination with partial pivoting, same operation order as the oracle's solve_lu (no FMA contraction)
__device__ bool solve_lu(double *A, double *b, int n)
{
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
            if (fabs(A[j * n + i]) > fabs(A[k * n + i])) k = j;
        if (fabs(A[k * n + i]) < DBL_EPSILON * 100) return false;
        if (k != i) {
            for (int j = i; j < n; j++) { double t = A[i * n + j]; A[i * n + j] = A[k * n + j]; A[k * n + j] = t; }
            double t = b[i]; b[i] = b[k]; b[k] = t;
        }
        double d = __ddiv_rn(-1.0, A[i * n + i]);
        for (int j = i + 1; j < n; j++) {
            double alpha = __dmul_rn(A[j * n + i], d);
            for (int c = i + 1; c < n; c++) A[j * n + c] = __dadd_rn(A[j * n + c], __dmul_rn(alpha, A[i * n + c]));
            b[j] = __dadd_rn(b[j], __dmul_rn(alpha, b[i]));
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double s = b[i];
        for (int c = i + 1; c < n; c++) s = __dsub_rn(s, __dmul_rn(A[i * n + c], b[c]));
        b[i] = __ddiv_rn(s, A[i * n + i]);
    }
    return true;
}

__device__ bool perspective_4pt(const double *src, const double *dst, double *H)
{
    double A[64], b[8];
    for (int i = 0; i < 64; i++) A[i] = 0;
    for (int i = 0; i < 4; i++) {
        double sx = src[2 * i], sy = src[2 * i + 1], X = dst[2 * i], Y = dst[2 * i + 1];
        double *r0 = A + i * 8, *r1 = A + (i + 4) * 8;
        r0[0] = r1[3] = sx; r0[1] = r1[4] = sy; r0[2] = r1[5] = 1;
        r0[6] = __dmul_rn(-sx, X); r0[7] = __dmul_rn(-sy, X);
        r1[6] = __dmul_rn(-sx, Y); r1[7] = __dmul_rn(-sy, Y);
        b[i] = X; b[i + 4] = Y;
    }
    if (!solve_lu(A, b, 8)) return false;
    for (int i = 0; i < 8; i++) H[i] = b[i];
    H[8] = 1;
    return true;
}

__device__ bool affine_3pt(const double *src, const double *dst, double *H)
{
    for (int c = 0; c < 2; c++) {
        double A[9], b[3];
        for (int i = 0; i < 3; i++) {
            A[i * 3] = src[2 * i]; A[i * 3 + 1] = src[2 * i + 1]; A[i * 3 + 2] = 1;
            b[i] = dst[2 * i + c];
        }
        if (!solve_lu(A, b, 3)) return false;
        H[3 * c] = b[0]; H[3 * c + 1] = b[1]; H[3 * c + 2] = b[2];
    }
    H[6] = 0; H[7] = 0; H[8] = 1;
    return true;
}

// perspective_4pt with one matrix ROW per lane: 8 consecutive lanes (a shuffle group of width 8) solve one hypothesis.  Every
// element sees exactly the operations of solve_lu above, in the same order (the row updates of one elimination step are independent
// of each other; the pivot is the first row of maximal |a| like the serial search with its strict >), so the result is bit-identical
// to the serial routine -- but the 8 x 8 system lives in registers instead of a dynamically indexed local array and a step's seven
// row updates run side by side.  All 32 lanes of a warp must call this together (idle groups pass any finite data).
__device__ __forceinline__ bool solve8_rows(double (&a)[8], double b, int r, double (&x)[8])
{
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        // pivot: the first row >= i with the largest |a[i]|
        double best = r >= i ? fabs(a[i]) : -1.0;
        int k = r;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o, 8);
            const int ok_ = __shfl_xor_sync(0xffffffffu, k, o, 8);
            if (ob > best || (ob == best && ok_ < k)) { best = ob; k = ok_; }
        }
        if (best < DBL_EPSILON * 100) ok = false;
        // swap rows i and k (lane i <-> lane k)
#pragma unroll
        for (int c = 0; c < 8; c++) {
            const double fk = __shfl_sync(0xffffffffu, a[c], k, 8), fi = __shfl_sync(0xffffffffu, a[c], i, 8);
            if (k != i) { if (r == i) a[c] = fk; else if (r == k) a[c] = fi; }
        }
        {
            const double fk = __shfl_sync(0xffffffffu, b, k, 8), fi = __shfl_sync(0xffffffffu, b, i, 8);
            if (k != i) { if (r == i) b = fk; else if (r == k) b = fi; }
        }
        // eliminate column i below the pivot row
        const double pii = __shfl_sync(0xffffffffu, a[i], i, 8);
        const double d = __ddiv_rn(-1.0, pii);
        const double alpha = __dmul_rn(a[i], d);
#pragma unroll
        for (int c = i + 1; c < 8; c++) {
            const double pc = __shfl_sync(0xffffffffu, a[c], i, 8);
            if (r > i) a[c] = __dadd_rn(a[c], __dmul_rn(alpha, pc));
        }
        const double pb = __shfl_sync(0xffffffffu, b, i, 8);
        if (r > i) b = __dadd_rn(b, __dmul_rn(alpha, pb));
    }
    // back substitution: row i waits for x[i + 1 .. 7]
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        double sacc = b;
#pragma unroll
        for (int c = i + 1; c < 8; c++) sacc = __dsub_rn(sacc, __dmul_rn(a[c], x[c]));
        const double xi = __ddiv_rn(sacc, a[i]);
        x[i] = __shfl_sync(0xffffffffu, xi, i, 8);
    }
    return ok;
}
__device__ __forceinline__ bool perspective_4pt_rows(double sx, double sy, double X, double Y, int r, double (&x)[8])
{
    // lane r holds the source / target point r & 3: rows 0-3 are the x equations, rows 4-7 the y equations of the four points
    const double T = r < 4 ? X : Y;
    double a[8];
    a[0] = r < 4 ? sx : 0.0; a[1] = r < 4 ? sy : 0.0; a[2] = r < 4 ? 1.0 : 0.0;
    a[3] = r < 4 ? 0.0 : sx; a[4] = r < 4 ? 0.0 : sy; a[5] = r < 4 ? 0.0 : 1.0;
    a[6] = __dmul_rn(-sx, T); a[7] = __dmul_rn(-sy, T);
    return solve8_rows(a, T, r, x);
}

#define HYP_THREADS 512
__global__ void __launch_bounds__(HYP_THREADS) k_hypotheses(const EgoParams p)
{
    const int b = blockIdx.x;
    __shared__ int s_idx[MD_MAX_HYP * 4 + 31];
    const int M = p.M[b];
    const int nh = p.mode == MD_EGO_FIRST4 ? 1 : p.iters;
    const int ndraw = nh * p.minimal;
    if (threadIdx.x == 0 && M >= p.minimal) {
        if (p.mode == MD_EGO_FIRST4) {
            for (int k = 0; k < 4; k++) s_idx[k] = k;
        } else {
            // srand(seed), 310 discarded outputs (10 rounds), then the draws, a round of 31 at a time
            uint32_t seed = p.seed0 + (uint32_t)b + (p.pair_ctr ? (uint32_t)*p.pair_ctr : 0u);
            if (seed == 0) seed = 1;
            uint32_t r[31];
            r[0] = seed;
#pragma unroll
            for (int i = 1; i < 31; i++) {
                const int32_t prev = (int32_t)r[i - 1];
                const long long hi = prev / 127773, lo = prev % 127773;
                long long word = 16807 * lo - 2836 * hi;
                if (word < 0) word += 2147483647;
                r[i] = (uint32_t)(int32_t)word;
            }
            for (int i = 0; i < 10; i++) glibc_rand_round31(r, nullptr);
            for (int i = 0; i < ndraw; i += 31) glibc_rand_round31(r, s_idx + i);
        }
    }
    __syncthreads();
    if (p.mode != MD_EGO_FIRST4 && M >= p.minimal)
        for (int i = threadIdx.x; i < ndraw; i += blockDim.x) s_idx[i] = s_idx[i] % M;       // rand() % M, all threads
    __syncthreads();
    if (p.mode == MD_EGO_RANSAC_AFFINE) {                    // two 3 x 3 solves per hypothesis: one thread each
        for (int j = threadIdx.x; j < nh; j += blockDim.x) {
            double *H = p.hyp + ((size_t)b * p.iters + j) * 9;
            bool ok = false;
            if (M >= p.minimal) {
                double s[8], d[8];
                for (int k = 0; k < p.minimal; k++) {
                    int c = p.kept_idx[(size_t)b * p.P + s_idx[j * p.minimal + k]];
                    float2 a = src_point(p, b, c), q = p.next[(size_t)b * p.P + c];
                    s[2 * k] = a.x; s[2 * k + 1] = a.y; d[2 * k] = q.x; d[2 * k + 1] = q.y;
                }
                double Hk[9];
                ok = affine_3pt(s, d, Hk);
                if (ok) for (int i = 0; i < 9; i++) H[i] = Hk[i];
            }
            if (!ok) for (int i = 0; i < 9; i++) H[i] = 0;
            p.hyp_valid[b * p.iters + j] = ok ? 1 : 0;
        }
        return;
    }
    // homography: 8 lanes per hypothesis, one matrix row each (whole warps run the solver together)
    const int r = threadIdx.x & 7;
    for (int j0 = 0; j0 < nh; j0 += HYP_THREADS / 8) {
        const int j = j0 + (threadIdx.x >> 3);
        const bool act = j < nh && M >= p.minimal;
        double sx = (r & 1) ? 1.0 : 0.0, sy = (r & 2) ? 1.0 : 0.0, X = sx, Y = sy;      // idle groups: a harmless unit square
        if (act) {
            const int c = p.kept_idx[(size_t)b * p.P + s_idx[j * 4 + (r & 3)]];
            const float2 a = src_point(p, b, c), q = p.next[(size_t)b * p.P + c];
            sx = a.x; sy = a.y; X = q.x; Y = q.y;
        }
        double x[8];
        const bool ok = perspective_4pt_rows(sx, sy, X, Y, r, x) && act;
        if (j < nh) {
            double *H = p.hyp + ((size_t)b * p.iters + j) * 9;
            double hr = 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) if (r == i) hr = x[i];
            H[r] = ok ? hr : 0.0;
            if (r == 0) { H[8] = ok ? 1.0 : 0.0; p.hyp_valid[b * p.iters + j] = ok ? 1 : 0; }
        }
    }
}

// division-free inlier test, oracle operation order: |H x - w X|^2 < thr^2 w^2
__device__ __forceinline__ bool is_inlier(const double *H, double x, double y, double X, double Y, double thr2)
{
    double w = __dadd_rn(__dadd_rn(__dmul_rn(H[6], x), __dmul_rn(H[7], y)), H[8]);
    double ex = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(H[0], x), __dmul_rn(H[1], y)), H[2]), __dmul_rn(X, w));
    double ey = __dsub_rn(__dadd_rn(__dadd_rn(__dmul_rn(H[3], x), __dmul_rn(H[4], y)), H[5]), __dmul_rn(Y, w));
    return __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)) < __dmul_rn(thr2, __dmul_rn(w, w));
}

__global__ void __launch_bounds__(256) k_score(const EgoParams p)
{
    const int b = blockIdx.y;
    __shared__ double sH[MD_MAX_HYP * 9];
    __shared__ int sV[MD_MAX_HYP];
    __shared__ int sC[MD_MAX_HYP];
    const int M = p.M[b];
    if (M < p.minimal) return;
    for (int i = threadIdx.x; i < p.iters * 9; i += blockDim.x) sH[i] = p.hyp[(size_t)b * p.iters * 9 + i];
    for (int i = threadIdx.x; i < p.iters; i += blockDim.x) { sV[i] = p.hyp_valid[b * p.iters + i]; sC[i] = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int Mr = (M + 31) & ~31;      // whole warps iterate together (ballot)
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < Mr; m += gridDim.x * blockDim.x) {
        double x = 0, y = 0, X = 0, Y = 0;
        const bool act = m < M;
        if (act) {
            int c = p.kept_idx[(size_t)b * p.P + m];
            float2 a = src_point(p, b, c), q = p.next[(size_t)b * p.P + c];
            x = a.x; y = a.y; X = q.x; Y = q.y;
        }
        for (int j = 0; j < p.iters; j++) {
            if (!sV[j]) continue;
            bool inl = act && is_inlier(sH + j * 9, x, y, X, Y, p.thr2);
            unsigned bal = __ballot_sync(0xffffffffu, inl);
            if (lane == 0 && bal) atomicAdd(&sC[j], __popc(bal));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < p.iters; i += blockDim.x)
        if (sC[i]) atomicAdd(&p.counts[b * p.iters + i], sC[i]);
}

__device__ __forceinline__ int best_hypothesis(const int *counts, int iters, int &best_cnt)
{
    int best = -1;
    best_cnt = 0;
    for (int j = 0; j < iters; j++)
        if (counts[j] > best_cnt) { best_cnt = counts[j]; best = j; }   // first best wins (outlier_detector.cpp:300)
    return best;
}

// normal-equation sums over the winner's inliers (normalised coordinates), one partial vector per block
__global__ void __launch_bounds__(256) k_accum(const EgoParams p)
{
    const int b = blockIdx.y;
    const int M = p.M[b];
    double S[NSUM];
#pragma unroll
    for (int i = 0; i < NSUM; i++) S[i] = 0;
    int best_cnt;
    const int best = M >= p.minimal ? best_hypothesis(p.counts + b * p.iters, p.iters, best_cnt) : -1;
    if (best >= 0 && best_cnt >= p.minimal) {
        double H[9];
        for (int i = 0; i < 9; i++) H[i] = p.hyp[((size_t)b * p.iters + best) * 9 + i];
        const double cx = 0.5 * p.w, cy = 0.5 * p.h, is = 1.0 / (0.5 * (p.w > p.h ? p.w : p.h));
        for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
            int c = p.kept_idx[(size_t)b * p.P + m];
            float2 a = src_point(p, b, c), q = p.next[(size_t)b * p.P + c];
            bool inl = is_inlier(H, a.x, a.y, q.x, q.y, p.thr2);
            if (p.inlier_mask) p.inlier_mask[(size_t)b * p.P + c] = inl ? 1 : 0;
            if (!inl) continue;
            double x = (a.x - cx) * is, y = (a.y - cy) * is, X = (q.x - cx) * is, Y = (q.y - cy) * is;
            double xx = x * x, xy = x * y, yy = y * y, R = X * X + Y * Y;
            S[0] += xx; S[1] += xy; S[2] += x; S[3] += yy; S[4] += y; S[5] += 1.0;
            S[6] += xx * X; S[7] += xy * X; S[8] += x * X; S[9] += yy * X; S[10] += y * X; S[11] += X;
            S[12] += xx * Y; S[13] += xy * Y; S[14] += x * Y; S[15] += yy * Y; S[16] += y * Y; S[17] += Y;
            S[18] += xx * R; S[19] += xy * R; S[20] += yy * R; S[21] += x * R; S[22] += y * R;
        }
    }
    __shared__ double sw[8][NSUM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NSUM; i++) {
        double v = S[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sw[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < NSUM) {
        double v = 0;
        for (int w = 0; w < 8; w++) v += sw[w][threadIdx.x];
        p.partial[((size_t)b * p.nblk_acc + blockIdx.x) * NSUM + threadIdx.x] = v;
    }
}

__device__ bool invert3(const double *S, double *D)
{
    double c0 = __dsub_rn(__dmul_rn(S[4], S[8]), __dmul_rn(S[5], S[7]));
    double c1 = __dsub_rn(__dmul_rn(S[3], S[8]), __dmul_rn(S[5], S[6]));
    double c2 = __dsub_rn(__dmul_rn(S[3], S[7]), __dmul_rn(S[4], S[6]));
    double d = __dadd_rn(__dsub_rn(__dmul_rn(S[0], c0), __dmul_rn(S[1], c1)), __dmul_rn(S[2], c2));
    if (d == 0) { for (int i = 0; i < 9; i++) D[i] = 0; return false; }
    d = __ddiv_rn(1.0, d);
    double t[9];
    t[0] = __dmul_rn(c0, d);
    t[1] = __dmul_rn(__dsub_rn(__dmul_rn(S[2], S[7]), __dmul_rn(S[1], S[8])), d);
    t[2] = __dmul_rn(__dsub_rn(__dmul_rn(S[1], S[5]), __dmul_rn(S[2], S[4])), d);
    t[3] = __dmul_rn(__dsub_rn(__dmul_rn(S[5], S[6]), __dmul_rn(S[3], S[8])), d);
    t[4] = __dmul_rn(__dsub_rn(__dmul_rn(S[0], S[8]), __dmul_rn(S[2], S[6])), d);
    t[5] = __dmul_rn(__dsub_rn(__dmul_rn(S[2], S[3]), __dmul_rn(S[0], S[5])), d);
    t[6] = __dmul_rn(c2, d);
    t[7] = __dmul_rn(__dsub_rn(__dmul_rn(S[1], S[6]), __dmul_rn(S[0], S[7])), d);
    t[8] = __dmul_rn(__dsub_rn(__dmul_rn(S[0], S[4]), __dmul_rn(S[1], S[3])), d);
    for (int i = 0; i < 9; i++) D[i] = t[i];
    return true;
}

__global__ void __launch_bounds__(32) k_solve(const EgoParams p)
{
    const int b = blockIdx.x;
    __shared__ double S[NSUM];
    if (threadIdx.x < NSUM) {
        double v = 0;
        const double *pp = p.partial + (size_t)b * p.nblk_acc * NSUM + threadIdx.x;
        int k = 0;
        for (; k + 8 <= p.nblk_acc; k += 8) {                 // eight loads in flight, added in the same fixed order
            double t[8];
#pragma unroll
            for (int u = 0; u < 8; u++) t[u] = __ldg(pp + (size_t)(k + u) * NSUM);
#pragma unroll
            for (int u = 0; u < 8; u++) v += t[u];
        }
        for (; k < p.nblk_acc; k++) v += __ldg(pp + (size_t)k * NSUM);
        S[threadIdx.x] = v;
    }
    __syncwarp();
    const int M = p.M[b];
    // the 8 x 8 normal equations of the homography refit, one row per lane (lanes 0-7; the other lanes solve an identity): same
    // arithmetic as solve_lu on the matrix thread 0 used to assemble
    double xr[8];
    bool ok_rows = false;
    if (p.mode == MD_EGO_RANSAC_HOMOGRAPHY) {                 // warp-uniform
        const int r = threadIdx.x & 7;
        double a[8], rhs;
#pragma unroll
        for (int c = 0; c < 8; c++) a[c] = r == c ? 1.0 : 0.0;
        rhs = 0.0;
        if (threadIdx.x < 8) {
            const double u[6] = {S[0], S[1], S[2], S[3], S[4], S[5]};
            const double c6[6] = {-S[6], -S[7], -S[8], -S[12], -S[13], -S[14]};
            const double c7[6] = {-S[7], -S[9], -S[10], -S[13], -S[15], -S[16]};
            const double rr[8] = {S[8], S[10], S[11], S[14], S[16], S[17], -S[21], -S[22]};
            const int sym[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                if (r != q) continue;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double v = 0.0;
                    if (q < 6) {
                        if (c < 6) v = (q / 3 == c / 3) ? u[sym[q % 3][c % 3]] : 0.0;
                        else v = c == 6 ? c6[q] : c7[q];
                    } else if (q == 6) v = c < 6 ? c6[c] : (c == 6 ? S[18] : S[19]);
                    else v = c < 6 ? c7[c] : (c == 6 ? S[19] : S[20]);
                    a[c] = v;
                }
                rhs = rr[q];
            }
        }
        ok_rows = solve8_rows(a, rhs, r, xr);
    }
    if (threadIdx.x != 0) return;
    double H[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    int inl = 0, valid = 0;
    if (p.mode == MD_EGO_FIRST4) {
        if (M >= 4 && p.hyp_valid[b * p.iters]) {
            for (int i = 0; i < 9; i++) H[i] = p.hyp[(size_t)b * p.iters * 9 + i];
            inl = 4; valid = 1;
        }
    } else if (M >= p.minimal) {
        int best_cnt;
        const int best = best_hypothesis(p.counts + b * p.iters, p.iters, best_cnt);
        if (best >= 0 && best_cnt >= p.minimal) {
            inl = best_cnt; valid = 1;
            double Hn[9];
            bool ok;
            if (p.mode == MD_EGO_RANSAC_AFFINE) {
                double N[9] = {S[0], S[1], S[2], S[1], S[3], S[4], S[2], S[4], S[5]}, N2[9];
                for (int i = 0; i < 9; i++) N2[i] = N[i];
                double r[3] = {S[8], S[10], S[11]}, r2[3] = {S[14], S[16], S[17]};
                ok = solve_lu(N, r, 3) && solve_lu(N2, r2, 3);
                Hn[0] = r[0]; Hn[1] = r[1]; Hn[2] = r[2]; Hn[3] = r2[0]; Hn[4] = r2[1]; Hn[5] = r2[2];
                Hn[6] = 0; Hn[7] = 0; Hn[8] = 1;
            } else {
                ok = ok_rows;
                for (int i = 0; i < 8; i++) Hn[i] = xr[i];
                Hn[8] = 1;
            }
            if (ok) {
                const double cx = 0.5 * p.w, cy = 0.5 * p.h, sc = 0.5 * (p.w > p.h ? p.w : p.h), is = 1.0 / sc;
                double A[9], R[9];
                for (int i = 0; i < 3; i++) {
                    A[i * 3 + 0] = Hn[i * 3 + 0] * is;
                    A[i * 3 + 1] = Hn[i * 3 + 1] * is;
                    A[i * 3 + 2] = Hn[i * 3 + 2] - (Hn[i * 3 + 0] * cx + Hn[i * 3 + 1] * cy) * is;
                }
                for (int j = 0; j < 3; j++) {
                    R[0 * 3 + j] = sc * A[0 * 3 + j] + cx * A[2 * 3 + j];
                    R[1 * 3 + j] = sc * A[1 * 3 + j] + cy * A[2 * 3 + j];
                    R[2 * 3 + j] = A[2 * 3 + j];
                }
                for (int i = 0; i < 9; i++) H[i] = R[i] / R[8];
            } else {
                for (int i = 0; i < 9; i++) H[i] = p.hyp[((size_t)b * p.iters + best) * 9 + i];
            }
        }
    }
    double Hi[9];
    invert3(H, Hi);
    for (int i = 0; i < 9; i++) { p.H[b * 9 + i] = H[i]; p.Hinv[b * 9 + i] = Hi[i]; }
    p.inliers[b] = inl;
    p.valid[b] = valid;
    if (p.stat_inliers && inl) atomicAdd(p.stat_inliers, (unsigned long long)inl);
}

cudaError_t launch_ego(const EgoParams &p, int pairs, cudaStream_t s)
{
    dim3 gs(p.nblk_scan, pairs);
    k_keep_count<<<gs, 256, 0, s>>>(p);
    k_scan<<<pairs, 1024, 0, s>>>(p);
    k_compact<<<gs, 256, 0, s>>>(p);
    k_hypotheses<<<pairs, HYP_THREADS, 0, s>>>(p);
    if (p.mode != MD_EGO_FIRST4) {
        int nb = (p.P + 255) / 256;
        if (nb > 592) nb = 592;
        k_score<<<dim3(nb, pairs), 256, 0, s>>>(p);
        k_accum<<<dim3(p.nblk_acc, pairs), 256, 0, s>>>(p);
    }
    k_solve<<<pairs, 32, 0, s>>>(p);
    MD_COUNT_LAUNCH(p.mode != MD_EGO_FIRST4 ? 7 : 5);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------
// Trajectory bookkeeping of calculateOpticalFlowTrajectory (common/src/optical_flow_calculator.cpp:178-241):
// a tracked point that stays inside the 10 px margin extends its trajectory and moves on; otherwise it is frozen.
__global__ void __launch_bounds__(256) k_traj_init(float2 *pts_cur, float2 *traj, int32_t *len, int P, int F, int ps, int gy)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    float2 pt = make_float2((float)(ps * (k / gy)), (float)(ps * (k % gy)));
    pts_cur[k] = pt;
    traj[(size_t)k * F] = pt;
    for (int f = 1; f < F; f++) traj[(size_t)k * F + f] = make_float2(0.f, 0.f);
    len[k] = 1;
}

__global__ void __launch_bounds__(256) k_traj_step(float2 *pts_cur, const float2 *next, const uint8_t *status, float2 *traj,
                                                   int32_t *len, int P, int F, int w, int h)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= P) return;
    if (!status[k]) return;
    float2 q = next[k];
    if (q.x > 10.0f && q.y > 10.0f && q.x < (float)(w - 10) && q.y < (float)(h - 10)) {
        int n = len[k];
        if (n < F) { traj[(size_t)k * F + n] = q; len[k] = n + 1; }
        pts_cur[k] = q;
    }
}

cudaError_t launch_traj_init(float2 *pts_cur, float2 *traj, int32_t *len, int P, int F, int ps, int gy, cudaStream_t s)
{
    k_traj_init<<<(P + 255) / 256, 256, 0, s>>>(pts_cur, traj, len, P, F, ps, gy);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
cudaError_t launch_traj_step(float2 *pts_cur, const float2 *next, const uint8_t *status, float2 *traj, int32_t *len,
                             int P, int F, int w, int h, cudaStream_t s)
{
    k_traj_step<<<(P + 255) / 256, 256, 0, s>>>(pts_cur, next, status, traj, len, P, F, w, h);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

// the context's pair counter lives on the device: bumped at the end of every batch (inside the captured graph as well)
__global__ void k_advance_pairs(unsigned long long *ctr, int pairs) { *ctr += (unsigned long long)pairs; }
__global__ void k_set_pairs(unsigned long long *ctr, unsigned long long v) { *ctr = v; }
cudaError_t launch_set_pairs(unsigned long long *ctr, unsigned long long v, cudaStream_t s)
{
    k_set_pairs<<<1, 1, 0, s>>>(ctr, v);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
cudaError_t launch_advance_pairs(unsigned long long *ctr, int pairs, cudaStream_t s)
{
    k_advance_pairs<<<1, 1, 0, s>>>(ctr, pairs);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}
