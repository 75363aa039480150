// k_live.cu -- device side of the node's live path after the trajectories exist
// (MotionDetectionNode::imageCallback, ros/src/motion_detection_node.cpp:294-392):
//   * ordered compaction of the complete trajectories (optical_flow_calculator.cpp:244-254) and of the outlier points
//     (outlier_detector.cpp:318-324: the SECOND-TO-LAST point of every trajectory whose residual exceeds the threshold);
//   * FlowClusterer::clusterEuclidean (common/src/flow_clusterer.cpp:227-269) -- the greedy, order-dependent grouping --
//     restated as a parallel nearest-EARLIER-neighbour search plus a short sequential label pass;
//   * OpticalFlowVisualizer::showBoundingBoxes (common/src/optical_flow_visualizer.cpp:223-240): cv::boundingRect of the
//     rounded points of every cluster with more than 5 members, in cluster creation order.
#include <float.h>

#include "md_internal.h"

// ---- ordered compaction of flagged items: idx[0..total) = ascending indices i with flag(i) != 0 --------------------------
#define CMP_ITEMS 2048        // items per block (256 threads x 8)

struct FlagLenEq { const int32_t *len; int F; __device__ bool operator()(int i) const { return len[i] == F; } };
struct FlagU8 { const uint8_t *f; __device__ bool operator()(int i) const { return f[i] != 0; } };

template <class Flag>
__global__ void __launch_bounds__(256) k_cmp_count(Flag flag, int n, int *blockcnt)
{
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int base = blockIdx.x * CMP_ITEMS;
    int c = 0;
    for (int i = threadIdx.x; i < CMP_ITEMS; i += 256) {
        const int k = base + i;
        if (k < n && flag(k)) c++;
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) blockcnt[blockIdx.x] = s_cnt;
}

// exclusive scan of the block counts in place (one block), total -> *total
__global__ void __launch_bounds__(1024) k_cmp_scan(int *blockcnt, int nblk, int *total)
{
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nblk; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblk ? blockcnt[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if ((threadIdx.x & 31) >= o) x += y;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_warp[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, w, o);
                if (threadIdx.x >= o) w += y;
            }
            s_warp[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? s_warp[(threadIdx.x >> 5) - 1] : 0;
        const int carry = s_carry;
        if (i < nblk) blockcnt[i] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = carry + warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

template <class Flag>
__global__ void __launch_bounds__(256) k_cmp_scatter(Flag flag, int n, const int *blockoff, int *idx)
{
    // items are visited in ascending order: chunk c of the block holds items base + 256 c + t
    __shared__ int s_warp[8];
    __shared__ int s_run;
    if (threadIdx.x == 0) s_run = blockoff[blockIdx.x];
    __syncthreads();
    const int base = blockIdx.x * CMP_ITEMS, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < CMP_ITEMS / 256; c++) {
        const int k = base + c * 256 + threadIdx.x;
        const bool f = k < n && flag(k);
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_run;
        for (int w2 = 0; w2 < warp; w2++) off += s_warp[w2];
        if (f) idx[off + __popc(bal & ((1u << lane) - 1))] = k;
        __syncthreads();
        if (threadIdx.x == 0) {
            int t = 0;
            for (int w2 = 0; w2 < 8; w2++) t += s_warp[w2];
            s_run += t;
        }
        __syncthreads();
    }
}

template <class Flag>
static cudaError_t ordered_compact(Flag flag, int n, int *blockcnt, int *idx, int *total, cudaStream_t s)
{
    const int nblk = (n + CMP_ITEMS - 1) / CMP_ITEMS;
    k_cmp_count<<<nblk, 256, 0, s>>>(flag, n, blockcnt);
    k_cmp_scan<<<1, 1024, 0, s>>>(blockcnt, nblk, total);
    k_cmp_scatter<<<nblk, 256, 0, s>>>(flag, n, blockcnt, idx);
    MD_COUNT_LAUNCH(3);
    return cudaGetLastError();
}

// complete trajectories, in grid order (optical_flow_calculator.cpp:244-254)
__global__ void __launch_bounds__(256) k_traj_gather(const float2 *traj, const int *idx, const int *total, int F, float2 *out)
{
    const int T = *total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * F; i += gridDim.x * blockDim.x)
        out[i] = traj[(size_t)idx[i / F] * F + (i % F)];
}

// outlier_points.push_back(trajectories.at(i).at(size - 2)) (outlier_detector.cpp:322)
__global__ void __launch_bounds__(256) k_outlier_points(const float2 *traj_c, const int *oidx, const int *total, int F, float2 *pts)
{
    const int n = *total;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        pts[i] = traj_c[(size_t)oidx[i] * F + (F - 2)];
}

cudaError_t launch_compact_trajectories(const float2 *traj, const int32_t *len, int P, int F, int *blockcnt, int *idx, int *total,
                                        float2 *traj_c, cudaStream_t s)
{
    cudaError_t e = ordered_compact(FlagLenEq{len, F}, P, blockcnt, idx, total, s);
    if (e != cudaSuccess) return e;
    k_traj_gather<<<296, 256, 0, s>>>(traj, idx, total, F, traj_c);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

cudaError_t launch_compact_outliers(const float2 *traj_c, const uint8_t *outlier, int T, int F, int *blockcnt, int *oidx, int *total,
                                    float2 *pts, cudaStream_t s)
{
    cudaError_t e = ordered_compact(FlagU8{outlier}, T, blockcnt, oidx, total, s);
    if (e != cudaSuccess) return e;
    k_outlier_points<<<148, 256, 0, s>>>(traj_c, oidx, total, F, pts);
    MD_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

// ---- FlowClusterer::clusterEuclidean -----------------------------------------------------------------------------------
// Reference loop (flow_clusterer.cpp:230-259): point i joins the cluster whose closest member is nearest, if that distance
// is below the threshold; clusters are scanned in creation order with a strict '<', so among equally near clusters the
// oldest wins; otherwise the point founds a new cluster.  "Distance to a cluster" is the minimum over its members
// (point_cluster.cpp:26-38), hence: the winning cluster is the cluster of the nearest EARLIER point, ties resolved to the
// smallest cluster id.  The distance search is independent per point (parallel); only the label lookup is sequential.
#define CL_MAXC 8             // tie candidates kept per point; more -> the label pass rescans that point

// PointCluster::getDistance (point_cluster.cpp:62-65): float differences, float products and sum (no FMA on the
// reference's x86-64 build); the unqualified sqrt() on that float resolves to the float overload, so the distance the
// reference compares (against the threshold and between clusters, strict '<') is the f32-rounded root.  sqrt is monotonic:
// the minimum is taken on the squared distance; ties are decided on the ROUNDED root (two different squares can collapse to
// one root, and the reference then keeps the older cluster).
__device__ __forceinline__ float cl_d2(float2 a, float2 b)
{
    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
__device__ __forceinline__ float cl_dist(float2 a, float2 b) { return __fsqrt_rn(cl_d2(a, b)); }

// one warp per point i: min over j < i, then the (ascending) list of j attaining it
__global__ void __launch_bounds__(256) k_cl_nearest(const float2 *pts, const int *n_ptr, float *m2, int *cand, int *ncand)
{
    const int n = *n_ptr;
    const int lane = threadIdx.x & 31;
    for (int i = blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += gridDim.x * 8) {
        const float2 a = pts[i];
        float m = FLT_MAX;
        for (int j = lane; j < i; j += 32) m = fminf(m, cl_d2(a, pts[j]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
        m = __fsqrt_rn(m);                          // i == 0: root of FLT_MAX, unused (no candidates)
        int cnt = 0;
        for (int j0 = 0; j0 < i; j0 += 32) {
            const int j = j0 + lane;
            const bool hit = j < i && cl_dist(a, pts[j]) == m;
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (hit) {
                const int slot = cnt + __popc(bal & ((1u << lane) - 1));
                if (slot < CL_MAXC) cand[(size_t)i * CL_MAXC + slot] = j;
            }
            cnt += __popc(bal);
        }
        if (lane == 0) { m2[i] = m; ncand[i] = cnt; }
    }
}

// sequential label pass: one warp walks the points in order.  Each lane prefetches the search result of one point of a
// 32-point chunk; lane t then resolves point i0 + t from the labels of its tie candidates (shared memory when they fit)
// and the running cluster count travels from lane to lane.
__global__ void __launch_bounds__(32) k_cl_label(const float2 *pts, const int *n_ptr, const float *m2, const int *cand, const int *ncand,
                                                 double thr, int *label, int *nclusters, int use_smem)
{
    extern __shared__ int s_label[];
    const int n = *n_ptr;
    const int lane = threadIdx.x;
    int *lab = use_smem ? s_label : label;
    int ncl = 0;
    for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        int nc = 0, c[CL_MAXC];
        float m = FLT_MAX;
        bool join = false;
        if (i < n) {
            nc = ncand[i]; m = m2[i];
            const int4 *cp = reinterpret_cast<const int4 *>(cand + (size_t)i * CL_MAXC);
            const int4 a = cp[0], b = cp[1];
            c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
            join = nc > 0 && (double)m < thr;
        }
        const int cnt = min(32, n - i0);
        for (int t = 0; t < cnt; t++) {
            const int nct = __shfl_sync(0xffffffffu, nc, t);
            const bool jt = __shfl_sync(0xffffffffu, (int)join, t) != 0;
            int best = -1;
            if (jt && nct > CL_MAXC) {
                // more ties than candidate slots (rare): the warp rescans the earlier points of point i0 + t
                const int it = i0 + t;
                const float2 a = pts[it];
                const float mt = __shfl_sync(0xffffffffu, m, t);
                int bb = 0x7fffffff;
                for (int j = lane; j < it; j += 32)
                    if (cl_dist(a, pts[j]) == mt) bb = min(bb, lab[j]);
                best = __reduce_min_sync(0xffffffffu, bb);
            }
            if (lane == t) {
                if (jt && nct <= CL_MAXC) {
                    best = 0x7fffffff;
#pragma unroll
                    for (int q = 0; q < CL_MAXC; q++)
                        if (q < nc) best = min(best, lab[c[q]]);
                }
                if (!jt) best = ncl++;
                lab[i] = best;
            }
            ncl = __shfl_sync(0xffffffffu, ncl, t);
            if (!use_smem) __threadfence_block();
            __syncwarp();
        }
    }
    if (use_smem)
        for (int i = lane; i < n; i += 32) label[i] = s_label[i];
    if (lane == 0) *nclusters = ncl;
}

__global__ void __launch_bounds__(256) k_cl_init(const int *n_ptr, int *sizes, int *box)
{
    const int n = *n_ptr;       // at most n clusters
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        sizes[c] = 0;
        box[4 * c] = 0x7fffffff; box[4 * c + 1] = 0x7fffffff; box[4 * c + 2] = (int)0x80000000; box[4 * c + 3] = (int)0x80000000;
    }
}

// sizes + cv::boundingRect over cvRound()ed members (Mat(Point2f).copyTo(vector<Point>) converts with saturate_cast<int>)
__global__ void __launch_bounds__(256) k_cl_accumulate(const float2 *pts, const int *n_ptr, const int *label, int *sizes, int *box)
{
    const int n = *n_ptr;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = label[i];
        const int x = __float2int_rn(pts[i].x), y = __float2int_rn(pts[i].y);
        atomicAdd(&sizes[c], 1);
        atomicMin(&box[4 * c], x); atomicMin(&box[4 * c + 1], y);
        atomicMax(&box[4 * c + 2], x); atomicMax(&box[4 * c + 3], y);
    }
}

// clusters with more than 5 members, in creation order (flow_clusterer.cpp:262-267); rect = (tl.x, tl.y, br.x, br.y) with
// br = tl + size = max + 1 (the CSV row of MotionLogger::writeBoundingBox, motion_logger.cpp:43-47)
__global__ void __launch_bounds__(32) k_cl_select(const int *nclusters, const int *sizes, const int *box, int min_size, int *nout,
                                                  int32_t *out_box, int32_t *out_size, int32_t *out_id)
{
    if (threadIdx.x) return;
    const int nc = *nclusters;
    int k = 0;
    for (int c = 0; c < nc; c++)
        if (sizes[c] > min_size) {
            out_box[4 * k] = box[4 * c]; out_box[4 * k + 1] = box[4 * c + 1];
            out_box[4 * k + 2] = box[4 * c + 2] + 1; out_box[4 * k + 3] = box[4 * c + 3] + 1;
            out_size[k] = sizes[c];
            out_id[k] = c;
            k++;
        }
    *nout = k;
}

// ---- FlowClusterer::getClusters (common/src/flow_clusterer.cpp:178-227) ---------------------------------------------------
// Greedy FIRST-FIT grouping of flow vectors: a vector joins the first cluster (creation order) that has a member nearer than
// distance_threshold (VectorCluster::getClosestDistance, vector_cluster.cpp:25-37) AND a member -- not necessarily the same
// one -- whose orientation differs by less than angular_threshold (getClosestOrientation :38-50, getAngularDistance
// :121-127).  The choice depends on the labels of all earlier vectors, so the vectors are walked in order by ONE block; for
// every vector the block's threads test all earlier vectors in parallel and raise two flags per cluster in shared memory,
// then the smallest cluster id with both flags wins.  Orientations (getAngle :129-136) come from a parallel pre-pass.
// All arithmetic in f64 in the reference's operation order.
__global__ void __launch_bounds__(256) k_vc_angle(const double4 *vec, int n, double *ang)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double a = atan2(vec[i].w, vec[i].z);
        if (a < 0.0) a = __dadd_rn(a, 2 * 3.14159265358979323846);
        ang[i] = a;
    }
}

#define VC_THREADS 1024
__global__ void __launch_bounds__(VC_THREADS) k_vc_cluster(const double4 *vec, const double *ang, int n, double dthr, double athr,
                                                           int *label, int *nclusters)
{
    extern __shared__ uint8_t vc_flags[];          // [n] bit 0: a member is near, bit 1: a member is aligned
    __shared__ int s_best[VC_THREADS / 32];
    __shared__ int s_ncl;
    const int t = threadIdx.x;
    for (int k = t; k < n; k += VC_THREADS) vc_flags[k] = 0;
    if (t == 0) s_ncl = 0;
    __syncthreads();
    for (int i = 0; i < n; i++) {
        const double4 v = vec[i];
        const double ai = ang[i];
        for (int j = t; j < i; j += VC_THREADS) {
            const double4 u = vec[j];
            const double dx = __dsub_rn(v.x, u.x), dy = __dsub_rn(v.y, u.y);
            const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            const double da = __dsub_rn(ai, ang[j]);
            const double ad = fabs(atan2(sin(da), cos(da)));
            const int f = (d < dthr ? 1 : 0) | (ad < athr ? 2 : 0);
            if (f) atomicOr(reinterpret_cast<unsigned *>(vc_flags + (label[j] & ~3)), (unsigned)f << (8 * (label[j] & 3)));
        }
        __syncthreads();
        const int ncl = s_ncl;
        int best = 0x7fffffff;
        for (int k = t; k < ncl; k += VC_THREADS) {
            if (vc_flags[k] == 3 && k < best) best = k;
            vc_flags[k] = 0;
        }
        best = __reduce_min_sync(0xffffffffu, best);
        if ((t & 31) == 0) s_best[t >> 5] = best;
        __syncthreads();
        if (t < 32) {
            int b = s_best[t];
            b = __reduce_min_sync(0xffffffffu, b);
            if (t == 0) {
                if (b == 0x7fffffff) b = s_ncl++;
                label[i] = b;
            }
        }
        __syncthreads();
    }
    if (t == 0) *nclusters = s_ncl;
}

cudaError_t launch_cluster_vectors(const double *vec4, int n, double dthr, double athr, double *ang, int *label, int *nclusters,
                                   cudaStream_t s)
{
    const size_t smem = ((size_t)n + 3) / 4 * 4 + 16;
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(k_vc_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return e;
    k_vc_angle<<<(n + 255) / 256, 256, 0, s>>>(reinterpret_cast<const double4 *>(vec4), n, ang);
    k_vc_cluster<<<1, VC_THREADS, smem, s>>>(reinterpret_cast<const double4 *>(vec4), ang, n, dthr, athr, label, nclusters);
    MD_COUNT_LAUNCH(2);
    return cudaGetLastError();
}

cudaError_t launch_cluster(const float2 *pts, const int *n_ptr, int n_max, double thr, int min_size, float *m2, int *cand, int *ncand,
                           int *label, int *nclusters, int *sizes, int *box, int *nout, int32_t *out_box, int32_t *out_size,
                           int32_t *out_id, cudaStream_t s)
{
    int nb = (n_max + 7) / 8;
    if (nb > 148 * 8) nb = 148 * 8;
    if (nb < 1) nb = 1;
    k_cl_nearest<<<nb, 256, 0, s>>>(pts, n_ptr, m2, cand, ncand);
    const int use_smem = (size_t)n_max * sizeof(int) <= 200 * 1024 ? 1 : 0;
    const size_t lsm = use_smem ? (size_t)n_max * sizeof(int) : 0;
    cudaError_t e = cudaFuncSetAttribute(k_cl_label, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
    if (e != cudaSuccess) return e;
    k_cl_label<<<1, 32, lsm, s>>>(pts, n_ptr, m2, cand, ncand, thr, label, nclusters, use_smem);
    k_cl_init<<<148, 256, 0, s>>>(n_ptr, sizes, box);
    k_cl_accumulate<<<148, 256, 0, s>>>(pts, n_ptr, label, sizes, box);
    k_cl_select<<<1, 32, 0, s>>>(nclusters, sizes, box, min_size, nout, out_box, out_size, out_id);
    MD_COUNT_LAUNCH(5);
    return cudaGetLastError();
}
