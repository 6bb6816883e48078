// K6: k-medoid insertion costs (kmedoids.py:32-60).  For every candidate point c:
//   cost[c] = sum_j min(curmin[j], ||p_c - p_j||_2)
// where curmin[j] is the distance from point j to its nearest already-chosen medoid.
// This is distancemin(kappa + [c], ba) of the reference without rebuilding the full
// pairwise distance matrix for every candidate.
//
// One CTA per candidate.  Up to kKmSequentialMax points the sum runs strictly left to right
// in one thread per candidate (the reference's builtin ``sum`` order, bit-exact); beyond that
// a fixed-shape tree (deterministic; exact whenever the distances are integers).
#pragma once
#include "vmat_common.cuh"

namespace vmat {

constexpr int kKmThreads = 256;
constexpr int64_t kKmSequentialMax = 4096;

__device__ __forceinline__ double km_dist(const double* __restrict__ Pc, const double* __restrict__ P, int dim,
                                          int64_t c, int64_t j) {
    double s = 0.0;
    for (int d = 0; d < dim; ++d) {
        const double df = __dsub_rn(Pc[c * dim + d], P[j * dim + d]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    return __dsqrt_rn(s);
}

// candidates Pc[gridDim.x] against the points P[n] (a rank's shard of the point set in the
// multi-GPU use; the caller adds the per-rank costs)
__global__ void __launch_bounds__(kKmThreads) k6_kmedoid_cost(int64_t n, int dim, const double* __restrict__ Pc,
                                                             const double* __restrict__ P,
                                                             const double* __restrict__ curmin,
                                                             double* __restrict__ cost) {
    __shared__ double sh[kKmThreads];
    const int64_t c = blockIdx.x;
    const int tid = threadIdx.x;
    if (n <= kKmSequentialMax) {
        // distances in parallel into shared memory, then one thread adds them in index order
        double total = 0.0;
        for (int64_t j0 = 0; j0 < n; j0 += kKmThreads) {
            const int64_t j = j0 + tid;
            sh[tid] = (j < n) ? fmin(curmin[j], km_dist(Pc, P, dim, c, j)) : 0.0;
            __syncthreads();
            if (tid == 0) {
                const int lim = (int)min((int64_t)kKmThreads, n - j0);
                for (int k = 0; k < lim; ++k) total = __dadd_rn(total, sh[k]);
            }
            __syncthreads();
        }
        if (tid == 0) cost[c] = total;
        return;
    }
    double v = 0.0;
    for (int64_t j = tid; j < n; j += kKmThreads) v += fmin(curmin[j], km_dist(Pc, P, dim, c, j));
    sh[tid] = v;
    __syncthreads();
    for (int o = kKmThreads / 2; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    if (tid == 0) cost[c] = sh[0];
}

// Large point sets (n > kKmSequentialMax): kKmTile candidates per CTA, so a point and its curmin
// are fetched once for kKmTile distances (the one-candidate form is bound by L2 traffic, n^2 x 32 B),
// and the square root is skipped wherever the squared distance already reaches curmin^2 (rounded
// up, so the test can only err towards taking the root): min(curmin, sqrt(s)) is curmin there,
// bit for bit.  Same thread <-> point mapping and tree as the one-candidate form: same sums.
constexpr int kKmTile = 8;

template <int DIM>
__global__ void __launch_bounds__(kKmThreads) k6_kmedoid_cost_tiled(int64_t n, int64_t nc, const double* __restrict__ Pc,
                                                                   const double* __restrict__ P,
                                                                   const double* __restrict__ curmin,
                                                                   double* __restrict__ cost) {
    __shared__ double sc[kKmTile][DIM];
    __shared__ double sh[kKmTile][kKmThreads];
    const int64_t c0 = (int64_t)blockIdx.x * kKmTile;
    const int tid = threadIdx.x;
    if (tid < kKmTile * DIM) {
        const int64_t c = min(c0 + tid / DIM, nc - 1);
        sc[tid / DIM][tid % DIM] = Pc[c * DIM + tid % DIM];
    }
    __syncthreads();
    double v[kKmTile], cc[kKmTile][DIM];       // the tile's candidates live in registers
#pragma unroll
    for (int t = 0; t < kKmTile; ++t) {
        v[t] = 0.0;
#pragma unroll
        for (int d = 0; d < DIM; ++d) cc[t][d] = sc[t][d];
    }
    for (int64_t j = tid; j < n; j += kKmThreads) {
        double pj[DIM];
#pragma unroll
        for (int d = 0; d < DIM; ++d) pj[d] = P[j * DIM + d];
        const double cm = curmin[j];
        const double cm2 = __dmul_ru(cm, cm);
#pragma unroll
        for (int t = 0; t < kKmTile; ++t) {
            double s = 0.0;
#pragma unroll
            for (int d = 0; d < DIM; ++d) {
                const double df = __dsub_rn(cc[t][d], pj[d]);
                s = __dadd_rn(s, __dmul_rn(df, df));
            }
            v[t] += (s >= cm2) ? cm : fmin(cm, __dsqrt_rn(s));
        }
    }
#pragma unroll
    for (int t = 0; t < kKmTile; ++t) sh[t][tid] = v[t];
    __syncthreads();
    for (int o = kKmThreads / 2; o > 0; o >>= 1) {
        if (tid < o) {
#pragma unroll
            for (int t = 0; t < kKmTile; ++t) sh[t][tid] += sh[t][tid + o];
        }
        __syncthreads();
    }
    if (tid < kKmTile && c0 + tid < nc) cost[c0 + tid] = sh[tid][0];
}

// curmin[j] = min(curmin[j], ||m - p_j||): one more medoid (one more row of D[kappa, :], kmedoids.py:38-41).  Same
// operation order as km_dist and as numpy's sqrt(((d0*d0) + d1*d1) + d2*d2): the resident session (points and curmin
// stay on the device between insertion steps) and the host bookkeeping agree bit for bit.
__global__ void k6_update_curmin(int64_t n, int dim, const double* __restrict__ P, const double* __restrict__ medoid,
                                 double* __restrict__ curmin) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int d = 0; d < dim; ++d) {
        const double df = __dsub_rn(P[j * dim + d], medoid[d]);
        s = __dadd_rn(s, __dmul_rn(df, df));
    }
    curmin[j] = fmin(curmin[j], __dsqrt_rn(s));
}

// ---------------------------------------------------------------------------------
// K7: dose-volume histograms (printresults, greedyVMATsampling.py:888-912).  Integer counts with
// integer atomics: exact whatever the order.  Bin k holds edges[k] <= d < edges[k+1]; the last bin
// also takes d == edges[last] (np.histogram with explicit edges).
// ---------------------------------------------------------------------------------
template <typename AT>
__global__ void k7_dose_max(int64_t V, const AT* __restrict__ dose, unsigned long long* __restrict__ out) {
    double m = 0.0;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x)
        m = fmax(m, (double)dose[v]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    // non-negative doubles order like their bit patterns
    if ((threadIdx.x & 31) == 0) atomicMax(out, (unsigned long long)__double_as_longlong(m));
}

template <typename AT>
__global__ void k7_dvh_hist(int64_t V, const AT* __restrict__ dose, const unsigned long long* __restrict__ full_mask,
                            int nstructs, const double* __restrict__ edges, int nedges,
                            unsigned long long* __restrict__ hist) {
    const int nbins = nedges - 1;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const double d = (double)dose[v];
        if (!(d >= edges[0]) || d > edges[nedges - 1]) continue;
        int lo = 0, hi = nedges;                 // first edge > d
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (edges[mid] <= d) lo = mid + 1; else hi = mid; }
        int k = lo - 1;
        if (k >= nbins) k = nbins - 1;
        unsigned long long m = full_mask[v];
        while (m) {
            const int s = __ffsll((long long)m) - 1;
            m &= m - 1;
            if (s < nstructs) atomicAdd(&hist[(size_t)s * nbins + k], 1ull);
        }
    }
}

}  // namespace vmat
