// Restricted-master-problem fast path: while scipy's L-BFGS-B re-optimises the intensities of a
// FIXED set of apertures (solveRMC, greedyVMATsampling.py:858-885), D_i^T s_i (dose of control
// point i at unit intensity) and D_i^T diagmaker_i (its dZdK column, :251) do not change.  They
// are computed once per RMP with the sparse dose kernel and kept as dense columns; an evaluation
// is then  d = sum_k y_k A_k,  penalty,  g_k = <G_k, vg>  - two passes over 2*V*ncols values
// instead of two passes over the whole sparse matrix.  ("algorithmic shortcut".)
// Same penalty arithmetic as K1, fixed reduction shapes: deterministic.
#pragma once
#include "vmat_common.cuh"

namespace vmat {

template <typename AT>
struct RmpArgs {
    int64_t V, ld;                // voxels, column stride
    int32_t ncols, nchunks, chunk;
    const AT* A;                  // [ncols][ld] dose columns
    const AT* G;                  // [ncols][ld] gradient columns
    const double* yk;             // [ncols] intensities of the cached control points
    const int32_t* beams;         // [ncols] control point of every column
    const int32_t* slots;         // [ncols] where that control point's columns live in A / G
    const AT* thr; const AT* over; const AT* under;   // natural voxel order
    AT* d; AT* vg;                // [V]
    double* fpart;                // [gridDim.x of the dose kernel * warps] objective partials
    int32_t nfpart;
    double* gpart;                // [ncols][nchunks]
    double* result;               // [nb+1]: f then g
    unsigned int* done;
};

template <typename AT>
__global__ void __launch_bounds__(256) k8_rmp_dose(const RmpArgs<AT> a) {
    __shared__ double ys[512];
    __shared__ int32_t sl[512];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < a.ncols; k += 256) { ys[k] = a.yk[k]; sl[k] = a.slots[k]; }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * 256 + tid;
    double fterm = 0.0;
    if (v < a.V) {
        AT acc = (AT)0;
        const AT* col = a.A + v;
        for (int k = 0; k < a.ncols; ++k) acc = fma_acc(col[(int64_t)sl[k] * a.ld], (AT)ys[k], acc);
        const AT t = a.thr[v], ov = a.over[v], un = a.under[v];
        AT o = sub_rn(acc, t);  o = o > (AT)0 ? o : (AT)0;
        AT u = sub_rn(t, acc);  u = u > (AT)0 ? u : (AT)0;
        fterm = (double)add_rn(mul_rn(mul_rn(o, o), ov), mul_rn(mul_rn(u, u), un));
        a.d[v] = acc;
        a.vg[v] = mul_rn((AT)2, sub_rn(mul_rn(mul_rn((AT)2, o), ov), mul_rn(mul_rn((AT)2, u), un)));
    }
    fterm = warp_sum(fterm);
    if (lane == 0) a.fpart[(int64_t)blockIdx.x * 8 + warp] = fterm;
}

template <typename AT>
__global__ void __launch_bounds__(256) k8_rmp_grad(const RmpArgs<AT> a) {
    __shared__ double sh[256];
    __shared__ bool last;
    const int tid = threadIdx.x;
    const int c = blockIdx.x, k = blockIdx.y;
    const int64_t v0 = (int64_t)c * a.chunk, v1 = min(a.V, v0 + a.chunk);
    const AT* col = a.G + (int64_t)a.slots[k] * a.ld;
    double s = 0.0;
    for (int64_t v = v0 + tid; v < v1; v += 256) s += (double)col[v] * (double)a.vg[v];
    sh[tid] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) sh[tid] += sh[tid + o]; __syncthreads(); }
    if (tid == 0) {
        a.gpart[(int64_t)k * a.nchunks + c] = sh[0];
        __threadfence();
        last = atomicAdd(a.done, 1u) == gridDim.x * gridDim.y - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    // the last CTA adds the chunk partials of every column and the objective partials, in index order
    for (int kk = tid; kk < a.ncols; kk += 256) {
        double g = 0.0;
        for (int cc = 0; cc < a.nchunks; ++cc) g += __ldcg(a.gpart + (int64_t)kk * a.nchunks + cc);
        a.result[1 + a.beams[kk]] = g;
    }
    double f = 0.0;
    for (int i = tid; i < a.nfpart; i += 256) f += __ldcg(a.fpart + i);
    __syncthreads();
    sh[tid] = f;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (tid < o) sh[tid] += sh[tid + o]; __syncthreads(); }
    if (tid == 0) { a.result[0] = sh[0]; *a.done = 0u; }
}

}  // namespace vmat
