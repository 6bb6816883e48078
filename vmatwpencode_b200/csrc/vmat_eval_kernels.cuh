// Evaluation kernels of libvmat_b200 (sm_100a): dose + penalty (K1), transposed
// product (K2), deterministic reductions (K3), layout construction.
//
// Data layout in HBM ("row-sorted sliced ELL", SELL-32-4):
//   * rows (voxels for K1, (tile, beamlet) segments for K2) are sorted by length and cut
//     into slices of 32; lane i of a warp owns row i of its slice;
//   * a slice is a run of chunks; a chunk holds 4 nonzeros of each of the 32 rows: first
//     the 32 int4 index vectors, then the 32 value vectors (vmat_common.cuh: Store<>), so
//     every warp load is one fully coalesced 128-bit-per-lane transaction and the whole
//     slice is one contiguous stream;
//   * the vector being gathered (x for K1, a tile of vg for K2) is staged in shared memory
//     so the irregular accesses never leave the SM.
// Each lane accumulates its row sequentially in storage order: no shuffles, no atomics,
// bit-reproducible.
#pragma once
#include "vmat_common.cuh"

namespace vmat {

// ---------------------------------------------------------------------------------
// generic SELL row dot product: acc = sum_k val[k] * gather(idx[k]) over `nch` chunks
// ---------------------------------------------------------------------------------
// Software pipelined: two register stages of U chunks each; the loads of stage k+1 are in
// flight while stage k is consumed, so a warp keeps 2*U KB of the stream outstanding.
template <typename VT, int U>
struct Stage {
    uint4 idx[U];
    typename Store<VT>::Raw raw[U];
};

template <typename VT, int U>
__device__ __forceinline__ void stage_load(Stage<VT, U>& st, const uint4* __restrict__ p, int lane, int nvalid) {
    using S = Store<VT>;
    if (nvalid >= U) {                   // warp-uniform fast path: a full stage
#pragma unroll
        for (int u = 0; u < U; ++u) {
            st.idx[u] = ldg_stream16(p + u * S::kUnits + lane);
            st.raw[u] = S::load(p + u * S::kUnits, lane);
        }
        return;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (u < nvalid) {
            st.idx[u] = ldg_stream16(p + u * S::kUnits + lane);
            st.raw[u] = S::load(p + u * S::kUnits, lane);
        } else {                         // past the end of the slice: (index 0, value 0)
            st.idx[u] = make_uint4(0u, 0u, 0u, 0u);
            st.raw[u] = typename S::Raw{};
        }
    }
}

template <typename VT, typename AT, int U, typename Gather>
__device__ __forceinline__ void stage_fma(const Stage<VT, U>& st, AT& acc, Gather g) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
        AT v[4];
        Store<VT>::unpack(st.raw[u], v);
        acc = fma_acc(v[0], g(st.idx[u].x), acc);
        acc = fma_acc(v[1], g(st.idx[u].y), acc);
        acc = fma_acc(v[2], g(st.idx[u].z), acc);
        acc = fma_acc(v[3], g(st.idx[u].w), acc);
    }
}

template <typename VT, typename AT, int U, typename Gather>
__device__ __forceinline__ AT sell_row_dot(const uint4* __restrict__ p, int nch, int lane, Gather g) {
    constexpr int kStep = U * Store<VT>::kUnits;
    AT acc = (AT)0;
    const int ngrp = (nch + U - 1) / U;          // the last group may be partial
    Stage<VT, U> A, B;
    if (ngrp > 0) stage_load<VT, U>(A, p, lane, nch);
    int it = 0;
    for (; it + 2 <= ngrp; it += 2) {
        stage_load<VT, U>(B, p + kStep, lane, nch - (it + 1) * U);
        stage_fma<VT, AT, U>(A, acc, g);
        if (it + 2 < ngrp) stage_load<VT, U>(A, p + 2 * kStep, lane, nch - (it + 2) * U);
        stage_fma<VT, AT, U>(B, acc, g);
        p += 2 * kStep;
    }
    if (it < ngrp) stage_fma<VT, AT, U>(A, acc, g);
    return acc;
}

// ---------------------------------------------------------------------------------
// K4 (only when x does not fit in shared memory): x_j = y[beam(j)] * w_dose[j]
// ---------------------------------------------------------------------------------
template <typename AT>
__global__ void k4_expand_x(int32_t B, const double* __restrict__ y, const int32_t* __restrict__ beam_of,
                            const AT* __restrict__ w_dose, AT* __restrict__ x) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < B) x[j] = mul_rn((AT)y[beam_of[j]], w_dose[j]);
}

// ---------------------------------------------------------------------------------
// K1: d = D x, fused penalty epilogue (objective partials, voxel gradient)
//     replaces calcDose + calcGradientandObjValue (greedyVMATsampling.py:243-271)
// ---------------------------------------------------------------------------------
template <typename AT>
struct K1Args {
    const uint4* sell;            // SELL stream (voxel-major)
    const int64_t* slice_off;     // [nslices+1] in 16-byte units
    const int32_t* slice_rows;    // [nslices*32] voxel id of each lane, -1 = padding lane
    int32_t nslices;
    int32_t B;
    const double* y;              // [nb] intensities
    const int32_t* beam_of;       // [B]
    const AT* w_dose;             // [B]
    const AT* x_global;           // [B] (only read when XS == false)
    const AT* thr_s;              // penalty tables in slice order [nslices*32]
    const AT* over_s;
    const AT* under_s;
    AT* d;                        // [V] dose, natural voxel order
    AT* vg;                       // [V] voxel gradient
    double* fpart;                // [nslices] objective partial of every slice (fixed order -> deterministic)
    unsigned int* counter;        // dynamic slice scheduler
};

template <typename VT, typename AT, bool XS, int U, int TPB>
__global__ void __launch_bounds__(TPB, 1) k1_dose_penalty(const K1Args<AT> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AT* xs = reinterpret_cast<AT*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31;
    if (XS) {   // x_j = y[beam(j)] * w_dose[j]   (K4 folded into the prologue)
        for (int j = tid; j < a.B; j += blockDim.x) xs[j] = mul_rn((AT)a.y[a.beam_of[j]], a.w_dose[j]);
        __syncthreads();
    }
    auto gather = [&](uint32_t j) -> AT { return XS ? xs[j] : __ldg(a.x_global + j); };

    // dynamic scheduling, longest slices first.  Slice ids are claimed two slices ahead and the
    // slice's metadata one slice ahead, so neither the atomic nor the offset loads are exposed.
    const unsigned int ns = (unsigned)a.nslices;
    unsigned int s = 0, s1 = 0, s2 = 0;
    if (lane == 0) { s = atomicAdd(a.counter, 1u); s1 = atomicAdd(a.counter, 1u); }
    s = __shfl_sync(0xffffffffu, s, 0);
    s1 = __shfl_sync(0xffffffffu, s1, 0);
    int64_t off = 0, off_end = 0;
    if (s < ns) { off = a.slice_off[s]; off_end = a.slice_off[s + 1]; }
    while (s < ns) {
        if (lane == 0) s2 = atomicAdd(a.counter, 1u);
        int64_t noff = 0, noff_end = 0;
        if (s1 < ns) { noff = a.slice_off[s1]; noff_end = a.slice_off[s1 + 1]; }
        const int nch = (int)((off_end - off) / Store<VT>::kUnits);
        const int64_t slot = (int64_t)s * 32 + lane;
        const int32_t row = a.slice_rows[slot];
        const AT t = a.thr_s[slot], ov = a.over_s[slot], un = a.under_s[slot];
        const AT dose = sell_row_dot<VT, AT, U>(a.sell + off, nch, lane, gather);
        double fterm = 0.0;
        if (row >= 0) {
            // (d-t)+ and (t-d)+, objective term, vg = 2*(2*o*over - 2*u*under)  (:255-271)
            AT o = sub_rn(dose, t);  o = o > (AT)0 ? o : (AT)0;
            AT u = sub_rn(t, dose);  u = u > (AT)0 ? u : (AT)0;
            const AT fo = mul_rn(mul_rn(o, o), ov);
            const AT fu = mul_rn(mul_rn(u, u), un);
            fterm = (double)add_rn(fo, fu);
            const AT go = mul_rn(mul_rn((AT)2, o), ov);
            const AT gu = mul_rn(mul_rn((AT)2, u), un);
            a.d[row] = dose;
            a.vg[row] = mul_rn((AT)2, sub_rn(go, gu));
        }
        fterm = warp_sum(fterm);      // butterfly: same value and order in every lane, every run
        if (lane == 0) a.fpart[s] = fterm;
        s = s1; off = noff; off_end = noff_end;
        s1 = __shfl_sync(0xffffffffu, s2, 0);
    }
}

// ---------------------------------------------------------------------------------
// K2: partial[tile][beamlet] = sum over the tile's voxels of D[v, beamlet] * vg[v]
//     (transposed product over the pre-built beamlet-major layout; PPsubroutine :646)
// ---------------------------------------------------------------------------------
struct K2Task { int32_t tile, slice_begin, slice_end, pad; };

template <typename AT>
struct K2Args {
    const uint4* sell;            // SELL stream (beamlet-major, column-tiled)
    const int64_t* slice_off;     // [nslices+1]
    const int32_t* slice_rows;    // [nslices*32] beamlet id, -1 padding
    const K2Task* tasks;          // (tile, slice range), ordered by tile
    const int32_t* cta_task_ptr;  // [gridDim.x+1]: CTA c runs tasks [ptr[c], ptr[c+1]) - equal chunk counts
    const AT* vg;                 // [V + T] (zero padded so a tile copy may run past V)
    AT* partial;                  // [ntiles][B]
    int64_t V;
    int32_t B;
    int32_t T;                    // voxels per tile
    unsigned int* k1_counter;     // reset here for the next evaluation
};

// Persistent CTAs; the host cut the (tile, length)-ordered slice sequence into gridDim.x pieces of
// equal work, so there is no tail.  The vg tile is staged by one TMA bulk copy per tile change.
template <typename VT, typename AT, int U>
__global__ void __launch_bounds__(256) k2_beamlet_grad(const K2Args<AT> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AT* vt = reinterpret_cast<AT*>(smem_raw);
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    if (blockIdx.x == 0 && tid == 0) *a.k1_counter = 0u;
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    auto gather = [&](uint32_t j) -> AT { return vt[j]; };
    uint32_t parity = 0;
    int cur_tile = -1;
    const int t_begin = a.cta_task_ptr[blockIdx.x], t_end = a.cta_task_ptr[blockIdx.x + 1];
    for (int ti = t_begin; ti < t_end; ++ti) {
        const K2Task task = a.tasks[ti];
        if (task.tile != cur_tile) {
            __syncthreads();                         // everybody is done with the previous tile
            if (tid == 0) {
                const int64_t base = (int64_t)task.tile * a.T;
                const int n = (int)min((int64_t)a.T, a.V - base);
                const uint32_t bytes = (uint32_t)(((size_t)n * sizeof(AT) + 15) & ~(size_t)15);
                mbar_expect_tx(&bar, bytes);
                tma_bulk_g2s(vt, a.vg + base, bytes, &bar);
            }
            mbar_wait(&bar, parity);
            parity ^= 1u;
            cur_tile = task.tile;
        }
        AT* out = a.partial + (int64_t)task.tile * a.B;
        for (int s = task.slice_begin + warp; s < task.slice_end; s += nwarps) {
            const int64_t off = a.slice_off[s];
            const int nch = (int)((a.slice_off[s + 1] - off) / Store<VT>::kUnits);
            const int32_t row = a.slice_rows[(int64_t)s * 32 + lane];
            const AT acc = sell_row_dot<VT, AT, U>(a.sell + off, nch, lane, gather);
            if (row >= 0) out[row] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------
// K3: gb[j] = sum_t partial[t][j] (fixed order), f = sum fpart (fixed order),
//     g_i = sum_{j in beam i} w_grad[j] gb[j]   (aperturegradient, :272)
// phase: 0 = all, 1 = reductions into gbf only (multi-GPU, before the all-reduce),
//        2 = aperture gradient from the (all-reduced) gbf
// ---------------------------------------------------------------------------------
template <typename AT>
struct K3Args {
    const AT* partial;            // [ntiles][B]
    int32_t ntiles, B, nb;
    const int32_t* beam_ptr;      // [nb+1]
    const double* w_grad;         // [B]
    const double* fpart;
    int32_t nfpart;
    double* gbf;                  // [B+1]: gb then f
    double* result;               // [nb+1]: f then g
};

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
    // fixed-shape tree: deterministic for a given blockDim (256)
    const int tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
#pragma unroll
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) sh[tid] += sh[tid + o];
        __syncthreads();
    }
    const double r = sh[0];
    __syncthreads();
    return r;
}

template <typename AT>
__global__ void __launch_bounds__(256) k3_reduce(const K3Args<AT> a, int phase) {
    __shared__ double sh[256];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x == a.nb) {      // objective
        if (phase != 2) {
            double v = 0.0;
            for (int k = tid; k < a.nfpart; k += 256) v += a.fpart[k];
            v = block_sum_256(v, sh);
            if (tid == 0) a.gbf[a.B] = v;
            if (tid == 0 && phase == 0) a.result[0] = v;
        } else if (tid == 0) {
            a.result[0] = a.gbf[a.B];
        }
        return;
    }
    const int lo = a.beam_ptr[blockIdx.x], hi = a.beam_ptr[blockIdx.x + 1];
    double contrib = 0.0;
    for (int j0 = lo; j0 < hi; j0 += 64) {
        // 4 threads per beamlet: thread q sums tiles t = q, q+4, ...; combined (q0+q1)+(q2+q3)
        const int j = j0 + (tid >> 2), q = tid & 3;
        double s = 0.0;
        if (phase != 2) {
            if (j < hi)
                for (int t = q; t < a.ntiles; t += 4) s += (double)a.partial[(int64_t)t * a.B + j];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (j < hi && q == 0) a.gbf[j] = s;
        } else if (j < hi) {
            s = a.gbf[j];
        }
        if (phase != 1 && j < hi && q == 0) contrib += a.w_grad[j] * s;
    }
    if (phase != 1) {
        const double g = block_sum_256(contrib, sh);
        if (tid == 0) a.result[1 + blockIdx.x] = g;
    }
}

// ---------------------------------------------------------------------------------
// layout construction (plan build time)
// ---------------------------------------------------------------------------------
// first position in each beamlet row whose voxel id is >= t*T, for t = 0..ntiles
__global__ void k_segment_bounds(int32_t B, int32_t ntiles, int32_t T, const int64_t* __restrict__ browptr,
                                 const int32_t* __restrict__ bcol, int64_t* __restrict__ segstart) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)B * (ntiles + 1);
    if (gid >= total) return;
    const int32_t b = (int32_t)(gid / (ntiles + 1)), t = (int32_t)(gid % (ntiles + 1));
    int64_t lo = browptr[b], hi = browptr[b + 1];
    const int64_t key = (int64_t)t * T;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)bcol[mid] < key) lo = mid + 1; else hi = mid;
    }
    segstart[gid] = lo;
}

template <typename T> struct StoreCvt;
template <> struct StoreCvt<float>  { template <typename S> __device__ static float  cvt(S v) { return (float)v; } };
template <> struct StoreCvt<double> { template <typename S> __device__ static double cvt(S v) { return (double)v; } };
template <> struct StoreCvt<__nv_bfloat16> {
    template <typename S> __device__ static __nv_bfloat16 cvt(S v) { return __float2bfloat16_rn((float)v); }
};

// one warp per slice: lane copies its segment (CSR positions seg_start .. +seg_len) into
// the chunked layout, indices rebased by seg_colbase; padding = (index 0, value 0)
template <typename VT, typename ST>
__global__ void k_fill_sell(int32_t nslices, const int64_t* __restrict__ slice_off,
                            const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                            const int32_t* __restrict__ seg_colbase,
                            const int32_t* __restrict__ col, const ST* __restrict__ val,
                            uint4* __restrict__ sell) {
    using S = Store<VT>;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int64_t off = slice_off[gw];
    const int nch = (int)((slice_off[gw + 1] - off) / S::kUnits);
    const int64_t slot = gw * 32 + lane;
    const int64_t st = seg_start[slot];
    const int32_t len = seg_len[slot], cb = seg_colbase[slot];
    for (int c = 0; c < nch; ++c) {
        uint4* chunk = sell + off + (int64_t)c * S::kUnits;
        int32_t ix[4];
        VT vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = c * 4 + k;
            if (e < len) { ix[k] = col[st + e] - cb; vv[k] = StoreCvt<VT>::cvt(val[st + e]); }
            else         { ix[k] = 0;                vv[k] = StoreCvt<VT>::cvt(0.0f); }
        }
        chunk[lane] = make_uint4((uint32_t)ix[0], (uint32_t)ix[1], (uint32_t)ix[2], (uint32_t)ix[3]);
        if constexpr (sizeof(VT) == 4) {
            float4 f = make_float4(vv[0], vv[1], vv[2], vv[3]);
            chunk[32 + lane] = *reinterpret_cast<uint4*>(&f);
        } else if constexpr (sizeof(VT) == 2) {
            __nv_bfloat16 h[4] = {vv[0], vv[1], vv[2], vv[3]};
            reinterpret_cast<uint2*>(chunk + 32)[lane] = *reinterpret_cast<uint2*>(h);
        } else {
            double2 a2 = make_double2(vv[0], vv[1]), b2 = make_double2(vv[2], vv[3]);
            chunk[32 + lane] = *reinterpret_cast<uint4*>(&a2);
            chunk[64 + lane] = *reinterpret_cast<uint4*>(&b2);
        }
    }
}

template <typename AT>
__global__ void k_cast_from_f64(int64_t n, const double* __restrict__ src, AT* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (AT)src[i];
}
template <typename AT>
__global__ void k_cast_to_f64(int64_t n, const AT* __restrict__ src, double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = (double)src[i];
}

}  // namespace vmat
