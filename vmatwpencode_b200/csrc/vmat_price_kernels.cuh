// K5: aperture pricing - the layered shortest path over (leaf row, l, r) nodes of
// PPsubroutine (greedyVMATsampling.py:618-789), one CTA per candidate control point.
//
// Every cost is formed in float64 with the reference's operation order and without FMA
// contraction (the DP compares sums that differ by the 1e-9*(r-l) tie-break term, :693,744),
// the per-node dose term is summed in numpy's pairwise order, interior rows keep the FIRST
// minimum (np.argmin, :748) and the sink keeps the LAST (``<=``, :767).  A row of the
// network lives in shared memory; each node's minimum over the previous row is taken by one
// warp with a shuffle (value, index) reduction.
#pragma once
#include "vmat_common.cuh"
#include "../../include/vmat_b200.h"

namespace vmat {

constexpr int kPriceThreads = 512;

__host__ __device__ inline int32_t price_max_nodes(int32_t N) {
    // l in [-1, N-1] -> N+1 values, r in (l, N]; plus up to two fallback nodes per l
    return (N + 1) * (N + 2) / 2 + 2 * (N + 2);
}

struct PriceNodeSmem {   // structure-of-arrays views carved out of dynamic shared memory
    double* w_prev; double* w_cur;
    double* lf; double* rf; double* dose; double* c3s; double* tie; double* side;
    int32_t* l_prev; int32_t* r_prev; int32_t* l_cur; int32_t* r_cur;
};
__host__ __device__ inline size_t price_smem_bytes(int32_t kmax) {
    return (size_t)kmax * (8 * 8 + 4 * 4) + 1024;
}

struct PriceArgs {
    const double* gb;               // [B] beamlet gradient of the last evaluation
    const int32_t* beam_ptr;        // [nb+1]
    const int32_t* cand;            // [ncand]
    const vmat_price_row_t* rows;   // [ncand*M]
    double C, C2, C3, b;
    int32_t M, N, kmax;
    double* p;                      // [ncand]
    int32_t* l;                     // [ncand*M]
    int32_t* r;                     // [ncand*M]
    int32_t* scratch;               // [ncand*M*kmax*3]: dad, l, r of every node
};

// numpy's contiguous float64 add-reduce order for n <= 128 elements (see
// oracle/vmat_oracle.py: numpy_pairwise_sum): the n selected beamGrad entries are the set
// bits of `sel` (y indices), read at bg[base + y].
__device__ __forceinline__ double numpy_order_sum(const double* __restrict__ bg, int base, uint64_t sel, int n) {
    if (n < 8) {
        double s = 0.0;
        while (sel) { const int y = __ffsll((long long)sel) - 1; sel &= sel - 1; s = __dadd_rn(s, bg[base + y]); }
        return s;
    }
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { const int y = __ffsll((long long)sel) - 1; sel &= sel - 1; acc[k] = bg[base + y]; }
    int done = 8;
    const int full = n - (n & 7);
    while (done < full) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { const int y = __ffsll((long long)sel) - 1; sel &= sel - 1; acc[k] = __dadd_rn(acc[k], bg[base + y]); }
        done += 8;
    }
    double s = __dadd_rn(__dadd_rn(__dadd_rn(acc[0], acc[1]), __dadd_rn(acc[2], acc[3])),
                         __dadd_rn(__dadd_rn(acc[4], acc[5]), __dadd_rn(acc[6], acc[7])));
    while (sel) { const int y = __ffsll((long long)sel) - 1; sel &= sel - 1; s = __dadd_rn(s, bg[base + y]); }
    return s;
}

__global__ void __launch_bounds__(kPriceThreads) k5_price_dp(const PriceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int kmax = a.kmax;
    double* dbl = reinterpret_cast<double*>(smem_raw);
    PriceNodeSmem S;
    S.w_prev = dbl; S.w_cur = dbl + kmax; S.lf = dbl + 2 * kmax; S.rf = dbl + 3 * kmax;
    S.dose = dbl + 4 * kmax; S.c3s = dbl + 5 * kmax; S.tie = dbl + 6 * kmax; S.side = dbl + 7 * kmax;
    int32_t* ints = reinterpret_cast<int32_t*>(dbl + 8 * kmax);
    S.l_prev = ints; S.r_prev = ints + kmax; S.l_cur = ints + 2 * kmax; S.r_cur = ints + 3 * kmax;
    __shared__ int32_t node_base[66];
    __shared__ int32_t n_cur_s, n_prev_s;
    __shared__ double red_v[kPriceThreads / 32];
    __shared__ int32_t red_i[kPriceThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int c = blockIdx.x;
    const int beam = a.cand[c];
    const double* bg = a.gb + a.beam_ptr[beam];
    const int nbl = a.beam_ptr[beam + 1] - a.beam_ptr[beam];
    const double C = a.C, C2 = a.C2, C3b = __dmul_rn(a.C3, a.b);
    const int M = a.M;

    for (int m = 0; m < M; ++m) {
        const vmat_price_row_t R = a.rows[(size_t)c * M + m];
        // ---- enumerate the nodes of this row: for each l, r runs over an integer range or
        //      falls back to the midpoint (:672-680,715-724)
        if (tid == 0) {
            int total = 0;
            for (int k = 0; k < R.n_l; ++k) {
                node_base[k] = total;
                const double lval = R.l_first + (double)k;
                const double r0 = ceil(fmax(lval + 1.0, R.r_lo));
                const double r1 = floor(R.r_hi);
                const int cnt = (r1 >= r0) ? (int)(r1 - r0) + 1 : R.n_r_mid;
                total += cnt;
            }
            node_base[R.n_l] = total;
            n_cur_s = total;
        }
        __syncthreads();
        const int ncur = n_cur_s;
        const int nprev = (m == 0) ? 0 : n_prev_s;
        // ---- per-node terms that do not depend on the previous row
        for (int k = tid; k < ncur; k += blockDim.x) {
            int li = 0;   // which l this node belongs to: last index with node_base <= k
            { int lo = 0, hi = R.n_l; while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (node_base[mid] <= k) lo = mid; else hi = mid; } li = lo; }
            const double lval = R.l_first + (double)li;
            const double r0 = ceil(fmax(lval + 1.0, R.r_lo));
            const double r1 = floor(R.r_hi);
            const double rval = ((r1 >= r0) ? r0 : R.r_mid) + (double)(k - node_base[li]);
            const double lp1 = __dadd_rn(lval, 1.0);
            const int y_lo = (int)ceil(lp1), y_hi = (int)floor(rval);       // open y range [y_lo, y_hi)
            uint64_t sel = 0;
            if (y_hi > y_lo) {
                const uint64_t upto_hi = (y_hi >= 64) ? ~0ull : ((1ull << y_hi) - 1ull);
                const uint64_t upto_lo = (y_lo <= 0) ? 0ull : ((y_lo >= 64) ? ~0ull : ((1ull << y_lo) - 1ull));
                sel = R.vb_mask & upto_hi & ~upto_lo;
            }
            const int nsel = __popcll(sel);
            const double width = __dsub_rn(rval, lval);
            // DoseSide: raw leaf positions index beamGrad (quirk Q5); zero for integer leaves
            int i_l = (int)floor(lp1), i_r = (int)ceil(rval);
            i_l = min(max(i_l, 0), nbl - 1); i_r = min(max(i_r, 0), nbl - 1);
            const double side = -__dadd_rn(__dmul_rn(__dsub_rn(ceil(lp1), lp1), bg[i_l]),
                                           __dmul_rn(__dsub_rn(rval, floor(rval)), bg[i_r]));
            double dose = 0.0;
            if (nsel > 0) dose = -numpy_order_sum(bg, R.grad_offset - R.vb_min, sel, nsel);
            const double tie = __dmul_rn(10E-10, width);
            S.lf[k] = lval; S.rf[k] = rval;
            S.l_cur[k] = (int32_t)lval;      // int arrays truncate toward zero (quirk Q7)
            S.r_cur[k] = (int32_t)rval;
            if (m == 0) {
                double wgt = 0.0;
                if (lp1 < rval) {
                    const double pen = __dmul_rn(C, __dsub_rn(__dmul_rn(C2, width), __dmul_rn(C3b, width)));
                    if (nsel > 0) wgt = __dadd_rn(__dadd_rn(__dsub_rn(pen, dose), tie), side);
                    else          wgt = __dadd_rn(__dadd_rn(pen, tie), side);
                }
                S.w_cur[k] = wgt;
            } else {
                S.dose[k] = dose;
                S.c3s[k] = (nsel > 0) ? __dmul_rn(C3b, width) : 0.0;
                S.tie[k] = tie;
                S.side[k] = side;
            }
        }
        __syncthreads();
        int32_t* sc = a.scratch + ((size_t)c * M + m) * kmax * 3;
        if (m > 0) {
            // ---- relax: one warp per node, lanes stride over the previous row (:743-750)
            for (int k = warp; k < ncur; k += nwarps) {
                const double lval = S.lf[k], rval = S.rf[k];
                const double dose = S.dose[k], c3s = S.c3s[k], tie = S.tie[k], side = S.side[k];
                double best = INFINITY;
                int bestj = 0x7fffffff;
                for (int j = lane; j < nprev; j += 32) {
                    const double pl = (double)S.l_prev[j], pr = (double)S.r_prev[j];
                    double lam = __dadd_rn(fabs(__dsub_rn(pl, lval)), fabs(__dsub_rn(pr, rval)));
                    lam = __dsub_rn(lam, __dmul_rn(2.0, fmax(0.0, __dsub_rn(pl, rval))));
                    lam = __dsub_rn(lam, __dmul_rn(2.0, fmax(0.0, __dsub_rn(lval, fabs(pr)))));
                    double wgt = __dmul_rn(C, __dsub_rn(__dmul_rn(C2, lam), c3s));
                    wgt = __dadd_rn(__dadd_rn(__dsub_rn(wgt, dose), tie), side);
                    const double tot = __dadd_rn(S.w_prev[j], wgt);
                    if (bestj == 0x7fffffff || tot < best) { best = tot; bestj = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                    if (oj != 0x7fffffff && (bestj == 0x7fffffff || ov < best || (ov == best && oj < bestj))) { best = ov; bestj = oj; }
                }
                if (lane == 0) { S.w_cur[k] = best; sc[k * 3 + 0] = bestj; }
            }
        }
        __syncthreads();
        for (int k = tid; k < ncur; k += blockDim.x) {
            if (m == 0) sc[k * 3 + 0] = -1;
            sc[k * 3 + 1] = S.l_cur[k];
            sc[k * 3 + 2] = S.r_cur[k];
            S.w_prev[k] = S.w_cur[k];
            S.l_prev[k] = S.l_cur[k];
            S.r_prev[k] = S.r_cur[k];
        }
        if (tid == 0) n_prev_s = ncur;
        __syncthreads();
    }

    // ---- sink: minimum of w + C*(C2*(r-l)), LAST minimum wins (:763-770)
    const int nlast = n_prev_s;
    double best = INFINITY;
    int bestk = -1;
    for (int k = tid; k < nlast; k += blockDim.x) {
        const double tot = __dadd_rn(S.w_prev[k], __dmul_rn(C, __dmul_rn(C2, (double)(S.r_prev[k] - S.l_prev[k]))));
        if (bestk < 0 || tot <= best) { best = tot; bestk = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bestk, o);
        if (ok >= 0 && (bestk < 0 || ov < best || (ov == best && ok > bestk))) { best = ov; bestk = ok; }
    }
    if (lane == 0) { red_v[warp] = best; red_i[warp] = bestk; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < nwarps; ++w) {
            if (red_i[w] >= 0 && (bestk < 0 || red_v[w] < best || (red_v[w] == best && red_i[w] > bestk))) { best = red_v[w]; bestk = red_i[w]; }
        }
        a.p[c] = best;
        // ---- backtrack through the stored predecessors (:771-786)
        int k = bestk;
        for (int m = M - 1; m >= 0; --m) {
            const int32_t* sc = a.scratch + ((size_t)c * M + m) * kmax * 3;
            a.l[(size_t)c * M + m] = sc[k * 3 + 1];
            a.r[(size_t)c * M + m] = sc[k * 3 + 2];
            k = sc[k * 3 + 0];
        }
    }
}

}  // namespace vmat
