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
    return (size_t)(kmax + 2) * (8 * 8 + 4 * 4) + 1024;
}

struct PriceArgs {
    const double* gb;               // [B] beamlet gradient of the last evaluation
    const int32_t* beam_ptr;        // [nb+1]
    const int32_t* cand;            // [ncand]
    const vmat_price_row_t* rows;   // [ncand*M]
    double C, C2, C3, b;
    int32_t M, N, kmax;             // M: network rows (descriptors per candidate)
    int32_t variant;                // VMAT_PRICE_SAMPLING / _CLASSIC / _SHORT
    int32_t out_stride;             // entries per candidate in l / r
    int32_t sink_last_only;         // classic, after trailing rows without nodes: the sink scans the last node only
    double* p;                      // [ncand]
    int32_t* l;                     // [ncand*out_stride]
    int32_t* r;                     // [ncand*out_stride]
    int32_t* len;                   // [ncand] entries written (nullptr: not wanted)
    int32_t* scratch;               // [ncand*M*kmax*3]: predecessor (global node index), l, r of every node
    int32_t* gstart;                // [ncand*(M+1)]: global index of the first node of every row, then of the sink
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

__device__ __forceinline__ uint64_t bits_below(int n) { return n <= 0 ? 0ull : (n >= 64 ? ~0ull : ((1ull << n) - 1ull)); }

// The three generations of the reference differ in how the network is wired; the nodes of
// a candidate are numbered globally as there (0 = source, rows consecutive from 1, then the sink):
//   sampling  greedyVMATsampling.py:618-789  a row is relaxed over the previous row; untouched weights are inf;
//             tie-break and DoseSide terms; the sink entry is dropped from the result
//   classic   greedyVMAT.py:469-636          the relaxation window is the previous row shifted by one node (it starts
//             at the node BEFORE the row's first and drops its last), the recorded predecessor is the node AFTER the
//             window position, the sink scans the window plus the row's last node, untouched weights are 0.0
//   short     greedyVMATshort.py:456-616     shifted window as well, predecessor = window position, no tie-break, dose
//             terms indexed by raw position (row 0) or by width only (other rows), untouched weights are 0.0
// In the shifted variants the extra window element in front is the "carry": the source for the first relaxed row,
// afterwards the last node of the row before the previous one.
__global__ void __launch_bounds__(kPriceThreads) k5_price_dp(const PriceArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int kmax = a.kmax, kp = kmax + 2;              // previous-row arrays hold the carry at index -1
    double* dbl = reinterpret_cast<double*>(smem_raw);
    PriceNodeSmem S;
    S.w_prev = dbl + 1; S.w_cur = dbl + kp; S.lf = dbl + kp + kmax; S.rf = dbl + kp + 2 * kmax;
    S.dose = dbl + kp + 3 * kmax; S.c3s = dbl + kp + 4 * kmax; S.tie = dbl + kp + 5 * kmax; S.side = dbl + kp + 6 * kmax;
    int32_t* ints = reinterpret_cast<int32_t*>(dbl + kp + 7 * kmax);
    S.l_prev = ints + 1; S.r_prev = ints + kp + 1; S.l_cur = ints + 2 * kp; S.r_cur = ints + 2 * kp + kmax;
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
    const bool sampling = a.variant == VMAT_PRICE_SAMPLING, classic = a.variant == VMAT_PRICE_CLASSIC,
               shortv = a.variant == VMAT_PRICE_SHORT;
    const int shift = sampling ? 0 : 1, dad_bias = classic ? 1 : 0, sink_ext = classic ? 1 : 0;
    int32_t* gstart = a.gstart + (size_t)c * (M + 1);
    int gpos = 1, gprev = 1;                             // global index of the first node of the current / previous row
    if (tid == 0) { S.w_prev[-1] = 0.0; S.l_prev[-1] = 0; S.r_prev[-1] = 0; }      // the source node

    for (int m = 0; m < M; ++m) {
        const vmat_price_row_t R = a.rows[(size_t)c * M + m];
        // ---- enumerate the nodes of this row: for each l, r runs over an integer range or falls back
        //      (sampling: the midpoint, :672-680,715-724; classic: r0 .. the row's first l, greedyVMAT.py:539-540;
        //      short: nothing)
        if (tid == 0) {
            int total = 0;
            for (int k = 0; k < R.n_l; ++k) {
                node_base[k] = total;
                const double lval = R.l_first + (double)k;
                const double r0 = ceil(fmax(lval + 1.0, R.r_lo));
                const double r1 = floor(R.r_hi);
                int cnt;
                if (r1 >= r0) cnt = (int)(r1 - r0) + 1;
                else if (sampling) cnt = R.n_r_mid;
                else if (classic) cnt = max(0, (int)(R.r_mid + 1.0 - r0));
                else cnt = 0;
                total += cnt;
            }
            node_base[R.n_l] = total;
            n_cur_s = total;
            gstart[m] = gpos;
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
            const double rval = ((r1 >= r0 || !sampling) ? r0 : R.r_mid) + (double)(k - node_base[li]);
            const double lp1 = __dadd_rn(lval, 1.0);
            const double width = __dsub_rn(rval, lval);
            // beamGrad entries of the dose term: set bits of `sel`, read at bg[base + y]
            uint64_t sel = 0;
            int base = 0;
            bool open;                                   // the reference's condition for a dose / C3 term
            if (!shortv) {
                const int y_lo = (int)ceil(lp1), y_hi = (int)floor(rval);       // open y range [y_lo, y_hi)
                if (y_hi > y_lo) sel = R.vb_mask & bits_below(y_hi) & ~bits_below(y_lo);
                base = R.grad_offset - R.vb_min;
                open = sel != 0;
            } else if (m == 0) {                         // beamGrad[l+1:r]: raw positions (greedyVMATshort.py:534)
                const int y_lo = (int)lp1, y_hi = min((int)rval, nbl);
                if (y_hi > y_lo) sel = bits_below(y_hi) & ~bits_below(y_lo);
                open = lp1 <= rval - 1.0;
            } else {                                     // beamGrad[lml+1 : (r-l)+lml]: the width only (:574-579)
                const int n = (int)width - 1;
                base = R.grad_offset + 1;
                if (n > 0) sel = bits_below(min(n, nbl - base));
                open = n >= 1;
            }
            const int nsel = __popcll(sel);
            double side = 0.0;
            if (sampling) {      // DoseSide: raw leaf positions index beamGrad (quirk Q5); zero for integer leaves
                int i_l = (int)floor(lp1), i_r = (int)ceil(rval);
                i_l = min(max(i_l, 0), nbl - 1); i_r = min(max(i_r, 0), nbl - 1);
                side = -__dadd_rn(__dmul_rn(__dsub_rn(ceil(lp1), lp1), bg[i_l]), __dmul_rn(__dsub_rn(rval, floor(rval)), bg[i_r]));
            }
            double dose = 0.0;
            if (nsel > 0) dose = -numpy_order_sum(bg, base, sel, nsel);
            else if (shortv && open) dose = -0.0;        // the negated sum of an empty slice
            const double tie = __dmul_rn(10E-10, width);
            S.lf[k] = lval; S.rf[k] = rval;
            S.l_cur[k] = (int32_t)lval;      // int arrays truncate toward zero (quirk Q7)
            S.r_cur[k] = (int32_t)rval;
            if (m == 0) {
                double wgt = 0.0;
                const double pen = __dmul_rn(C, __dsub_rn(__dmul_rn(C2, width), __dmul_rn(C3b, width)));
                if (sampling) {
                    if (lp1 < rval) {
                        if (nsel > 0) wgt = __dadd_rn(__dadd_rn(__dsub_rn(pen, dose), tie), side);
                        else          wgt = __dadd_rn(__dadd_rn(pen, tie), side);
                    }
                } else if (classic) {                    // greedyVMAT.py:546-556
                    if (lp1 < rval && nsel > 0) wgt = __dadd_rn(__dsub_rn(pen, dose), tie);
                } else {                                 // greedyVMATshort.py:531-537
                    if (open) wgt = __dsub_rn(pen, dose);
                }
                S.w_cur[k] = wgt;
            } else {
                S.dose[k] = dose;
                S.c3s[k] = open ? __dmul_rn(C3b, width) : 0.0;
                S.tie[k] = tie;
                S.side[k] = side;
            }
        }
        __syncthreads();
        int32_t* sc = a.scratch + ((size_t)c * M + m) * kmax * 3;
        if (m > 0) {
            // ---- relax: one warp per node, lanes stride over the window (:743-750); window position j is the
            //      previous-row index j - shift (-1 = the carry), i.e. global node gprev + j - shift
            for (int k = warp; k < ncur; k += nwarps) {
                const double lval = S.lf[k], rval = S.rf[k];
                const double dose = S.dose[k], c3s = S.c3s[k], tie = S.tie[k], side = S.side[k];
                double best = INFINITY;
                int bestj = 0x7fffffff;
                for (int j = lane; j < nprev; j += 32) {
                    const int e = j - shift;
                    const double pl = (double)S.l_prev[e], pr = (double)S.r_prev[e];
                    double lam = __dadd_rn(fabs(__dsub_rn(pl, lval)), fabs(__dsub_rn(pr, rval)));
                    lam = __dsub_rn(lam, __dmul_rn(2.0, fmax(0.0, __dsub_rn(pl, rval))));
                    lam = __dsub_rn(lam, __dmul_rn(2.0, fmax(0.0, __dsub_rn(lval, fabs(pr)))));
                    double wgt = __dmul_rn(C, __dsub_rn(__dmul_rn(C2, lam), c3s));
                    wgt = __dsub_rn(wgt, dose);
                    if (!shortv) wgt = __dadd_rn(wgt, tie);
                    if (sampling) wgt = __dadd_rn(wgt, side);
                    const double tot = __dadd_rn(S.w_prev[e], wgt);
                    if (bestj == 0x7fffffff || tot < best) { best = tot; bestj = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
                    if (oj != 0x7fffffff && (bestj == 0x7fffffff || ov < best || (ov == best && oj < bestj))) { best = ov; bestj = oj; }
                }
                if (lane == 0) { S.w_cur[k] = best; sc[k * 3 + 0] = gprev + bestj - shift + dad_bias; }
            }
        }
        __syncthreads();
        // the next carry is the last node of the row that is about to be replaced (for row 0 it stays the source)
        double cw = 0.0; int32_t cl = 0, cr = 0;
        if (tid == 0 && m > 0 && nprev > 0) { cw = S.w_prev[nprev - 1]; cl = S.l_prev[nprev - 1]; cr = S.r_prev[nprev - 1]; }
        __syncthreads();
        for (int k = tid; k < ncur; k += blockDim.x) {
            if (m == 0) sc[k * 3 + 0] = 0;               // dadnetwork stays 0 for the first row: the path ends there
            sc[k * 3 + 1] = S.l_cur[k];
            sc[k * 3 + 2] = S.r_cur[k];
            S.w_prev[k] = S.w_cur[k];
            S.l_prev[k] = S.l_cur[k];
            S.r_prev[k] = S.r_cur[k];
        }
        if (tid == 0) {
            n_prev_s = ncur;
            if (m > 0 && nprev > 0) { S.w_prev[-1] = cw; S.l_prev[-1] = cl; S.r_prev[-1] = cr; }
        }
        gprev = gpos;
        gpos += ncur;
        __syncthreads();
    }

    // ---- sink: minimum of w + C*(C2*(r-l)) over its window, LAST minimum wins (:763-770); the sink starts at
    //      the untouched weight (inf, or 0.0 in the earlier generations: then only paths <= 0 are taken)
    const int nlast = n_prev_s;
    const int nwin = nlast + sink_ext;
    double best = INFINITY;
    int bestk = -1;
    for (int k = (a.sink_last_only ? nwin - 1 : 0) + tid; k < nwin; k += blockDim.x) {
        const int e = k - shift;
        const double tot = __dadd_rn(S.w_prev[e], __dmul_rn(C, __dmul_rn(C2, (double)(S.r_prev[e] - S.l_prev[e]))));
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
        gstart[M] = gpos;                                // the sink
        int t = 0;                                       // predecessor of the sink (0: none)
        const double init = sampling ? INFINITY : 0.0;
        if (bestk >= 0 && best <= init) { a.p[c] = best; t = gprev + bestk - shift; }
        else a.p[c] = nan("");                           // the reference leaves p unassigned here (UnboundLocalError)
        // ---- backtrack through the recorded predecessors until the source (:771-786); the earlier generations
        //      keep the sink's (0, 0) entry at the end of the leaf vectors
        int32_t* lo_ = a.l + (size_t)c * a.out_stride;
        int32_t* ro_ = a.r + (size_t)c * a.out_stride;
        int n = 0;
        if (!sampling) { lo_[n] = 0; ro_[n] = 0; ++n; }
        while (t != 0 && n < a.out_stride) {
            int row = M - 1;
            while (row > 0 && gstart[row] > t) --row;
            const int32_t* scr = a.scratch + ((size_t)c * M + row) * kmax * 3 + (size_t)(t - gstart[row]) * 3;
            lo_[n] = scr[1]; ro_[n] = scr[2]; ++n;
            t = scr[0];
        }
        for (int i = 0; i < n / 2; ++i) {                // collected from the sink backwards
            const int32_t tl = lo_[i], tr = ro_[i];
            lo_[i] = lo_[n - 1 - i]; ro_[i] = ro_[n - 1 - i];
            lo_[n - 1 - i] = tl; ro_[n - 1 - i] = tr;
        }
        if (a.len) a.len[c] = n;
    }
}

}  // namespace vmat
