/*
 * TEST INFRASTRUCTURE - C restatement (float64) of the reference evaluation, used as the
 * checker at sizes numpy is slow at and as the CPU baseline of bench.py.  Not product code.
 *
 * Follows greedyVMATsampling.py: calcDose :243-251 (d = sum_i y_i D_i^T s_i, here in the
 * global form d = D x with x_j = y[beam(j)] w_dose[j]), calcGradientandObjValue :254-272
 * (objective, voxelgradient with the reference's factor 4) and PPsubroutine :646
 * (beamGrad = D_i vg, for every control point: gb = D^T vg), aperture gradient :272.
 * Parity pinning: checked against oracle/vmat_oracle.py, which is itself pinned against the
 * reference's own functions (see its header) - tests/test_oracle_c.py.
 *
 * D is given twice, like scipy would hold DlistT (voxel-major gather) and Dlist
 * (beamlet-major gather), so both products parallelise over rows without atomics.
 */
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* returns the objective; writes d[V], vg[V], gb[B], g[nb] */
double oracle_eval(int64_t V, int64_t B, int32_t nb, const int32_t* beam_ptr,
                   const int64_t* vrowptr, const int32_t* vcol, const double* vval,
                   const int64_t* browptr, const int32_t* bcol, const double* bval,
                   const double* thr, const double* over, const double* under,
                   const double* y, const double* w_dose, const double* w_grad,
                   double* x, double* d, double* vg, double* fterm, double* gb, double* g,
                   int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    for (int32_t i = 0; i < nb; ++i)
        for (int32_t j = beam_ptr[i]; j < beam_ptr[i + 1]; ++j) x[j] = y[i] * w_dose[j];

#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t v = 0; v < V; ++v) {
        double s = 0.0;
        for (int64_t k = vrowptr[v]; k < vrowptr[v + 1]; ++k) s += vval[k] * x[vcol[k]];
        d[v] = s;
        double o = s - thr[v];  o = o > 0 ? o : 0.0;
        double u = thr[v] - s;  u = u > 0 ? u : 0.0;
        fterm[v] = o * o * over[v] + u * u * under[v];
        vg[v] = 2 * (2 * o * over[v] - 2 * u * under[v]);
    }
    double f = 0.0;                       /* left-to-right, like Python's sum() (:265) */
    for (int64_t v = 0; v < V; ++v) f += fterm[v];

#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t j = 0; j < B; ++j) {
        double s = 0.0;
        for (int64_t k = browptr[j]; k < browptr[j + 1]; ++k) s += bval[k] * vg[bcol[k]];
        gb[j] = s;
    }
    for (int32_t i = 0; i < nb; ++i) {
        double s = 0.0;
        for (int32_t j = beam_ptr[i]; j < beam_ptr[i + 1]; ++j) s += w_grad[j] * gb[j];
        g[i] = s;
    }
    return f;
}
