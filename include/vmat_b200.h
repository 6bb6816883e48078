/*
 * vmat_b200.h - C ABI of libvmat_b200.so, the B200 (sm_100a) implementation of the
 * dose / objective / gradient evaluation and aperture pricing of VMATwPenCode.
 *
 * The reference is pure Python and has no FFI; this is the boundary its scripts
 * would bind with ctypes.  Each entry point names the
 * reference code it replaces (paths under the reference root).
 *
 * Conventions: every function returns 0 on success or a negative VMAT_E_* code and
 * never throws; the plan handle is opaque; one host thread per plan; one plan per
 * GPU (rank).  Host arrays are caller-owned; the library keeps its own device
 * copies.  Index arrays are int32 except row pointers (int64).  "f64" host
 * vectors are what the reference's numpy code holds.
 */
#ifndef VMAT_B200_H
#define VMAT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vmat_plan vmat_plan_t;

enum { VMAT_F32 = 0, VMAT_BF16 = 1, VMAT_F64 = 2 };          /* value / accumulate dtypes */
enum { VMAT_HOST = 0, VMAT_DEVICE = 1 };                     /* where the CSR arrays live */

enum {
    VMAT_OK = 0,
    VMAT_E_ARG = -1,        /* bad argument */
    VMAT_E_CUDA = -2,       /* CUDA runtime error (vmat_last_error has the text) */
    VMAT_E_NOMEM = -3,
    VMAT_E_UNSUPPORTED = -4,
    VMAT_E_STATE = -5       /* call order (e.g. pricing before any evaluation) */
};

typedef struct vmat_options {
    int32_t store_dtype;    /* VMAT_F32 | VMAT_BF16 | VMAT_F64: how D is kept in HBM          */
    int32_t acc_dtype;      /* VMAT_F32 | VMAT_F64: accumulation, and dtype of x, dose, vg     */
    int32_t tile_voxels;    /* voxels per column tile of the transposed layout (0 = default)   */
    int32_t k1_threads;     /* dose kernel: 1..511 -> 4 consumer warps per CTA, else 16 (default) */
    int32_t k2_threads;     /* transposed kernel: >= 512 -> 16 consumer warps per CTA, >= 256 -> 8, else 4 (default) */
    int32_t k2_tasks;       /* CTAs of the transposed kernel (0 = SM count x resident CTAs per SM) */
    int32_t use_graph;      /* 1: capture the evaluation as a CUDA graph (used by vmat_eval when y
                               cannot travel in the launch parameters)                          */
    /* tuning, 0 = default everywhere */
    int32_t k1_stages;      /* ring stages of the dose kernel (and of the fused kernel)                              */
    int32_t k2_stages;      /* ring stages of the transposed kernel                                                  */
    int32_t chunk_cap;      /* chunk cap of the dose layout: chunks of 4 nonzeros per lane and super-slice; longer
                               rows get more lanes (default: 48, less on small slabs)                                */
    int32_t no_pdl;         /* 1 = no programmatic dependent launch between the evaluation kernels                   */
    int32_t host_path;      /* host-buffer path of vmat_eval, bits: 1 y read from mapped host memory, 2 results +
                               completion tags written to mapped host memory, 4 y inside the launch parameters
                               (default), 8 none (copy nodes)                                                        */
    int32_t index_bytes;    /* 4 = int32 column indices, 2 = uint16 whenever they fit, 1 = byte deltas (needs sorted rows);
                               0 = byte deltas for fp32 storage + fp32 accumulation from 5*10^6 nonzeros on, else as 2  */
    int32_t claim_lead;     /* stages before a super-slice's end at which the dose kernel claims the next            */
    int32_t k2_chunk_cap;   /* chunk cap of the transposed layout (default: 24, less on small slabs)                 */
    int32_t eval_kernels;   /* 0 or 1 = the three kernels K1/K2/K3 chained by programmatic dependent launch
                               (default), 2 = one fused persistent kernel (dose, transposed product, reductions and
                               exchange in a single launch; fails if x does not fit in shared memory); with the
                               fused kernel k2_tasks is the grid size (default: one CTA per SM)                      */
    int32_t k2_pool_pct;    /* share (percent) of the transposed layout that is not statically partitioned but claimed
                               in small tasks by CTAs that have finished their piece (default 12; -1 = none)          */
    int32_t consume_batch;  /* ring stages a consumer warp takes per pass (their barrier tests, loads and gathers in
                               flight together): 4, 2 or 1; 0 = 2 (four measured no faster); tuning: bits 4..7, when
                               set, are the transposed kernel's own value                                            */
    int32_t reserved[2];
} vmat_options_t;

/* Library / build identification. */
int         vmat_version(void);
const char* vmat_last_error(void);
void        vmat_default_options(vmat_options_t* opt);

/*
 * Build a plan from the dose-deposition matrix D (V voxels x B beamlets) given twice:
 * voxel-major CSR (rows = voxels, cols = global beamlet ids) and beamlet-major CSR
 * (rows = beamlets, cols = voxel ids ascending).  B = beam_ptr[nb]; control point i owns
 * beamlets [beam_ptr[i], beam_ptr[i+1]).  Replaces the per-control-point lists
 * data.Dlist[i] / DlistT[i] (greedyVMATsampling.py:482-483,524-525) and the per-voxel
 * penalty tables quadHelperThresh/Over/Under (greedyVMATsampling.py:564-576).
 * `src_dtype` is the dtype of vval/bval (VMAT_F32 or VMAT_F64); `location` says whether
 * the six CSR arrays are host or device pointers.  thr/over/under are host f64[V].
 * In a multi-GPU run each rank passes its own voxel slab (V = slab size).
 */
int vmat_plan_create(vmat_plan_t** plan, int device, int64_t V, int64_t B, int32_t nb,
                     const int32_t* beam_ptr,
                     const int64_t* vrowptr, const int32_t* vcol, const void* vval,
                     const int64_t* browptr, const int32_t* bcol, const void* bval,
                     int src_dtype, int location,
                     const double* thr, const double* over, const double* under,
                     const vmat_options_t* opt);
int vmat_plan_destroy(vmat_plan_t* plan);

/* Bytes of HBM one evaluation must move for this plan's layouts (padding and headers included),
 * the SURVEY 8(d) algorithmic figure 2*nnz*(s_v+s_i) + 28V + 12B for the same matrix with this
 * plan's index width s_i, kernel launches per evaluation, and s_i itself (2 = uint16, 4 = int32). */
int vmat_plan_info(const vmat_plan_t* plan, int64_t* nnz, int64_t* layout_bytes_per_eval,
                   int64_t* algorithmic_bytes_per_eval, int32_t* launches_per_eval, int32_t* index_bytes);

/*
 * Set the two beamlet weight vectors of control point `beam` (length beam_ptr[beam+1]-
 * beam_ptr[beam]): w_dose = openApertureStrength scattered over openApertureMaps, w_grad =
 * diagmaker - the outputs of updateOpenAperture (greedyVMATsampling.py:1070-1130) as used by
 * calcDose (:250-251).  NULL pointers close the control point (weights 0).
 */
int vmat_set_aperture(vmat_plan_t* plan, int32_t beam, const double* w_dose, const double* w_grad);

/*
 * One evaluation = calcObjGrad(x) (greedyVMATsampling.py:578-582): dose d = D x with
 * x_j = y[beam(j)] w_dose[j] (calcDose :243-251), objective f and voxel gradient
 * (calcGradientandObjValue :254-271), beamlet gradient gb = D^T vg (PPsubroutine :646) and
 * aperture gradient g_i = sum_j w_grad[j] gb[j] (:272).  y, g are host f64[nb].
 * Copies y in, runs the kernels, copies f and g out, synchronises.
 */
int vmat_eval(vmat_plan_t* plan, const double* y, double* f, double* g);

/* Resident variants: y already on the device (vmat_set_intensities), results stay on the
 * device until fetched.  `stream` is a cudaStream_t (NULL = the plan's own stream).
 * vmat_eval_async enqueues the evaluation and returns without synchronising.  In a
 * multi-GPU run phase 0 = everything (single GPU), 1 = up to the local [gb | f] partial
 * (to be all-reduced by the caller: vmat_reduce_buffer), 2 = after the all-reduce. */
int vmat_set_intensities(vmat_plan_t* plan, const double* y);
int vmat_eval_async(vmat_plan_t* plan, void* stream, int phase);
int vmat_fetch_result(vmat_plan_t* plan, double* f, double* g);
int vmat_sync(vmat_plan_t* plan);

/* Per-kernel device timing: while enabled, every vmat_eval_async records CUDA events around
 * its kernels on the launch stream; vmat_get_timing synchronises and returns the mean
 * milliseconds of [dose kernel (K1, with the x expansion), transposed kernel (K2),
 * reductions (K3), whole evaluation] over the last n (<= 1024) evaluations. */
int vmat_set_timing(vmat_plan_t* plan, int enable);
int vmat_get_timing(vmat_plan_t* plan, double* ms /* 4 */, int32_t* n);

/* Diagnostics: run one evaluation (current intensities) with per-CTA time stamps. out receives
 * 8 uint64 per CTA, the dose kernel's CTAs first, then the transposed kernel's: globaltimer ns at
 * [0] kernel entry, [1] inputs staged and predecessor finished, [2] first stage landed,
 * [3] consumers done, [4] producer done.  Behind them, one block of 8 per control point (the reduction kernel's
 * CTAs): [0] predecessor finished, [1] this rank's partials reduced and (multi-GPU) stored into the peers' buffers,
 * [2] all ranks' values summed, [3] done.  cap = capacity of out in elements. */
int vmat_trace_eval(vmat_plan_t* plan, uint64_t* out, int64_t cap, int32_t* k1_ctas, int32_t* k2_ctas);

/* Index guard (diagnostics).  With byte-delta indices (index_bytes 1) a column index is the result of arithmetic on
 * streamed bytes; a kernel that computes one outside the gathered vector writes a record into pinned host memory and
 * traps (the call fails with VMAT_E_CUDA) instead of reading wild shared memory.  out[0] != 0: tripped; out[1] kernel
 * (1 dose, 2 transposed) << 32 | CTA, out[2] thread << 32 | ring slot, out[3] super-slice << 32 | data stage,
 * out[4] running index on entry << 32 | after the stages, out[5] limit.  Valid after a failed call as well.
 * (Absolute indices, index_bytes 2 / 4, are range-checked once when the plan is built.)  No reference counterpart. */
int vmat_guard_status(const vmat_plan_t* plan, uint64_t* out, int32_t nwords);

/* Device pointer and element count (B+1 doubles: gb then f) of the buffer a multi-GPU
 * caller must sum across ranks between phase 1 and phase 2. */
int vmat_reduce_buffer(vmat_plan_t* plan, void** dev_ptr, int64_t* count);

/*
 * Multi-GPU exchange without NCCL: every rank exports the CUDA IPC handle of its comm buffer
 * (vmat_comm_handle, 64 bytes), the caller gathers the handles of all ranks (any transport) and
 * hands them to vmat_comm_attach (at most 16 ranks).  From then on vmat_eval_async(phase 0) /
 * vmat_eval run the reductions fused with a push all-reduce over peer memory (NVLink): the CTA of
 * control point i stores its partial beamlet gradient into every rank's buffer, each value carrying
 * the evaluation's epoch (two 8-byte words {32 bits of the value | epoch}), and the thread that needs
 * a value polls its own buffer until every rank's entry carries the epoch, then adds the ranks in rank
 * order - identical bits on every rank, no fence and no separate flag on the path; phases 1/2 are not
 * needed.  All ranks must attach before any of them evaluates, and must evaluate in lock step.  A peer
 * that does not deliver within the time-out (default 20 s, vmat_comm_set_timeout) makes the evaluation
 * return VMAT_E_CUDA with NaN results; the peers then fail the same way in their next evaluation, and
 * the ranks have to attach again before they can continue.
 * This replaces the reference's only parallelism, the process pool of greedyVMATsampling.py:836-841.
 *
 * vmat_comm_attach_local is the same for plans that live in ONE process (plans[q] = rank q's plan,
 * plans[rank] == plan): on different devices with peer access, or - for tests - on the same device
 * (the reduction kernels of all "ranks" must then be able to run concurrently: small plans only).
 */
int vmat_comm_handle(vmat_plan_t* plan, void* handle_out /* 64 bytes */);
int vmat_comm_attach(vmat_plan_t* plan, int32_t rank, int32_t nranks, const void* handles /* nranks*64 bytes */);
int vmat_comm_attach_local(vmat_plan_t* plan, int32_t rank, int32_t nranks, vmat_plan_t* const* plans /* nranks */);
int vmat_comm_set_timeout(vmat_plan_t* plan, double seconds);

/* Device buffers for callers that keep the optimisation loop on the GPU: the intensities the
 * next vmat_eval_async reads (f64[nb]) and the result the last one wrote (f64[nb+1]: f then g). */
int vmat_intensities_buffer(vmat_plan_t* plan, void** dev_ptr, int64_t* count);
int vmat_result_buffer(vmat_plan_t* plan, void** dev_ptr, int64_t* count);

/* Read back state of the last evaluation as f64: data.currentDose (:244-250),
 * data.voxelgradient (:271), and beamGrad of every control point (:646) concatenated. */
int vmat_get_dose(vmat_plan_t* plan, double* d_out /* V */);
int vmat_get_voxelgrad(vmat_plan_t* plan, double* vg_out /* V */);
int vmat_get_beamlet_grad(vmat_plan_t* plan, double* gb_out /* B */);

/*
 * Pricing: PPsubroutine (greedyVMATsampling.py:618-789) for `ncand` candidate control
 * points in one launch, on the beamlet gradient of the last evaluation.
 */
typedef struct vmat_price_row {     /* one leaf row of one candidate (host-built, :664-680,713-724) */
    double l_first;                 /* first left-leaf position (integer valued, or the fractional midpoint) */
    int32_t n_l;                    /* number of consecutive l values (step 1)                */
    int32_t vb_min;                 /* min(validbeamlets) of this row (fvalidbeamlets :589-602) */
    double r_lo;                    /* max(rcm - travel-, rcp - travel+)  (before max with l+1) */
    double r_hi;                    /* min(N, rcm + travel-, rcp + travel+)                    */
    double r_mid;                   /* first r of the midpoint fallback used when the r range is empty */
    uint64_t vb_mask;               /* bit n set: beamlet n exists in this row                 */
    int32_t grad_offset;            /* beamGrad index of y index vb_min in this row, quirk Q1 applied */
    int32_t n_r_mid;                /* how many r values the midpoint fallback yields: len(np.arange(mid, mid+1)), 1 or 2 */
} vmat_price_row_t;

typedef struct vmat_price_params {
    double C, C2, C3, b;
    int32_t M, N;                   /* network rows (= row descriptors per candidate), beamlets per leaf row */
    int32_t variant;                /* which generation of the reference's network (VMAT_PRICE_*), 0 = canonical */
    int32_t reserved[5];            /* [0] = 1 (CLASSIC only): the sink scans the last node only - what the reference's
                                       sink loop covers when the rows after the last described one had no nodes */
} vmat_price_params_t;

/* The generations of PPsubroutine.  The row descriptors are built by the caller with the
 * generation's own range expressions; the kernel wires the network accordingly:
 *   SAMPLING  greedyVMATsampling.py:618-789 (canonical): M rows; empty r range -> midpoint (r_mid, n_r_mid);
 *             tie-break and DoseSide terms; result = M leaf pairs.
 *   CLASSIC   greedyVMAT.py:469-636: rows 0..M-2; empty r range -> r from its start up to r_mid (the row's first l);
 *             relaxation window = previous row shifted by one node, predecessor recorded one node later, the sink
 *             also scans the row's last node, untouched weights 0.0; result = one pair per visited row plus the
 *             sink's (0, 0), as the reference returns it.
 *   SHORT     greedyVMATshort.py:456-616: row 0 and rows 2..M-1 (with the beamlets of row m-1); no fallback
 *             (n_r_mid = 0); shifted window, no tie-break; dose terms read beamGrad by raw position (row 0, vb_mask =
 *             all ones, grad_offset = vb_min = 0) or from grad_offset + 1 over r - l - 1 entries (other rows). */
enum { VMAT_PRICE_SAMPLING = 0, VMAT_PRICE_CLASSIC = 1, VMAT_PRICE_SHORT = 2 };

int vmat_price(vmat_plan_t* plan, int32_t ncand, const int32_t* cand_beam,
               const vmat_price_params_t* params, const vmat_price_row_t* rows /* ncand*M */,
               double* p /* ncand */, int32_t* l /* ncand*M */, int32_t* r /* ncand*M */);
/* Any generation: l and r hold params->M + 1 entries per candidate, len[c] says how many are used (the earlier
 * generations return paths of varying length).  p[c] is NaN where the reference would leave p unassigned (no
 * path into the sink at or below its initial weight: it raises UnboundLocalError there). */
int vmat_price_ex(vmat_plan_t* plan, int32_t ncand, const int32_t* cand_beam,
                  const vmat_price_params_t* params, const vmat_price_row_t* rows /* ncand*M */,
                  double* p /* ncand */, int32_t* l /* ncand*(M+1) */, int32_t* r /* ncand*(M+1) */,
                  int32_t* len /* ncand */);

/*
 * k-medoid insertion costs (kmedoids.py:32-60): for every candidate point c, the sum over
 * all n points j of min(curmin[j], dist(c, j)); points are n x dim f64 (host).  The caller
 * keeps curmin (distance to the nearest already-chosen medoid) and picks the first minimum.
 */
int vmat_kmedoid_costs(int device, int64_t n, int32_t dim, const double* points,
                       const double* curmin, double* cost /* n */);
/* Same with the candidates given separately (n_cand x dim) from the points summed over (n x dim):
 * a rank's shard of the point set in the multi-GPU use; the caller adds the per-rank costs. */
int vmat_kmedoid_costs_shard(int device, int64_t n_cand, const double* cand, int64_t n, int32_t dim,
                             const double* points, const double* curmin, double* cost /* n_cand */);

/* The same with the point set resident on the device between insertion steps (a whole ordering is thousands of
 * steps over the same points): vmat_kmedoid_open uploads this rank's points (n x dim), their curmin and the candidate
 * coordinates once; vmat_kmedoid_costs_resident computes cost[n_cand] (to the host, and/or leaves it on the device for
 * the caller's all-reduce: *cost_dev); vmat_kmedoid_insert(medoid) lowers curmin by one more medoid, in the operation
 * order of the host bookkeeping (bit-identical). */
typedef struct vmat_kmedoid vmat_kmedoid_t;
int vmat_kmedoid_open(vmat_kmedoid_t** session, int device, int64_t n, int32_t dim, const double* points,
                      const double* curmin, int64_t n_cand, const double* cand);
int vmat_kmedoid_insert(vmat_kmedoid_t* session, const double* medoid /* dim */);
int vmat_kmedoid_costs_resident(vmat_kmedoid_t* session, double* cost_host /* n_cand or NULL */, void** cost_dev /* or NULL */);
int vmat_kmedoid_close(vmat_kmedoid_t* session);

/*
 * Restricted-master-problem fast path (solveRMC, greedyVMATsampling.py:858-885): while the
 * apertures of the listed control points stay fixed, vmat_rmp_begin caches their dose columns
 * D_i^T w_dose_i and gradient columns D_i^T w_grad_i (dense, V each); vmat_rmp_eval then evaluates
 * (f, g) for intensities y[nb] from the cache (control points outside the list contribute nothing
 * and get g = 0, as they have zero weights in calcDose anyway).  vmat_rmp_end ends the RMP; the
 * columns stay cached, and the next vmat_rmp_begin only rebuilds control points whose aperture was
 * set again (vmat_set_aperture) in between - one new aperture per greedy iteration.
 * The beamlet gradient used by pricing is NOT refreshed by vmat_rmp_eval.
 */
int vmat_rmp_begin(vmat_plan_t* plan, int32_t ncols, const int32_t* beams);
int vmat_rmp_eval(vmat_plan_t* plan, const double* y, double* f, double* g);
int vmat_rmp_end(vmat_plan_t* plan);

/*
 * Dose-volume histograms of the last evaluation's dose (printresults, greedyVMATsampling.py:
 * 888-912): vmat_dose_max returns max(dose); vmat_dvh_histogram counts, for every structure s
 * (bit s of full_mask[v], data.fullMaskValue), the voxels whose dose falls in
 * [edges[k], edges[k+1]) - np.histogram with explicit edges, last bin closed.  Host arrays.
 */
int vmat_dose_max(vmat_plan_t* plan, double* max_dose);
int vmat_dvh_histogram(vmat_plan_t* plan, const uint64_t* full_mask /* V */, int32_t nstructs,
                       const double* edges, int32_t nedges, int64_t* hist /* nstructs*(nedges-1) */);

#ifdef __cplusplus
}
#endif
#endif /* VMAT_B200_H */
