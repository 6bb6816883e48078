// libvmat_b200.so - C ABI (include/vmat_b200.h) over the sm_100a kernels.
// Plan construction (CSR -> row-sorted sliced-ELL streams), evaluation launch sequence
// (K1 -> K2 -> K3 chained by programmatic dependent launch, replayed as one CUDA graph on the
// host-buffer path), result transfer, cross-GPU exchange set-up, pricing, RMP column cache, DVH.
#include "../../include/vmat_b200.h"
#include "vmat_eval_kernels.cuh"
#include "vmat_fused_kernel.cuh"
#include "vmat_price_kernels.cuh"
#include "vmat_kmedoid_kernels.cuh"
#include "vmat_rmp_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

using namespace vmat;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            return fail(VMAT_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                 \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) {
        if (p) { cudaFree(p); p = nullptr; }
        bytes = n;
        return cudaMalloc(&p, n ? n : 16);
    }
    template <typename T> T* as() const { return static_cast<T*>(p); }
};

struct vmat_plan {
    int device = 0;
    int64_t V = 0, B = 0, nnz = 0;
    int32_t nb = 0;
    vmat_options_t opt{};
    size_t acc_size = 4;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int iw = 4;                    // bytes per stored column index: 2 (uint16) when every index fits, else 4
    int smem_optin = 0;

    std::vector<int32_t> beam_ptr;
    // K1 layout (stream of super-slices of W1 slices)
    DevBuf sell1, off1, tpr1;
    int32_t nslices1 = 0, nss1 = 0, W1 = 16, nstages1 = 8, stage_bytes1 = 0;
    int64_t units1 = 0;
    // K2 layout
    DevBuf sell2, off2, tpr2, tasks2, cta_ptr2, pool2;
    int32_t npool2 = 0;
    int k2_grid = 1;
    int32_t nslices2 = 0, nss2 = 0, ntasks2 = 0, ntiles = 0, T = 0, W2 = 4, nstages2 = 8, stage_bytes2 = 0;
    int64_t units2 = 0;
    // fused single-kernel evaluation (vmat_fused_kernel.cuh): both streams ragged, W = 16, one item queue
    bool fused = false;
    DevBuf itemsF, syncF, offF2;           // offF2[q]: stream offset of the q-th transposed item of the queue
    int32_t nF1 = 0, nF = 0, nstagesF = 0, stage_bytesF = 0, ngroupsF = 0, gridF = 0;
    size_t smemF = 0;
    // state
    DevBuf y, beam_of, beam_ptr_d, w_dose, w_grad, xg, d, vg, fpart, fblock, partial, gbf, result, counter, scratch64;
    void* h_w = nullptr;         // pinned staging for vmat_set_aperture: [max_beamlets f64 | max_beamlets f64]
    int32_t max_beamlets = 0;
    bool weights_in_flight = false;    // vmat_set_aperture's copies may still be running on the plan's stream
    double* h_y = nullptr;       // pinned + mapped: K1 reads it directly on the host-buffer path
    unsigned long long* h_guard = nullptr;   // pinned + mapped: the index guard's record (Guard, vmat_eval_kernels.cuh)
    unsigned long long* d_guard = nullptr;
    double* h_res = nullptr;     // pinned + mapped: K3 writes results here ...
    unsigned long long* h_tags = nullptr;   // ... each followed by a completion tag the host polls
    double *d_hy = nullptr, *d_hres = nullptr;          // device views of the mapped host buffers
    unsigned long long* d_htags = nullptr;
    unsigned long long host_seq = 0;
    // fused cross-GPU exchange (vmat_comm_attach)
    DevBuf comm_buf, comm_ptrs, comm_state;
    std::vector<void*> comm_peers;        // opened IPC mappings (own slot = nullptr)
    int comm_rank = 0, comm_nranks = 1;
    unsigned long long comm_epoch = 0;                 // fused evaluations launched so far (same on every rank)
    unsigned long long comm_timeout_ns = 20000000000ull;
    bool comm_ipc = false;                // peers were opened with cudaIpcOpenMemHandle (else: plain pointers)
    int host_path = 4;             // bit0: K1 reads y from mapped host memory; bit1: K3 publishes results + tags to host memory
    bool x_in_smem = true;
    bool pdl = true;               // programmatic dependent launch between the evaluation kernels
    int k1_threads = 1024, k1_grid = 148, k2_threads = 256;
    size_t k1_smem = 0, k2_smem = 0;
    bool evaluated = false;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    // pricing scratch
    DevBuf price_rows, price_cand, price_p, price_l, price_r, price_dad, price_len, price_gstart;
    size_t price_cap_cand = 0, price_cap_rows = 0, price_cap_dad = 0;
    // natural-order penalty tables and the RMP column cache (vmat_rmp_begin)
    DevBuf thr_n, over_n, under_n, colA, colG, col_y, col_w, rmp_yk, rmp_beams, rmp_fpart, rmp_gpart;
    std::vector<int32_t> rmp_cols;                 // control points of the active RMP
    std::vector<int32_t> col_slot, col_built;      // per control point: slot in colA/colG (-1 none), aperture epoch it was built from
    std::vector<int32_t> ap_epoch;                 // per control point: bumped by vmat_set_aperture
    DevBuf rmp_slots;
    int32_t col_cap = 0, col_used = 0;
    int64_t rmp_ld = 0;
    bool col_mode = false;
    const void* col_w_ptr = nullptr;
    // per-kernel timing (vmat_set_timing)
    bool timing = false;
    DevBuf trace;                      // vmat_trace_eval: kTraceSlots timestamps per CTA, K1 then K2
    bool tracing = false;
    std::vector<cudaEvent_t> tev;      // 4 events per evaluation slot
    int64_t tcount = 0;
    static constexpr int kTimingSlots = 1024;

    ~vmat_plan() {
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (graph) cudaGraphDestroy(graph);
        for (auto e : tev) cudaEventDestroy(e);
        if (comm_ipc) for (void* q : comm_peers) if (q) cudaIpcCloseMemHandle(q);
        if (h_w) cudaFreeHost(h_w);
        if (h_y) cudaFreeHost(h_y);
        if (h_res) cudaFreeHost(h_res);
        if (h_guard) cudaFreeHost(h_guard);
        if (h_tags) cudaFreeHost(h_tags);
        if (stream) cudaStreamDestroy(stream);
    }
};

extern "C" int vmat_version(void) { return 100; }
extern "C" const char* vmat_last_error(void) { return g_err.c_str(); }

extern "C" void vmat_default_options(vmat_options_t* o) {
    std::memset(o, 0, sizeof(*o));
    o->store_dtype = VMAT_F32;
    o->acc_dtype = VMAT_F32;
    o->use_graph = 1;
}

// ---------------------------------------------------------------------------------
// dispatch helpers
// ---------------------------------------------------------------------------------
#define DISPATCH_STORE(dt, VT, ...)                                          \
    switch (dt) {                                                            \
        case VMAT_F32:  { using VT = float; __VA_ARGS__; break; }            \
        case VMAT_BF16: { using VT = __nv_bfloat16; __VA_ARGS__; break; }    \
        case VMAT_F64:  { using VT = double; __VA_ARGS__; break; }           \
        default: return fail(VMAT_E_ARG, "bad store dtype");                 \
    }
#define DISPATCH_IW(IW, ...)                                                 \
    if (P->iw == 2) { constexpr int IW = 2; __VA_ARGS__; } else if (P->iw == 1) { constexpr int IW = 1; __VA_ARGS__; } else { constexpr int IW = 4; __VA_ARGS__; }
#define DISPATCH_ACC(dt, AT, ...)                                            \
    switch (dt) {                                                            \
        case VMAT_F32: { using AT = float; __VA_ARGS__; break; }             \
        case VMAT_F64: { using AT = double; __VA_ARGS__; break; }            \
        default: return fail(VMAT_E_ARG, "bad accumulate dtype");            \
    }

static int store_units(int dt, int iw) { return 8 * iw + (dt == VMAT_F32 ? 32 : dt == VMAT_BF16 ? 16 : 64); }
static int store_bytes(int dt) { return dt == VMAT_F32 ? 4 : dt == VMAT_BF16 ? 2 : 8; }

template <typename AT>
static int upload_cast(DevBuf& dst, const std::vector<double>& src, cudaStream_t st) {
    std::vector<AT> tmp(src.size());
    for (size_t i = 0; i < src.size(); ++i) tmp[i] = (AT)src[i];
    CU(dst.alloc(tmp.size() * sizeof(AT)));
    CU(cudaMemcpyAsync(dst.p, tmp.data(), tmp.size() * sizeof(AT), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

template <typename Tv>
static int upload_vec(DevBuf& dst, const std::vector<Tv>& src, cudaStream_t st) {
    CU(dst.alloc(src.size() * sizeof(Tv)));
    if (!src.empty()) {
        CU(cudaMemcpyAsync(dst.p, src.data(), src.size() * sizeof(Tv), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
    }
    return 0;
}

// Build one stream layout from segment descriptors (already grouped in slices of 32, W slices per
// super-slice).  `tables` (thr/over/under in slice order, element size es) go into the headers.
static int build_stream(vmat_plan* P, int store_dtype, int src_dtype, int32_t W, int32_t HU,
                        const std::vector<int64_t>& ss_off, const std::vector<uint8_t>& ss_tpr,
                        const std::vector<int64_t>& seg_start,
                        const std::vector<int32_t>& seg_len, const std::vector<int32_t>& seg_colbase,
                        const std::vector<int32_t>& rows, const DevBuf* thr, const DevBuf* over, const DevBuf* under,
                        int32_t es, const int32_t* d_col, const void* d_val, DevBuf& sell, DevBuf& off_d, DevBuf& tpr_d,
                        const std::vector<uint8_t>* ss_nch = nullptr) {
    const int32_t nslices = (int32_t)(rows.size() / 32);
    const int64_t units = ss_off.back();
    CU(sell.alloc((size_t)units * 16));
    CU(cudaMemsetAsync(sell.p, 0, (size_t)units * 16, P->stream));
    if (upload_vec(off_d, ss_off, P->stream)) return VMAT_E_CUDA;
    if (upload_vec(tpr_d, ss_tpr, P->stream)) return VMAT_E_CUDA;
    if (nslices == 0) return 0;
    DevBuf dstart, dlen, dcb, drows, dnch;
    if (ss_nch && upload_vec(dnch, *ss_nch, P->stream)) return VMAT_E_CUDA;
    const uint8_t* nchp = ss_nch ? dnch.as<uint8_t>() : nullptr;
    if (upload_vec(dstart, seg_start, P->stream)) return VMAT_E_CUDA;
    if (upload_vec(dlen, seg_len, P->stream)) return VMAT_E_CUDA;
    if (upload_vec(dcb, seg_colbase, P->stream)) return VMAT_E_CUDA;
    if (upload_vec(drows, rows, P->stream)) return VMAT_E_CUDA;
    const int threads = 256;
    const int64_t blocks = ((int64_t)nslices * 32 + threads - 1) / threads;
    const unsigned char* t0 = thr ? thr->as<unsigned char>() : nullptr;
    const unsigned char* t1 = over ? over->as<unsigned char>() : nullptr;
    const unsigned char* t2 = under ? under->as<unsigned char>() : nullptr;
    DISPATCH_STORE(store_dtype, VT, { DISPATCH_IW(IW, {
        if (src_dtype == VMAT_F32)
            k_fill_stream<VT, float, IW><<<(unsigned)blocks, threads, 0, P->stream>>>(
                nslices, W, HU, off_d.as<int64_t>(), dstart.as<int64_t>(), dlen.as<int32_t>(), dcb.as<int32_t>(),
                tpr_d.as<uint8_t>(), drows.as<int32_t>(), t0, t1, t2, es, d_col, static_cast<const float*>(d_val), sell.as<uint4>(), nchp);
        else
            k_fill_stream<VT, double, IW><<<(unsigned)blocks, threads, 0, P->stream>>>(
                nslices, W, HU, off_d.as<int64_t>(), dstart.as<int64_t>(), dlen.as<int32_t>(), dcb.as<int32_t>(),
                tpr_d.as<uint8_t>(), drows.as<int32_t>(), t0, t1, t2, es, d_col, static_cast<const double*>(d_val), sell.as<uint4>(), nchp);
    }); });
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(P->stream));
    return 0;
}

// index_bytes = 0: fp32 / fp32 plans with at least this many nonzeros store byte deltas (measured down to 7.3 M
// nonzeros - a quarter of config 2, both layouts in L2 - where they still gain 1.5 %)
constexpr int64_t kAutoDeltaNnz = 5000000;

// byte-delta layout: bridge entries each segment needs (k_count_pads), for the plan's storage type
static int count_pads(vmat_plan* P, int src_dtype, int64_t nseg, const int64_t* d_bounds, const int32_t* d_col,
                      const void* d_val, int32_t* d_pads) {
    if (nseg <= 0) return 0;
    const unsigned blocks = (unsigned)((nseg + 255) / 256);
    DISPATCH_STORE(P->opt.store_dtype, VT, {
        if (src_dtype == VMAT_F32)
            k_count_pads<VT, float><<<blocks, 256, 0, P->stream>>>(nseg, d_bounds, d_col, static_cast<const float*>(d_val), d_pads);
        else
            k_count_pads<VT, double><<<blocks, 256, 0, P->stream>>>(nseg, d_bounds, d_col, static_cast<const double*>(d_val), d_pads);
    });
    CU(cudaGetLastError());
    return 0;
}

// Host-side description of a stream: super-slices of W slices; a slice holds 32/tpr rows with tpr
// lanes per row, where tpr is the smallest power of two that keeps the slice within `cap` chunks.
struct StreamDesc {
    int32_t W, HU, units, cap;
    bool ragged = false;                          // fused kernel: every slice keeps its own chunk count
    bool too_long = false;                        // ragged: a row needs more than 255 chunks per lane
    bool record_nch = false;                      // fused kernel with uniform stages: ss_nch = the item's chunk count for every warp
    std::vector<uint8_t> ss_nch;                  // [nss * W] (ragged only)
    std::vector<int64_t> ss_off{0}, seg_start;
    std::vector<uint8_t> ss_tpr;
    std::vector<int32_t> seg_len, seg_cb, rows, ss_tile;
    std::vector<int64_t> ss_cost;                 // ring stages (header + chunks) of each super-slice
    std::vector<double> t_s, o_s, u_s;            // penalty tables in slot order (K1 only)
    struct Item { int32_t len, id; int64_t start; int32_t src_len; };   // len: slots the row needs (nonzeros + bridge entries), src_len: nonzeros
    // `items` sorted by len descending; appends the super-slices of one group (all voxels / one tile)
    void add_group(const std::vector<Item>& items, int32_t colbase, int32_t tile,
                   const double* thr, const double* over, const double* under) {
        size_t pos = 0;
        while (pos < items.size()) {
            const int32_t maxlen = items[pos].len;
            const int64_t nch1 = (maxlen + 3) / 4;
            int tpr = 1;
            while ((nch1 + tpr - 1) / tpr > cap && tpr < 32) tpr *= 2;
            const int64_t nch = (nch1 + tpr - 1) / tpr;
            const size_t rows_per_slice = 32 / tpr, rows_per_ss = rows_per_slice * W;
            const size_t base_slot = seg_start.size(), first_item = pos;
            seg_start.resize(base_slot + (size_t)32 * W, 0);
            seg_len.resize(base_slot + (size_t)32 * W, 0);
            seg_cb.resize(base_slot + (size_t)32 * W, 0);
            rows.resize(base_slot + (size_t)32 * W, -1);
            if (thr) { t_s.resize(base_slot + (size_t)32 * W, 0.0); o_s.resize(base_slot + (size_t)32 * W, 0.0); u_s.resize(base_slot + (size_t)32 * W, 0.0); }
            for (size_t r = 0; r < rows_per_ss && pos < items.size(); ++r, ++pos) {
                const Item& it = items[pos];
                const size_t w = r / rows_per_slice, q = r % rows_per_slice;
                for (int part = 0; part < tpr; ++part) {
                    const size_t slot = base_slot + w * 32 + q * tpr + part;
                    seg_start[slot] = it.start; seg_len[slot] = it.src_len; seg_cb[slot] = colbase;
                    if (part == 0) {
                        rows[slot] = it.id;
                        if (thr) { t_s[slot] = thr[it.id]; o_s[slot] = over[it.id]; u_s[slot] = under[it.id]; }
                    }
                }
            }
            ss_tpr.push_back((uint8_t)tpr);
            ss_tile.push_back(tile);
            ss_cost.push_back(nch + 1);
            if (ragged) {     // the first row of a slice is its longest (items are sorted by length)
                int64_t chunks = 0;
                for (size_t w = 0; w < (size_t)W; ++w) {
                    const size_t p0 = first_item + w * rows_per_slice;
                    const int64_t n1w = p0 < items.size() ? (items[p0].len + 3) / 4 : 0;
                    const int64_t nw = (n1w + tpr - 1) / tpr;
                    if (nw > 255) too_long = true;         // chunk counts travel as bytes
                    ss_nch.push_back((uint8_t)nw);
                    chunks += nw;
                }
                ss_off.push_back(ss_off.back() + (int64_t)W * HU + chunks * (int64_t)units);
            } else {
                if (record_nch) { if (nch > 255) too_long = true; for (int w = 0; w < W; ++w) ss_nch.push_back((uint8_t)nch); }
                ss_off.push_back(ss_off.back() + (int64_t)W * HU + nch * (int64_t)W * units);
            }
        }
    }
};

// consumer warps (slices per super-slice) of the dose kernel: 16 with one CTA per SM by default,
// 4 (several CTAs per SM, each with its own copy of x) on request; chosen per plan (vmat_plan::W1)
// ... of the transposed kernel: 4 by default (five CTAs per SM), 8 or 16 on request (vmat_plan::W2)
#define DISPATCH_W2(W, ...)                                                  \
    if (P->W2 == 16) { constexpr int W = 16; __VA_ARGS__; } else if (P->W2 == 8) { constexpr int W = 8; __VA_ARGS__; } else { constexpr int W = 4; __VA_ARGS__; }
// chunks per lane and super-slice (more lanes per row beyond that).  Measured (configs 2 and 3): the
// dynamically scheduled dose kernel likes long super-slices (less padding: 140 -> 134 us on config 3
// from 24 to 48), the statically partitioned transposed kernel short ones (finer piece boundaries)
constexpr int kChunkCapK1 = 48, kChunkCapK2 = 24;
// voxels per tile of the transposed layout: larger tiles mean longer beamlet segments (less padding)
// and fewer tile partials for K3, but one CTA less per SM.  Measured: config 2 (300k voxels) 65.2 us
// with 4096 against 66.2 us with 6144; config 3 (1M) 273 against 268 us; config 4 (4M) 434 against 446 evals/s
// The same tile bytes with double accumulation (half the voxels): config 2 93 -> 88 us, config 3 344 -> 330 us.
static int32_t default_tile_voxels(int64_t V, size_t acc_size) { return (int32_t)((V > 500000 ? 6144 : 4096) * 4 / acc_size); }
constexpr int kClaimLeadDefault = 8;

// dynamic shared memory limit of a kernel = what the device allows next to the kernel's static part
template <typename Kern>
static cudaError_t allow_max_smem(Kern kern, int smem_optin) {
    cudaFuncAttributes fa{};
    cudaError_t e = cudaFuncGetAttributes(&fa, kern);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin - (int)fa.sharedSizeBytes);
}

template <typename VT, typename AT>
static int k2_occupancy_t(vmat_plan* P, int* occ) {
    DISPATCH_IW(IW, { DISPATCH_W2(W2c, {
        auto kern = k2_beamlet_grad<VT, AT, W2c, IW>;
        // the device maximum, set once per plan: every plan asks for the same value, so plans of
        // different shapes never shrink each other's limit and launches need no further set-up
        CU(allow_max_smem(kern, P->smem_optin));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, (W2c + 1) * 32, P->k2_smem));
    }); });
    if (*occ < 1) return fail(VMAT_E_ARG, "transposed kernel does not fit on an SM with these options");
    return 0;
}
#define DISPATCH_W1(W, ...)                                                  \
    if (P->W1 == 4) { constexpr int W = 4; __VA_ARGS__; } else { constexpr int W = 16; __VA_ARGS__; }

template <typename VT, typename AT>
static int k1_occupancy_t(vmat_plan* P, int* occ) {
    DISPATCH_IW(IW, { DISPATCH_W1(W, {
        if (P->x_in_smem) {
            auto kern = k1_dose_penalty<VT, AT, true, W, IW>;
            CU(allow_max_smem(kern, P->smem_optin));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, P->k1_threads, P->k1_smem));
        } else {
            auto kern = k1_dose_penalty<VT, AT, false, W, IW>;
            CU(allow_max_smem(kern, P->smem_optin));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, P->k1_threads, P->k1_smem));
        }
    }); });
    if (*occ < 1) return fail(VMAT_E_ARG, "dose kernel does not fit on an SM with these options");
    return 0;
}
static int k1_occupancy(vmat_plan* P, int* occ) {
    DISPATCH_STORE(P->opt.store_dtype, VT, {
        DISPATCH_ACC(P->opt.acc_dtype, AT, { return k1_occupancy_t<VT, AT>(P, occ); });
    });
    return 0;
}
static int k2_occupancy(vmat_plan* P, int* occ) {
    DISPATCH_STORE(P->opt.store_dtype, VT, {
        DISPATCH_ACC(P->opt.acc_dtype, AT, { return k2_occupancy_t<VT, AT>(P, occ); });
    });
    return 0;
}

// launch with programmatic stream serialization: the kernel may begin (prologue up to its
// griddepcontrol.wait) while its predecessor in the stream is still draining
template <typename Kern, typename Arg, typename... Rest>
static cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t st, bool pdl, Arg arg, Rest... rest) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, arg, rest...);
}

#define DISPATCH_IWF(IW, ...)                                                \
    if (P->iw == 2) { constexpr int IW = 2; __VA_ARGS__; } else { constexpr int IW = 4; __VA_ARGS__; }

template <typename VT, typename AT>
static int fused_occupancy_t(vmat_plan* P, int* occ) {
    DISPATCH_IWF(IW, {
        auto kern = kf_eval<VT, AT, IW>;
        CU(allow_max_smem(kern, P->smem_optin));
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kFThreads, P->smemF));
    });
    if (*occ < 1) return fail(VMAT_E_ARG, "fused evaluation kernel does not fit on an SM with these options");
    return 0;
}
static int fused_occupancy(vmat_plan* P, int* occ) {
    DISPATCH_STORE(P->opt.store_dtype, VT, {
        DISPATCH_ACC(P->opt.acc_dtype, AT, { return fused_occupancy_t<VT, AT>(P, occ); });
    });
    return 0;
}

static int capture_graph(vmat_plan* P);
static int launch_eval(vmat_plan* P, cudaStream_t st, int phase, bool host_mode = false, bool inline_y = false);

extern "C" int vmat_plan_create(vmat_plan_t** out, int device, int64_t V, int64_t B, int32_t nb,
                                const int32_t* beam_ptr,
                                const int64_t* vrowptr, const int32_t* vcol, const void* vval,
                                const int64_t* browptr, const int32_t* bcol, const void* bval,
                                int src_dtype, int location,
                                const double* thr, const double* over, const double* under,
                                const vmat_options_t* opt_in) {
    if (!out || V <= 0 || B <= 0 || nb <= 0 || !beam_ptr || !vrowptr || !browptr || !thr || !over || !under)
        return fail(VMAT_E_ARG, "vmat_plan_create: null or non-positive argument");
    if (src_dtype != VMAT_F32 && src_dtype != VMAT_F64) return fail(VMAT_E_ARG, "src_dtype must be f32 or f64");
    if (beam_ptr[0] != 0 || beam_ptr[nb] != B) return fail(VMAT_E_ARG, "beam_ptr must run from 0 to B");
    if (V >= (int64_t)1 << 31 || B >= (int64_t)1 << 31) return fail(VMAT_E_UNSUPPORTED, "V and B must fit int32");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(VMAT_E_ARG, "no such CUDA device");
    CU(cudaSetDevice(device));

    vmat_plan* P = new (std::nothrow) vmat_plan();
    if (!P) return fail(VMAT_E_NOMEM, "out of host memory");
    struct Guard { vmat_plan* p; ~Guard() { delete p; } } guard{P};
    P->device = device; P->V = V; P->B = B; P->nb = nb;
    if (opt_in) P->opt = *opt_in; else vmat_default_options(&P->opt);
    if (P->opt.acc_dtype != VMAT_F32 && P->opt.acc_dtype != VMAT_F64) return fail(VMAT_E_ARG, "bad acc_dtype");
    if (P->opt.store_dtype < 0 || P->opt.store_dtype > 2) return fail(VMAT_E_ARG, "bad store_dtype");
    P->acc_size = P->opt.acc_dtype == VMAT_F64 ? 8 : 4;
    P->pdl = P->opt.no_pdl == 0;     // no_pdl = 1 disables programmatic dependent launch
    // host_path: host-buffer path bits - 1 = K1 reads y from mapped host memory, 2 = K3 publishes
    // results + completion tags to mapped host memory, 4 = y travels in the launch parameters (plain
    // launches, no graph; default - measured on config 2: 86 us per call against 90-95 us for the
    // graph replays with a copy node in front, tools/latency.py)
    P->host_path = P->opt.host_path > 0 ? (P->opt.host_path & 7) : 4;     // 8 = none of the bits
    P->beam_ptr.assign(beam_ptr, beam_ptr + nb + 1);
    CU(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));
    CU(cudaDeviceGetAttribute(&P->sm_count, cudaDevAttrMultiProcessorCount, device));
    CU(cudaDeviceGetAttribute(&P->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cudaStream_t st = P->stream;
    bool iw_automatic = false;      // byte deltas chosen by default: given up again if the voxel-major rows are not sorted
    int iw_absolute = 2;
    {   // index width: uint16 when every stored index fits (x index < B; K2 indices are tile-local, < T)
        const int32_t T0 = P->opt.tile_voxels > 0 ? P->opt.tile_voxels : default_tile_voxels(V, P->acc_size);
        const bool fits16 = B <= 65536 && ((T0 + 63) / 64 * 64) <= 65536;
        // index_bytes: 4 forces int32 indices, 2 absolute indices (uint16 when they fit), 1 byte deltas (any B: a
        // lane's start index is a u32 in the header); 0 = byte deltas for fp32 storage with 4-byte accumulation
        // from kAutoDeltaNnz nonzeros on (5 instead of 6 bytes per nonzero; measured, DESIGN.md section 2: config 2
        // 63.9 -> 58.7 us, config 3 264.7 -> 230.5 us; bf16 storage and 8-byte accumulation, whose kernels are
        // bound by the consumers' instruction and shared-memory rates, are slower with them), else absolute
        const int ib = P->opt.index_bytes;
        int64_t nnz_peek = 0;
        CU(cudaMemcpy(&nnz_peek, vrowptr + V, 8, location == VMAT_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost));
        const bool automatic = ib == 0 && P->acc_size == 4 && (P->opt.eval_kernels & 3) != 2 &&
                               P->opt.store_dtype == VMAT_F32 && nnz_peek >= kAutoDeltaNnz;
        iw_absolute = (ib == 4 || !fits16) ? 4 : 2;
        P->iw = (ib == 1 || automatic) ? 1 : iw_absolute;
        iw_automatic = automatic;
    }
    auto dose_shape = [&]() {   // dose-kernel CTA shape
        const size_t as0 = P->acc_size;
        const size_t xy = (((size_t)B * as0 + 127) & ~(size_t)127) + (((size_t)nb * 8 + 127) & ~(size_t)127);
        const int hu1 = 8 + 3 * (32 * (int)as0 / 16) + header_base_units(P->iw);
        const size_t stage4 = (size_t)4 * std::max<int>(store_units(P->opt.store_dtype, P->iw), hu1) * 16;
        const bool fits2 = 2 * (xy + 8 * stage4 + 2048) <= (size_t)P->smem_optin;
        // measured on config 2: one 16-warp CTA per SM (48.8 us) beats three 4-warp CTAs (52.0 us)
        P->W1 = (P->opt.k1_threads > 0 && P->opt.k1_threads < 512 && fits2) ? 4 : 16;
    };
    dose_shape();

    // ---- one fused kernel or the three-kernel path --------------------------------------------------
    // eval_kernels & 3: 0 / 1 = the three kernels K1/K2/K3 chained by programmatic dependent launch (default:
    // measured faster on every configuration), 2 = the single fused kernel (needs x + y or
    // two vg tiles next to at least 6 ring stages of 16 warps in shared memory); bit 2: uniform stages (experiment)
    int32_t T_plan = P->opt.tile_voxels > 0 ? P->opt.tile_voxels : default_tile_voxels(V, P->acc_size);
    T_plan = (T_plan + 63) / 64 * 64;
    {
        const size_t as0 = P->acc_size;
        const size_t xy = (((size_t)B * as0 + 127) & ~(size_t)127) + (((size_t)nb * 8 + 127) & ~(size_t)127);
        const size_t tb = ((size_t)T_plan * as0 + 127) & ~(size_t)127;
        const size_t U = std::max(xy, 2 * tb);
        const int hu1 = 8 + 3 * (32 * (int)as0 / 16), hu2 = 8;
        const size_t SB = (size_t)kFW * std::max<int>(store_units(P->opt.store_dtype, P->iw), std::max(hu1, hu2)) * 16;
        const size_t room = (size_t)P->smem_optin > U + 2048 ? (size_t)P->smem_optin - U - 2048 : 0;
        int nst = (int)std::min<size_t>(kFMaxStages, room / SB);
        if (P->opt.k1_stages > 0) nst = std::min(nst, std::max(2, P->opt.k1_stages));
        const bool can = P->iw != 1 && nst >= (P->opt.k1_stages > 0 ? 2 : 6);
        const int mode = P->opt.eval_kernels & 3;
        if (mode == 2 && !can) return fail(VMAT_E_UNSUPPORTED, "fused evaluation kernel does not fit this plan (beamlet vector too large for shared memory, or byte-delta indices)");
        P->fused = can && mode == 2;
        if (P->fused) {
            P->W1 = kFW;
            P->nstagesF = nst; P->stage_bytesF = (int32_t)SB;
            P->smemF = U + (size_t)nst * SB;
        }
    }

    // ---- bring row pointers to the host, CSR payload to the device ----------------
    std::vector<int64_t> h_vrow(V + 1), h_brow(B + 1);
    const cudaMemcpyKind to_host = location == VMAT_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost;
    CU(cudaMemcpy(h_vrow.data(), vrowptr, (V + 1) * sizeof(int64_t), to_host));
    CU(cudaMemcpy(h_brow.data(), browptr, (B + 1) * sizeof(int64_t), to_host));
    const int64_t nnz = h_vrow[V];
    if (h_vrow[0] != 0 || h_brow[0] != 0) return fail(VMAT_E_ARG, "CSR row pointers must start at 0");
    for (int64_t v = 0; v < V; ++v)
        if (h_vrow[v + 1] < h_vrow[v]) return fail(VMAT_E_ARG, "voxel-major row pointers must not decrease");
    for (int64_t j = 0; j < B; ++j)
        if (h_brow[j + 1] < h_brow[j]) return fail(VMAT_E_ARG, "beamlet-major row pointers must not decrease");
    if (nnz != h_brow[B]) return fail(VMAT_E_ARG, "voxel-major and beamlet-major CSR disagree on nnz");
    if (nnz > 0 && (!vcol || !vval || !bcol || !bval)) return fail(VMAT_E_ARG, "null CSR payload");
    P->nnz = nnz;
    const size_t sv = src_dtype == VMAT_F32 ? 4 : 8;
    DevBuf up_vcol, up_vval, up_bcol, up_bval, up_brow;
    const int32_t *d_vcol = vcol, *d_bcol = bcol;
    const void *d_vval = vval, *d_bval = bval;
    const int64_t* d_brow = browptr;
    if (location == VMAT_HOST) {
        CU(up_vcol.alloc(nnz * 4)); CU(up_vval.alloc(nnz * sv));
        // on the plan's stream: a synchronous cudaMemcpy from pageable memory may return while the DMA is
        // still in flight, and the (non-blocking) plan stream is not ordered behind the default stream
        CU(cudaMemcpyAsync(up_vcol.p, vcol, nnz * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(up_vval.p, vval, nnz * sv, cudaMemcpyHostToDevice, st));
        d_vcol = up_vcol.as<int32_t>(); d_vval = up_vval.p;
    }

    DevBuf csr_flag, up_vrow;
    std::vector<uint8_t> fused_nch1, fused_tpr1;          // fused plans: per dose item chunk counts, lanes per row,
    std::vector<int32_t> fused_grp1, fused_grp_items;     // group, and the number of dose items per group
    int32_t fused_grp_tiles = 1;
    const bool fused_uniform = (P->opt.eval_kernels & 4) != 0;     // experiment: uniform (padded) stages in the fused layout
    std::vector<int32_t> vpads;            // per voxel row (byte-delta layout only)
    CU(csr_flag.alloc(4)); CU(cudaMemsetAsync(csr_flag.p, 0, 4, st));
    auto check_csr = [&](int64_t nrows, int64_t ncols, int sorted, const int64_t* d_rowptr, const int32_t* d_col,
                         const char* what, bool* unsorted = nullptr) -> int {
        if (nnz == 0) return 0;
        CU(cudaMemsetAsync(csr_flag.p, 0, 4, st));
        k_check_csr<<<(unsigned)((nrows + 255) / 256), 256, 0, st>>>(nrows, ncols, sorted, d_rowptr, d_col, csr_flag.as<int>());
        CU(cudaGetLastError());
        int bad = 0;
        CU(cudaMemcpyAsync(&bad, csr_flag.p, 4, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (bad & 1) return fail(VMAT_E_ARG, std::string(what) + " CSR: index out of range");
        if ((bad & 2) && unsorted) { *unsorted = true; return 0; }
        if (bad & 2) return fail(VMAT_E_ARG, std::string(what) + " CSR: indices must be sorted within a row");
        return 0;
    };
    {
        const int64_t* d_vrow = vrowptr;
        if (location == VMAT_HOST) {
            CU(up_vrow.alloc((V + 1) * 8));
            CU(cudaMemcpyAsync(up_vrow.p, vrowptr, (V + 1) * 8, cudaMemcpyHostToDevice, st));
            d_vrow = up_vrow.as<int64_t>();
        }
        bool unsorted = false;
        if (int rc = check_csr(V, B, P->iw == 1 ? 1 : 0, d_vrow, d_vcol, "voxel-major", iw_automatic ? &unsorted : nullptr)) return rc;
        if (unsorted) {                    // byte deltas were this library's choice, not the caller's: absolute indices then
            P->iw = iw_absolute;
            dose_shape();
        }
        if (P->iw == 1 && nnz > 0) {      // byte deltas: slots a row needs = nonzeros + bridge entries over gaps > 255
            DevBuf d_pads;
            CU(d_pads.alloc(V * 4));
            if (int rc = count_pads(P, src_dtype, V, d_vrow, d_vcol, d_vval, d_pads.as<int32_t>())) return rc;
            vpads.resize(V);
            CU(cudaMemcpyAsync(vpads.data(), d_pads.p, V * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        up_vrow.alloc(0);
    }

    // ---- K1 layout: voxels sorted by row length (descending, stable) ----------------
    const int units = store_units(P->opt.store_dtype, P->iw);
    const int es = (int)P->acc_size;
    {
        std::vector<int32_t> order(V), slots(V);      // slots: entries the row occupies in the stream
        for (int64_t v = 0; v < V; ++v) slots[v] = (int32_t)(h_vrow[v + 1] - h_vrow[v]) + (vpads.empty() ? 0 : vpads[v]);
        {
            int64_t maxlen = 0;
            for (int64_t v = 0; v < V; ++v) maxlen = std::max<int64_t>(maxlen, slots[v]);
            std::vector<int64_t> cnt(maxlen + 2, 0);
            for (int64_t v = 0; v < V; ++v) cnt[maxlen - slots[v] + 1]++;
            for (size_t k = 1; k < cnt.size(); ++k) cnt[k] += cnt[k - 1];
            for (int64_t v = 0; v < V; ++v) order[cnt[maxlen - slots[v]]++] = (int32_t)v;
        }
        const int32_t W = P->W1, HU = 8 + 3 * (32 * es / 16) + header_base_units(P->iw);
        // 16-warp stages per CTA if the work were spread evenly: small slabs (a case sharded over several GPUs)
        // get shorter super-slices, so that the longest one is not most of the kernel (config 2 on 8 GPUs:
        // 73 super-slices of up to 49 stages for 148 CTAs, the dose kernel's CTAs finished between 2 and 12.5 us)
        const int64_t per_cta = nnz / 2048 / std::max(1, P->sm_count);
        const int32_t cap = P->opt.chunk_cap > 0 ? P->opt.chunk_cap
                          : (int32_t)std::min<int64_t>(kChunkCapK1, std::max<int64_t>(12, per_cta / 2));
        StreamDesc sd{W, HU, units, P->fused ? std::min(cap, 255) : cap};
        sd.ragged = P->fused && !fused_uniform;
        sd.record_nch = P->fused;
        std::vector<StreamDesc::Item> items(V);
        for (int64_t pos = 0; pos < V; ++pos) {
            const int32_t v = order[pos];
            items[pos] = StreamDesc::Item{slots[v], v, h_vrow[v], (int32_t)(h_vrow[v + 1] - h_vrow[v])};
        }
        if (!P->fused) {
            sd.add_group(items, 0, 0, thr, over, under);
        } else {
            // groups of consecutive tiles: a transposed item of tile t can start as soon as the dose items of
            // group(t) are done; inside a group the rows stay sorted by length
            const int32_t nt = (int32_t)((V + T_plan - 1) / T_plan);
            const int32_t gt = (nt + 11) / 12;                          // tiles per group: about 12 groups
            const int32_t ng = (nt + gt - 1) / gt;
            std::vector<std::vector<StreamDesc::Item>> by_group(ng);
            for (const auto& it : items) by_group[it.id / T_plan / gt].push_back(it);
            fused_grp_tiles = gt; P->ngroupsF = ng;
            for (int32_t g = 0; g < ng; ++g) {
                const size_t before = sd.ss_tpr.size();
                sd.add_group(by_group[g], 0, g, thr, over, under);       // ss_tile carries the group here
                fused_grp_items.push_back((int32_t)(sd.ss_tpr.size() - before));
            }
            fused_nch1 = sd.ss_nch; fused_tpr1 = sd.ss_tpr; fused_grp1 = sd.ss_tile;
            if (sd.too_long) return fail(VMAT_E_UNSUPPORTED, "fused evaluation kernel: a voxel row has more than 32 640 nonzeros (use eval_kernels = 1)");
        }
        const int32_t nss = (int32_t)sd.ss_tpr.size();
        P->nslices1 = nss * W; P->nss1 = nss; P->W1 = W; P->units1 = sd.ss_off.back();
        DevBuf thr_s, over_s, under_s;
        int rc = 0;
        DISPATCH_ACC(P->opt.acc_dtype, AT, {
            rc |= upload_cast<AT>(thr_s, sd.t_s, st);
            rc |= upload_cast<AT>(over_s, sd.o_s, st);
            rc |= upload_cast<AT>(under_s, sd.u_s, st);
        });
        if (rc) return VMAT_E_CUDA;
        if (int rc2 = build_stream(P, P->opt.store_dtype, src_dtype, W, HU, sd.ss_off, sd.ss_tpr, sd.seg_start, sd.seg_len,
                                   sd.seg_cb, sd.rows, &thr_s, &over_s, &under_s, es, d_vcol, d_vval,
                                   P->sell1, P->off1, P->tpr1, sd.ragged ? &sd.ss_nch : nullptr)) return rc2;
        P->stage_bytes1 = W * std::max<int>(units, HU) * 16;
    }
    up_vcol.alloc(0); up_vval.alloc(0);   // release the voxel-major CSR copy

    if (location == VMAT_HOST) {
        CU(up_bcol.alloc(nnz * 4)); CU(up_bval.alloc(nnz * sv)); CU(up_brow.alloc((B + 1) * 8));
        CU(cudaMemcpyAsync(up_bcol.p, bcol, nnz * 4, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(up_bval.p, bval, nnz * sv, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(up_brow.p, browptr, (B + 1) * 8, cudaMemcpyHostToDevice, st));
        d_bcol = up_bcol.as<int32_t>(); d_bval = up_bval.p; d_brow = up_brow.as<int64_t>();
    }
    if (int rc = check_csr(B, V, 1, d_brow, d_bcol, "beamlet-major")) return rc;

    // ---- K2 layout: column tiles of T voxels; per tile, beamlet segments sorted by length
    {
        const int32_t T = T_plan;
        const int32_t W = P->fused ? kFW : P->opt.k2_threads >= 512 ? 16 : P->opt.k2_threads >= 256 ? 8 : 4, HU = header_units_k2(P->iw);
        P->W2 = W;
        P->stage_bytes2 = W * std::max<int>(units, HU) * 16;
        // ring depth: with byte deltas a stage is 2.5 KB and consumers take four per pass - 12 measured best on configs 2
        // and 3 (8 / 10 / 12 / 16 stages: 236.2 / 232.1 / 230.5 / 235.0 us on config 3)
        P->nstages2 = P->opt.k2_stages > 0 ? std::max(2, std::min(16, P->opt.k2_stages)) : (P->iw == 1 ? 12 : 8);
        const size_t tile_bytes = ((size_t)T * P->acc_size + 127) & ~(size_t)127;
        P->k2_smem = tile_bytes + (size_t)P->nstages2 * P->stage_bytes2;
        if (!P->fused && P->k2_smem + 2048 > (size_t)P->smem_optin) return fail(VMAT_E_ARG, "tile_voxels too large for shared memory");
        const int32_t nt = (int32_t)((V + T - 1) / T);
        P->T = T; P->ntiles = nt;
        DevBuf d_seg;
        const int64_t nseg = (int64_t)B * (nt + 1);
        CU(d_seg.alloc(nseg * 8));
        k_segment_bounds<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>((int32_t)B, nt, T, d_brow, d_bcol, d_seg.as<int64_t>());
        CU(cudaGetLastError());
        std::vector<int64_t> segstart(nseg);
        CU(cudaMemcpyAsync(segstart.data(), d_seg.p, nseg * 8, cudaMemcpyDeviceToHost, st));
        std::vector<int32_t> spads;            // per (beamlet, tile) segment (byte-delta layout only)
        if (P->iw == 1 && nnz > 0) {
            DevBuf d_pads;
            CU(d_pads.alloc(nseg * 4));
            CU(cudaMemsetAsync(d_pads.p, 0, nseg * 4, st));
            if (int rc = count_pads(P, src_dtype, nseg - 1, d_seg.as<int64_t>(), d_bcol, d_bval, d_pads.as<int32_t>())) return rc;
            spads.resize(nseg);
            CU(cudaMemcpyAsync(spads.data(), d_pads.p, nseg * 4, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        CU(cudaStreamSynchronize(st));
        d_seg.alloc(0);

        const int64_t per_cta2 = nnz / 2048 / std::max(1, P->sm_count);
        const int32_t cap = P->opt.k2_chunk_cap > 0 ? P->opt.k2_chunk_cap
                          : (int32_t)std::min<int64_t>(P->fused ? 48 : kChunkCapK2, std::max<int64_t>(8, per_cta2 / 4));
        StreamDesc sd{W, HU, units, P->fused ? std::min(cap, 255) : cap};
        sd.ragged = P->fused && !fused_uniform;
        sd.record_nch = P->fused;
        std::vector<StreamDesc::Item> items;
        for (int32_t t = 0; t < nt; ++t) {
            items.clear();
            for (int32_t b = 0; b < (int32_t)B; ++b) {
                const int64_t s0 = segstart[(int64_t)b * (nt + 1) + t], s1 = segstart[(int64_t)b * (nt + 1) + t + 1];
                if (s1 > s0) items.push_back(StreamDesc::Item{(int32_t)(s1 - s0) + (spads.empty() ? 0 : spads[(int64_t)b * (nt + 1) + t]), b, s0, (int32_t)(s1 - s0)});
            }
            std::stable_sort(items.begin(), items.end(),
                             [](const StreamDesc::Item& a, const StreamDesc::Item& b) { return a.len > b.len; });
            sd.add_group(items, t * T, t, nullptr, nullptr, nullptr);
        }
        const std::vector<int64_t>& ss_cost = sd.ss_cost;
        const std::vector<int32_t>& ss_tile = sd.ss_tile;
        P->nss2 = (int32_t)ss_cost.size(); P->nslices2 = P->nss2 * W; P->units2 = sd.ss_off.back();
        if (P->fused) {
            if (sd.too_long) return fail(VMAT_E_UNSUPPORTED, "fused evaluation kernel: a beamlet segment has more than 32 640 nonzeros (use eval_kernels = 1)");
            // the queue: dose items (group order), then transposed items (tile order)
            std::vector<FItem> fi((size_t)P->nss1 + P->nss2);
            for (int32_t i = 0; i < P->nss1; ++i) {
                FItem& f = fi[i];
                std::memset(&f, 0, sizeof(f));
                for (int w = 0; w < kFW; ++w) f.nch[w] = fused_nch1[(size_t)i * kFW + w];
                f.tile = -1; f.grp = fused_grp1[i]; f.need = 0; f.tpr = fused_tpr1[i];
            }
            // Transposed items in three passes over the tiles - long, medium, short (by stage count) - so that the
            // queue ends with small items and the tail of the kernel is a short item, not a long one.  Every
            // item fetches its own tile, so the order is free; tile order inside a pass keeps the items whose
            // group finishes last at the end of each pass.
            std::vector<int32_t> perm;
            perm.reserve(P->nss2);
            for (int pass = 0; pass < 3; ++pass)
                for (int32_t i = 0; i < P->nss2; ++i) {
                    const int64_t stages = ss_cost[i];
                    const int cls = stages > cap / 2 ? 0 : stages > std::max(2, cap / 6) ? 1 : 2;
                    if (cls == pass) perm.push_back(i);
                }
            std::vector<int64_t> off2q(P->nss2 + 1, 0);
            for (int32_t q = 0; q < P->nss2; ++q) {
                const int32_t i = perm[q];
                FItem& f = fi[(size_t)P->nss1 + q];
                std::memset(&f, 0, sizeof(f));
                for (int w = 0; w < kFW; ++w) f.nch[w] = sd.ss_nch[(size_t)i * kFW + w];
                f.tile = ss_tile[i]; f.grp = ss_tile[i] / fused_grp_tiles;
                f.need = kFW * fused_grp_items[f.grp]; f.tpr = sd.ss_tpr[i];
                off2q[q] = sd.ss_off[i];
            }
            if (upload_vec(P->offF2, off2q, st)) return VMAT_E_CUDA;
            if (upload_vec(P->itemsF, fi, st)) return VMAT_E_CUDA;
            P->nF1 = P->nss1; P->nF = P->nss1 + P->nss2;
            CU(P->syncF.alloc((size_t)(kFSyncWords + P->ngroupsF) * 4));
            CU(cudaMemsetAsync(P->syncF.p, 0, P->syncF.bytes, st));
            if (int rc = build_stream(P, P->opt.store_dtype, src_dtype, W, HU, sd.ss_off, sd.ss_tpr, sd.seg_start, sd.seg_len,
                                      sd.seg_cb, sd.rows, nullptr, nullptr, nullptr, 0, d_bcol, d_bval,
                                      P->sell2, P->off2, P->tpr2, sd.ragged ? &sd.ss_nch : nullptr)) return rc;
        } else {
        // Persistent CTAs, static partition: every CTA gets the same number of ring stages.  Within a
        // tile the super-slices run from long rows (up to cap+1 stages each) to short ones (a few
        // stages), so a CTA takes long ones from the front of the tile's range while they fit and
        // tops up with short ones from the back: piece sizes then differ by a few stages instead of
        // by a whole long super-slice (60..99 stages around a mean of 85 on config 2 before).  A CTA
        // that exhausts a tile continues in the next one (one more vg tile load).
        int occ = 1;
        if (int rc = k2_occupancy(P, &occ)) return rc;
        const int64_t G = P->opt.k2_tasks > 0 ? P->opt.k2_tasks : (int64_t)P->sm_count * occ;
        P->k2_grid = (int)std::max<int64_t>(1, std::min<int64_t>(G, std::max<int64_t>(1, (int64_t)P->nss2)));
        // Dynamic tail: the shortest super-slices of every tile (its last ones) - about k2_pool_pct percent of the
        // tile's stages - are kept out of the static pieces and cut into small tasks that CTAs claim from a counter
        // once their own piece is done.  Equal stage counts do not mean equal times (the pieces of config 3 are
        // equal to 0.5 % and finish between 121 and 136 us), and a CTA that is done early has nothing else to do.
        const int32_t pool_pct = P->opt.k2_pool_pct < 0 ? 0 : (P->opt.k2_pool_pct > 0 ? std::min(P->opt.k2_pool_pct, 90) : 12);
        std::vector<int32_t> pool_begin(P->nss2 + 1, 0);      // per super-slice index of a tile's first one: where its pool part starts
        std::vector<K2Task> pool;
        {
            const int32_t nss = P->nss2;
            const int64_t task_cost = 8;                        // stages per pool task (one or a few short super-slices)
            int32_t t0 = 0;
            while (t0 < nss) {
                int32_t t1 = t0;
                int64_t tile_w = 0;
                while (t1 < nss && ss_tile[t1] == ss_tile[t0]) tile_w += ss_cost[t1++];
                int32_t pb = t1;
                int64_t w = 0;
                while (pb > t0 + 1 && (w + ss_cost[pb - 1]) * 100 <= tile_w * pool_pct) w += ss_cost[--pb];
                for (int32_t k = t0; k < t1; ++k) pool_begin[k] = pb;
                int32_t a0 = pb;
                int64_t acc = 0;
                for (int32_t k = pb; k < t1; ++k) {
                    acc += ss_cost[k];
                    if (acc >= task_cost || k == t1 - 1) { pool.push_back(K2Task{ss_tile[t0], a0, k + 1, 0}); a0 = k + 1; acc = 0; }
                }
                t0 = t1;
            }
        }
        int64_t total_w = 0;
        for (int32_t k = 0; k < P->nss2; ++k) if (k < pool_begin[k]) total_w += ss_cost[k];
        std::vector<K2Task> tasks;
        std::vector<int32_t> cta_ptr(P->k2_grid + 1, 0);
        {
            const int32_t nss = P->nss2;
            int32_t f = 0, bk = 0, tend = 0;         // unassigned super-slices of the current tile: [f, bk); the tile ends at tend
            auto next_tile = [&]() {
                f = tend;
                while (tend < nss && ss_tile[tend] == ss_tile[f]) ++tend;
                bk = f < nss ? pool_begin[f] : tend;  // the rest of the tile belongs to the pool
            };
            next_tile();
            int64_t assigned = 0;
            for (int c = 0; c < P->k2_grid; ++c) {
                cta_ptr[c] = (int32_t)tasks.size();
                const bool last = c == P->k2_grid - 1;
                const int64_t goal = total_w * (c + 1) / P->k2_grid;
                while (f < nss && (last || assigned < goal)) {
                    const int32_t tile = ss_tile[f], f0 = f, b0 = bk;
                    while (f < bk && (last || assigned + ss_cost[f] <= goal)) assigned += ss_cost[f++];
                    while (f < bk && (last || assigned + ss_cost[bk - 1] <= goal)) assigned += ss_cost[--bk];
                    // nothing fits any more: take the smallest remaining one only if that ends closer to the goal
                    if (f < bk && assigned < goal && 2 * (goal - assigned) > ss_cost[bk - 1]) assigned += ss_cost[--bk];
                    if (f > f0) tasks.push_back(K2Task{tile, f0, f, 0});
                    if (bk < b0) tasks.push_back(K2Task{tile, bk, b0, 0});
                    if (f < bk) break;               // goal reached inside this tile
                    next_tile();                     // tile exhausted: continue in the next one
                }
            }
            cta_ptr[P->k2_grid] = (int32_t)tasks.size();
        }
        if (upload_vec(P->cta_ptr2, cta_ptr, st)) return VMAT_E_CUDA;
        P->ntasks2 = (int32_t)tasks.size();
        if (int rc = build_stream(P, P->opt.store_dtype, src_dtype, W, HU, sd.ss_off, sd.ss_tpr, sd.seg_start, sd.seg_len,
                                  sd.seg_cb, sd.rows, nullptr, nullptr, nullptr, 0, d_bcol, d_bval,
                                  P->sell2, P->off2, P->tpr2)) return rc;
        if (upload_vec(P->tasks2, tasks, st)) return VMAT_E_CUDA;
        if (upload_vec(P->pool2, pool, st)) return VMAT_E_CUDA;
        P->npool2 = (int32_t)pool.size();
        }
    }
    up_bcol.alloc(0); up_bval.alloc(0); up_brow.alloc(0);

    // ---- evaluation state -----------------------------------------------------------
    {
        std::vector<int32_t> beam_of(B);
        for (int32_t i = 0; i < nb; ++i)
            for (int32_t j = beam_ptr[i]; j < beam_ptr[i + 1]; ++j) beam_of[j] = i;
        if (upload_vec(P->beam_of, beam_of, st)) return VMAT_E_CUDA;
        if (upload_vec(P->beam_ptr_d, P->beam_ptr, st)) return VMAT_E_CUDA;
    }
    const size_t as = P->acc_size;
    CU(P->y.alloc(nb * 8));            CU(cudaMemset(P->y.p, 0, nb * 8));
    CU(P->w_dose.alloc(B * as));       CU(cudaMemset(P->w_dose.p, 0, B * as));
    CU(P->w_grad.alloc(B * 8));        CU(cudaMemset(P->w_grad.p, 0, B * 8));
    CU(P->xg.alloc(B * as));           CU(cudaMemset(P->xg.p, 0, B * as));
    CU(P->d.alloc(V * as));            CU(cudaMemset(P->d.p, 0, V * as));
    CU(P->vg.alloc((V + P->T) * as));  CU(cudaMemset(P->vg.p, 0, (V + P->T) * as));   // padded: tile copies may run past V
    CU(P->partial.alloc((size_t)P->ntiles * B * as));
    CU(cudaMemset(P->partial.p, 0, (size_t)P->ntiles * B * as));
    CU(P->gbf.alloc((B + 1) * 8));     CU(cudaMemset(P->gbf.p, 0, (B + 1) * 8));
    CU(P->result.alloc((nb + 2) * 8)); CU(cudaMemset(P->result.p, 0, (nb + 2) * 8));   // f, g[nb], exchange status
    CU(P->counter.alloc(64));          CU(cudaMemset(P->counter.p, 0, 64));
    const size_t nfblock = (size_t)std::max(std::max(1, P->k2_grid), 4 * P->sm_count);
    CU(P->fblock.alloc(nfblock * 8)); CU(cudaMemset(P->fblock.p, 0, nfblock * 8));
    CU(P->scratch64.alloc(std::max<int64_t>(V, B) * 8));
    {
        std::vector<double> tt(thr, thr + V), oo(over, over + V), uu(under, under + V);
        int rc = 0;
        DISPATCH_ACC(P->opt.acc_dtype, AT, {
            rc |= upload_cast<AT>(P->thr_n, tt, st); rc |= upload_cast<AT>(P->over_n, oo, st); rc |= upload_cast<AT>(P->under_n, uu, st);
        });
        if (rc) return VMAT_E_CUDA;
    }
    CU(P->comm_state.alloc(64));         CU(cudaMemset(P->comm_state.p, 0, 64));
    for (int32_t i = 0; i < nb; ++i) P->max_beamlets = std::max(P->max_beamlets, beam_ptr[i + 1] - beam_ptr[i]);
    CU(cudaHostAlloc(&P->h_w, (size_t)std::max(1, P->max_beamlets) * 16, cudaHostAllocDefault));
    CU(cudaHostAlloc(&P->h_y, nb * 8, cudaHostAllocMapped));
    CU(cudaHostAlloc(&P->h_res, (nb + 2) * 8, cudaHostAllocMapped));
    CU(cudaHostAlloc(&P->h_tags, (nb + 1) * 8, cudaHostAllocMapped));
    std::memset(P->h_y, 0, nb * 8); std::memset(P->h_res, 0, (nb + 2) * 8); std::memset(P->h_tags, 0, (nb + 1) * 8);
    CU(cudaHostGetDevicePointer(&P->d_hy, P->h_y, 0));
    CU(cudaHostGetDevicePointer(&P->d_hres, P->h_res, 0));
    CU(cudaHostGetDevicePointer(&P->d_htags, P->h_tags, 0));
    CU(cudaHostAlloc(&P->h_guard, kGuardWords * 8, cudaHostAllocMapped));
    std::memset(P->h_guard, 0, kGuardWords * 8);
    CU(cudaHostGetDevicePointer(&P->d_guard, P->h_guard, 0));

    // K1 launch shape: persistent CTAs of W1 consumer warps + a producer warp; x (and y) staged
    // once per CTA, the rest of the CTA's shared memory is the TMA ring
    if (P->fused) {
        int occ = 0;
        if (int rc = fused_occupancy(P, &occ)) return rc;
        // one CTA per SM (the grid barrier needs every CTA resident; k2_tasks = explicit grid size, for tests)
        P->gridF = P->opt.k2_tasks > 0 ? std::min(P->opt.k2_tasks, P->sm_count * occ) : P->sm_count;
        P->x_in_smem = true;
        P->host_path &= 4;               // no zero-copy y / tagged host results on this path
    } else {
        const size_t xbytes = (((size_t)B * as + 127) & ~(size_t)127) + (((size_t)nb * 8 + 127) & ~(size_t)127);   // x and y
        P->x_in_smem = xbytes + 2 * (size_t)P->stage_bytes1 + 4096 <= (size_t)P->smem_optin;
        const size_t avail = (size_t)P->smem_optin - 4096 - (P->x_in_smem ? xbytes : 0);
        const size_t want = P->opt.k1_stages > 0 ? (size_t)P->opt.k1_stages : (P->W1 == 4 ? 8 : 16);
        P->nstages1 = (int)std::max<size_t>(2, std::min<size_t>({(size_t)16, want, avail / P->stage_bytes1}));
        P->k1_smem = (P->x_in_smem ? xbytes : 0) + (size_t)P->nstages1 * P->stage_bytes1;
        P->k1_threads = (P->W1 + 1) * 32;
        int occ = 1;
        if (int rc = k1_occupancy(P, &occ)) return rc;
        P->k1_grid = P->sm_count * occ;
    }
    CU(P->fpart.alloc((size_t)P->nslices1 * 8)); CU(cudaMemset(P->fpart.p, 0, (size_t)P->nslices1 * 8));
    CU(cudaDeviceSynchronize());

    if (P->opt.use_graph) { if (int rc = capture_graph(P)) return rc; }
    guard.p = nullptr;
    *out = P;
    return VMAT_OK;
}

extern "C" int vmat_plan_destroy(vmat_plan_t* P) {
    if (!P) return VMAT_OK;
    cudaSetDevice(P->device);
    cudaDeviceSynchronize();
    delete P;
    return VMAT_OK;
}

extern "C" int vmat_plan_info(const vmat_plan_t* P, int64_t* nnz, int64_t* layout_bytes, int64_t* algo_bytes,
                              int32_t* launches, int32_t* index_bytes) {
    if (!P) return fail(VMAT_E_ARG, "null plan");
    const int64_t as = (int64_t)P->acc_size, sv = store_bytes(P->opt.store_dtype);
    if (nnz) *nnz = P->nnz;
    if (layout_bytes) {
        // what the three kernels stream: both SELL layouts (padding included), slice metadata,
        // penalty tables, d/vg writes, vg tile reads, tile partials (write + read), x inputs, gb.
        int64_t b = (P->units1 + P->units2) * 16;
        b += (int64_t)(P->nss1 + P->nss2) * 8;            // super-slice offsets (headers are inside the streams)
        b += 2 * P->V * as + P->V * as;                     // d, vg written; vg read
        b += 2 * (int64_t)P->ntiles * P->B * as;
        b += P->B * (as + 4) + P->B * 8 * 2;
        *layout_bytes = b;
    }
    if (algo_bytes) *algo_bytes = 2 * P->nnz * (sv + P->iw) + 28 * P->V + 12 * P->B;    // SURVEY 8(d) with this plan's index width
    if (launches) *launches = P->fused ? 1 : P->x_in_smem ? 3 : 4;
    if (index_bytes) *index_bytes = P->iw;
    return VMAT_OK;
}

// ---------------------------------------------------------------------------------
// evaluation
// ---------------------------------------------------------------------------------
// ring stages a consumer warp takes per pass (their barrier tests, loads and gathers in flight together).  Two by
// default: four measured the same on configs 2 and 3 (59.5 / 59.0 us and 242.4 / 241.6 us) and, the warp holding
// them until all have arrived, lost 30 % on a shallow ring (config 3 with 8-byte accumulation, 7 stages)
static int consume_batch_for(const vmat_plan* P, int nstages, int kernel) {
    // option: low 4 bits = both kernels, or (tuning) bits 4..7 = the transposed kernel's own value
    int o = P->opt.consume_batch > 0 ? P->opt.consume_batch : 0;
    if (kernel == 2 && (o >> 4)) o >>= 4;
    o &= 15;
    int b = o > 0 ? o : 2;
    b = b >= 4 ? 4 : b >= 2 ? 2 : 1;
    while (b > 1 && b > nstages) b >>= 1;
    return b;
}

// the fused path: one launch of kf_eval (phase 0 or 1: everything up to gbf/result, with the exchange when
// ranks are attached and phase == 0; phase 2 is the aperture gradient from an externally all-reduced gbf)
template <typename VT, typename AT>
static int launch_fused_t(vmat_plan* P, cudaStream_t st, int phase, bool inline_y, cudaEvent_t* ev) {
    FArgs<AT> a{};
    a.sell = P->sell1.as<uint4>(); a.sell2 = P->sell2.as<uint4>();
    a.it_off = P->off1.as<int64_t>(); a.it_off2 = P->offF2.as<int64_t>();
    a.items = P->itemsF.as<FItem>(); a.n1 = P->nF1; a.nitems = P->col_mode ? P->nF1 : P->nF;
    a.nstages = P->nstagesF; a.stage_bytes = P->stage_bytesF;
    a.B = (int32_t)P->B; a.nb = P->nb; a.T = P->T; a.ntiles = P->ntiles; a.ngroups = P->ngroupsF; a.V = P->V;
    a.y = P->col_mode ? P->col_y.as<double>() : P->y.as<double>();
    a.beam_of = P->beam_of.as<int32_t>();
    a.w_dose = P->col_mode ? static_cast<const AT*>(P->col_w_ptr) : P->w_dose.as<AT>();
    a.d = P->d.as<AT>(); a.vg = P->vg.as<AT>(); a.fpart = P->fpart.as<double>(); a.partial = P->partial.as<AT>();
    a.sync = P->syncF.as<unsigned int>(); a.fblock = P->fblock.as<double>();
    a.beam_ptr = P->beam_ptr_d.as<int32_t>(); a.w_grad = P->w_grad.as<double>();
    a.gbf = P->gbf.as<double>(); a.result = P->result.as<double>();
    a.claim_lead = P->opt.claim_lead > 0 ? P->opt.claim_lead : kClaimLeadDefault;
    a.zero = 0u;
    a.dose_only = P->col_mode ? 1 : 0;
    a.trace = P->tracing ? P->trace.as<unsigned long long>() : nullptr;
    if (inline_y) {
        a.y_mirror = P->y.as<double>();
        a.y_inline_n = P->nb;
        std::memcpy(a.y_inline, P->h_y, (size_t)P->nb * 8);
    }
    CommArgs c{};
    c.nranks = 1;
    if (P->comm_nranks > 1 && phase == 0 && !P->col_mode) {
        c.peer_bufs = P->comm_ptrs.as<double*>(); c.rank = P->comm_rank; c.nranks = P->comm_nranks;
        c.epoch = ++P->comm_epoch; c.timeout_ns = P->comm_timeout_ns;
        c.error = P->comm_state.as<int32_t>() + 8; c.status = P->result.as<double>() + P->nb + 1;
    }
    DISPATCH_IWF(IW, {
        auto kern = kf_eval<VT, AT, IW>;
        CU(launch_pdl(kern, P->gridF, kFThreads, P->smemF, st, P->pdl, a, c));
    });
    CU(cudaGetLastError());
    if (P->col_mode) CU(cudaMemsetAsync(P->syncF.p, 0, P->syncF.bytes, st));    // dose-only launches skip the barrier that resets the queue
    if (ev) { CU(cudaEventRecord(ev[1], st)); CU(cudaEventRecord(ev[2], st)); CU(cudaEventRecord(ev[3], st)); }
    return 0;
}

template <typename VT, typename AT>
static int launch_eval_t(vmat_plan* P, cudaStream_t st, int phase, bool host_mode, bool inline_y) {
    cudaEvent_t* ev = nullptr;
    if (P->timing && phase == 0) {
        ev = &P->tev[(P->tcount % vmat_plan::kTimingSlots) * 4];
        P->tcount++;
        CU(cudaEventRecord(ev[0], st));
    }
    if (P->fused && phase != 2) return launch_fused_t<VT, AT>(P, st, phase, inline_y, ev);
    if (phase != 2) {
        if (!P->x_in_smem) {
            if (host_mode) CU(cudaMemcpyAsync(P->y.p, P->h_y, P->nb * 8, cudaMemcpyHostToDevice, st));
            k4_expand_x<AT><<<(unsigned)((P->B + 255) / 256), 256, 0, st>>>(
                (int32_t)P->B, P->col_mode ? P->col_y.as<double>() : P->y.as<double>(), P->beam_of.as<int32_t>(),
                P->col_mode ? static_cast<const AT*>(P->col_w_ptr) : P->w_dose.as<AT>(), P->xg.as<AT>());
        }
        K1Args<AT> a1{};
        a1.sell = P->sell1.as<uint4>(); a1.ss_off = P->off1.as<int64_t>(); a1.ss_tpr = P->tpr1.as<uint8_t>(); a1.nss = P->nss1;
        a1.nstages = P->nstages1; a1.stage_bytes = P->stage_bytes1;
        a1.B = (int32_t)P->B; a1.y = P->y.as<double>(); a1.beam_of = P->beam_of.as<int32_t>();
        a1.w_dose = P->w_dose.as<AT>(); a1.x_global = P->xg.as<AT>();
        a1.d = P->d.as<AT>(); a1.vg = P->vg.as<AT>(); a1.fpart = P->fpart.as<double>();
        a1.counter = P->counter.as<unsigned int>();
        a1.nb = P->nb;
        a1.claim_lead = P->opt.claim_lead > 0 ? P->opt.claim_lead : kClaimLeadDefault;
        a1.trace = P->tracing ? P->trace.as<unsigned long long>() : nullptr;
        a1.k2_pool_counter = (P->npool2 > 0 && !P->col_mode) ? P->counter.as<unsigned int>() + 2 : nullptr;
        a1.batch = consume_batch_for(P, P->nstages1, 1);
        a1.guard = Guard{P->d_guard, (uint32_t)P->B, 1};
        a1.zero = 0u;
        if (inline_y) {              // the intensities travel in the launch parameters; CTA 0 mirrors them into P->y
            a1.y_mirror = P->y.as<double>();
            a1.y_inline_n = P->nb;
            std::memcpy(a1.y_inline, P->h_y, (size_t)P->nb * 8);
        }
        if (P->col_mode) { a1.y = P->col_y.as<double>(); a1.w_dose = static_cast<const AT*>(P->col_w_ptr); }
        if (host_mode && (P->host_path & 1)) { a1.y = P->d_hy; a1.y_mirror = P->y.as<double>(); a1.seq = P->counter.as<unsigned long long>() + 4; }
        else if (host_mode) { a1.seq = P->counter.as<unsigned long long>() + 4; a1.y_mirror = P->y.as<double>(); }
        DISPATCH_IW(IW, { DISPATCH_W1(W, {
            if (P->x_in_smem) {
                auto kern = k1_dose_penalty<VT, AT, true, W, IW>;
                CU(launch_pdl(kern, P->k1_grid, P->k1_threads, P->k1_smem, st, P->pdl, a1));
            } else {
                auto kern = k1_dose_penalty<VT, AT, false, W, IW>;
                CU(launch_pdl(kern, P->k1_grid, P->k1_threads, P->k1_smem, st, P->pdl, a1));
            }
        }); });
        if (P->col_mode) { CU(cudaGetLastError()); return 0; }      // column build: the dose kernel only
        if (ev) CU(cudaEventRecord(ev[1], st));
        K2Args<AT> a2{};
        a2.sell = P->sell2.as<uint4>(); a2.ss_off = P->off2.as<int64_t>(); a2.ss_tpr = P->tpr2.as<uint8_t>();
        a2.tasks = P->tasks2.as<K2Task>(); a2.cta_task_ptr = P->cta_ptr2.as<int32_t>();
        a2.vg = P->vg.as<AT>(); a2.partial = P->partial.as<AT>();
        a2.V = P->V; a2.B = (int32_t)P->B; a2.T = P->T; a2.nstages = P->nstages2; a2.stage_bytes = P->stage_bytes2;
        a2.k1_counter = P->counter.as<unsigned int>();
        a2.fpart = P->fpart.as<double>(); a2.fblock = P->fblock.as<double>(); a2.nfpart = P->nslices1;
        a2.trace = P->tracing ? P->trace.as<unsigned long long>() + (size_t)P->k1_grid * kTraceSlots : nullptr;
        a2.pool_tasks = P->pool2.as<K2Task>(); a2.npool = P->npool2; a2.pool_counter = P->counter.as<unsigned int>() + 2;
        a2.batch = consume_batch_for(P, P->nstages2, 2);
        a2.guard = Guard{P->d_guard, (uint32_t)P->T, 2};
        a2.zero = 0u;
        DISPATCH_IW(IW, {   // always launched (even with an empty task list): it also folds the objective partials
            DISPATCH_W2(W2c, {
                auto kern2 = k2_beamlet_grad<VT, AT, W2c, IW>;
                CU(launch_pdl(kern2, P->k2_grid, (W2c + 1) * 32, P->k2_smem, st, P->pdl, a2));
            });
        });
    }
    if (ev) CU(cudaEventRecord(ev[2], st));
    K3Args<AT> a3{};
    a3.partial = P->partial.as<AT>(); a3.ntiles = P->ntiles; a3.B = (int32_t)P->B; a3.nb = P->nb;
    a3.beam_ptr = P->beam_ptr_d.as<int32_t>(); a3.w_grad = P->w_grad.as<double>();
    a3.fpart = P->fpart.as<double>(); a3.nfpart = P->nslices1; a3.gbf = P->gbf.as<double>();
    a3.result = P->result.as<double>();
    a3.fblock = P->fblock.as<double>(); a3.nfblock = P->k2_grid;
    a3.trace = P->tracing ? P->trace.as<unsigned long long>() + ((size_t)P->k1_grid + P->k2_grid) * kTraceSlots : nullptr;
    if (P->comm_nranks > 1 && phase == 0) {
        CommArgs c{};
        c.peer_bufs = P->comm_ptrs.as<double*>(); c.rank = P->comm_rank; c.nranks = P->comm_nranks;
        c.epoch = ++P->comm_epoch; c.timeout_ns = P->comm_timeout_ns;
        c.error = P->comm_state.as<int32_t>() + 8; c.status = P->result.as<double>() + P->nb + 1;
        CU(launch_pdl(k3_reduce_exchange<AT>, P->nb, 256, 0, st, P->pdl, a3, c));
        CU(cudaGetLastError());
        if (ev) CU(cudaEventRecord(ev[3], st));
        return 0;
    }
    if (host_mode && (P->host_path & 2)) { a3.host_result = P->d_hres; a3.host_tags = P->d_htags; a3.seq = P->counter.as<unsigned long long>() + 4; }
    CU(launch_pdl(k3_reduce<AT>, P->nb, 256, 0, st, P->pdl, a3, phase));
    CU(cudaGetLastError());
    if (ev) CU(cudaEventRecord(ev[3], st));
    return 0;
}

static int launch_eval(vmat_plan* P, cudaStream_t st, int phase, bool host_mode, bool inline_y) {
    DISPATCH_STORE(P->opt.store_dtype, VT, {
        DISPATCH_ACC(P->opt.acc_dtype, AT, { return launch_eval_t<VT, AT>(P, st, phase, host_mode, inline_y); });
    });
    return 0;
}

static int capture_graph(vmat_plan* P) {
    // warm the kernels (sets the shared-memory attributes outside capture)
    if (int rc = launch_eval(P, P->stream, 0)) return rc;
    CU(cudaStreamSynchronize(P->stream));
    CU(cudaStreamBeginCapture(P->stream, cudaStreamCaptureModeThreadLocal));
    if (!(P->host_path & 1)) cudaMemcpyAsync(P->y.p, P->h_y, P->nb * 8, cudaMemcpyHostToDevice, P->stream);
    int rc = launch_eval(P, P->stream, 0, true);       // host-buffer path: zero-copy y in, results + tags out
    if (!(P->host_path & 2)) cudaMemcpyAsync(P->h_res, P->result.p, (P->nb + 1) * 8, cudaMemcpyDeviceToHost, P->stream);
    cudaError_t e = cudaStreamEndCapture(P->stream, &P->graph);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(VMAT_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    CU(cudaGraphInstantiate(&P->graph_exec, P->graph, 0));
    return 0;
}

extern "C" int vmat_set_aperture(vmat_plan_t* P, int32_t beam, const double* w_dose, const double* w_grad) {
    if (!P || beam < 0 || beam >= P->nb) return fail(VMAT_E_ARG, "vmat_set_aperture: bad beam");
    CU(cudaSetDevice(P->device));
    const int32_t lo = P->beam_ptr[beam], n = P->beam_ptr[beam + 1] - lo;
    if (n == 0) return VMAT_OK;
    if (P->ap_epoch.empty()) P->ap_epoch.assign(P->nb, 0);
    P->ap_epoch[beam]++;                 // cached RMP columns of this control point are stale now
    const size_t as = P->acc_size;
    // One synchronisation, in front: earlier evaluations have read the old weights and the pinned staging buffer
    // is free again.  The weights are converted into the staging buffer (so the caller's arrays may go away) and
    // copied with stream-ordered asynchronous copies: later launches on the plan's stream see the new weights.
    CU(cudaStreamSynchronize(P->stream));
    char* stage = static_cast<char*>(P->h_w);
    if (w_dose) {
        if (as == 4) { float* t = reinterpret_cast<float*>(stage); for (int k = 0; k < n; ++k) t[k] = (float)w_dose[k]; }
        else std::memcpy(stage, w_dose, (size_t)n * 8);
        CU(cudaMemcpyAsync(P->w_dose.as<char>() + (size_t)lo * as, stage, (size_t)n * as, cudaMemcpyHostToDevice, P->stream));
    } else {
        CU(cudaMemsetAsync(P->w_dose.as<char>() + (size_t)lo * as, 0, (size_t)n * as, P->stream));
    }
    if (w_grad) {
        char* sg = stage + (size_t)P->max_beamlets * 8;
        std::memcpy(sg, w_grad, (size_t)n * 8);
        CU(cudaMemcpyAsync(P->w_grad.as<double>() + lo, sg, (size_t)n * 8, cudaMemcpyHostToDevice, P->stream));
    } else {
        CU(cudaMemsetAsync(P->w_grad.as<double>() + lo, 0, (size_t)n * 8, P->stream));
    }
    P->weights_in_flight = true;
    return VMAT_OK;
}

extern "C" int vmat_set_intensities(vmat_plan_t* P, const double* y) {
    if (!P || !y) return fail(VMAT_E_ARG, "vmat_set_intensities: null");
    CU(cudaSetDevice(P->device));
    std::memcpy(P->h_y, y, P->nb * 8);
    CU(cudaMemcpyAsync(P->y.p, P->h_y, P->nb * 8, cudaMemcpyHostToDevice, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    return VMAT_OK;
}

// A CTA of the fused exchange that gave up waiting for a peer sets the status word that travels
// behind the results (result[nb+1]); it is cleared here so that the next evaluation reports afresh.
// After a time-out this rank's caller sees an error while its peers carry on: they time out in
// their next evaluation (nobody delivers this rank's part any more) and fail the same way.
static int check_comm(vmat_plan* P) {
    if (P->comm_nranks <= 1 || P->h_res[P->nb + 1] == 0.0) return 0;
    P->h_res[P->nb + 1] = 0.0;
    CU(cudaMemsetAsync(P->result.as<double>() + P->nb + 1, 0, 8, P->stream));
    CU(cudaMemsetAsync(P->comm_state.as<int32_t>() + 8, 0, 4, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    return fail(VMAT_E_CUDA, "multi-GPU exchange timed out: a peer rank did not evaluate in lock step");
}

extern "C" int vmat_eval(vmat_plan_t* P, const double* y, double* f, double* g) {
    if (!P || !y) return fail(VMAT_E_ARG, "vmat_eval: null");
    CU(cudaSetDevice(P->device));
    std::memcpy(P->h_y, y, P->nb * 8);
    const bool multi = P->comm_nranks > 1;            // fused exchange: plain copies, no host tags
    const bool inline_y = (P->host_path & 4) && P->nb <= kInlineY && P->x_in_smem;
    const bool zc_y = ((P->host_path & 1) && !multi) || inline_y, zc_res = (P->host_path & 2) && !multi;
    const unsigned long long expect = ++P->host_seq;
    if (P->graph_exec && !multi && !inline_y) {
        CU(cudaGraphLaunch(P->graph_exec, P->stream));
    } else {
        if (!zc_y) CU(cudaMemcpyAsync(P->y.p, P->h_y, P->nb * 8, cudaMemcpyHostToDevice, P->stream));
        if (int rc = launch_eval(P, P->stream, 0, !multi, inline_y)) return rc;
        if (!zc_res) CU(cudaMemcpyAsync(P->h_res, P->result.p, (P->nb + (multi ? 2 : 1)) * 8, cudaMemcpyDeviceToHost, P->stream));
    }
    if (!zc_res) {
        CU(cudaStreamSynchronize(P->stream));
        if (int rc = check_comm(P)) return rc;
        P->evaluated = true;
        if (f) *f = P->h_res[0];
        if (g) std::memcpy(g, P->h_res + 1, P->nb * 8);
        return VMAT_OK;
    }
    // K3 tags every result slot with the evaluation's sequence number: poll the tags instead of
    // paying a stream synchronisation; fall back to the stream's status now and then for errors
    volatile unsigned long long* tags = P->h_tags;
    for (unsigned long long spins = 1;; ++spins) {
        int k = 0;
        while (k <= P->nb && tags[k] == expect) ++k;
        if (k > P->nb) break;
        if ((spins & 0xfff) == 0) {
            const cudaError_t q = cudaStreamQuery(P->stream);
            if (q != cudaSuccess && q != cudaErrorNotReady) return fail(VMAT_E_CUDA, std::string("vmat_eval: ") + cudaGetErrorString(q));
            if (q == cudaSuccess) {                     // stream drained: the tags must be there
                k = 0;
                while (k <= P->nb && tags[k] == expect) ++k;
                if (k > P->nb) break;
                return fail(VMAT_E_CUDA, "vmat_eval: evaluation finished without publishing its result");
            }
        }
    }
    P->evaluated = true;
    if (f) *f = P->h_res[0];
    if (g) std::memcpy(g, P->h_res + 1, P->nb * 8);
    return VMAT_OK;
}

extern "C" int vmat_eval_async(vmat_plan_t* P, void* stream, int phase) {
    if (!P || phase < 0 || phase > 2) return fail(VMAT_E_ARG, "vmat_eval_async: bad argument");
    CU(cudaSetDevice(P->device));
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : P->stream;
    if (P->weights_in_flight) {          // a caller's stream is not ordered behind the weight copies on the plan's stream
        if (st != P->stream) CU(cudaStreamSynchronize(P->stream));
        P->weights_in_flight = false;
    }
    P->evaluated = true;
    return launch_eval(P, st, phase);
}

extern "C" int vmat_fetch_result(vmat_plan_t* P, double* f, double* g) {
    if (!P) return fail(VMAT_E_ARG, "null plan");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(P->h_res, P->result.p, (P->nb + 2) * 8, cudaMemcpyDeviceToHost));
    if (int rc = check_comm(P)) return rc;
    if (f) *f = P->h_res[0];
    if (g) std::memcpy(g, P->h_res + 1, P->nb * 8);
    return VMAT_OK;
}

// The index guard's record (all zero while it has not tripped): see Guard / guard_trip in vmat_eval_kernels.cuh.
// Readable after a failed call too - the record lives in pinned host memory.
extern "C" int vmat_guard_status(const vmat_plan_t* P, uint64_t* out, int32_t nwords) {
    if (!P || !out || nwords < 1) return fail(VMAT_E_ARG, "vmat_guard_status: null");
    for (int32_t k = 0; k < nwords; ++k) out[k] = k < kGuardWords && P->h_guard ? (uint64_t)((volatile unsigned long long*)P->h_guard)[k] : 0;
    return VMAT_OK;
}

extern "C" int vmat_sync(vmat_plan_t* P) {
    if (!P) return fail(VMAT_E_ARG, "null plan");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    return VMAT_OK;
}

extern "C" int vmat_set_timing(vmat_plan_t* P, int enable) {
    if (!P) return fail(VMAT_E_ARG, "null plan");
    CU(cudaSetDevice(P->device));
    if (enable && P->tev.empty()) {
        P->tev.resize((size_t)vmat_plan::kTimingSlots * 4);
        for (auto& e : P->tev) CU(cudaEventCreate(&e));
    }
    P->timing = enable != 0;
    P->tcount = 0;
    return VMAT_OK;
}

extern "C" int vmat_get_timing(vmat_plan_t* P, double* ms, int32_t* n) {
    if (!P || !ms || !n) return fail(VMAT_E_ARG, "null");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    const int64_t cnt = std::min<int64_t>(P->tcount, vmat_plan::kTimingSlots);
    double acc[4] = {0, 0, 0, 0};
    for (int64_t k = 0; k < cnt; ++k) {
        cudaEvent_t* ev = &P->tev[k * 4];
        float t;
        CU(cudaEventElapsedTime(&t, ev[0], ev[1])); acc[0] += t;
        CU(cudaEventElapsedTime(&t, ev[1], ev[2])); acc[1] += t;
        CU(cudaEventElapsedTime(&t, ev[2], ev[3])); acc[2] += t;
        CU(cudaEventElapsedTime(&t, ev[0], ev[3])); acc[3] += t;
    }
    for (int k = 0; k < 4; ++k) ms[k] = cnt ? acc[k] / cnt : 0.0;
    *n = (int32_t)cnt;
    return VMAT_OK;
}

extern "C" int vmat_trace_eval(vmat_plan_t* P, uint64_t* out, int64_t cap, int32_t* k1_ctas, int32_t* k2_ctas) {
    if (!P || !out || !k1_ctas || !k2_ctas) return fail(VMAT_E_ARG, "null");
    const size_t n = P->fused ? (size_t)P->gridF * (kTraceSlots + kFTraceItems * 8)
                              : ((size_t)P->k1_grid + (size_t)P->k2_grid + (size_t)P->nb) * kTraceSlots;      // K1, K2, then K3's CTAs
    if (cap < (int64_t)n) return fail(VMAT_E_ARG, "vmat_trace_eval: buffer too small");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    CU(P->trace.alloc(n * 8));
    CU(cudaMemsetAsync(P->trace.p, 0, n * 8, P->stream));
    P->tracing = true;
    const int rc = launch_eval(P, P->stream, 0);
    P->tracing = false;
    if (rc) return rc;
    P->evaluated = true;
    CU(cudaMemcpyAsync(out, P->trace.p, n * 8, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    *k1_ctas = P->fused ? P->gridF : P->k1_grid; *k2_ctas = P->fused ? 0 : P->k2_grid;
    return VMAT_OK;
}

extern "C" int vmat_reduce_buffer(vmat_plan_t* P, void** dev_ptr, int64_t* count) {
    if (!P || !dev_ptr || !count) return fail(VMAT_E_ARG, "null");
    *dev_ptr = P->gbf.p; *count = P->B + 1;
    return VMAT_OK;
}

// comm buffer: recv[2 parities][kMaxRanks][B+1] entries of 16 bytes (value + epoch tags), allocated on first use
static int comm_buffer(vmat_plan* P) {
    if (P->comm_buf.p) return 0;
    const size_t bytes = 2 * (size_t)kMaxRanks * (size_t)(P->B + 1) * 16;
    CU(P->comm_buf.alloc(bytes));
    CU(cudaMemset(P->comm_buf.p, 0, bytes));
    return 0;
}

extern "C" int vmat_comm_handle(vmat_plan_t* P, void* handle_out) {
    if (!P || !handle_out) return fail(VMAT_E_ARG, "null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(P->device));
    if (int rc = comm_buffer(P)) return rc;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, P->comm_buf.p));
    std::memcpy(handle_out, &h, 64);
    return VMAT_OK;
}

// common tail of the two attach entry points: `ptrs[q]` = rank q's comm buffer as seen from this device
static int comm_finish_attach(vmat_plan* P, int32_t rank, int32_t nranks, const std::vector<double*>& ptrs) {
    // Every CTA of the reduction kernel first stores its part into the peers' buffers and then waits for the
    // peers' CTA of the same index, so it is enough that CTAs start in the same order on every rank; with all
    // of them resident at once not even that is needed, which is what is required here.
    if (!P->fused) {
        int occ = 0;
        if (P->acc_size == 8) CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_reduce_exchange<double>, 256, 0));
        else CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k3_reduce_exchange<float>, 256, 0));
        if ((int64_t)P->nb > (int64_t)occ * P->sm_count)
            return fail(VMAT_E_UNSUPPORTED, "vmat_comm_attach: more control points than reduction CTAs that can be resident at once");
    }
    if (upload_vec(P->comm_ptrs, ptrs, P->stream)) return VMAT_E_CUDA;
    CU(cudaMemset(P->comm_buf.p, 0, P->comm_buf.bytes));
    CU(cudaMemset(P->comm_state.p, 0, 64));
    CU(cudaMemset(P->result.as<double>() + P->nb + 1, 0, 8));
    P->comm_rank = rank; P->comm_nranks = nranks; P->comm_epoch = 0;
    // the captured graph (if any) holds the single-GPU reduction
    if (P->graph_exec) { cudaGraphExecDestroy(P->graph_exec); P->graph_exec = nullptr; }
    if (P->graph) { cudaGraphDestroy(P->graph); P->graph = nullptr; }
    CU(cudaDeviceSynchronize());
    return VMAT_OK;
}

extern "C" int vmat_comm_attach(vmat_plan_t* P, int32_t rank, int32_t nranks, const void* handles) {
    if (!P || !handles || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        return fail(VMAT_E_ARG, "vmat_comm_attach: bad argument (at most 16 ranks)");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    if (int rc = comm_buffer(P)) return rc;
    if (P->comm_ipc) for (void* q : P->comm_peers) if (q) cudaIpcCloseMemHandle(q);
    P->comm_peers.assign(nranks, nullptr);
    P->comm_ipc = true;
    std::vector<double*> ptrs(nranks, nullptr);
    for (int q = 0; q < nranks; ++q) {
        if (q == rank) { ptrs[q] = P->comm_buf.as<double>(); continue; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)q * 64, 64);
        void* mapped = nullptr;
        CU(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
        P->comm_peers[q] = mapped;
        ptrs[q] = static_cast<double*>(mapped);
    }
    return comm_finish_attach(P, rank, nranks, ptrs);
}

extern "C" int vmat_comm_attach_local(vmat_plan_t* P, int32_t rank, int32_t nranks, vmat_plan_t* const* plans) {
    if (!P || !plans || nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || plans[rank] != P)
        return fail(VMAT_E_ARG, "vmat_comm_attach_local: bad argument (at most 16 ranks; plans[rank] must be the plan itself)");
    for (int q = 0; q < nranks; ++q)
        if (!plans[q] || plans[q]->B != P->B || plans[q]->nb != P->nb)
            return fail(VMAT_E_ARG, "vmat_comm_attach_local: all plans must share the beamlet layout");
    std::vector<double*> ptrs(nranks, nullptr);
    for (int q = 0; q < nranks; ++q) {
        CU(cudaSetDevice(plans[q]->device));
        if (int rc = comm_buffer(plans[q])) return rc;
        ptrs[q] = plans[q]->comm_buf.as<double>();
    }
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    for (int q = 0; q < nranks; ++q) {
        if (plans[q]->device == P->device) continue;
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, P->device, plans[q]->device));
        if (!can) return fail(VMAT_E_UNSUPPORTED, "vmat_comm_attach_local: no peer access between the plans' devices");
        const cudaError_t e = cudaDeviceEnablePeerAccess(plans[q]->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU(e);
        (void)cudaGetLastError();
    }
    if (P->comm_ipc) for (void* q : P->comm_peers) if (q) cudaIpcCloseMemHandle(q);
    P->comm_peers.clear();
    P->comm_ipc = false;
    // comm_finish_attach clears this plan's buffer only: a peer attaching later must not wipe flags this
    // plan has already received, so all plans have to be attached before any of them evaluates
    return comm_finish_attach(P, rank, nranks, ptrs);
}

extern "C" int vmat_comm_set_timeout(vmat_plan_t* P, double seconds) {
    if (!P || !(seconds > 0.0)) return fail(VMAT_E_ARG, "vmat_comm_set_timeout: need a positive time");
    P->comm_timeout_ns = (unsigned long long)(seconds * 1e9);
    return VMAT_OK;
}

extern "C" int vmat_intensities_buffer(vmat_plan_t* P, void** dev_ptr, int64_t* count) {
    if (!P || !dev_ptr || !count) return fail(VMAT_E_ARG, "null");
    *dev_ptr = P->y.p; *count = P->nb;
    return VMAT_OK;
}
extern "C" int vmat_result_buffer(vmat_plan_t* P, void** dev_ptr, int64_t* count) {
    if (!P || !dev_ptr || !count) return fail(VMAT_E_ARG, "null");
    *dev_ptr = P->result.p; *count = P->nb + 1;      // (one more word behind: the exchange status)
    return VMAT_OK;
}

template <typename AT>
static int fetch_as_f64(vmat_plan* P, const void* src, int64_t n, double* out) {
    k_cast_to_f64<AT><<<(unsigned)((n + 255) / 256), 256, 0, P->stream>>>(n, static_cast<const AT*>(src), P->scratch64.as<double>());
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, P->scratch64.p, n * 8, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    return 0;
}

extern "C" int vmat_get_dose(vmat_plan_t* P, double* out) {
    if (!P || !out) return fail(VMAT_E_ARG, "null");
    CU(cudaSetDevice(P->device)); CU(cudaDeviceSynchronize());
    DISPATCH_ACC(P->opt.acc_dtype, AT, { return fetch_as_f64<AT>(P, P->d.p, P->V, out); });
    return 0;
}
extern "C" int vmat_get_voxelgrad(vmat_plan_t* P, double* out) {
    if (!P || !out) return fail(VMAT_E_ARG, "null");
    CU(cudaSetDevice(P->device)); CU(cudaDeviceSynchronize());
    DISPATCH_ACC(P->opt.acc_dtype, AT, { return fetch_as_f64<AT>(P, P->vg.p, P->V, out); });
    return 0;
}
extern "C" int vmat_get_beamlet_grad(vmat_plan_t* P, double* out) {
    if (!P || !out) return fail(VMAT_E_ARG, "null");
    CU(cudaSetDevice(P->device)); CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, P->gbf.p, P->B * 8, cudaMemcpyDeviceToHost));
    return VMAT_OK;
}

// ---------------------------------------------------------------------------------
// pricing
// ---------------------------------------------------------------------------------
static int price_impl(vmat_plan_t* P, int32_t ncand, const int32_t* cand_beam,
                      const vmat_price_params_t* prm, const vmat_price_row_t* rows,
                      double* p, int32_t* l, int32_t* r, int32_t* len, int32_t out_stride) {
    if (!P || ncand < 0 || (ncand > 0 && (!cand_beam || !prm || !rows || !p || !l || !r)))
        return fail(VMAT_E_ARG, "vmat_price: null argument");
    if (ncand == 0) return VMAT_OK;
    if (!P->evaluated) return fail(VMAT_E_STATE, "vmat_price: no evaluation has produced a beamlet gradient yet");
    const int32_t M = prm->M, N = prm->N, variant = prm->variant;
    if (M <= 0 || N <= 0 || N > 62) return fail(VMAT_E_UNSUPPORTED, "vmat_price: need 1 <= N <= 62, M >= 1");
    if (variant < VMAT_PRICE_SAMPLING || variant > VMAT_PRICE_SHORT) return fail(VMAT_E_ARG, "vmat_price: unknown variant");
    for (int32_t c = 0; c < ncand; ++c)
        if (cand_beam[c] < 0 || cand_beam[c] >= P->nb) return fail(VMAT_E_ARG, "vmat_price: bad candidate");
    CU(cudaSetDevice(P->device));
    const int32_t kmax = price_max_nodes(N);
    const size_t nrow = (size_t)ncand * M;
    // the kernel sizes its node tables from N: every caller-supplied row descriptor must stay inside them
    for (size_t k = 0; k < nrow; ++k) {
        const vmat_price_row_t& R = rows[k];
        if (R.n_l < 1 || R.n_l > N + 1 || !(R.l_first >= -1.0) || !(R.l_first + (double)(R.n_l - 1) <= (double)N))
            return fail(VMAT_E_ARG, "vmat_price: a row has no left-leaf positions or leaves the grid (need 1 <= n_l <= N+1, -1 <= l <= N)");
        if (R.n_r_mid < 0 || R.n_r_mid > 2 || (variant == VMAT_PRICE_SAMPLING && R.n_r_mid < 1) || !(R.r_hi <= (double)N) ||
            R.vb_min < 0 || R.vb_min >= 64)
            return fail(VMAT_E_ARG, "vmat_price: bad row descriptor (need n_r_mid in {1,2}, r_hi <= N, 0 <= vb_min < 64)");
        const int32_t beam = cand_beam[k / M], nbl = P->beam_ptr[beam + 1] - P->beam_ptr[beam];
        if (variant != VMAT_PRICE_SHORT && R.vb_mask) {      // beamGrad entries the dose term may touch: [grad_offset - vb_min + lowest, ... + highest existing y]
            const int ylo = __builtin_ctzll(R.vb_mask), yhi = 63 - __builtin_clzll(R.vb_mask);
            if ((int64_t)R.grad_offset - R.vb_min + ylo < 0 || (int64_t)R.grad_offset - R.vb_min + yhi >= nbl)
                return fail(VMAT_E_ARG, "vmat_price: a row's beamlet gradient window leaves its control point");
        }
        if (variant == VMAT_PRICE_SHORT && (R.grad_offset < 0 || R.grad_offset > nbl))      // (reads past the end are clipped, like a Python slice)
            return fail(VMAT_E_ARG, "vmat_price: a row's beamlet gradient offset leaves its control point");
        int64_t total = 0;
        for (int32_t q = 0; q < R.n_l; ++q) {
            const double lval = R.l_first + (double)q;
            const double r0 = std::ceil(std::fmax(lval + 1.0, R.r_lo)), r1 = std::floor(R.r_hi);
            if (r1 >= r0) total += (int64_t)(r1 - r0) + 1;
            else if (variant == VMAT_PRICE_SAMPLING) total += R.n_r_mid;
            else if (variant == VMAT_PRICE_CLASSIC) total += std::max<int64_t>(0, (int64_t)(R.r_mid + 1.0 - r0));
        }
        if (total < 1 || total > kmax) return fail(VMAT_E_ARG, "vmat_price: a row has no nodes, or more than the network size for this N");
    }
    if ((size_t)ncand > P->price_cap_cand || nrow > P->price_cap_rows || nrow * kmax > P->price_cap_dad) {
        CU(P->price_cand.alloc(ncand * 4)); CU(P->price_p.alloc(ncand * 8));
        CU(P->price_rows.alloc(nrow * sizeof(vmat_price_row_t)));
        CU(P->price_l.alloc((nrow + ncand) * 4)); CU(P->price_r.alloc((nrow + ncand) * 4));
        CU(P->price_len.alloc(ncand * 4)); CU(P->price_gstart.alloc((nrow + ncand) * 4));
        CU(P->price_dad.alloc(nrow * kmax * sizeof(int32_t) * 3));
        P->price_cap_cand = ncand; P->price_cap_rows = nrow; P->price_cap_dad = nrow * kmax;
    }
    cudaStream_t st = P->stream;
    CU(cudaMemcpyAsync(P->price_cand.p, cand_beam, ncand * 4, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(P->price_rows.p, rows, nrow * sizeof(vmat_price_row_t), cudaMemcpyHostToDevice, st));
    PriceArgs a{};
    a.gb = P->gbf.as<double>(); a.beam_ptr = P->beam_ptr_d.as<int32_t>(); a.cand = P->price_cand.as<int32_t>();
    a.rows = P->price_rows.as<vmat_price_row_t>(); a.C = prm->C; a.C2 = prm->C2; a.C3 = prm->C3; a.b = prm->b;
    a.M = M; a.N = N; a.kmax = kmax; a.variant = variant; a.out_stride = out_stride;
    a.sink_last_only = (variant == VMAT_PRICE_CLASSIC && prm->reserved[0] == 1) ? 1 : 0;
    a.p = P->price_p.as<double>(); a.l = P->price_l.as<int32_t>();
    a.r = P->price_r.as<int32_t>(); a.len = P->price_len.as<int32_t>(); a.scratch = P->price_dad.as<int32_t>();
    a.gstart = P->price_gstart.as<int32_t>();
    const size_t smem = price_smem_bytes(kmax);
    if (smem > (size_t)P->smem_optin) return fail(VMAT_E_UNSUPPORTED, "vmat_price: N too large for shared memory");
    CU(cudaFuncSetAttribute(k5_price_dp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k5_price_dp<<<ncand, kPriceThreads, smem, st>>>(a);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(p, P->price_p.p, ncand * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(l, P->price_l.p, (size_t)ncand * out_stride * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(r, P->price_r.p, (size_t)ncand * out_stride * 4, cudaMemcpyDeviceToHost, st));
    if (len) CU(cudaMemcpyAsync(len, P->price_len.p, ncand * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return VMAT_OK;
}

extern "C" int vmat_price(vmat_plan_t* P, int32_t ncand, const int32_t* cand_beam,
                          const vmat_price_params_t* prm, const vmat_price_row_t* rows,
                          double* p, int32_t* l, int32_t* r) {
    if (prm && prm->variant != VMAT_PRICE_SAMPLING)
        return fail(VMAT_E_ARG, "vmat_price: the earlier generations return paths of varying length: use vmat_price_ex");
    return price_impl(P, ncand, cand_beam, prm, rows, p, l, r, nullptr, prm ? prm->M : 0);
}

extern "C" int vmat_price_ex(vmat_plan_t* P, int32_t ncand, const int32_t* cand_beam,
                             const vmat_price_params_t* prm, const vmat_price_row_t* rows,
                             double* p, int32_t* l, int32_t* r, int32_t* len) {
    if (!len) return fail(VMAT_E_ARG, "vmat_price_ex: null argument");
    return price_impl(P, ncand, cand_beam, prm, rows, p, l, r, len, prm ? prm->M + 1 : 0);
}

// ---------------------------------------------------------------------------------
// restricted-master-problem column cache
// ---------------------------------------------------------------------------------
template <typename AT>
static int rmp_begin_t(vmat_plan* P, int32_t ncols, const int32_t* beams) {
    cudaStream_t st = P->stream;
    const int64_t ld = (P->V + 63) / 64 * 64;
    if (P->ap_epoch.empty()) P->ap_epoch.assign(P->nb, 0);
    if (P->col_cap == 0) {
        // column store: room for every control point if that fits in a quarter of the free memory
        size_t free_b = 0, total_b = 0;
        CU(cudaMemGetInfo(&free_b, &total_b));
        const size_t per = 2 * (size_t)ld * sizeof(AT);
        const int32_t cap = (int32_t)std::min<size_t>((size_t)P->nb, free_b / 4 / per);
        if (cap < 1) return fail(VMAT_E_NOMEM, "vmat_rmp_begin: no memory for the column cache");
        CU(P->colA.alloc((size_t)cap * ld * sizeof(AT)));
        CU(P->colG.alloc((size_t)cap * ld * sizeof(AT)));
        CU(P->col_y.alloc(P->nb * 8));
        CU(P->col_w.alloc(P->B * sizeof(AT)));
        P->col_cap = cap; P->col_used = 0; P->rmp_ld = ld;
        P->col_slot.assign(P->nb, -1); P->col_built.assign(P->nb, -1);
        const int64_t nblk = (P->V + 255) / 256;
        CU(P->rmp_fpart.alloc(nblk * 8 * 8));    CU(cudaMemsetAsync(P->rmp_fpart.p, 0, nblk * 8 * 8, st));
    }
    // (re)build only the columns whose aperture changed since they were cached
    bool cast_done = false;
    const double one = 1.0;
    for (int32_t k = 0; k < ncols; ++k) {
        const int32_t bm = beams[k];
        if (P->col_slot[bm] >= 0 && P->col_built[bm] == P->ap_epoch[bm]) continue;
        if (P->col_slot[bm] < 0) {
            if (P->col_used >= P->col_cap) return fail(VMAT_E_NOMEM, "vmat_rmp_begin: column cache is full");
            P->col_slot[bm] = P->col_used++;
        }
        if (!cast_done) {
            k_cast_from_f64<AT><<<(unsigned)((P->B + 255) / 256), 256, 0, st>>>(P->B, P->w_grad.as<double>(), P->col_w.as<AT>());
            CU(cudaGetLastError());
            cast_done = true;
        }
        CU(cudaMemsetAsync(P->col_y.p, 0, P->nb * 8, st));
        CU(cudaMemcpyAsync(P->col_y.as<double>() + bm, &one, 8, cudaMemcpyHostToDevice, st));
        for (int which = 0; which < 2; ++which) {
            P->col_mode = true;
            P->col_w_ptr = which == 0 ? P->w_dose.p : P->col_w.p;
            CU(cudaMemsetAsync(P->counter.p, 0, 4, st));
            const int rc = launch_eval(P, st, 0, false);
            P->col_mode = false;
            if (rc) return rc;
            AT* dst = (which == 0 ? P->colA.as<AT>() : P->colG.as<AT>()) + (size_t)P->col_slot[bm] * ld;
            CU(cudaMemcpyAsync(dst, P->d.p, P->V * sizeof(AT), cudaMemcpyDeviceToDevice, st));
        }
        P->col_built[bm] = P->ap_epoch[bm];
    }
    CU(cudaMemsetAsync(P->counter.p, 0, 4, st));
    const int32_t chunk = 16384;
    const int32_t nchunks = (int32_t)((P->V + chunk - 1) / chunk);
    CU(P->rmp_yk.alloc(ncols * 8));
    CU(P->rmp_gpart.alloc((size_t)ncols * nchunks * 8));
    std::vector<int32_t> hb(beams, beams + ncols), hs(ncols);
    for (int32_t k = 0; k < ncols; ++k) hs[k] = P->col_slot[beams[k]];
    if (upload_vec(P->rmp_beams, hb, st)) return VMAT_E_CUDA;
    if (upload_vec(P->rmp_slots, hs, st)) return VMAT_E_CUDA;
    CU(cudaMemsetAsync(P->result.p, 0, (P->nb + 1) * 8, st));
    CU(cudaMemsetAsync(P->counter.as<unsigned int>() + 12, 0, 4, st));
    CU(cudaStreamSynchronize(st));
    P->rmp_cols = hb;
    return 0;
}

extern "C" int vmat_rmp_begin(vmat_plan_t* P, int32_t ncols, const int32_t* beams) {
    if (!P || ncols < 1 || ncols > 512 || !beams) return fail(VMAT_E_ARG, "vmat_rmp_begin: need 1..512 control points");
    for (int32_t k = 0; k < ncols; ++k)
        if (beams[k] < 0 || beams[k] >= P->nb) return fail(VMAT_E_ARG, "vmat_rmp_begin: bad control point");
    if (P->comm_nranks > 1) return fail(VMAT_E_UNSUPPORTED, "vmat_rmp_begin: not available on the sharded path");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    DISPATCH_ACC(P->opt.acc_dtype, AT, { return rmp_begin_t<AT>(P, ncols, beams); });
    return 0;
}

template <typename AT>
static int rmp_eval_t(vmat_plan* P) {
    cudaStream_t st = P->stream;
    const int32_t ncols = (int32_t)P->rmp_cols.size();
    RmpArgs<AT> a{};
    a.V = P->V; a.ld = P->rmp_ld; a.ncols = ncols; a.chunk = 16384; a.nchunks = (int32_t)((P->V + a.chunk - 1) / a.chunk);
    a.A = P->colA.as<AT>(); a.G = P->colG.as<AT>(); a.yk = P->rmp_yk.as<double>(); a.beams = P->rmp_beams.as<int32_t>(); a.slots = P->rmp_slots.as<int32_t>();
    a.thr = P->thr_n.as<AT>(); a.over = P->over_n.as<AT>(); a.under = P->under_n.as<AT>();
    a.d = P->d.as<AT>(); a.vg = P->vg.as<AT>();
    const int64_t nblk = (P->V + 255) / 256;
    a.fpart = P->rmp_fpart.as<double>(); a.nfpart = (int32_t)(nblk * 8);
    a.gpart = P->rmp_gpart.as<double>(); a.result = P->result.as<double>();
    a.done = P->counter.as<unsigned int>() + 12;
    k8_rmp_dose<AT><<<(unsigned)nblk, 256, 0, st>>>(a);
    k8_rmp_grad<AT><<<dim3((unsigned)a.nchunks, (unsigned)ncols), 256, 0, st>>>(a);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vmat_rmp_eval(vmat_plan_t* P, const double* y, double* f, double* g) {
    if (!P || !y) return fail(VMAT_E_ARG, "vmat_rmp_eval: null");
    if (P->rmp_cols.empty()) return fail(VMAT_E_STATE, "vmat_rmp_eval: no column cache (call vmat_rmp_begin)");
    CU(cudaSetDevice(P->device));
    const int32_t ncols = (int32_t)P->rmp_cols.size();
    std::memcpy(P->h_y, y, P->nb * 8);
    std::vector<double> yk(ncols);
    for (int32_t k = 0; k < ncols; ++k) yk[k] = y[P->rmp_cols[k]];
    std::memcpy(P->h_res, yk.data(), std::min<size_t>(ncols, P->nb + 1) * 8);       // pinned staging
    if ((size_t)ncols <= (size_t)P->nb + 1) CU(cudaMemcpyAsync(P->rmp_yk.p, P->h_res, ncols * 8, cudaMemcpyHostToDevice, P->stream));
    else { CU(cudaMemcpyAsync(P->rmp_yk.p, yk.data(), ncols * 8, cudaMemcpyHostToDevice, P->stream)); CU(cudaStreamSynchronize(P->stream)); }
    CU(cudaMemcpyAsync(P->y.p, P->h_y, P->nb * 8, cudaMemcpyHostToDevice, P->stream));
    DISPATCH_ACC(P->opt.acc_dtype, AT, { if (int rc = rmp_eval_t<AT>(P)) return rc; });
    CU(cudaMemcpyAsync(P->h_res, P->result.p, (P->nb + 1) * 8, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    if (f) *f = P->h_res[0];
    if (g) std::memcpy(g, P->h_res + 1, P->nb * 8);
    return VMAT_OK;
}

extern "C" int vmat_rmp_end(vmat_plan_t* P) {
    if (!P) return fail(VMAT_E_ARG, "null plan");
    CU(cudaSetDevice(P->device));
    CU(cudaDeviceSynchronize());
    P->rmp_cols.clear();            // the columns stay cached: the next RMP only builds what changed
    return VMAT_OK;
}

// ---------------------------------------------------------------------------------
// dose-volume histograms
// ---------------------------------------------------------------------------------
extern "C" int vmat_dose_max(vmat_plan_t* P, double* out) {
    if (!P || !out) return fail(VMAT_E_ARG, "null");
    CU(cudaSetDevice(P->device)); CU(cudaDeviceSynchronize());
    DevBuf m;
    CU(m.alloc(8)); CU(cudaMemsetAsync(m.p, 0, 8, P->stream));
    DISPATCH_ACC(P->opt.acc_dtype, AT, {
        k7_dose_max<AT><<<P->sm_count * 4, 256, 0, P->stream>>>(P->V, P->d.as<AT>(), m.as<unsigned long long>());
    });
    CU(cudaGetLastError());
    unsigned long long bits = 0;
    CU(cudaMemcpyAsync(&bits, m.p, 8, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    std::memcpy(out, &bits, 8);
    return VMAT_OK;
}

extern "C" int vmat_dvh_histogram(vmat_plan_t* P, const uint64_t* full_mask, int32_t nstructs,
                                  const double* edges, int32_t nedges, int64_t* hist) {
    if (!P || !full_mask || !edges || !hist || nstructs < 1 || nstructs > 64 || nedges < 2)
        return fail(VMAT_E_ARG, "vmat_dvh_histogram: bad argument");
    CU(cudaSetDevice(P->device)); CU(cudaDeviceSynchronize());
    DevBuf dm, de, dh;
    const size_t nh = (size_t)nstructs * (nedges - 1);
    CU(dm.alloc(P->V * 8)); CU(de.alloc((size_t)nedges * 8)); CU(dh.alloc(nh * 8));
    CU(cudaMemcpyAsync(dm.p, full_mask, P->V * 8, cudaMemcpyHostToDevice, P->stream));
    CU(cudaMemcpyAsync(de.p, edges, (size_t)nedges * 8, cudaMemcpyHostToDevice, P->stream));
    CU(cudaMemsetAsync(dh.p, 0, nh * 8, P->stream));
    DISPATCH_ACC(P->opt.acc_dtype, AT, {
        k7_dvh_hist<AT><<<P->sm_count * 4, 256, 0, P->stream>>>(P->V, P->d.as<AT>(), dm.as<unsigned long long>(), nstructs,
                                                                 de.as<double>(), nedges, dh.as<unsigned long long>());
    });
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(hist, dh.p, nh * 8, cudaMemcpyDeviceToHost, P->stream));
    CU(cudaStreamSynchronize(P->stream));
    return VMAT_OK;
}

// ---------------------------------------------------------------------------------
// k-medoid insertion costs
// ---------------------------------------------------------------------------------
extern "C" int vmat_kmedoid_costs_shard(int device, int64_t nc, const double* cand, int64_t n, int32_t dim,
                                        const double* points, const double* curmin, double* cost) {
    if (nc <= 0 || n < 0 || dim <= 0 || dim > 8 || !cand || !cost || (n > 0 && (!points || !curmin)))
        return fail(VMAT_E_ARG, "vmat_kmedoid_costs: bad argument");
    CU(cudaSetDevice(device));
    DevBuf dc, dp, dm, dout;
    CU(dc.alloc(nc * dim * 8)); CU(dp.alloc(n * dim * 8)); CU(dm.alloc(n * 8)); CU(dout.alloc(nc * 8));
    CU(cudaMemcpy(dc.p, cand, nc * dim * 8, cudaMemcpyHostToDevice));
    if (n > 0) {
        CU(cudaMemcpy(dp.p, points, n * dim * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(dm.p, curmin, n * 8, cudaMemcpyHostToDevice));
    }
    if (n <= kKmSequentialMax) {
        k6_kmedoid_cost<<<(unsigned)nc, kKmThreads>>>(n, dim, dc.as<double>(), dp.as<double>(), dm.as<double>(), dout.as<double>());
    } else {
        const unsigned grid = (unsigned)((nc + kKmTile - 1) / kKmTile);
#define VMAT_KM_CASE(D) case D: k6_kmedoid_cost_tiled<D><<<grid, kKmThreads>>>(n, nc, dc.as<double>(), dp.as<double>(), dm.as<double>(), dout.as<double>()); break;
        switch (dim) { VMAT_KM_CASE(1) VMAT_KM_CASE(2) VMAT_KM_CASE(3) VMAT_KM_CASE(4) VMAT_KM_CASE(5) VMAT_KM_CASE(6) VMAT_KM_CASE(7) VMAT_KM_CASE(8) }
#undef VMAT_KM_CASE
    }
    CU(cudaGetLastError());
    CU(cudaMemcpy(cost, dout.p, nc * 8, cudaMemcpyDeviceToHost));
    return VMAT_OK;
}

// ---- resident session: points, candidates and curmin stay on the device between insertion steps ----------------
struct vmat_kmedoid {
    int device = 0;
    int64_t n = 0, nc = 0;
    int32_t dim = 0;
    DevBuf points, cand, curmin, cost, medoid;
};

static int km_launch_costs(int64_t n, int64_t nc, int32_t dim, const double* dc, const double* dp, const double* dm, double* dout) {
    if (n <= kKmSequentialMax) {
        k6_kmedoid_cost<<<(unsigned)nc, kKmThreads>>>(n, dim, dc, dp, dm, dout);
    } else {
        const unsigned grid = (unsigned)((nc + kKmTile - 1) / kKmTile);
#define VMAT_KM_CASE(D) case D: k6_kmedoid_cost_tiled<D><<<grid, kKmThreads>>>(n, nc, dc, dp, dm, dout); break;
        switch (dim) { VMAT_KM_CASE(1) VMAT_KM_CASE(2) VMAT_KM_CASE(3) VMAT_KM_CASE(4) VMAT_KM_CASE(5) VMAT_KM_CASE(6) VMAT_KM_CASE(7) VMAT_KM_CASE(8) }
#undef VMAT_KM_CASE
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int vmat_kmedoid_open(vmat_kmedoid_t** out, int device, int64_t n, int32_t dim, const double* points,
                                 const double* curmin, int64_t n_cand, const double* cand) {
    if (!out || n < 0 || n_cand <= 0 || dim <= 0 || dim > 8 || !cand || (n > 0 && (!points || !curmin)))
        return fail(VMAT_E_ARG, "vmat_kmedoid_open: bad argument");
    CU(cudaSetDevice(device));
    vmat_kmedoid* K = new (std::nothrow) vmat_kmedoid();
    if (!K) return fail(VMAT_E_NOMEM, "out of host memory");
    struct Guard { vmat_kmedoid* p; ~Guard() { delete p; } } guard{K};
    K->device = device; K->n = n; K->nc = n_cand; K->dim = dim;
    CU(K->points.alloc((size_t)n * dim * 8)); CU(K->curmin.alloc((size_t)n * 8));
    CU(K->cand.alloc((size_t)n_cand * dim * 8)); CU(K->cost.alloc((size_t)n_cand * 8)); CU(K->medoid.alloc(8 * 8));
    if (n > 0) {
        CU(cudaMemcpy(K->points.p, points, (size_t)n * dim * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(K->curmin.p, curmin, (size_t)n * 8, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(K->cand.p, cand, (size_t)n_cand * dim * 8, cudaMemcpyHostToDevice));
    guard.p = nullptr;
    *out = K;
    return VMAT_OK;
}

extern "C" int vmat_kmedoid_insert(vmat_kmedoid_t* K, const double* medoid) {
    if (!K || !medoid) return fail(VMAT_E_ARG, "vmat_kmedoid_insert: null");
    CU(cudaSetDevice(K->device));
    CU(cudaMemcpy(K->medoid.p, medoid, (size_t)K->dim * 8, cudaMemcpyHostToDevice));
    if (K->n > 0) {
        k6_update_curmin<<<(unsigned)((K->n + 255) / 256), 256>>>(K->n, K->dim, K->points.as<double>(), K->medoid.as<double>(), K->curmin.as<double>());
        CU(cudaGetLastError());
    }
    return VMAT_OK;
}

extern "C" int vmat_kmedoid_costs_resident(vmat_kmedoid_t* K, double* cost_host, void** cost_dev) {
    if (!K) return fail(VMAT_E_ARG, "vmat_kmedoid_costs_resident: null");
    CU(cudaSetDevice(K->device));
    if (K->n > 0) { if (int rc = km_launch_costs(K->n, K->nc, K->dim, K->cand.as<double>(), K->points.as<double>(), K->curmin.as<double>(), K->cost.as<double>())) return rc; }
    else CU(cudaMemset(K->cost.p, 0, (size_t)K->nc * 8));
    if (cost_host) CU(cudaMemcpy(cost_host, K->cost.p, (size_t)K->nc * 8, cudaMemcpyDeviceToHost));
    else CU(cudaDeviceSynchronize());
    if (cost_dev) *cost_dev = K->cost.p;
    return VMAT_OK;
}

extern "C" int vmat_kmedoid_close(vmat_kmedoid_t* K) {
    if (!K) return VMAT_OK;
    cudaSetDevice(K->device);
    cudaDeviceSynchronize();
    delete K;
    return VMAT_OK;
}

extern "C" int vmat_kmedoid_costs(int device, int64_t n, int32_t dim, const double* points,
                                  const double* curmin, double* cost) {
    return vmat_kmedoid_costs_shard(device, n, points, n, dim, points, curmin, cost);
}
