// KF: the whole evaluation in ONE persistent kernel (sm_100a).
//
//   x = y * w_dose  ->  d = D x, penalty, voxel gradient  ->  gb = D^T vg  ->  aperture gradient
//   (calcObjGrad, greedyVMATsampling.py:578-582 = calcDose :243-251 + calcGradientandObjValue
//   :254-272, plus the beamlet gradient PPsubroutine :646 needs), optionally with the cross-GPU
//   exchange of the beamlet gradient.
//
// One CTA per SM: 16 consumer warps + 1 producer warp, as in the dose kernel K1.  The work of both
// sparse products is ONE queue of super-slices ("items"): first the dose items (voxel-major stream),
// then the transposed items (beamlet-major, column-tiled stream).  CTAs claim items from an atomic
// counter, so there is a single tail at the very end instead of one per kernel, no launch/drain gap
// between the two products, and a CTA that runs out of dose items simply continues with transposed
// ones.  What orders the two products is data, not a kernel boundary:
//   * the voxels are cut into GROUPS of consecutive tiles; the dose stream is sorted by row length
//     inside each group, and every consumer warp bumps grp_done[group] when its slice of a dose item
//     has stored d and vg;
//   * a transposed item needs the vg of one tile; its producer waits until grp_done[group(tile)] has
//     reached the number of dose slices of that group, then fetches the tile with one TMA bulk copy.
//     Transposed items come after all dose items in the queue, so the wait is normally already
//     satisfied when the item is claimed (it bites only on small slabs, where CTAs start on
//     transposed items right away and prefetch their D stream while the dose items finish).
// The region of shared memory that holds x during the dose items is reused for two vg tile buffers
// once the CTA has seen its first transposed item (the queue order guarantees it never gets a dose
// item again), so the next item's tile is fetched while the current item is still being consumed.
//
// Super-slices are RAGGED: the 16 slices of an item keep their own chunk counts nch[w] (non-increasing
// in w, the rows are sorted by length); stage c holds chunk c of the warps 0..k_c-1 that still have
// one.  A 16-warp CTA therefore streams no more padding than a 1-warp layout would.
//
// When the queue is empty the CTAs meet at a grid barrier (arrive counter + generation word), then
// CTA i reduces the tile partials of control points i, i+G, ... (fixed order), forms the aperture
// gradient and - multi-GPU - runs the push exchange of vmat_eval_kernels.cuh for its control points.
// The objective is summed in two fixed-shape levels (per-CTA shares of the per-slice partials, then
// CTA 0).  No floating-point atomics; results are bit-reproducible although the item->CTA assignment
// is dynamic (which CTA computes a row changes, how it is computed does not).
#pragma once
#include "vmat_eval_kernels.cuh"

namespace vmat {

constexpr int kFW = 16;                  // consumer warps per CTA
constexpr int kFWarps = kFW + 2;         // + producer warp + signal warp
constexpr int kFThreads = kFWarps * 32;
constexpr int kFMaxStages = 16;
constexpr int kFTraceItems = 32;         // items per CTA covered by the per-item diagnostics (vmat_trace_eval)
constexpr int kFSyncWords = 16;          // sync[0] queue, [1] barrier arrivals, [2] barrier generation, [3] objective shares done; groups follow

struct __align__(16) FItem {             // one queue item (32 bytes)
    uint8_t nch[kFW];                    // chunks of warp w, non-increasing in w
    int32_t tile;                        // transposed item: voxel tile; dose item: -1
    int32_t grp;                         // dose item: group to signal; transposed item: group to wait for
    int32_t need;                        // transposed item: value grp_done[grp] must have reached
    uint8_t tpr, pad[3];
};
static_assert(sizeof(FItem) == 32, "FItem layout");

struct __align__(16) FStage {            // what the producer tells the consumers about a ring slot
    int32_t item;                        // -1: no more work
    int32_t tile;                        // -1 for dose items
    int32_t grp;
    uint8_t flags, kact, tpr, tbuf;      // kact: warps 0..kact-1 have a chunk in this stage; tbuf: tile buffer and its phase parity (bit 1)
    uint8_t nch[kFW];
};
static_assert(sizeof(FStage) == 32, "FStage layout");

template <typename AT>
struct FArgs {
    const uint4* sell;                   // dose stream (voxel-major), items 0 .. n1-1
    const uint4* sell2;                  // transposed stream (beamlet-major, column-tiled), items n1 .. nitems-1
    const int64_t* it_off;               // [n1+1] offsets into sell, 16-byte units
    const int64_t* it_off2;              // [nitems-n1+1] offsets into sell2
    const FItem* items;                  // [nitems]
    int32_t n1, nitems;                  // dose items, all items (dose-only launches pass nitems = n1)
    int32_t nstages, stage_bytes;
    int32_t B, nb, T, ntiles, ngroups;
    int64_t V;
    const double* y;                     // [nb]
    const int32_t* beam_of;              // [B]
    const AT* w_dose;                    // [B]
    AT* d;                               // [V]
    AT* vg;                              // [V + T]
    double* fpart;                       // [n1 * 16] objective partial of every dose slice
    AT* partial;                         // [ntiles][B]
    unsigned int* sync;                  // [kFSyncWords + ngroups]
    double* fblock;                      // [gridDim.x]
    const int32_t* beam_ptr;             // [nb+1]
    const double* w_grad;                // [B]
    double* gbf;                         // [B+1]
    double* result;                      // [nb+2]
    double* y_mirror;                    // host-buffer path: y is mirrored into the device copy by CTA 0
    int32_t y_inline_n;
    int32_t claim_lead;
    uint32_t zero;                // 0 (mbar_arrive_after)
    int32_t dose_only;                   // 1: dose items only, no reductions (RMP column build)
    unsigned long long* trace;           // [gridDim.x * kTraceSlots] or nullptr
    double y_inline[kInlineY];
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// one bounded hardware wait (the thread sleeps until the phase completes or a time limit expires)
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int atom_add_release_gpu(unsigned int* p, unsigned int v) {
    unsigned int old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ double ld_cg_f64(const double* p) { return __ldcg(p); }
__device__ __forceinline__ float ld_cg_f64(const float* p) { return __ldcg(p); }

// sum of src[first], src[first+stride], ... (< n) in that order, loads bypassing L1 (the data was
// written by other SMs in this same kernel), batched 8 deep
template <typename T>
__device__ __forceinline__ double strided_sum_cg(const T* __restrict__ src, int64_t first, int64_t stride, int64_t n) {
    double s = 0.0;
    int64_t k = first;
    for (; k + 7 * stride < n; k += 8 * stride) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + k + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k + u * stride < n) ? __ldcg(src + k + u * stride) : (T)0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (k + u * stride < n) s += (double)v[u];
    }
    return s;
}

// fixed-shape sum over the kFThreads threads of the CTA: warp butterflies, then the 18 warp sums in order
__device__ __forceinline__ double block_sum_f(double v, double* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < kFWarps; ++w) r += sh[w];
    return r;
}

template <typename VT, typename AT, int IW>
__global__ void __launch_bounds__(kFThreads) kf_eval(const __grid_constant__ FArgs<AT> a, const CommArgs c) {
    constexpr int W = kFW, KU = chunk_units<VT, IW>(), HU1 = header_units_k1<AT>(IW), HU2 = header_units_k2(IW);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kFMaxStages], empty[kFMaxStages], tile_full[2], tile_empty[2], x_free;
    __shared__ __align__(16) FStage sinfo[kFMaxStages];
    __shared__ double red[kFWarps];
    // dose-slice announcements (see the signal warp): one barrier per CTA-local dose item, modulo 16
    __shared__ __align__(8) uint64_t sbar[16];
    __shared__ int scmd[16];
    const int xbytes = (a.B * (int)sizeof(AT) + 127) & ~127;
    const int ybytes = (a.nb * 8 + 127) & ~127;
    const int tbytes = (a.T * (int)sizeof(AT) + 127) & ~127;
    const int ubytes = max(xbytes + ybytes, 2 * tbytes);
    AT* xs = reinterpret_cast<AT*>(smem_raw);
    double* ys = reinterpret_cast<double*>(smem_raw + xbytes);
    unsigned char* ring = smem_raw + ubytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nst = a.nstages, SB = a.stage_bytes;
    const unsigned int G = gridDim.x;
    unsigned long long* const trace = a.trace ? a.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
    // diagnostics: per CTA, 8 stamps for each of its first kFTraceItems items (after the per-CTA block):
    // producer [0] item known, [1] header pushed, [2] all stages pushed, [3] next item claimed;
    // consumer warp 0 [4] header seen, [5] header done (tile there), [6] last chunk done, [7] epilogue done
    unsigned long long* const itrace = a.trace ? a.trace + (size_t)gridDim.x * kTraceSlots + (size_t)blockIdx.x * kFTraceItems * 8 : nullptr;
    unsigned int* const grp_done = a.sync + kFSyncWords;

    if (tid == 0) {
        if (trace) trace[0] = global_timer_ns();
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], W); }
        mbar_init(&tile_full[0], 1); mbar_init(&tile_full[1], 1);
        mbar_init(&tile_empty[0], W); mbar_init(&tile_empty[1], W);
        mbar_init(&x_free, W);
        for (int k = 0; k < 16; ++k) mbar_init(&sbar[k], W);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();
    if (warp < W) {
        // x_j = y[beam(j)] * w_dose[j] into shared memory (as in K1; the producer streams D meanwhile)
        constexpr int kPre = 4;
        const int ngroups4 = a.B >> 2;
        int4 bo[kPre];
        AT wd[kPre][4];
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int gidx = tid + k * W * 32;
            if (gidx < ngroups4) {
                bo[k] = reinterpret_cast<const int4*>(a.beam_of)[gidx];
                load4(a.w_dose + 4 * gidx, wd[k]);
            }
        }
        if (a.y_inline_n > 0) { for (int j = tid; j < a.nb; j += W * 32) ys[j] = a.y_inline[j]; }
        else { for (int j = tid; j < a.nb; j += W * 32) ys[j] = a.y[j]; }
        named_bar_sync(1, W * 32);
#pragma unroll
        for (int k = 0; k < kPre; ++k) {
            const int gidx = tid + k * W * 32;
            if (gidx < ngroups4) {
                store4(xs + 4 * gidx, mul_rn((AT)ys[bo[k].x], wd[k][0]), mul_rn((AT)ys[bo[k].y], wd[k][1]),
                       mul_rn((AT)ys[bo[k].z], wd[k][2]), mul_rn((AT)ys[bo[k].w], wd[k][3]));
            }
        }
        for (int gidx = tid + kPre * W * 32; gidx < ngroups4; gidx += W * 32) {
            const int4 b4 = reinterpret_cast<const int4*>(a.beam_of)[gidx];
            AT w4[4];
            load4(a.w_dose + 4 * gidx, w4);
            store4(xs + 4 * gidx, mul_rn((AT)ys[b4.x], w4[0]), mul_rn((AT)ys[b4.y], w4[1]),
                   mul_rn((AT)ys[b4.z], w4[2]), mul_rn((AT)ys[b4.w], w4[3]));
        }
        for (int j = 4 * ngroups4 + tid; j < a.B; j += W * 32) xs[j] = mul_rn((AT)ys[a.beam_of[j]], a.w_dose[j]);
        named_bar_sync(1, W * 32);
        // d, vg, fpart, partial and the sync words belong to the previous evaluation until it is complete
        pdl_wait();
        if (blockIdx.x == 0 && a.y_mirror != nullptr) { for (int j = tid; j < a.nb; j += W * 32) a.y_mirror[j] = ys[j]; }
    }
    if (trace && tid == 0) trace[1] = global_timer_ns();

    if (warp == W) {
        // ---------------- producer warp: streams items into the ring ---------------------------------
        // All lanes walk the item loop together; the stages of an item are pushed in batches of NB, one
        // lane per stage, so the per-stage bookkeeping (active warps, stream offset, ring slot) is done
        // once per batch in SIMD instead of once per stage by a single thread - a one-thread producer
        // needed ~0.5 us per stage and starved the consumers (measured: 102 us per evaluation on config 2
        // against 65 us for the three kernels, the consumers waiting on `full` a third of the time).  A lane
        // issues its copy as soon as its own slot is free; slots free in ring order, so lanes finish in
        // order and a batch never holds back a copy that could have been issued.  Lane 0 takes the LAST
        // stage of a batch: it waits the longest and is the one that polls a pending tile fetch meanwhile.
        {
            const uint64_t pol = l2_evict_first_policy();
            const int NB = min(8, nst);
            const unsigned int unst = (unsigned)nst;
            unsigned int abs_pos = 0;          // ring stages pushed so far (header and data stages)
            const unsigned int ni = (unsigned)a.nitems;
            const int lead = a.claim_lead;
            unsigned int cur = blockIdx.x;
            bool first = true;                 // the first item is assigned statically; nothing written by the previous evaluation has been touched yet
            int ntile_use = 0;                 // transposed items of this CTA so far
            // tile fetch in flight for the current transposed item (lane 0 issues it as soon as buffer and group allow)
            bool tile_pending = false, tile_buf_free = false, tile_grp_ready = false;
            int tb = 0; uint32_t tp = 0; int tgrp = 0; unsigned int tneed = 0; int ttile = 0;
            unsigned long long ready_mask = 0ull;       // groups whose dose items were seen complete
            // Every barrier arrival is anonymous, so each consumer warp must arrive exactly once per phase:
            // once on x_free (it has read x for the last time), once on tile_empty[b] per item that used
            // buffer b - and it can reach that item's header only after the tile has landed, i.e. after the
            // phase of the buffer's previous item is complete.  The first use of a buffer waits for x_free.
            auto poll_tile = [&]() {
                if (!tile_buf_free) tile_buf_free = ntile_use <= 2 ? mbar_test(&x_free, 0u) : mbar_test(&tile_empty[tb], tp ^ 1u);
                if (tile_buf_free && !tile_grp_ready) {
                    if (tgrp < 64 && ((ready_mask >> tgrp) & 1ull)) tile_grp_ready = true;       // seen complete before
                    else if (ld_acquire_gpu_u32(grp_done + tgrp) >= tneed) { tile_grp_ready = true; if (tgrp < 64) ready_mask |= 1ull << tgrp; }
                }
                if (tile_buf_free && tile_grp_ready) {
                    fence_proxy_async_all();           // vg was written through the generic proxy by other SMs
                    const int64_t base = (int64_t)ttile * a.T;
                    const int n = (int)min((int64_t)a.T, a.V - base);
                    const uint32_t bytes = (uint32_t)(((size_t)n * sizeof(AT) + 15) & ~(size_t)15);
                    mbar_expect_tx(&tile_full[tb], bytes);
                    tma_bulk_g2s(smem_raw + (size_t)tb * tbytes, a.vg + base, bytes, &tile_full[tb]);
                    tile_pending = false;
                }
            };
            // push ring stage number `at` (slot at % nst, lap at / nst): wait for the slot, describe it, start the copy
            auto push_at = [&](unsigned int at, const uint4 w0, const uint4 w1, const void* src, uint32_t bytes, bool poller) {
                const unsigned int lap = at / unst, slot = at - lap * unst;
                for (;;) {
                    // try_wait sleeps in hardware and returns now and then, which is when a pending tile is polled
                    if (mbar_try_wait(&empty[slot], (lap & 1u) ^ 1u)) {
                        reinterpret_cast<uint4*>(&sinfo[slot])[0] = w0;
                        reinterpret_cast<uint4*>(&sinfo[slot])[1] = w1;
                        if (bytes) {
                            mbar_expect_tx(&full[slot], bytes);
                            tma_bulk_g2s_hint(ring + (size_t)slot * SB, src, bytes, &full[slot], pol);
                        } else {
                            mbar_arrive(&full[slot]);
                        }
                        break;
                    }
                    if (poller && tile_pending) poll_tile();
                }
            };
            // an item's metadata as two 16-byte loads: {nch[16]} and {tile, grp, need, tpr}
            uint4 nchw = make_uint4(0u, 0u, 0u, 0u);
            int4 meta = make_int4(-1, 0, 0, 1);
            const uint4* o0 = nullptr;
            const unsigned int n1u = (unsigned)a.n1;
            auto item_base = [&](unsigned int q) -> const uint4* {
                return q < n1u ? a.sell + a.it_off[q] : a.sell2 + a.it_off2[q - n1u];
            };
            auto load_item = [&](unsigned int q, uint4& n, int4& m) {
                const uint4* ip = reinterpret_cast<const uint4*>(a.items + q);
                n = __ldg(ip);
                const uint4 v1 = __ldg(ip + 1);
                m = make_int4((int)v1.x, (int)v1.y, (int)v1.z, (int)(v1.w & 0xffu));
            };
            if (cur < ni) { load_item(cur, nchw, meta); o0 = item_base(cur); }
            int pk = 0;                        // items of this CTA so far (diagnostics)
            while (cur < ni) {
                const int it_tile = meta.x, it_grp = meta.y, it_need = meta.z, it_tpr = meta.w;
                unsigned long long* const its = (itrace && pk < kFTraceItems && lane == 0) ? itrace + pk * 8 : nullptr;
                ++pk;
                if (its) its[0] = global_timer_ns();
                const bool dose = it_tile < 0;
                const int HU = dose ? HU1 : HU2;
                const int nch0 = (int)(nchw.x & 0xffu);
                if (!dose) {
                    if (first) { pdl_wait(); first = false; }      // grp_done is about to be read
                    tb = ntile_use & 1; tp = (uint32_t)(ntile_use >> 1) & 1u; ++ntile_use;
                    tgrp = it_grp; tneed = (unsigned)it_need; ttile = it_tile;
                    tile_pending = true; tile_buf_free = false; tile_grp_ready = false;
                }
                // FStage as two 16-byte words: {item, tile, grp, flags | kact << 8 | tpr << 16 | tbuf << 24}, {nch[16]}
                const uint32_t tail = ((uint32_t)it_tpr << 16) | ((uint32_t)(tb | (tp << 1)) << 24);
                if (lane == 0) {
                    push_at(abs_pos, make_uint4(cur, (uint32_t)it_tile, (uint32_t)it_grp,
                                                (uint32_t)(kStageHeader | (nch0 == 0 ? kStageLast : 0)) | ((uint32_t)W << 8) | tail),
                            nchw, o0, (uint32_t)(W * HU * 16), true);
                    if (tile_pending) poll_tile();
                    if (its) its[1] = global_timer_ns();
                }
                __syncwarp();
                const uint4* data = o0 + W * HU;
                unsigned int nxt = 0u;
                uint4 nnch = make_uint4(0u, 0u, 0u, 0u);
                int4 nmeta = make_int4(-1, 0, 0, 1);
                const uint4* n0 = nullptr;
                bool claimed = false;
                const int c_claim = max(0, nch0 - lead);
                for (int c0 = 0; c0 < nch0; c0 += NB) {
                    if (!first && !claimed && c0 + NB > c_claim) {      // claim the next item while the last stages go out
                        if (lane == 0) nxt = G + atomicAdd(a.sync, 1u);
                        nxt = __shfl_sync(0xffffffffu, nxt, 0);
                        claimed = true;
                        if (nxt < ni) { load_item(nxt, nnch, nmeta); n0 = item_base(nxt); }
                    }
                    const int nact = min(NB, nch0 - c0);
                    if (lane < nact) {
                        const int cc = c0 + (lane == 0 ? nact - 1 : lane - 1);
                        // warps with a chunk in stage cc, and where the stage starts: KU * sum_w min(nch_w, cc)
                        const uint32_t crep = (uint32_t)cc * 0x01010101u;
                        const int kact = (__popc(__vcmpgtu4(nchw.x, crep)) + __popc(__vcmpgtu4(nchw.y, crep)) +
                                          __popc(__vcmpgtu4(nchw.z, crep)) + __popc(__vcmpgtu4(nchw.w, crep))) >> 3;
                        const uint32_t before = __dp4a(__vminu4(nchw.x, crep), 0x01010101u, 0u) + __dp4a(__vminu4(nchw.y, crep), 0x01010101u, 0u) +
                                                __dp4a(__vminu4(nchw.z, crep), 0x01010101u, 0u) + __dp4a(__vminu4(nchw.w, crep), 0x01010101u, 0u);
                        push_at(abs_pos + 1u + (unsigned)cc,
                                make_uint4(cur, (uint32_t)it_tile, (uint32_t)it_grp, ((uint32_t)kact << 8) | tail), nchw,
                                data + (int64_t)before * KU, (uint32_t)(kact * KU * 16), lane == 0);
                    }
                    __syncwarp();
                    if (lane == 0 && tile_pending) poll_tile();
                }
                if (lane == 0) { while (tile_pending) poll_tile(); }      // (only when the item had fewer stages than the ring)
                __syncwarp();
                if (its) its[2] = global_timer_ns();
                abs_pos += 1u + (unsigned)nch0;
                if (first) { pdl_wait(); first = false; }    // the queue counter was reset by the previous evaluation
                if (!claimed) {
                    if (lane == 0) nxt = G + atomicAdd(a.sync, 1u);
                    nxt = __shfl_sync(0xffffffffu, nxt, 0);
                    if (nxt < ni) { load_item(nxt, nnch, nmeta); n0 = item_base(nxt); }
                }
                cur = nxt; nchw = nnch; meta = nmeta; o0 = n0;
                if (its) its[3] = global_timer_ns() | (unsigned long long)(o0 != nullptr);     // (uses the loaded values)
            }
            if (lane == 0) push_at(abs_pos, make_uint4(0xffffffffu, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u), nullptr, 0, false);
        }
    } else if (warp == W + 1) {
        // ---------------- signal warp: announces finished dose items to the other CTAs ----------------------
        // Barrier k (mod 16) completes when all 16 consumer warps have stored d / vg of this CTA's k-th dose
        // item (they cannot be 16 items apart: the ring has at most 16 slots and every item has a header
        // stage).  The barrier orders their stores before this thread at CTA scope; its gpu-scope fence then
        // orders them before the counter update for everybody (cumulativity), so a producer that sees
        // grp_done[g] complete may fetch the group's vg.
        if (lane == 0) {
            pdl_wait();
            for (int k = 0;; ++k) {
                mbar_wait(&sbar[k & 15], (uint32_t)(k >> 4) & 1u);
                const int g = scmd[k & 15];
                if (g < 0) break;
                __threadfence();
                atomicAdd(grp_done + g, (unsigned)W);
            }
        }
    } else {
        // ---------------- consumers: warp w owns slice w of every item -----------------------------
        RingPos pos;
        AT acc = (AT)0, t = (AT)0, ov = (AT)0, un = (AT)0;
        int32_t row = -1;
        uint32_t cur = 0;
        int my_nch = 0, cdone = 0, tpr = 1, item = 0, grp = 0, tile = -1, tbuf = 0;
        bool dose = true, x_released = false;
        // A finished dose slice is announced to the signal warp through a CTA-local barrier (cheap); the
        // gpu-scope release fence in front of the grp_done update costs about 1 us under load (measured), so
        // it is taken by the signal warp, off the consumers' path.  dk counts this CTA's dose items.
        int dk = 0;
        bool dose_over = false;
        int ck = -1;                           // items seen so far (diagnostics)
        unsigned long long* its = nullptr;
        const AT* vt = reinterpret_cast<const AT*>(smem_raw);
        if (trace) { mbar_wait(&full[0], 0); if (tid == 0) trace[2] = global_timer_ns(); }
        while (true) {
            mbar_wait(&full[pos.stage], pos.phase);
            const FStage* sp = &sinfo[pos.stage];
            const uint4 sw = *reinterpret_cast<const uint4*>(sp);
            const int s_item = (int)sw.x;
            if (s_item < 0) break;
            const int s_flags = (int)(sw.w & 0xffu), s_kact = (int)((sw.w >> 8) & 0xffu);
            bool finished = false;
            if (s_flags & kStageHeader) {
                ++ck;
                its = (itrace && ck < kFTraceItems && tid == 0) ? itrace + ck * 8 : nullptr;
                if (its) its[4] = global_timer_ns();
                item = s_item; tile = (int)sw.y; grp = (int)sw.z; tpr = (int)((sw.w >> 16) & 0xffu); my_nch = sp->nch[warp];
                dose = tile < 0;
                cdone = 0;
                acc = (AT)0;
                if (dose) {
                    const unsigned char* hp = ring + (size_t)pos.stage * SB + warp * (HU1 * 16);
                    row = reinterpret_cast<const int32_t*>(hp)[lane];
                    t  = reinterpret_cast<const AT*>(hp + 128)[lane];
                    ov = reinterpret_cast<const AT*>(hp + 128 + 32 * sizeof(AT))[lane];
                    un = reinterpret_cast<const AT*>(hp + 128 + 64 * sizeof(AT))[lane];
                } else {
                    const int tb2 = (int)(sw.w >> 24);
                    tbuf = tb2 & 1;
                    if (!x_released) {       // first transposed item: no dose item follows, the x region becomes the tile buffers
                        x_released = true;
                        if (tid == 0) scmd[dk & 15] = -1;              // no more dose items: the signal warp can go
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sbar[dk & 15]);
                        dose_over = true;
                        if (trace && tid == 0) trace[3] = global_timer_ns();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&x_free);
                    }
                    row = reinterpret_cast<const int32_t*>(ring + (size_t)pos.stage * SB + warp * (HU2 * 16))[lane];
                    vt = reinterpret_cast<const AT*>(smem_raw + (size_t)tbuf * tbytes);
                    mbar_wait(&tile_full[tbuf], (uint32_t)(tb2 >> 1));      // every warp, also one without chunks (see the producer)
                }
                __syncwarp();
                // the slot goes back only when the header words have arrived in registers (mbar_arrive_after)
                if (lane == 0) mbar_arrive_after(&empty[pos.stage], dose ? ((uint32_t)row | dep_bits(t) | dep_bits(ov) | dep_bits(un)) : (uint32_t)row, a.zero);
                if (its) its[5] = global_timer_ns();
                finished = my_nch == 0;
            } else if (warp < s_kact) {
                const int warp_off = warp * (KU * 16);
                ChunkRegs<AT> A;
                chunk_load<VT, AT, IW>(ring + (size_t)pos.stage * SB + warp_off, lane, A, cur);
                if (cdone + 1 < my_nch) {         // the next stage belongs to this item and has a chunk of this warp too
                    RingPos nx = pos;
                    nx.advance(nst);
                    mbar_wait(&full[nx.stage], nx.phase);
                    ChunkRegs<AT> Bc;
                    chunk_load<VT, AT, IW>(ring + (size_t)nx.stage * SB + warp_off, lane, Bc, cur);
                    AT x0, x1, x2, x3, y0, y1, y2, y3;
                    if (dose) {
                        x0 = xs[A.i[0]]; x1 = xs[A.i[1]]; x2 = xs[A.i[2]]; x3 = xs[A.i[3]];
                        y0 = xs[Bc.i[0]]; y1 = xs[Bc.i[1]]; y2 = xs[Bc.i[2]]; y3 = xs[Bc.i[3]];
                    } else {
                        x0 = vt[A.i[0]]; x1 = vt[A.i[1]]; x2 = vt[A.i[2]]; x3 = vt[A.i[3]];
                        y0 = vt[Bc.i[0]]; y1 = vt[Bc.i[1]]; y2 = vt[Bc.i[2]]; y3 = vt[Bc.i[3]];
                    }
                    __syncwarp();
                    if (lane == 0) {
                        const uint32_t dep = dep_bits(A.v[0]) | dep_bits(A.v[3]) | dep_bits(Bc.v[0]) | dep_bits(Bc.v[3]);
                        mbar_arrive_after(&empty[pos.stage], dep, a.zero); mbar_arrive_after(&empty[nx.stage], dep, a.zero);
                    }
                    acc = fma_acc(A.v[0], x0, acc); acc = fma_acc(A.v[1], x1, acc);
                    acc = fma_acc(A.v[2], x2, acc); acc = fma_acc(A.v[3], x3, acc);
                    acc = fma_acc(Bc.v[0], y0, acc); acc = fma_acc(Bc.v[1], y1, acc);
                    acc = fma_acc(Bc.v[2], y2, acc); acc = fma_acc(Bc.v[3], y3, acc);
                    pos = nx;
                    cdone += 2;
                } else {
                    AT x0, x1, x2, x3;
                    if (dose) { x0 = xs[A.i[0]]; x1 = xs[A.i[1]]; x2 = xs[A.i[2]]; x3 = xs[A.i[3]]; }
                    else      { x0 = vt[A.i[0]]; x1 = vt[A.i[1]]; x2 = vt[A.i[2]]; x3 = vt[A.i[3]]; }
                    __syncwarp();
                    if (lane == 0) mbar_arrive_after(&empty[pos.stage], dep_bits(A.v[0]) | dep_bits(A.v[3]), a.zero);
                    acc = fma_acc(A.v[0], x0, acc); acc = fma_acc(A.v[1], x1, acc);
                    acc = fma_acc(A.v[2], x2, acc); acc = fma_acc(A.v[3], x3, acc);
                    cdone += 1;
                }
                finished = cdone == my_nch;
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[pos.stage]);
            }
            if (finished) {
                if (its) its[6] = global_timer_ns();
                for (int o = 1; o < tpr; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // tpr lanes of a row
                if (dose) {
                    double fterm = 0.0;
                    if (row >= 0) {
                        // (d-t)+ and (t-d)+, objective term, vg = 2*(2*o*over - 2*u*under)  (:255-271)
                        AT o = sub_rn(acc, t);  o = o > (AT)0 ? o : (AT)0;
                        AT u = sub_rn(t, acc);  u = u > (AT)0 ? u : (AT)0;
                        const AT fo = mul_rn(mul_rn(o, o), ov);
                        const AT fu = mul_rn(mul_rn(u, u), un);
                        fterm = (double)add_rn(fo, fu);
                        const AT go = mul_rn(mul_rn((AT)2, o), ov);
                        const AT gu = mul_rn(mul_rn((AT)2, u), un);
                        a.d[row] = acc;
                        a.vg[row] = mul_rn((AT)2, sub_rn(go, gu));
                    }
                    fterm = warp_sum(fterm);
                    if (lane == 0) a.fpart[(int64_t)item * W + warp] = fterm;
                    if (tid == 0) scmd[dk & 15] = grp;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sbar[dk & 15]);                   // d, vg of this slice are stored
                    ++dk;
                } else {
                    if (row >= 0) a.partial[(int64_t)tile * a.B + row] = acc;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tile_empty[tbuf]);                // done with this tile buffer
                }
                if (its) its[7] = global_timer_ns();
            }
            pos.advance(nst);
        }
        if (!dose_over) {
            if (tid == 0) scmd[dk & 15] = -1;
            __syncwarp();
            if (lane == 0) mbar_arrive(&sbar[dk & 15]);
        }
        __threadfence();      // this warp's partial / fpart stores, while the other warps are still working
    }
    if (trace && tid == 0) trace[4] = global_timer_ns();
    if (a.dose_only) return;

    // ---------------- all transposed partials of all CTAs are written: grid barrier -----------------
    __syncthreads();
    if (tid == 0) {
        // sync[1] counts arrivals of all launches (never reset): this launch's arrivals take it from k*G to (k+1)*G
        const unsigned int old = atom_add_release_gpu(a.sync + 1, 1u);
        const unsigned int target = (old / G + 1u) * G;
        while ((int)(ld_acquire_gpu_u32(a.sync + 1) - target) < 0) { }
    }
    __syncthreads();
    if (trace && tid == 0) trace[5] = global_timer_ns();

    // ---------------- objective: this CTA's fixed share of the per-slice partials --------------------
    const int nfpart = a.n1 * W;
    if (warp == W) {
        const int per = (nfpart + (int)G - 1) / (int)G;
        const int f0 = blockIdx.x * per, f1 = min(nfpart, f0 + per);
        double v = 0.0;
        for (int k = f0 + lane; k < f1; k += 32) v += __ldcg(a.fpart + k);
        v = warp_sum(v);
        if (lane == 0) { a.fblock[blockIdx.x] = v; red_release_gpu_add(a.sync + 3, 1u); }
    }
    // ---------------- control points blockIdx.x, blockIdx.x + G, ...: tile partials -> gb -> g ------
    // (the objective's second level is done by CTA nb % G, one with the fewest control points)
    const bool multi = c.nranks > 1;
    const int parity = (int)(c.epoch & 1ull);
    const uint32_t ep = (uint32_t)c.epoch;
    const bool fcta = (int)blockIdx.x == a.nb % (int)G;
    double fv = 0.0;
    if (fcta) {     // the per-CTA shares in index order
        if (tid == 0) {
            while (ld_acquire_gpu_u32(a.sync + 3) < G) { }
            a.sync[3] = 0u;
            a.sync[0] = 0u;                                   // queue and group counters for the next evaluation:
            for (int g = 0; g < a.ngroups; ++g) grp_done[g] = 0u;     // everybody is past the barrier, nobody reads them any more
        }
        __syncthreads();
        fv = strided_sum_cg(a.fblock, (int64_t)tid, (int64_t)kFThreads, (int64_t)G);
        fv = block_sum_f(fv, red);
        __syncthreads();
        if (multi) { if (tid < c.nranks) ll_store(c.peer_bufs[tid], comm_entry(parity, c.rank, c.nranks, a.B, (size_t)a.B), fv, ep); }
        else if (tid == 0) { a.gbf[a.B] = fv; a.result[0] = fv; }
    }
    // A: this rank's sums over the tiles; single GPU: that is gb; multi-GPU: stored into every rank's buffer
    for (int cp = blockIdx.x; cp < a.nb; cp += (int)G) {
        const int lo = a.beam_ptr[cp], hi = a.beam_ptr[cp + 1];
        double contrib = 0.0;
        for (int j0 = lo; j0 < hi; j0 += kFThreads / 8) {
            // 8 threads per beamlet: thread q sums tiles t = q, q+8, ...; combined ((q0+q1)+(q2+q3))+((q4+q5)+(q6+q7))
            const int j = j0 + (tid >> 3), q = tid & 7;
            double s = 0.0;
            if (j < hi) s = strided_sum_cg(a.partial + j, (int64_t)q * a.B, (int64_t)8 * a.B, (int64_t)a.ntiles * a.B);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (j < hi) {
                if (multi) { for (int p = q; p < c.nranks; p += 8) ll_store(c.peer_bufs[p], comm_entry(parity, c.rank, c.nranks, a.B, (size_t)j), s, ep); }
                else if (q == 0) { a.gbf[j] = s; contrib += a.w_grad[j] * s; }
            }
        }
        if (!multi) {
            const double g = block_sum_f(contrib, red);
            if (tid == 0) a.result[1 + cp] = g;
            __syncthreads();
        }
    }
    if (multi) {
        // B/C: all of this CTA's values are on their way; add the ranks' parts in rank order as they arrive in
        // this rank's own buffer (identical bits on every rank)
        int ok = 1;
        for (int cp = blockIdx.x; cp < a.nb; cp += (int)G) {
            const int lo = a.beam_ptr[cp], hi = a.beam_ptr[cp + 1];
            double contrib = 0.0;
            for (int j = lo + tid; j < hi && ok; j += kFThreads) {
                double s = 0.0;
                if (!comm_sum_ranks(c, a.B, (size_t)j, s)) { ok = 0; break; }
                a.gbf[j] = s;
                contrib += a.w_grad[j] * s;
            }
            ok = __syncthreads_and(ok);
            const double g = block_sum_f(contrib, red);
            if (tid == 0) a.result[1 + cp] = ok ? g : nan("");
            __syncthreads();
        }
        if (fcta) {
            if (tid == 0) {
                double f = 0.0;
                if (ok && !comm_sum_ranks(c, a.B, (size_t)a.B, f)) ok = 0;
                a.gbf[a.B] = ok ? f : nan("");
                a.result[0] = ok ? f : nan("");
            }
        }
        if (tid == 0 && !ok) { atomicExch(c.error, 1); *c.status = 1.0; }
    }
    if (trace && tid == 0) trace[6] = global_timer_ns();
}

}  // namespace vmat
