// Evaluation kernels of libvmat_b200 (sm_100a): dose + penalty (K1), transposed
// product (K2), deterministic reductions (K3), layout construction.
//
// Data layout in HBM ("row-sorted sliced ELL stream"):
//   * rows (voxels for K1, (tile, beamlet) segments for K2) are sorted by length and cut into
//     slices that one warp processes: 32/tpr rows with tpr lanes per row (tpr = 1 for short rows,
//     a power of two for long ones, so that no slice is longer than kChunkCap chunks and work
//     items stay small against an SM's share); W consecutive slices form a super-slice that one
//     CTA processes with W consumer warps in lock step;
//   * a super-slice is one contiguous run of the stream: a header (per warp: the 32 row ids and,
//     for K1, the rows' penalty tables) followed by nch stages; stage c holds chunk c of every
//     warp; a chunk is 4 nonzeros of each of the warp's 32 lanes: the 32 index vectors (4 x int32,
//     or 4 x uint16 when every index fits - x has fewer than 65 536 entries and K2's indices are
//     tile-local), then the 32 value vectors;
//   * a producer thread streams the CTA's super-slices with TMA bulk copies into a shared-memory
//     ring (full/empty mbarriers); consumers read their chunk with conflict-free 128-bit shared
//     loads.  The stream is never touched by the LSU global path and no slice boundary stalls it;
//   * the vector being gathered (x for K1, a tile of vg for K2) is staged in shared memory so
//     the irregular accesses never leave the SM.
// Each lane accumulates its share of a row sequentially in storage order; the tpr lanes of a row
// are combined by a fixed butterfly.  No atomics anywhere: bit-reproducible.
#pragma once
#include "vmat_common.cuh"

namespace vmat {

constexpr int kStageHeader = 1;     // stage carries the super-slice header
constexpr int kStageLast = 2;       // last stage of the super-slice (K1/K2 consumers count the stages from the header's flags >> 8 instead)
struct StageInfo { int32_t ss, flags, tile, tpr; };

// header units (16 B) per warp: 32 row ids (+ thr, over, under as AT for K1)
// (+ with IW == 1 the 32 lanes' start indices as u32: the chunks then hold byte deltas)
__host__ __device__ constexpr int header_base_units(int IW) { return IW == 1 ? 8 : 0; }
template <typename AT> __host__ __device__ constexpr int header_units_k1(int IW) { return 8 + 3 * (32 * (int)sizeof(AT) / 16) + header_base_units(IW); }
__host__ __device__ constexpr int header_units_k2(int IW) { return 8 + header_base_units(IW); }

// one chunk of one warp, as it sits in a ring stage: 4 indices and 4 values per lane
template <typename AT> struct ChunkRegs { uint32_t i[4]; AT v[4]; };

// IW == 1: the index word holds four byte deltas; `cur` is the lane's running index (set from the
// header's start index, advanced here).  An entry whose stored value is zero contributes nothing, and
// its delta counts units of 256: one such entry bridges a gap of up to 65 280 (k_fill_stream puts
// them in front of every nonzero that is 256 or more away from its predecessor).
template <typename VT, typename AT, int IW>
__device__ __forceinline__ void chunk_load(const unsigned char* __restrict__ cp, int lane, ChunkRegs<AT>& c, uint32_t& cur) {
    const unsigned char* vp = cp + idx_units(IW) * 16;
    bool z0, z1, z2, z3;                     // stored value is (+-)0 (used by IW == 1 only)
    if constexpr (sizeof(VT) == 4) {
        const uint4 r = *reinterpret_cast<const uint4*>(vp + lane * 16);
        c.v[0] = (AT)__uint_as_float(r.x); c.v[1] = (AT)__uint_as_float(r.y);
        c.v[2] = (AT)__uint_as_float(r.z); c.v[3] = (AT)__uint_as_float(r.w);
        z0 = (r.x << 1) == 0u; z1 = (r.y << 1) == 0u; z2 = (r.z << 1) == 0u; z3 = (r.w << 1) == 0u;
    } else if constexpr (sizeof(VT) == 2) {
        const uint2 r = *reinterpret_cast<const uint2*>(vp + lane * 8);
        c.v[0] = (AT)__uint_as_float(r.x << 16); c.v[1] = (AT)__uint_as_float(r.x & 0xffff0000u);
        c.v[2] = (AT)__uint_as_float(r.y << 16); c.v[3] = (AT)__uint_as_float(r.y & 0xffff0000u);
        z0 = (r.x & 0x7fffu) == 0u; z1 = (r.x & 0x7fff0000u) == 0u; z2 = (r.y & 0x7fffu) == 0u; z3 = (r.y & 0x7fff0000u) == 0u;
    } else {
        const double2 a = *reinterpret_cast<const double2*>(vp + lane * 16);
        const double2 b = *reinterpret_cast<const double2*>(vp + 512 + lane * 16);
        c.v[0] = (AT)a.x; c.v[1] = (AT)a.y; c.v[2] = (AT)b.x; c.v[3] = (AT)b.y;
        z0 = a.x == 0.0; z1 = a.y == 0.0; z2 = b.x == 0.0; z3 = b.y == 0.0;
    }
    if constexpr (IW == 4) {
        const uint4 idx = *reinterpret_cast<const uint4*>(cp + lane * 16);
        c.i[0] = idx.x; c.i[1] = idx.y; c.i[2] = idx.z; c.i[3] = idx.w;
    } else if constexpr (IW == 2) {
        const uint2 idx = *reinterpret_cast<const uint2*>(cp + lane * 8);
        c.i[0] = idx.x & 0xffffu; c.i[1] = idx.x >> 16; c.i[2] = idx.y & 0xffffu; c.i[3] = idx.y >> 16;
    } else {
        const uint32_t d = *reinterpret_cast<const uint32_t*>(cp + lane * 4);
        const uint32_t d0 = d & 0xffu, d1 = (d >> 8) & 0xffu, d2 = (d >> 16) & 0xffu, d3 = d >> 24;
        c.i[0] = cur + (z0 ? d0 << 8 : d0);
        c.i[1] = c.i[0] + (z1 ? d1 << 8 : d1);
        c.i[2] = c.i[1] + (z2 ? d2 << 8 : d2);
        c.i[3] = c.i[2] + (z3 ? d3 << 8 : d3);
        cur = c.i[3];
    }
}

// ring bookkeeping shared by producer and consumers
struct RingPos {
    int stage = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance(int nstages) { if (++stage == nstages) { stage = 0; phase ^= 1u; } }
};

// producer side: publish one stage (header or data) of `bytes` from `src`
__device__ __forceinline__ void ring_push(unsigned char* ring, int stage_bytes, uint64_t* full, uint64_t* empty,
                                          StageInfo* sinfo, RingPos& pos, int nstages, const StageInfo& info,
                                          const void* src, uint32_t bytes, uint64_t policy) {
    mbar_wait(&empty[pos.stage], pos.phase ^ 1u);
    sinfo[pos.stage] = info;
    if (bytes) {
        mbar_expect_tx(&full[pos.stage], bytes);
        tma_bulk_g2s_hint(ring + (size_t)pos.stage * stage_bytes, src, bytes, &full[pos.stage], policy);
    } else {
        __threadfence_block();        // the descriptor store is performed before the arrive that publishes it (no copy, no latency, in between)
        mbar_arrive(&full[pos.stage]);
    }
    pos.advance(nstages);
}

// mbarrier.try_wait once (it suspends the thread for a hardware-defined time; the answer takes ~90 cycles even
// when the phase is complete, so several are issued back to back before any is looked at)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0u;
}

// Index guard of the byte-delta layout (IW == 1), where an index is the result of arithmetic on streamed bytes: a
// running index past the gathered vector means the stream, or the synchronisation around the ring, is broken.  The
// thread that sees one writes a record into pinned host memory (it survives the context) and traps; the host
// reports it (vmat_guard_status).  Absolute indices (IW 2 / 4) are range-checked when the plan is built.
constexpr int kGuardWords = 8;
struct Guard {
    unsigned long long* host;     // [kGuardWords] mapped pinned host memory; word 0 != 0: tripped
    uint32_t limit;               // valid indices are < limit
    int32_t kernel;               // 1 = dose kernel, 2 = transposed kernel
};
__device__ __forceinline__ void guard_check(const Guard& gd, uint32_t cur_in, uint32_t cur, int ss, int c, int slot) {
    if (cur >= gd.limit) {        // indices only grow along a lane: the last one bounds them all
        volatile unsigned long long* h = gd.host;
        h[1] = ((unsigned long long)(uint32_t)gd.kernel << 32) | blockIdx.x;
        h[2] = ((unsigned long long)threadIdx.x << 32) | (uint32_t)slot;
        h[3] = ((unsigned long long)(uint32_t)ss << 32) | (uint32_t)c;
        h[4] = ((unsigned long long)cur_in << 32) | cur;
        h[5] = gd.limit;
        __threadfence_system();
        h[0] = 0x6775617264ull;   // "guard"
        __threadfence_system();
        __trap();
    }
}

// Release of a ring stage (or of anything else the producer refills).  SYNCS.ARRIVE does not wait for shared-memory
// loads that are still in flight - ptxas gives it no dependency on them - so a load whose result is not needed
// before the arrive can execute after the producer has refilled the slot (seen on config 4 with byte deltas, where
// the producer is always waiting: a header word read after the refill, then a wild gather).  `dep` is any value
// computed from every such load, `zero` a kernel argument that is 0: the address then depends on the loads.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint32_t dep, uint32_t zero) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar) + (dep & zero)) : "memory");
}
__device__ __forceinline__ uint32_t dep_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t dep_bits(double v) { return (uint32_t)__double2loint(v) ^ (uint32_t)__double2hiint(v); }

// consumer side: U data stages starting at `pos` (all of one super-slice): acc += this warp's chunks . gathered
// vector, in chunk order.  The U barrier tests, the U sets of shared-memory loads and the 4U gathers are in
// flight together; the stages go back to the producer with one warp synchronisation.  `pos` ends past them.
template <typename VT, typename AT, int IW, int U, typename Gather>
__device__ __forceinline__ void consume_stages(const unsigned char* ring, int SB, int warp_off, uint64_t* full,
                                               uint64_t* empty, RingPos& pos, int nst, int lane, AT& acc,
                                               uint32_t& cur, Gather g, const Guard& gd, uint32_t zero, int ss, int c) {
    int st[U];
    bool ok[U];
    uint32_t ph[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { st[u] = pos.stage; ph[u] = pos.phase; pos.advance(nst); }
#pragma unroll
    for (int u = 0; u < U; ++u) ok[u] = mbar_try(&full[st[u]], ph[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) if (!ok[u]) mbar_wait(&full[st[u]], ph[u]);
    ChunkRegs<AT> R[U];
    const uint32_t cur_in = cur;
#pragma unroll
    for (int u = 0; u < U; ++u) chunk_load<VT, AT, IW>(ring + (size_t)st[u] * SB + warp_off, lane, R[u], cur);
    if constexpr (IW == 1) guard_check(gd, cur_in, cur, ss, c, st[0]);
    AT x[U][4];
    uint32_t dep = 0;             // every load from the ring: the indices feed the gathers, the values feed `dep`
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) x[u][k] = g(R[u].i[k]);
        dep |= dep_bits(R[u].v[0]) | dep_bits(R[u].v[3]);
    }
    __syncwarp();
    if (lane == 0) {
#pragma unroll
        for (int u = 0; u < U; ++u) mbar_arrive_after(&empty[st[u]], dep, zero);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int k = 0; k < 4; ++k) acc = fma_acc(R[u].v[k], x[u][k], acc);
    }
}

// the nch data stages of a super-slice, `batch` (4, 2 or 1) at a time, then two, then one.  The consumers hold a
// batch until all of its stages have arrived, so the batch may not exceed the ring depth - and on a shallow ring
// a large batch leaves the producer too few free stages (the host picks it, consume_batch_for).  The row sum order
// is the chunk order whatever the grouping.
template <typename VT, typename AT, int IW, typename Gather>
__device__ __forceinline__ void consume_run(const unsigned char* ring, int SB, int warp_off, uint64_t* full,
                                            uint64_t* empty, RingPos& pos, int nst, int nch, int lane, AT& acc,
                                            uint32_t& cur, Gather g, int batch, const Guard& gd, uint32_t zero, int ss) {
    int c = 0;
    if (batch >= 4)
        for (; c + 4 <= nch; c += 4) consume_stages<VT, AT, IW, 4>(ring, SB, warp_off, full, empty, pos, nst, lane, acc, cur, g, gd, zero, ss, c);
    if (batch >= 2)
        for (; c + 2 <= nch; c += 2) consume_stages<VT, AT, IW, 2>(ring, SB, warp_off, full, empty, pos, nst, lane, acc, cur, g, gd, zero, ss, c);
    for (; c < nch; ++c) consume_stages<VT, AT, IW, 1>(ring, SB, warp_off, full, empty, pos, nst, lane, acc, cur, g, gd, zero, ss, c);
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
// CTA = W consumer warps + 1 producer warp; persistent, one CTA per SM; super-slices are
// claimed from an atomic counter, longest first.
// ---------------------------------------------------------------------------------
constexpr int kInlineY = 384;

constexpr int kTraceSlots = 8;

template <typename AT>
struct K1Args {
    const uint4* sell;            // stream (voxel-major)
    const int64_t* ss_off;        // [nss+1] in 16-byte units
    const uint8_t* ss_tpr;        // [nss] lanes per row in this super-slice
    int32_t nss;
    int32_t nstages;              // ring depth
    int32_t stage_bytes;          // ring slot size
    int32_t B;
    const double* y;              // [nb] intensities
    const int32_t* beam_of;       // [B]
    const AT* w_dose;             // [B]
    const AT* x_global;           // [B] (only read when XS == false)
    AT* d;                        // [V] dose, natural voxel order
    AT* vg;                       // [V] voxel gradient
    double* fpart;                // [nss*W] objective partial of every slice (fixed order -> deterministic)
    unsigned int* counter;        // dynamic super-slice scheduler
    // host-buffer path (vmat_eval): y is read straight from mapped pinned host memory, mirrored
    // into the device copy, and the evaluation sequence number is bumped for K3's completion tags
    double* y_mirror;             // [nb] or nullptr
    int32_t nb;
    unsigned long long* seq;      // nullptr on the resident path
    // host-buffer path, nb <= kInlineY: the intensities travel inside the launch parameters, so no
    // host-to-device copy sits in front of the kernel
    int32_t y_inline_n;
    double y_inline[kInlineY];
    // diagnostics (vmat_trace_eval): per CTA, globaltimer at [entry, x staged + predecessor done,
    // first stage landed, consumers done, producer done]; nullptr = off
    unsigned long long* trace;
    int32_t claim_lead;           // stages before the end of a super-slice at which the next one is claimed
    unsigned int* k2_pool_counter; // the transposed kernel's dynamic-tail counter, reset here for this evaluation (or nullptr)
    int32_t batch;                // ring stages a consumer warp takes per pass (1, 2 or 4; at most the ring depth)
    Guard guard;                  // index guard of the byte-delta layout
    uint32_t zero;                // 0 (mbar_arrive_after)
};

// four consecutive elements with one (float) or two (double) 16-byte loads
__device__ __forceinline__ void load4(const float* p, float (&o)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
__device__ __forceinline__ void load4(const double* p, double (&o)[4]) {
    const double2 v0 = reinterpret_cast<const double2*>(p)[0], v1 = reinterpret_cast<const double2*>(p)[1];
    o[0] = v0.x; o[1] = v0.y; o[2] = v1.x; o[3] = v1.y;
}

__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d) {
    reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
    reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
}

template <typename VT, typename AT, bool XS, int W, int IW>
__global__ void __launch_bounds__((W + 1) * 32) k1_dose_penalty(const __grid_constant__ K1Args<AT> a) {
    constexpr int KU = chunk_units<VT, IW>(), HU = header_units_k1<AT>(IW);
    constexpr int kMaxStages = 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages];
    __shared__ StageInfo sinfo[kMaxStages];
    const int xbytes = XS ? ((a.B * (int)sizeof(AT) + 127) & ~127) : 0;
    const int ybytes = XS ? ((a.nb * 8 + 127) & ~127) : 0;
    AT* xs = reinterpret_cast<AT*>(smem_raw);
    double* ys = reinterpret_cast<double*>(smem_raw + xbytes);
    unsigned char* ring = smem_raw + xbytes + ybytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nst = a.nstages, SB = a.stage_bytes;
    unsigned long long* const trace = a.trace ? a.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;

    if (tid == 0) {
        if (trace) trace[0] = global_timer_ns();
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], W); }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    __syncthreads();                                   // barriers initialised
    if (warp < W) {
        if (XS) {   // x_j = y[beam(j)] * w_dose[j]   (K4 folded into the prologue; consumers only -
                    // the producer is already streaming D meanwhile)
            // The beamlet tables of this thread's first groups of four are requested before y is
            // waited for, so the two round trips to L2 overlap.
            constexpr int kPre = 4;
            const int ngroups = a.B >> 2;                  // whole groups of four beamlets
            int4 bo[kPre];
            AT wd[kPre][4];
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int gidx = tid + k * W * 32;
                if (gidx < ngroups) {
                    bo[k] = reinterpret_cast<const int4*>(a.beam_of)[gidx];
                    load4(a.w_dose + 4 * gidx, wd[k]);
                }
            }
            // y (possibly in mapped host memory) is read once per CTA into shared memory
            if (a.y_inline_n > 0) { for (int j = tid; j < a.nb; j += W * 32) ys[j] = a.y_inline[j]; }
            else { for (int j = tid; j < a.nb; j += W * 32) ys[j] = a.y[j]; }
            named_bar_sync(1, W * 32);
#pragma unroll
            for (int k = 0; k < kPre; ++k) {
                const int gidx = tid + k * W * 32;
                if (gidx < ngroups) {
                    store4(xs + 4 * gidx, mul_rn((AT)ys[bo[k].x], wd[k][0]), mul_rn((AT)ys[bo[k].y], wd[k][1]),
                           mul_rn((AT)ys[bo[k].z], wd[k][2]), mul_rn((AT)ys[bo[k].w], wd[k][3]));
                }
            }
            for (int gidx = tid + kPre * W * 32; gidx < ngroups; gidx += W * 32) {
                const int4 b4 = reinterpret_cast<const int4*>(a.beam_of)[gidx];
                AT w4[4];
                load4(a.w_dose + 4 * gidx, w4);
                store4(xs + 4 * gidx, mul_rn((AT)ys[b4.x], w4[0]), mul_rn((AT)ys[b4.y], w4[1]),
                       mul_rn((AT)ys[b4.z], w4[2]), mul_rn((AT)ys[b4.w], w4[3]));
            }
            for (int j = 4 * ngroups + tid; j < a.B; j += W * 32) xs[j] = mul_rn((AT)ys[a.beam_of[j]], a.w_dose[j]);
            named_bar_sync(1, W * 32);
        }
    }
    // d, vg, fpart and the slice counter belong to the previous evaluation until its kernels are
    // done: consumers wait here; the producer first streams its statically assigned super-slice
    // (read-only data into its own ring) and waits only before it touches the counter
    if (warp < W) pdl_wait();
    if (trace && tid == 0) trace[1] = global_timer_ns();
    if (blockIdx.x == 0 && warp < W) {
        if (XS && a.y_mirror != nullptr) { for (int j = tid; j < a.nb; j += W * 32) a.y_mirror[j] = ys[j]; }
        if (a.seq != nullptr && tid == 0) atomicAdd(a.seq, 1ull);
        if (a.k2_pool_counter != nullptr && tid == 0) *a.k2_pool_counter = 0u;      // (the previous evaluation is complete)
    }

    if (warp == W) {
        // ---------------- producer: one thread streams super-slices into the ring ----------
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            RingPos pos;
            const unsigned int ns = (unsigned)a.nss;
            // super-slices 0 .. gridDim-1 (the longest) are assigned statically, the rest are
            // claimed from the counter while the last `claim_lead` stages of the current one are
            // being pushed: late enough that no CTA sits on work others could take (the traced
            // spread of CTA finish times halves against claiming two super-slices ahead), early
            // enough that the claim and the offset loads are back before the producer needs them
            const unsigned int G = gridDim.x;
            const int lead = a.claim_lead;
            unsigned int cur = blockIdx.x;
            bool first = true;
            int64_t o0 = 0, o1 = 0;
            int tpr = 1;
            if (cur < ns) { o0 = a.ss_off[cur]; o1 = a.ss_off[cur + 1]; tpr = a.ss_tpr[cur]; }
            while (cur < ns) {
                const int nch = (int)((o1 - o0 - W * HU) / (W * KU));
                unsigned int nxt = 0u;
                int64_t n0 = 0, n1 = 0;
                int ntpr = 1;
                bool claimed = false, loaded = false;
                const int c_claim = max(0, nch - lead), c_load = min(c_claim + 4, max(0, nch - 1));
                ring_push(ring, SB, full, empty, sinfo, pos, nst,
                          StageInfo{(int32_t)cur, kStageHeader | (nch == 0 ? kStageLast : 0) | (nch << 8), 0, tpr},
                          a.sell + o0, W * HU * 16, pol);
                const uint4* data = a.sell + o0 + W * HU;
                for (int c = 0; c < nch; ++c) {
                    if (!first) {
                        if (c == c_claim) { nxt = G + atomicAdd(a.counter, 1u); claimed = true; }
                        if (c == c_load && claimed) {
                            if (nxt < ns) { n0 = a.ss_off[nxt]; n1 = a.ss_off[nxt + 1]; ntpr = a.ss_tpr[nxt]; }
                            loaded = true;
                        }
                    }
                    ring_push(ring, SB, full, empty, sinfo, pos, nst,
                              StageInfo{(int32_t)cur, c == nch - 1 ? kStageLast : 0, 0, tpr},
                              data + (int64_t)c * W * KU, W * KU * 16, pol);
                }
                if (first) {
                    pdl_wait();                       // the counter was reset by the previous evaluation's K2
                    first = false;
                }
                if (!claimed) nxt = G + atomicAdd(a.counter, 1u);
                if (!loaded && nxt < ns) { n0 = a.ss_off[nxt]; n1 = a.ss_off[nxt + 1]; ntpr = a.ss_tpr[nxt]; }
                cur = nxt; o0 = n0; o1 = n1; tpr = ntpr;
            }
            ring_push(ring, SB, full, empty, sinfo, pos, nst, StageInfo{-1, 0, 0, 0}, nullptr, 0, pol);
            if (trace) trace[4] = global_timer_ns();
        }
        return;
    }

    // ---------------- consumers: warp w owns slice w of every super-slice -------------------
    auto gather = [&](uint32_t j) -> AT { return XS ? xs[j] : __ldg(a.x_global + j); };
    RingPos pos;
    AT acc = (AT)0, t = (AT)0, ov = (AT)0, un = (AT)0;
    int32_t row = -1;
    uint32_t cur = 0;
    if (trace) { mbar_wait(&full[0], 0); if (tid == 0) trace[2] = global_timer_ns(); }    // outside the hot loop
    while (true) {
        // a super-slice: its header stage (which says how many data stages follow), the data stages, the epilogue
        mbar_wait(&full[pos.stage], pos.phase);      // every lane polls: a single polling lane + __syncwarp measured 2.5x slower
        const StageInfo info = sinfo[pos.stage];
        if (info.ss < 0) break;
        {
            const unsigned char* hp = ring + (size_t)pos.stage * SB + warp * (HU * 16);
            row = reinterpret_cast<const int32_t*>(hp)[lane];
            t  = reinterpret_cast<const AT*>(hp + 128)[lane];
            ov = reinterpret_cast<const AT*>(hp + 128 + 32 * sizeof(AT))[lane];
            un = reinterpret_cast<const AT*>(hp + 128 + 64 * sizeof(AT))[lane];
            if constexpr (IW == 1) cur = reinterpret_cast<const uint32_t*>(hp + (HU - 8) * 16)[lane];
            acc = (AT)0;
            __syncwarp();
            // the slot goes back only when the header words have arrived in registers (mbar_arrive_after)
            if (lane == 0) mbar_arrive_after(&empty[pos.stage], (uint32_t)row | cur | dep_bits(t) | dep_bits(ov) | dep_bits(un), a.zero);
            pos.advance(nst);
        }
        consume_run<VT, AT, IW>(ring, SB, warp * (KU * 16), full, empty, pos, nst, info.flags >> 8, lane, acc, cur, gather, a.batch, a.guard, a.zero, info.ss);
        {
            for (int o = 1; o < info.tpr; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // tpr lanes of a row
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
            fterm = warp_sum(fterm);      // butterfly: same value and order in every lane, every run
            if (lane == 0) a.fpart[(int64_t)info.ss * W + warp] = fterm;
        }
    }
    if (trace && tid == 0) trace[3] = global_timer_ns();
}

// ---------------------------------------------------------------------------------
// K2: partial[tile][beamlet] = sum over the tile's voxels of D[v, beamlet] * vg[v]
//     (transposed product over the pre-built beamlet-major layout; PPsubroutine :646)
// Persistent CTAs (W consumer warps + producer); the host cut the (tile, length)-ordered
// super-slice sequence into gridDim.x pieces of equal work, so there is no tail.  The vg tile
// is staged by one TMA bulk copy per tile change.
// ---------------------------------------------------------------------------------
struct K2Task { int32_t tile, ss_begin, ss_end, pad; };

template <typename AT>
struct K2Args {
    const uint4* sell;            // stream (beamlet-major, column-tiled)
    const int64_t* ss_off;        // [nss+1]
    const uint8_t* ss_tpr;        // [nss]
    const K2Task* tasks;          // (tile, super-slice range), ordered by tile
    const int32_t* cta_task_ptr;  // [gridDim.x+1]: CTA c runs tasks [ptr[c], ptr[c+1])
    const AT* vg;                 // [V + T] (zero padded so a tile copy may run past V)
    AT* partial;                  // [ntiles][B]
    int64_t V;
    int32_t B;
    int32_t T;                    // voxels per tile
    int32_t nstages, stage_bytes;
    unsigned int* k1_counter;     // reset here for the next evaluation
    const double* fpart;          // [nfpart] K1's per-slice objective partials ...
    double* fblock;               // [gridDim.x] ... folded here into one fixed-order sum per CTA
    int32_t nfpart;
    unsigned long long* trace;    // diagnostics, as in K1Args: [entry, K1 done, first stage, consumers done, producer done]
    // dynamic tail: the shortest super-slices of every tile are not part of the static pieces but of a pool of small
    // tasks that CTAs claim when their own piece is done - it absorbs the spread of the pieces' finish times
    const K2Task* pool_tasks;     // [npool], ordered by tile
    int32_t npool;
    unsigned int* pool_counter;   // reset by the dose kernel of the same evaluation
    int32_t batch;                // ring stages a consumer warp takes per pass (1, 2 or 4; at most the ring depth)
    Guard guard;                  // index guard of the byte-delta layout
    uint32_t zero;                // 0 (mbar_arrive_after)
};

template <typename VT, typename AT, int W, int IW>
__global__ void __launch_bounds__((W + 1) * 32) k2_beamlet_grad(const K2Args<AT> a) {
    constexpr int KU = chunk_units<VT, IW>(), HU = header_units_k2(IW);
    constexpr int kMaxStages = 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kMaxStages], empty[kMaxStages], tilebar;
    __shared__ StageInfo sinfo[kMaxStages];
    const int tbytes = (a.T * (int)sizeof(AT) + 127) & ~127;
    AT* vt = reinterpret_cast<AT*>(smem_raw);
    unsigned char* ring = smem_raw + tbytes;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nst = a.nstages, SB = a.stage_bytes;
    unsigned long long* const trace = a.trace ? a.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
    pdl_launch_dependents();
    if (tid == 0) {
        if (trace) trace[0] = global_timer_ns();
        for (int s = 0; s < nst; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], W); }
        mbar_init(&tilebar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    const int t_begin = a.cta_task_ptr[blockIdx.x], t_end = a.cta_task_ptr[blockIdx.x + 1];

    if (warp == W) {     // the producer streams D (static data) right away; only consumers need K1's vg
        if (lane == 0) {
            const uint64_t pol = l2_evict_first_policy();
            RingPos pos;
            bool dynamic = false;
            for (int ti = t_begin;; ++ti) {
                K2Task task;
                if (!dynamic && ti < t_end) {
                    task = a.tasks[ti];
                } else {
                    if (a.npool == 0) break;
                    if (!dynamic) { pdl_wait(); dynamic = true; }      // the dose kernel has reset the pool counter
                    const unsigned int q = atomicAdd(a.pool_counter, 1u);
                    if (q >= (unsigned)a.npool) break;
                    task = a.pool_tasks[q];
                }
                int64_t o0 = a.ss_off[task.ss_begin];
                for (int ss = task.ss_begin; ss < task.ss_end; ++ss) {
                    const int64_t o1 = a.ss_off[ss + 1];
                    const int nch = (int)((o1 - o0 - W * HU) / (W * KU));
                    const int tpr = a.ss_tpr[ss];
                    ring_push(ring, SB, full, empty, sinfo, pos, nst,
                              StageInfo{ss, kStageHeader | (nch == 0 ? kStageLast : 0) | (nch << 8), task.tile, tpr},
                              a.sell + o0, W * HU * 16, pol);
                    const uint4* data = a.sell + o0 + W * HU;
                    for (int c = 0; c < nch; ++c)
                        ring_push(ring, SB, full, empty, sinfo, pos, nst,
                                  StageInfo{ss, c == nch - 1 ? kStageLast : 0, task.tile, tpr},
                                  data + (int64_t)c * W * KU, W * KU * 16, pol);
                    o0 = o1;
                }
            }
            ring_push(ring, SB, full, empty, sinfo, pos, nst, StageInfo{-1, 0, 0, 0}, nullptr, 0, pol);
            if (trace) trace[4] = global_timer_ns();
        }
        return;
    }

    pdl_wait();          // vg is complete, the previous K3 has consumed `partial`
    if (trace && tid == 0) trace[1] = global_timer_ns();
    if (blockIdx.x == 0 && tid == 0) *a.k1_counter = 0u;
    int cur_tile = -1;
    bool tile_pending = false;
    if (t_begin < t_end) {   // the first tile is requested right away, not when the first header arrives
        cur_tile = a.tasks[t_begin].tile;
        if (tid == 0) {
            const int64_t base = (int64_t)cur_tile * a.T;
            const int n = (int)min((int64_t)a.T, a.V - base);
            const uint32_t bytes = (uint32_t)(((size_t)n * sizeof(AT) + 15) & ~(size_t)15);
            mbar_expect_tx(&tilebar, bytes);
            tma_bulk_g2s(vt, a.vg + base, bytes, &tilebar);
        }
        tile_pending = true;
    }
    if (warp == 0) {     // objective: this CTA's fixed share of the per-slice partials (order fixed -> reproducible)
        const int per = (a.nfpart + gridDim.x - 1) / gridDim.x;
        const int f0 = blockIdx.x * per, f1 = min(a.nfpart, f0 + per);
        double v = 0.0;
        for (int k = f0 + lane; k < f1; k += 32) v += a.fpart[k];
        v = warp_sum(v);
        if (lane == 0) a.fblock[blockIdx.x] = v;
    }
    auto gather = [&](uint32_t j) -> AT { return vt[j]; };
    RingPos pos;
    AT acc = (AT)0;
    int32_t row = -1;
    uint32_t cur = 0;
    uint32_t tparity = 0;
    if (trace) { mbar_wait(&full[0], 0); if (tid == 0) trace[2] = global_timer_ns(); }    // outside the hot loop
    while (true) {
        // a super-slice: its header stage (which says how many data stages follow), the data stages, the store
        mbar_wait(&full[pos.stage], pos.phase);      // every lane polls: a single polling lane + __syncwarp measured 2.5x slower
        const StageInfo info = sinfo[pos.stage];
        if (info.ss < 0) break;
        {
            const unsigned char* sp = ring + (size_t)pos.stage * SB;
            if (info.tile != cur_tile) {
                named_bar_sync(1, W * 32);            // all consumers are done with the previous tile
                if (tid == 0) {
                    const int64_t base = (int64_t)info.tile * a.T;
                    const int n = (int)min((int64_t)a.T, a.V - base);
                    const uint32_t bytes = (uint32_t)(((size_t)n * sizeof(AT) + 15) & ~(size_t)15);
                    mbar_expect_tx(&tilebar, bytes);
                    tma_bulk_g2s(vt, a.vg + base, bytes, &tilebar);
                }
                mbar_wait(&tilebar, tparity);
                tparity ^= 1u;
                cur_tile = info.tile;
                tile_pending = false;
            } else if (tile_pending) {
                mbar_wait(&tilebar, tparity);
                tparity ^= 1u;
                tile_pending = false;
            }
            row = reinterpret_cast<const int32_t*>(sp + warp * (HU * 16))[lane];
            if constexpr (IW == 1) cur = reinterpret_cast<const uint32_t*>(sp + warp * (HU * 16) + (HU - 8) * 16)[lane];
            acc = (AT)0;
            __syncwarp();
            // the slot goes back only when the header words have arrived in registers (mbar_arrive_after)
            if (lane == 0) mbar_arrive_after(&empty[pos.stage], (uint32_t)row | cur, a.zero);
            pos.advance(nst);
        }
        consume_run<VT, AT, IW>(ring, SB, warp * (KU * 16), full, empty, pos, nst, info.flags >> 8, lane, acc, cur, gather, a.batch, a.guard, a.zero, info.ss);
        for (int o = 1; o < info.tpr; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (row >= 0) a.partial[(int64_t)cur_tile * a.B + row] = acc;
    }
    if (trace && tid == 0) trace[3] = global_timer_ns();
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
    const double* fblock;         // [nfblock] objective sums of K2's CTAs
    int32_t nfblock;
    // host-buffer path: results also go to mapped pinned host memory, each slot followed by a tag
    // (the evaluation sequence number) that the host polls instead of synchronising the stream
    double* host_result;          // [nb+1] or nullptr
    volatile unsigned long long* host_tags;   // [nb+1]
    const unsigned long long* seq;
    // diagnostics (vmat_trace_eval): per CTA, globaltimer at [0] predecessor finished, [1] this rank's partials reduced
    // (and stored into the peers' buffers), [2] all ranks' values summed (= every peer's values have arrived), [3] done
    unsigned long long* trace;
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

// sum of src[first], src[first+stride], ... (< n) in that order; loads are issued in batches of 8
// so the chain costs one memory round trip per 8 terms instead of one per term
template <typename T>
__device__ __forceinline__ double strided_sum(const T* __restrict__ src, int64_t first, int64_t stride, int64_t n) {
    double s = 0.0;
    int64_t k = first;
    for (; k + 7 * stride < n; k += 8 * stride) {
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[k + u * stride];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += (double)v[u];
    }
    {   // tail (< 8 terms): one more batch of predicated loads instead of a chain of single loads
        T v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (k + u * stride < n) ? src[k + u * stride] : (T)0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (k + u * stride < n) s += (double)v[u];
    }
    return s;
}

template <typename AT>
__global__ void __launch_bounds__(256) k3_reduce(const K3Args<AT> a, int phase) {
    __shared__ double sh[256];
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    const int lo = a.beam_ptr[blockIdx.x], hi = a.beam_ptr[blockIdx.x + 1];
    pdl_wait();
    unsigned long long* const trace = a.trace ? a.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
    if (trace && tid == 0) trace[0] = global_timer_ns();
    // objective (block 0): the per-CTA sums K2 left in fblock, added in index order; the loads are
    // issued before the beamlet loop so that they share its memory round trip
    double fv = 0.0;
    if (blockIdx.x == 0 && phase != 2) fv = strided_sum(a.fblock, tid, 256, a.nfblock);
    double contrib = 0.0;
    for (int j0 = lo; j0 < hi; j0 += 64) {
        // 4 threads per beamlet: thread q sums tiles t = q, q+4, ...; combined (q0+q1)+(q2+q3)
        const int j = j0 + (tid >> 2), q = tid & 3;
        double s = 0.0;
        const double wg = (phase != 1 && j < hi && q == 0) ? a.w_grad[j] : 0.0;
        if (phase != 2) {
            if (j < hi) s = strided_sum(a.partial + j, (int64_t)q * a.B, (int64_t)4 * a.B, (int64_t)a.ntiles * a.B);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            if (j < hi && q == 0) a.gbf[j] = s;
        } else if (j < hi) {
            s = a.gbf[j];
        }
        contrib += wg * s;
    }
    if (trace && tid == 0) trace[1] = trace[2] = global_timer_ns();
    const bool to_host = a.host_result != nullptr && phase == 0;
    unsigned long long seq = 0;
    if (to_host && tid == 0) seq = *a.seq;
    if (phase != 1) {
        const double g = block_sum_256(contrib, sh);
        if (tid == 0) {
            a.result[1 + blockIdx.x] = g;
            if (to_host) { a.host_result[1 + blockIdx.x] = g; __threadfence_system(); a.host_tags[1 + blockIdx.x] = seq; }
        }
    }
    if (trace && tid == 0) trace[3] = global_timer_ns();
    if (blockIdx.x == 0) {
        if (phase == 2) {
            if (tid == 0) a.result[0] = a.gbf[a.B];
        } else {
            fv = block_sum_256(fv, sh);
            if (tid == 0) {
                a.gbf[a.B] = fv;
                if (phase == 0) a.result[0] = fv;
                if (to_host) { a.host_result[0] = fv; __threadfence_system(); a.host_tags[0] = seq; }
            }
        }
    }
}

// ---------------------------------------------------------------------------------
// K3 fused with the cross-GPU exchange (multi-GPU, one rank per GPU): a push all-reduce inside
// the reduction kernel, with the flag carried by the data (the "LL" idea of NCCL's low-latency
// protocol).  Every rank owns a "comm buffer" in its HBM that all peers have mapped (CUDA IPC over
// NVLink, or plain peer pointers inside one process):
//     recv[parity 0..1][source rank 0..nranks-1][gb partial (B) | f partial]   16 bytes per value
// A value travels as two 8-byte words {low 32 bits | epoch << 32, high 32 bits | epoch << 32}
// written with one 16-byte store (each 8-byte half is atomic).  CTA i (control point i) reduces this
// rank's tile partials for its beamlets and stores them into slot [parity][my rank] of EVERY rank's
// buffer - fire-and-forget remote stores, no fence, no separate flag.  The thread that needs value
// j then polls its OWN buffer (local memory) until both words of every rank's entry carry this
// evaluation's epoch, and adds the ranks in rank order: identical bits on every rank, one NVLink
// one-way latency, no NCCL launch, no rank-wide step (a CTA depends on the peers' CTA of the same
// control point only).  Two parities make a slot safe to overwrite in the next evaluation: a rank
// can only be one evaluation ahead of a peer, because finishing evaluation e needs every peer's
// values of e, which the peer stores after its own reduction of e-1 has completed (stream order).
// Epochs are counted by the host (1, 2, ...: the same on all ranks; 0 is the cleared buffer).
// ---------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;

struct CommArgs {
    double* const* peer_bufs;     // [nranks] device pointers to every rank's comm buffer (own included)
    int32_t rank, nranks;
    unsigned long long epoch;     // this evaluation's epoch
    unsigned long long timeout_ns;
    int32_t* error;               // set when a peer did not deliver within the time-out
    double* status;               // result[nb+1]: set to 1.0 by a CTA that timed out
};

// entry (parity, source rank, j) of a comm buffer, in 16-byte units
__device__ __forceinline__ size_t comm_entry(int parity, int src, int nranks, int B, size_t j) {
    return ((size_t)parity * nranks + src) * ((size_t)B + 1) + j;
}
__device__ __forceinline__ void ll_store(double* buf, size_t entry, double v, uint32_t ep) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v), tag = (unsigned long long)ep << 32;
    const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};"
                 :: "l"(reinterpret_cast<ulonglong2*>(buf) + entry), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ ulonglong2 ll_load_raw(const double* buf, size_t entry) {
    ulonglong2 w;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];"
                 : "=l"(w.x), "=l"(w.y) : "l"(reinterpret_cast<const ulonglong2*>(buf) + entry) : "memory");
    return w;
}
// Sum over the ranks, in rank order, of entry j of this rank's own buffer; waits for entries that have not
// arrived yet.  Returns false on time-out.  (Eight ranks at a time: their loads are issued together, late
// ones are retried.)
__device__ __forceinline__ bool comm_sum_ranks(const CommArgs& c, int B, size_t j, double& out) {
    const double* mybuf = c.peer_bufs[c.rank];
    const int parity = (int)(c.epoch & 1ull);
    const uint32_t ep = (uint32_t)c.epoch;
    double s = 0.0;
    for (int q0 = 0; q0 < c.nranks; q0 += 8) {
        ulonglong2 w[8];
        uint32_t missing = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q0 + q < c.nranks) w[q] = ll_load_raw(mybuf, comm_entry(parity, q0 + q, c.nranks, B, j));
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q0 + q < c.nranks && ((uint32_t)(w[q].x >> 32) != ep || (uint32_t)(w[q].y >> 32) != ep)) missing |= 1u << q;
        if (missing) {
            const unsigned long long t0 = global_timer_ns();
            unsigned int spins = 0;
            while (missing) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if ((missing >> q) & 1u) {
                        w[q] = ll_load_raw(mybuf, comm_entry(parity, q0 + q, c.nranks, B, j));
                        if ((uint32_t)(w[q].x >> 32) == ep && (uint32_t)(w[q].y >> 32) == ep) missing &= ~(1u << q);
                    }
                if (missing && (++spins & 0xffu) == 0 && global_timer_ns() - t0 > c.timeout_ns) return false;
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q0 + q < c.nranks) s += __longlong_as_double((long long)((w[q].x & 0xffffffffull) | (w[q].y << 32)));
    }
    out = s;
    return true;
}

template <typename AT>
__global__ void __launch_bounds__(256, 4) k3_reduce_exchange(const K3Args<AT> a, const CommArgs c) {
    __shared__ double sh[256];
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    const int lo = a.beam_ptr[blockIdx.x], hi = a.beam_ptr[blockIdx.x + 1];
    pdl_wait();
    unsigned long long* const trace = a.trace ? a.trace + (size_t)blockIdx.x * kTraceSlots : nullptr;
    if (trace && tid == 0) trace[0] = global_timer_ns();
    const int parity = (int)(c.epoch & 1ull);
    const uint32_t ep = (uint32_t)c.epoch;
    // ---- A: this rank's partial for my beamlets (and the objective), stored into every rank's buffer
    if (blockIdx.x == 0) {
        double fv = strided_sum(a.fblock, tid, 256, a.nfblock);
        fv = block_sum_256(fv, sh);
        if (tid < c.nranks) ll_store(c.peer_bufs[tid], comm_entry(parity, c.rank, c.nranks, a.B, (size_t)a.B), fv, ep);
    }
    for (int j0 = lo; j0 < hi; j0 += 64) {
        const int j = j0 + (tid >> 2), q = tid & 3;
        double s = 0.0;
        if (j < hi) s = strided_sum(a.partial + j, (int64_t)q * a.B, (int64_t)4 * a.B, (int64_t)a.ntiles * a.B);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);       // all four lanes hold the sum: lane q serves ranks q, q+4, ...
        if (j < hi) for (int p = q; p < c.nranks; p += 4) ll_store(c.peer_bufs[p], comm_entry(parity, c.rank, c.nranks, a.B, (size_t)j), s, ep);
    }
    if (trace) { __syncthreads(); if (tid == 0) trace[1] = global_timer_ns(); }
    // ---- B/C: add the ranks' partials in rank order as they arrive in this rank's own buffer
    double contrib = 0.0;
    int ok = 1;
    for (int j = lo + tid; j < hi; j += 256) {
        double s = 0.0;
        if (!comm_sum_ranks(c, a.B, (size_t)j, s)) { ok = 0; break; }
        a.gbf[j] = s;
        contrib += a.w_grad[j] * s;
    }
    double f = 0.0;
    if (blockIdx.x == 0 && tid == 0 && ok) ok = comm_sum_ranks(c, a.B, (size_t)a.B, f) ? 1 : 0;
    ok = __syncthreads_and(ok);
    if (trace && tid == 0) trace[2] = global_timer_ns();
    const double g = block_sum_256(contrib, sh);
    if (tid == 0) {
        if (trace) trace[3] = global_timer_ns();
        a.result[1 + blockIdx.x] = ok ? g : nan("");
        if (blockIdx.x == 0) { a.gbf[a.B] = ok ? f : nan(""); a.result[0] = ok ? f : nan(""); }
        if (!ok) { atomicExch(c.error, 1); *c.status = 1.0; }
    }
}

// ---------------------------------------------------------------------------------
// layout construction (plan build time)
// ---------------------------------------------------------------------------------
// CSR payload check: every index in [0, ncols); with `sorted`, non-decreasing inside a row (the
// tile segmentation below bisects beamlet rows).  One thread per row; flag bit 0 = range, bit 1 = order.
__global__ void k_check_csr(int64_t nrows, int64_t ncols, int sorted, const int64_t* __restrict__ rowptr,
                            const int32_t* __restrict__ col, int* __restrict__ flag) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    int bad = 0;
    int32_t prev = 0;
    for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) {
        const int32_t c = col[k];
        if (c < 0 || c >= ncols) bad |= 1;
        if (sorted && c < prev) bad |= 2;
        prev = c;
    }
    if (bad) atomicOr(flag, bad);
}

template <typename T> struct StoreCvt;
template <> struct StoreCvt<float>  { template <typename S> __device__ static float  cvt(S v) { return (float)v; } };
template <> struct StoreCvt<double> { template <typename S> __device__ static double cvt(S v) { return (double)v; } };
template <> struct StoreCvt<__nv_bfloat16> {
    template <typename S> __device__ static __nv_bfloat16 cvt(S v) { return __float2bfloat16_rn((float)v); }
};
template <typename VT> __device__ __forceinline__ bool stored_zero(VT v) { return (float)v == 0.0f; }
template <> __device__ __forceinline__ bool stored_zero<double>(double v) { return v == 0.0; }

// byte-delta layout: a nonzero `gap` (>= 0) away from its predecessor is reached by entries of value 0 whose
// delta counts units of 256 (at most 255 of them each), then its own delta gap & 255
__host__ __device__ __forceinline__ int32_t bridge_entries(int32_t gap) { return ((gap >> 8) + 254) / 255; }

// byte-delta layout: bridge entries a segment needs.  Entries whose STORED value is zero stay where the running
// index is (delta 0; the decoder would scale their delta), so gaps are measured between the entries that remain
// nonzero after conversion to the storage type.  Segment k spans CSR positions bounds[k] .. bounds[k + 1].
template <typename VT, typename ST>
__global__ void k_count_pads(int64_t nseg, const int64_t* __restrict__ bounds, const int32_t* __restrict__ col,
                             const ST* __restrict__ val, int32_t* __restrict__ pads) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nseg) return;
    const int64_t p0 = bounds[k], p1 = bounds[k + 1];
    int32_t n = 0, idx = p0 < p1 ? col[p0] : 0;
    for (int64_t p = p0 + 1; p < p1; ++p) {
        if (stored_zero<VT>(StoreCvt<VT>::cvt(val[p]))) continue;
        n += bridge_entries(col[p] - idx);
        idx = col[p];
    }
    pads[k] = n;
}

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


// one warp per slice s = ss*W + w: writes the warp's part of the super-slice header (row ids and,
// when `tables` is set, thr/over/under of elem size `es`) and copies the lane's segment (CSR
// positions seg_start .. +seg_len) into the chunked stages; padding = (index 0, value 0)
template <typename VT, typename ST, int IW>
__global__ void k_fill_stream(int32_t nslices, int32_t W, int32_t HU, const int64_t* __restrict__ ss_off,
                              const int64_t* __restrict__ seg_start, const int32_t* __restrict__ seg_len,
                              const int32_t* __restrict__ seg_colbase, const uint8_t* __restrict__ ss_tpr,
                              const int32_t* __restrict__ rows,
                              const unsigned char* __restrict__ thr, const unsigned char* __restrict__ over,
                              const unsigned char* __restrict__ under, int32_t es,
                              const int32_t* __restrict__ col, const ST* __restrict__ val,
                              uint4* __restrict__ sell, const uint8_t* __restrict__ ss_nch) {
    // ss_nch != nullptr: ragged super-slices (fused kernel) - warp w has ss_nch[ss*W + w] chunks
    // (non-increasing in w) and stage c holds chunk c of the warps that still have one
    constexpr int KU = chunk_units<VT, IW>(), IU = idx_units(IW);
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (gw >= nslices) return;
    const int64_t ss = gw / W;
    const int w = (int)(gw % W);
    const int64_t off = ss_off[ss];
    const int nch = ss_nch ? (int)ss_nch[ss * W + w] : (int)((ss_off[ss + 1] - off - (int64_t)W * HU) / ((int64_t)W * KU));
    const int64_t slot = gw * 32 + lane;
    unsigned char* hp = reinterpret_cast<unsigned char*>(sell + off + (int64_t)w * HU);
    reinterpret_cast<int32_t*>(hp)[lane] = rows[slot];
    if (thr) {
        for (int k = 0; k < es; ++k) {
            hp[128 + lane * es + k] = thr[slot * es + k];
            hp[128 + 32 * es + lane * es + k] = over[slot * es + k];
            hp[128 + 64 * es + lane * es + k] = under[slot * es + k];
        }
    }
    const int64_t st = seg_start[slot];
    const int32_t len = seg_len[slot], cb = seg_colbase[slot];
    const int tpr = ss_tpr[ss], part = lane % tpr;
    uint4* data = sell + off + (int64_t)W * HU;
    if constexpr (IW == 1) {
        // Byte-delta indices.  The row is first expanded: in front of a nonzero that is 256 or more past its
        // predecessor go bridge entries (value 0, delta in units of 256, k_count_pads sized the row for them);
        // entries whose stored value is zero keep the running index (delta 0).  Lane `part` of the row then
        // owns the expanded entries [part*4*nch, (part+1)*4*nch): a contiguous block; the running index at its
        // first entry goes into the header, that entry's delta is 0.  Entries past the row's end: delta 0, value 0.
        const int64_t first = (int64_t)part * 4 * nch, last = first + (int64_t)4 * nch;
        int32_t e = 0;                               // next source element
        int32_t jump_left = 0;                       // units of 256 still to bridge before element e
        int32_t idx = 0;                             // running index after the entry emitted last
        uint32_t base = 0, word = 0;
        VT vv[4];
        for (int64_t s = 0; s < last; ++s) {
            VT v = StoreCvt<VT>::cvt(0.0f);
            uint32_t dbyte = 0;
            if (e < len) {
                if (jump_left > 0) {
                    const int32_t piece = jump_left < 255 ? jump_left : 255;
                    dbyte = (uint32_t)piece; idx += piece << 8; jump_left -= piece;
                } else {
                    const VT sv = StoreCvt<VT>::cvt(val[st + e]);
                    const int32_t c = col[st + e] - cb;
                    if (e == 0) idx = c;
                    if (!stored_zero<VT>(sv)) { dbyte = (uint32_t)(c - idx); idx = c; v = sv; }
                    ++e;
                    if (e < len && !stored_zero<VT>(StoreCvt<VT>::cvt(val[st + e])))
                        jump_left = ((col[st + e] - cb) - idx) >> 8;
                }
            }
            if (s < first) continue;
            const int k = (int)((s - first) & 3);
            if (s == first) { base = (uint32_t)idx; dbyte = 0; }
            if (k == 0) word = 0;
            word |= dbyte << (8 * k);
            vv[k] = v;
            if (k == 3) {
                uint4* chunk = data + ((s - first) >> 2) * W * KU + (int64_t)w * KU;
                reinterpret_cast<uint32_t*>(chunk)[lane] = word;
                if constexpr (sizeof(VT) == 4) {
                    float4 f = make_float4(vv[0], vv[1], vv[2], vv[3]);
                    chunk[IU + lane] = *reinterpret_cast<uint4*>(&f);
                } else if constexpr (sizeof(VT) == 2) {
                    __nv_bfloat16 h[4] = {vv[0], vv[1], vv[2], vv[3]};
                    reinterpret_cast<uint2*>(chunk + IU)[lane] = *reinterpret_cast<uint2*>(h);
                } else {
                    double2 a2 = make_double2(vv[0], vv[1]), b2 = make_double2(vv[2], vv[3]);
                    chunk[IU + lane] = *reinterpret_cast<uint4*>(&a2);
                    chunk[IU + 32 + lane] = *reinterpret_cast<uint4*>(&b2);
                }
            }
        }
        reinterpret_cast<uint32_t*>(hp + (HU - 8) * 16)[lane] = base;
    } else {
    // absolute indices: the row's nonzeros go round-robin (4 at a time) over its lanes
    int64_t stage_off = 0;               // ragged: start of stage c = KU * sum over warps of min(nch_w, c)
    for (int c = 0; c < nch; ++c) {
        uint4* chunk = ss_nch ? data + stage_off + (int64_t)w * KU : data + ((int64_t)c * W + w) * KU;
        if (ss_nch) {
            int kact = 0;
            for (int q = 0; q < W; ++q) kact += ss_nch[ss * W + q] > c ? 1 : 0;
            stage_off += (int64_t)kact * KU;
        }
        int32_t ix[4];
        VT vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = (c * tpr + part) * 4 + k;
            if (e < len) { ix[k] = col[st + e] - cb; vv[k] = StoreCvt<VT>::cvt(val[st + e]); }
            else         { ix[k] = 0;                vv[k] = StoreCvt<VT>::cvt(0.0f); }
        }
        if constexpr (IW == 4) {
            chunk[lane] = make_uint4((uint32_t)ix[0], (uint32_t)ix[1], (uint32_t)ix[2], (uint32_t)ix[3]);
        } else {
            reinterpret_cast<uint2*>(chunk)[lane] = make_uint2((uint32_t)ix[0] | ((uint32_t)ix[1] << 16),
                                                              (uint32_t)ix[2] | ((uint32_t)ix[3] << 16));
        }
        if constexpr (sizeof(VT) == 4) {
            float4 f = make_float4(vv[0], vv[1], vv[2], vv[3]);
            chunk[IU + lane] = *reinterpret_cast<uint4*>(&f);
        } else if constexpr (sizeof(VT) == 2) {
            __nv_bfloat16 h[4] = {vv[0], vv[1], vv[2], vv[3]};
            reinterpret_cast<uint2*>(chunk + IU)[lane] = *reinterpret_cast<uint2*>(h);
        } else {
            double2 a2 = make_double2(vv[0], vv[1]), b2 = make_double2(vv[2], vv[3]);
            chunk[IU + lane] = *reinterpret_cast<uint4*>(&a2);
            chunk[IU + 32 + lane] = *reinterpret_cast<uint4*>(&b2);
        }
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
