// Microbenchmark: what read bandwidth do different access patterns reach on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o streambench streambench.cu && ./streambench
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldg16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned mix(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }

// 1. classic grid-stride streaming read
template <int U>
__global__ void k_gridstride(const uint4* __restrict__ p, size_t n, unsigned* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg16(p + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) acc ^= mix(v[u]);
    }
    if (acc == 0x12345678u) out[0] = acc;
}

// 2. warp-private streams: each warp pulls a slice (slice_units 16B-units) and reads it with two register
//    stages of U x 1 KB (two 512 B pieces per KB); optional smem gather per word
template <int U, bool GATHER>
__global__ void k_warpstream(const uint4* __restrict__ p, int nslices, int slice_units, unsigned* counter, unsigned* out, int table) {
    extern __shared__ float xs[];
    for (int j = threadIdx.x; j < table; j += blockDim.x) xs[j] = (float)j;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    unsigned s = 0;
    float facc = 0.f;
    unsigned acc = 0;
    if (lane == 0) s = atomicAdd(counter, 1u);
    s = __shfl_sync(0xffffffffu, s, 0);
    const int nch = slice_units / 64;
    while (s < (unsigned)nslices) {
        unsigned sn = 0;
        if (lane == 0) sn = atomicAdd(counter, 1u);
        const uint4* q = p + (size_t)s * slice_units;
        uint4 A[2 * U], B[2 * U];
        const int ngrp = nch / U;
#pragma unroll
        for (int u = 0; u < U; ++u) { A[2 * u] = ldg16(q + u * 64 + lane); A[2 * u + 1] = ldg16(q + u * 64 + 32 + lane); }
        for (int it = 0; it + 2 <= ngrp; it += 2) {
#pragma unroll
            for (int u = 0; u < U; ++u) { B[2 * u] = ldg16(q + (U + u) * 64 + lane); B[2 * u + 1] = ldg16(q + (U + u) * 64 + 32 + lane); }
#pragma unroll
            for (int u = 0; u < 2 * U; ++u) {
                if (GATHER) { facc += xs[A[u].x % table] + xs[A[u].y % table] + xs[A[u].z % table] + xs[A[u].w % table]; }
                else acc ^= mix(A[u]);
            }
            if (it + 2 < ngrp) {
#pragma unroll
                for (int u = 0; u < U; ++u) { A[2 * u] = ldg16(q + (2 * U + u) * 64 + lane); A[2 * u + 1] = ldg16(q + (2 * U + u) * 64 + 32 + lane); }
            }
#pragma unroll
            for (int u = 0; u < 2 * U; ++u) {
                if (GATHER) { facc += xs[B[u].x % table] + xs[B[u].y % table] + xs[B[u].z % table] + xs[B[u].w % table]; }
                else acc ^= mix(B[u]);
            }
            q += 2 * U * 64;
        }
        s = __shfl_sync(0xffffffffu, sn, 0);
    }
    if (acc == 0x12345678u || facc == 1.2345f) out[0] = acc;
}

// 3. CTA-contiguous: a CTA pulls a block of `blk_units` and its warps read adjacent KBs (warp w reads KB w, w+nw, ...)
template <int U>
__global__ void k_ctastream(const uint4* __restrict__ p, int nblocks, int blk_units, unsigned* counter, unsigned* out) {
    __shared__ unsigned sb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned acc = 0;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) sb = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned b = sb;
        if (b >= (unsigned)nblocks) break;
        const uint4* q = p + (size_t)b * blk_units;
        const int nkb = blk_units / 64;
        for (int k = warp * U; k < nkb; k += nw * U) {
            uint4 v[2 * U];
#pragma unroll
            for (int u = 0; u < U; ++u) { v[2 * u] = ldg16(q + (k + u) * 64 + lane); v[2 * u + 1] = ldg16(q + (k + u) * 64 + 32 + lane); }
#pragma unroll
            for (int u = 0; u < 2 * U; ++u) acc ^= mix(v[u]);
        }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
static float timeit(F f, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main() {
    const size_t bytes = (size_t)240 << 20;
    const size_t n16 = bytes / 16;
    uint4* buf; cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    uint4* buf2; cudaMalloc(&buf2, bytes); cudaMemset(buf2, 2, bytes);   // alternated so L2 never helps
    unsigned *counter, *out; cudaMalloc(&counter, 4); cudaMalloc(&out, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int reps = 20;
    int flip = 0;
    auto src = [&]() { flip ^= 1; return flip ? buf : buf2; };
    for (int tpb : {256, 512, 1024}) for (int occ : {1, 2, 4, 8}) {
        if (tpb * occ > 2048) continue;
        float ms = timeit([&] { k_gridstride<4><<<sms * occ, tpb>>>(src(), n16, out); }, reps);
        printf("gridstride U4 tpb %4d ctas/sm %d : %.1f GB/s\n", tpb, occ, bytes / ms / 1e6);
    }
    for (int tpb : {256, 512}) {
        float ms = timeit([&] { k_gridstride<8><<<sms * (2048 / tpb), tpb>>>(src(), n16, out); }, reps);
        printf("gridstride U8 tpb %4d full occ : %.1f GB/s\n", tpb, bytes / ms / 1e6);
    }
    cudaFuncSetAttribute(k_warpstream<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int slice_kb : {8, 24, 96}) for (int tpb : {256, 512, 1024}) {
        const int slice_units = slice_kb * 64, nsl = (int)(n16 / slice_units);
        float ms = timeit([&] { cudaMemsetAsync(counter, 0, 4); k_warpstream<4, false><<<sms, tpb, 16>>>(src(), nsl, slice_units, counter, out, 1); }, reps);
        printf("warpstream U4x2 slice %3d KB tpb %4d 1cta/sm : %.1f GB/s\n", slice_kb, tpb, (double)nsl * slice_units * 16 / ms / 1e6);
        if (tpb <= 512) {
            ms = timeit([&] { cudaMemsetAsync(counter, 0, 4); k_warpstream<4, false><<<sms * 2, tpb, 16>>>(src(), nsl, slice_units, counter, out, 1); }, reps);
            printf("warpstream U4x2 slice %3d KB tpb %4d 2cta/sm : %.1f GB/s\n", slice_kb, tpb, (double)nsl * slice_units * 16 / ms / 1e6);
        }
        ms = timeit([&] { cudaMemsetAsync(counter, 0, 4); k_warpstream<4, true><<<sms, tpb, 40 * 1024>>>(src(), nsl, slice_units, counter, out, 10080); }, reps);
        printf("warpstream+gather  slice %3d KB tpb %4d 1cta/sm : %.1f GB/s\n", slice_kb, tpb, (double)nsl * slice_units * 16 / ms / 1e6);
    }
    for (int blk_kb : {32, 128}) for (int tpb : {256, 512, 1024}) {
        const int blk_units = blk_kb * 64, nb = (int)(n16 / blk_units);
        float ms = timeit([&] { cudaMemsetAsync(counter, 0, 4); k_ctastream<4><<<sms * (2048 / tpb), tpb>>>(src(), nb, blk_units, counter, out); }, reps);
        printf("ctastream U4 block %3d KB tpb %4d full occ : %.1f GB/s\n", blk_kb, tpb, (double)nb * blk_units * 16 / ms / 1e6);
    }
    return 0;
}
