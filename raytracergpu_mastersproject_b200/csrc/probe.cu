// probe.cu -- rtb_probe_gather: a live micro-benchmark of the fetch pattern the trace kernels are made of, for bench.py's
// "node-fetch roofline".  Every lane of a persistent grid (the trace kernel's shape: 128-thread CTAs, 7 per SM) fetches 64-byte
// records (two 256-bit loads, as wave_step_u does) at pseudo-random indices of a buffer of the given footprint, several
// INDEPENDENT fetches in flight per lane.  The result is the rate at which this GPU can deliver divergent 64-byte record fetches
// at that working-set size (an L2 / HBM mix that depends on the footprint): an upper bound for a traversal, whose fetches are
// additionally dependent on each other.
#include "common.cuh"
#include "ctx.h"

using namespace rtb;

namespace {

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(128, 7) gather_probe_kernel(const uint4* __restrict__ buf, uint32_t records, uint32_t iters, uint32_t salt,
                                                              unsigned int* sink) {
    uint32_t s = mix((blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + salt);
    float acc = 0.f;
    for (uint32_t i = 0; i < iters; i++) {
        uint32_t idx[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { s = mix(s + 0x9e3779b9u); idx[k] = (uint32_t)(((uint64_t)s * records) >> 32); }
        f8 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { a[k] = ldg256(buf + 4ull * idx[k]); b[k] = ldg256(buf + 4ull * idx[k] + 2); }
#pragma unroll
        for (int k = 0; k < 4; k++) acc += a[k].lo.x + a[k].hi.w + b[k].lo.y + b[k].hi.z;
    }
    if (acc == 123.456f) atomicAdd(sink, 1u);     // keeps the loads alive
}

}  // namespace

extern "C" int rtb_probe_gather(rtb_ctx* c, size_t footprintBytes, float* gbPerSecond) {
    REQUIRE(c && gbPerSecond, "rtb_probe_gather: bad argument");
    Activate act(c);
    if (footprintBytes < (1u << 20)) footprintBytes = 1u << 20;
    const uint32_t records = (uint32_t)(footprintBytes / 64 > 0xFFFFFFFFull ? 0xFFFFFFFFull : footprintBytes / 64);
    void* buf = nullptr;
    CK(cudaMalloc(&buf, (size_t)records * 64));
    CK(cudaMemsetAsync(buf, 0, (size_t)records * 64, c->stream));
    if (ensure(c, c->errFlag, 16)) { cudaFree(buf); return 1; }
    const unsigned grid = (unsigned)c->smCount * 7u;
    const uint32_t iters = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 0.f;
    for (int rep = 0; rep < 4; rep++) {            // rep 0 warms up; best of the others
        cudaEventRecord(e0, c->stream);
        gather_probe_kernel<<<grid, 128, 0, c->stream>>>((const uint4*)buf, records, iters, 17u * rep, (unsigned int*)c->errFlag.p + 2);
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const float gbs = (float)((double)grid * 128.0 * iters * 4.0 * 64.0 / (ms * 1e-3) / 1e9);
        if (rep > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf);
    c->launches += 4;
    *gbPerSecond = best;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail("gather_probe_kernel", e);
    return 0;
}
