// radix_sort.cu -- K4: RadixSortSimple.comp:52-158 re-designed for a whole GPU.
//
// The reference sorts the MortonPrimitive array by `code` with a stable LSD radix sort, 4 passes x 8 bits, inside ONE
// 256-thread workgroup (RaytracerBVH.cpp:916), i.e. serially on one SM.  Same algorithm, same pass structure, same
// (stable) result -- but every pass is spread over the grid:
//     digit histogram per 4096-element tile  ->  exclusive scan of the [digit][tile] table  ->  stable scatter.
// Keys travel as SoA (code, global primitive id); the 12-byte records are (un)packed by bvh_build.cu.
#include "common.cuh"
#include "kernels.h"

namespace rtb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                                 uint32_t numTiles, uint32_t* counts) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t i = base + k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[threadIdx.x * numTiles + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of `total` counters in place, one block (total = 256 * numTiles, a few MB at most)
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t* counts, uint32_t total) {
    __shared__ uint32_t warpSums[32];
    const uint32_t chunk = (total + 1023u) / 1024u;
    const uint32_t begin = min(threadIdx.x * chunk, total), end = min(begin + chunk, total);
    uint32_t sum = 0;
    for (uint32_t i = begin; i < end; i++) sum += counts[i];
    // block exclusive scan of the 1024 partial sums
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += v;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warpSums[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += v;
        }
        warpSums[lane] = wi - w;
    }
    __syncthreads();
    uint32_t run = warpSums[warp] + incl - sum;
    for (uint32_t i = begin; i < end; i++) { const uint32_t c = counts[i]; counts[i] = run; run += c; }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                    uint32_t* keysOut, uint32_t* valsOut, uint32_t n, uint32_t shift,
                                                                    uint32_t numTiles, const uint32_t* __restrict__ offsets) {
    __shared__ uint32_t warpCnt[SORT_THREADS / 32][256];
    __shared__ uint32_t binBase[256];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    binBase[threadIdx.x] = offsets[threadIdx.x * numTiles + blockIdx.x];
    const uint32_t base = blockIdx.x * SORT_TILE;
    for (int k = 0; k < SORT_ITEMS; k++) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; w++) warpCnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + k * SORT_THREADS + threadIdx.x;
        const bool valid = i < n;
        uint32_t key = 0, val = 0, digit = 0xFFFFFFFFu;
        if (valid) { key = keysIn[i]; val = valsIn[i]; digit = (key >> shift) & 255u; }
        // stable rank inside the warp: lanes with the same digit, lower lanes first (RadixSortSimple.comp:133-145)
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warpCnt[warp][digit] = __popc(peers);
        __syncthreads();
        {   // per digit: exclusive prefix over the warps (warp order = element order), then advance the bin base
            uint32_t run = binBase[threadIdx.x];
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; w++) { const uint32_t c = warpCnt[w][threadIdx.x]; warpCnt[w][threadIdx.x] = run; run += c; }
            binBase[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const uint32_t dst = warpCnt[warp][digit] + rank;
            keysOut[dst] = key;
            valsOut[dst] = val;
        }
        __syncthreads();
    }
}

// 4 passes; keys/vals[0] hold the input and (after an even number of passes) the result.  Returns #launches.
int launch_radix_sort(cudaStream_t st, uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, uint32_t n, uint32_t* counts) {
    if (n == 0) return 0;
    const uint32_t numTiles = (n + SORT_TILE - 1) / SORT_TILE;
    uint32_t *kin = keys0, *vin = vals0, *kout = keys1, *vout = vals1;
    for (uint32_t pass = 0; pass < 4; pass++) {       // ITERATIONS 4 x BITS_PER_ITERATION 8 (RadixSortSimple.comp:11-12)
        const uint32_t shift = 8u * pass;
        sort_hist_kernel<<<numTiles, SORT_THREADS, 0, st>>>(kin, n, shift, numTiles, counts);
        sort_scan_kernel<<<1, 1024, 0, st>>>(counts, 256u * numTiles);
        sort_scatter_kernel<<<numTiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, numTiles, counts);
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    return 12;
}
size_t radix_sort_counts_bytes(uint32_t n) {
    const uint32_t numTiles = (n + SORT_TILE - 1) / SORT_TILE;
    return sizeof(uint32_t) * 256ull * (numTiles ? numTiles : 1);
}

}  // namespace rtb
