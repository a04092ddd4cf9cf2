// radix_sort.cu -- K4: RadixSortSimple.comp:52-158 re-designed for a whole GPU.
//
// The reference sorts the MortonPrimitive array by `code` with a stable LSD radix sort, 4 passes x 8 bits, inside ONE
// 256-thread workgroup (RaytracerBVH.cpp:916), i.e. serially on one SM.  Same algorithm, same pass structure, same
// (stable) result -- but every pass is spread over the grid:
//     digit histogram per 4096-element tile  ->  per-digit exclusive scan over tiles (one block per digit)  ->  stable scatter.
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

// Exclusive scan of the [digit][tile] table, one block per digit row (rows are contiguous): row d becomes the exclusive
// prefix over tiles and its total goes to totals[d].  The scatter kernel adds the exclusive prefix over digits itself.
__global__ void __launch_bounds__(256) sort_scan_kernel(uint32_t* counts, uint32_t numTiles, uint32_t* totals) {
    __shared__ uint32_t warpSums[8];
    __shared__ uint32_t carry;
    uint32_t* row = counts + (size_t)blockIdx.x * numTiles;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < numTiles; base += 256) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t c = i < numTiles ? row[i] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        uint32_t before = carry;
        for (uint32_t w = 0; w < warp; w++) before += warpSums[w];
        if (i < numTiles) row[i] = before + incl - c;
        __syncthreads();
        if (threadIdx.x == 255) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                    uint32_t* keysOut, uint32_t* valsOut, uint32_t n, uint32_t shift,
                                                                    uint32_t numTiles, const uint32_t* __restrict__ offsets,
                                                                    const uint32_t* __restrict__ totals) {
    __shared__ uint32_t warpCnt[SORT_THREADS / 32][256];
    __shared__ uint32_t binBase[256];
    __shared__ uint32_t wsum[8];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // global start of bin d = (exclusive prefix of the digit totals) + (prefix of this digit over the earlier tiles)
        const uint32_t t = totals[threadIdx.x];
        uint32_t incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; w++) before += wsum[w];
        binBase[threadIdx.x] = before + incl - t + offsets[threadIdx.x * numTiles + blockIdx.x];
    }
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
        sort_scan_kernel<<<256, 256, 0, st>>>(counts, numTiles, counts + 256ull * numTiles);
        sort_scatter_kernel<<<numTiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, n, shift, numTiles, counts, counts + 256ull * numTiles);
        uint32_t* t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
    }
    return 12;
}
size_t radix_sort_counts_bytes(uint32_t n) {
    const uint32_t numTiles = (n + SORT_TILE - 1) / SORT_TILE;
    return sizeof(uint32_t) * (256ull * (numTiles ? numTiles : 1) + 256ull);   // [digit][tile] table + 256 digit totals
}

}  // namespace rtb
