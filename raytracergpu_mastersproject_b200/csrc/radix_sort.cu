// radix_sort.cu -- K4: RadixSortSimple.comp:52-158 re-designed for a whole GPU.
//
// The reference sorts the MortonPrimitive array by `code` with a stable LSD radix sort, 4 passes x 8 bits, inside ONE
// 256-thread workgroup (RaytracerBVH.cpp:916), i.e. serially on one SM.  Same algorithm, same pass structure, same
// (stable) result -- but every pass is spread over the grid:
//     digit histogram per 4096-element tile  ->  per-digit exclusive scan over tiles (one block per digit)  ->  stable scatter
//     (warp-private ranking, tile staged in digit order in shared memory, coalesced runs out).
// Keys travel as SoA (code, global primitive id); the 12-byte records are (un)packed by bvh_build.cu.
#include "common.cuh"
#include "kernels.h"

namespace rtb {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

// Element order inside a tile: warp w of the block owns the 512 consecutive elements [w * 512, (w + 1) * 512) and reads them as 16
// coalesced rows of 32 (element = w * 512 + k * 32 + lane), so ascending element index = (warp, row, lane) in lexicographic order.
__device__ __forceinline__ uint32_t tile_element(uint32_t base, uint32_t warp, int k, uint32_t lane) {
    return base + warp * (SORT_ITEMS * 32u) + (uint32_t)k * 32u + lane;
}

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t shift,
                                                                 uint32_t numTiles, uint32_t* counts) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 4
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t i = tile_element(base, warp, k, lane);
        const uint32_t digit = i < n ? (keys[i] >> shift) & 255u : 0xFFFFFFFFu;
        // one shared-memory atomic per distinct digit of the row (sorted-ish input: the upper digits of a row are nearly all equal)
        const uint32_t peers = __match_any_sync(0xFFFFFFFFu, digit);
        if (digit != 0xFFFFFFFFu && lane == (uint32_t)__ffs(peers) - 1u) atomicAdd(&h[digit], (uint32_t)__popc(peers));
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

// Stable scatter of one tile.  Ranking needs no block-wide barrier per row: every warp keeps a private running count per digit
// (warpCnt[warp][digit]) for its own 512 consecutive elements -- the lanes of a row that share a digit are ranked with
// __match_any_sync (lower lanes first, RadixSortSimple.comp:133-145), the first of them advances the warp's counter.  The tile is then
// put in digit order in shared memory and written out in runs: consecutive threads write consecutive addresses of a bin (the old
// kernel wrote every element straight to its bin: 32 separate 4-byte sectors per warp store).
__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                                    uint32_t* keysOut, uint32_t* valsOut, uint32_t n, uint32_t shift,
                                                                    uint32_t numTiles, const uint32_t* __restrict__ offsets,
                                                                    const uint32_t* __restrict__ totals) {
    __shared__ uint32_t warpCnt[SORT_THREADS / 32][256];   // running count, then exclusive prefix over the warps, per digit
    __shared__ uint32_t binBase[256];                      // global start of this tile's run of digit d
    __shared__ uint32_t digitStart[256];                   // start of digit d inside the tile
    __shared__ uint32_t wsum[8];
    __shared__ uint32_t sKeys[SORT_TILE], sVals[SORT_TILE];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; w++) warpCnt[w][threadIdx.x] = 0;
    uint32_t globalBase;
    {   // global start of bin d = (exclusive prefix of the digit totals) + (prefix of this digit over the earlier tiles)
        const uint32_t t = totals[threadIdx.x];
        uint32_t incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();                                   // (also: warpCnt is zero for every warp)
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; w++) before += wsum[w];
        globalBase = before + incl - t + offsets[threadIdx.x * numTiles + blockIdx.x];
    }
    // ---- rank: 16 rows per warp, warp-private counters, no block barrier ----
    uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t i = tile_element(base, warp, k, lane);
        const bool valid = i < n;
        key[k] = valid ? keysIn[i] : 0u;
        const uint32_t digit = valid ? (key[k] >> shift) & 255u : 0xFFFFFFFFu;
        const uint32_t peers = __match_any_sync(FULL, digit);
        const int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && (int)lane == leader) { old = warpCnt[warp][digit]; warpCnt[warp][digit] = old + (uint32_t)__popc(peers); }
        __syncwarp();
        old = __shfl_sync(FULL, old, leader);
        rank[k] = old + (uint32_t)__popc(peers & ((1u << lane) - 1u));      // rank among the warp's elements of this digit
    }
    __syncthreads();
    {   // per digit: exclusive prefix over the warps (warp order = element order), tile total -> start of the digit inside the tile
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; w++) { const uint32_t c = warpCnt[w][threadIdx.x]; warpCnt[w][threadIdx.x] = run; run += c; }
        uint32_t incl = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(FULL, incl, o);
            if (lane >= (uint32_t)o) incl += v;
        }
        if (lane == 31) wsum[warp] = incl;                 // (its earlier readers are behind the barrier that ended the ranking)
        __syncthreads();
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; w++) before += wsum[w];
        digitStart[threadIdx.x] = before + incl - run;
        binBase[threadIdx.x] = globalBase;
    }
    __syncthreads();
    // ---- the tile in digit order in shared memory ----
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; k++) {
        const uint32_t i = tile_element(base, warp, k, lane);
        if (i < n) {
            const uint32_t digit = (key[k] >> shift) & 255u;
            const uint32_t p = digitStart[digit] + warpCnt[warp][digit] + rank[k];
            sKeys[p] = key[k];
            sVals[p] = valsIn[i];
        }
    }
    __syncthreads();
    // ---- runs out: position p of the sorted tile goes to binBase[digit] + (p - digitStart[digit]) ----
    const uint32_t count = n - base < (uint32_t)SORT_TILE ? n - base : (uint32_t)SORT_TILE;
#pragma unroll 4
    for (uint32_t p = threadIdx.x; p < count; p += SORT_THREADS) {
        const uint32_t k2 = sKeys[p];
        const uint32_t digit = (k2 >> shift) & 255u;
        const uint32_t dst = binBase[digit] + (p - digitStart[digit]);
        keysOut[dst] = k2;
        valsOut[dst] = sVals[p];
    }
}

// Small arrays (<= 2048 elements: complexScene-sized scenes, where the build is a chain of launch latencies): all four passes in ONE
// block, ping-ponging between two shared-memory copies -- 1 launch instead of 12.  Same stable ranking as sort_scatter_kernel (element
// = warp * 256 + row * 32 + lane), hence the same result.
constexpr int SORT_SMALL_N = 2048;
constexpr int SORT_SMALL_ROWS = SORT_SMALL_N / SORT_THREADS;

__global__ void __launch_bounds__(SORT_THREADS) sort_small_kernel(uint32_t* keys, uint32_t* vals, uint32_t n) {
    __shared__ uint32_t sk[2][SORT_SMALL_N], sv[2][SORT_SMALL_N];
    __shared__ uint32_t warpCnt[SORT_THREADS / 32][256];
    __shared__ uint32_t digitStart[256];
    __shared__ uint32_t wsum[8];
    const uint32_t FULL = 0xFFFFFFFFu;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) { sk[0][i] = keys[i]; sv[0][i] = vals[i]; }
    __syncthreads();
    for (uint32_t pass = 0; pass < 4; pass++) {
        const uint32_t src = pass & 1u, dst = src ^ 1u, shift = 8u * pass;
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; w++) warpCnt[w][threadIdx.x] = 0;
        __syncthreads();
        uint32_t key[SORT_SMALL_ROWS], rank[SORT_SMALL_ROWS];
#pragma unroll
        for (int k = 0; k < SORT_SMALL_ROWS; k++) {
            const uint32_t i = warp * (SORT_SMALL_ROWS * 32u) + (uint32_t)k * 32u + lane;
            const bool valid = i < n;
            key[k] = valid ? sk[src][i] : 0u;
            const uint32_t digit = valid ? (key[k] >> shift) & 255u : 0xFFFFFFFFu;
            const uint32_t peers = __match_any_sync(FULL, digit);
            const int leader = __ffs(peers) - 1;
            uint32_t old = 0;
            if (valid && (int)lane == leader) { old = warpCnt[warp][digit]; warpCnt[warp][digit] = old + (uint32_t)__popc(peers); }
            __syncwarp();
            old = __shfl_sync(FULL, old, leader);
            rank[k] = old + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; w++) { const uint32_t c = warpCnt[w][threadIdx.x]; warpCnt[w][threadIdx.x] = run; run += c; }
            uint32_t incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(FULL, incl, o);
                if (lane >= (uint32_t)o) incl += v;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (uint32_t w = 0; w < warp; w++) before += wsum[w];
            digitStart[threadIdx.x] = before + incl - run;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SORT_SMALL_ROWS; k++) {
            const uint32_t i = warp * (SORT_SMALL_ROWS * 32u) + (uint32_t)k * 32u + lane;
            if (i < n) {
                const uint32_t digit = (key[k] >> shift) & 255u;
                const uint32_t p = digitStart[digit] + warpCnt[warp][digit] + rank[k];
                sk[dst][p] = key[k];
                sv[dst][p] = sv[src][i];
            }
        }
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n; i += SORT_THREADS) { keys[i] = sk[0][i]; vals[i] = sv[0][i]; }
}

// 4 passes; keys/vals[0] hold the input and (after an even number of passes) the result.  Returns #launches.
int launch_radix_sort(cudaStream_t st, uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, uint32_t n, uint32_t* counts) {
    if (n == 0) return 0;
    if (n <= (uint32_t)SORT_SMALL_N) { sort_small_kernel<<<1, SORT_THREADS, 0, st>>>(keys0, vals0, n); return 1; }
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
