// traversal_tree.cu -- the hierarchy the order-free (nearest-first, t-culled) walk descends.
//
// The reference's tree (bvh_build.cu: 30-bit Morton codes of coord / span -- not (coord - min) / span --, leaves in input order) is what
// the API returns bit for bit, but it is a poor tree to WALK: at C2 ten bits per axis leave several triangles per Morton cell, whose order
// inside the cell is the input order; the y bit is spent as often as x and z however flat the scene is; and a handful of room-sized
// primitives inflate every ancestor box.  A CPU simulation (profiles/sim_tree2.py) put a plain longest-axis median tree at 10.0 4-ary steps
// per ray against 18.9 for the reference tree (C2) and 15.6 against 28.0 (C5).  The walk only needs a hierarchy of CONSERVATIVE boxes over
// the same leaves -- the leaf tests (exact leaf box, then the reference's primitive test with the order-free acceptance) decide the result,
// and they do not care where a leaf hangs -- so rtb_build_bvh builds a second tree for the walk, behind the reference's:
//   * BIG leaves (box above 1/256 of the surface area of the scene's box, at most MAX_BIG) are left out and presented to every ray by a short
//     chain of records in front of the root (three big leaves and a link each): a ray that starts inside a room-sized box can never drop it by t;
//   * the other leaves get a 32-bit code of their box centre relative to the bounds of those centres, the bits spent on the axis whose cell is
//     currently the longest (a flat height field splits x and z five times before it splits y once; a cube degenerates to plain Morton order);
//   * stable radix sort (radix_sort.cu), Karras topology with the index tie-break (build_shared.cuh), atomic bottom-up union of the exact leaf
//     boxes and of the hit-point slack (as refit_kernel), 4-ary greedy cuts quantised outward to 8 bits (as pack_wide_kernel).
// Everything is decided on the device (how many big leaves there are, hence how many leaves the tree has: flags[2], flags[3]); grids are sized
// for all N primitives.  Results are the reference's bit for bit (tests/test_gpu_fuzz.py, and every parity test on scenes of >= 8192 primitives).
#include "build_shared.cuh"
#include "common.cuh"
#include "kernels.h"

namespace rtb {

namespace {

// 32 bits, most significant first; each bit halves the cell along the axis on which the cell is currently the longest
__device__ __forceinline__ uint32_t adaptive_code(float ux, float uy, float uz, float ex, float ey, float ez) {
    uint32_t code = 0;
#pragma unroll 4
    for (int b = 0; b < 32; b++) {
        const bool px = ex >= ey && ex >= ez, py = !px && ey >= ez;
        float u = px ? ux : (py ? uy : uz);
        u *= 2.0f;
        const bool bit = u >= 1.0f;
        u -= bit ? 1.0f : 0.0f;
        if (px) { ux = u; ex *= 0.5f; } else if (py) { uy = u; ey *= 0.5f; } else { uz = u; ez *= 0.5f; }
        code = (code << 1) | (bit ? 1u : 0u);
    }
    return code;
}

// keys: big leaves sort to the very end (0xFFFFFFFF), everything else by its adaptive code; thread 0 publishes n2 = leaves in the tree
__global__ void __launch_bounds__(256) tt_keys_kernel(const float4* __restrict__ leafBox, uint32_t N, const uint32_t* __restrict__ primBounds,
                                                      const uint32_t* __restrict__ smallBounds, unsigned int* flags, uint32_t* keys, uint32_t* vals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int nb = flags[2];
    const bool hoist = nb > 0u && nb <= MAX_BIG && N - nb >= 2u;
    if (g == 0) flags[3] = hoist ? N - nb : N;
    if (g >= N) return;
    const float4 lo = leafBox[2ull * g], hi = leafBox[2ull * g + 1];
    uint32_t key = 0xFFFFFFFFu;
    if (!(hoist && box_is_big(lo.x, hi.x, lo.y, hi.y, lo.z, hi.z, hoist_threshold(primBounds)))) {
        const float bx = ord2f(smallBounds[0]), by = ord2f(smallBounds[1]), bz = ord2f(smallBounds[2]);
        const float ex = ord2f(smallBounds[3]) - bx, ey = ord2f(smallBounds[4]) - by, ez = ord2f(smallBounds[5]) - bz;
        // a leaf outside the bounds (a big one that is not hoisted after all) or a degenerate extent clamps to the border cell
        float ux = ex > 0.f ? (0.5f * (lo.x + hi.x) - bx) / ex : 0.f, uy = ey > 0.f ? (0.5f * (lo.y + hi.y) - by) / ey : 0.f,
              uz = ez > 0.f ? (0.5f * (lo.z + hi.z) - bz) / ez : 0.f;
        ux = fminf(fmaxf(ux, 0.f), 0.99999994f); uy = fminf(fmaxf(uy, 0.f), 0.99999994f); uz = fminf(fmaxf(uz, 0.f), 0.99999994f);
        if (!(ux == ux)) ux = 0.f; if (!(uy == uy)) uy = 0.f; if (!(uz == uz)) uz = 0.f;
        key = min(adaptive_code(ux, uy, uz, fmaxf(ex, 0.f), fmaxf(ey, 0.f), fmaxf(ez, 0.f)), 0xFFFFFFFEu);
    }
    keys[g] = key; vals[g] = g;
}

// Karras topology over the first n2 sorted keys: child[i] = (left, right) as node indices (a leaf is n2 - 1 + sorted position), parent links
__global__ void __launch_bounds__(256) tt_topology_kernel(const uint32_t* __restrict__ keys, const unsigned int* __restrict__ flags, uint2* child,
                                                          uint32_t* parent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n2 = (int)flags[3];
    if ((int)i >= n2 - 1) return;
    const Codes codes{ keys, 1, n2 };
    int l, r;
    karras_children(codes, (int)i, n2 - 1, l, r);
    child[i] = make_uint2((uint32_t)l, (uint32_t)r);
    parent[l] = i; parent[r] = i;
    if (i == 0) parent[0] = 0;
}

__device__ __forceinline__ void tt_node_box(const uint32_t c, const uint32_t leafOffset2, const uint32_t* __restrict__ vals, const float4* __restrict__ leafBox,
                                            const float* __restrict__ etaLeaf, const float4* box, const float* eta, float4& lo, float4& hi, float& e) {
    if (c >= leafOffset2) {
        const uint32_t g = vals[c - leafOffset2];
        lo = leafBox[2ull * g]; hi = leafBox[2ull * g + 1]; e = etaLeaf[g];
    } else {
        lo = __ldcg(&box[2ull * c]); hi = __ldcg(&box[2ull * c + 1]); e = __ldcg(&eta[c]);
    }
}

// bottom-up union of boxes and slack: one thread per leaf climbs, the second arrival at a node continues (as refit_kernel).  The climbing
// thread keeps the box of the subtree it comes from in registers and the (static) topology words of a node are fetched beside the arrival
// atomic, so a level costs two dependent memory round trips -- the atomic, then the sibling's box -- instead of four.
__global__ void __launch_bounds__(256) tt_refit_kernel(const unsigned int* __restrict__ flags, const uint2* __restrict__ child, const uint32_t* __restrict__ parent,
                                                       unsigned int* arrivals, const uint32_t* __restrict__ vals, const float4* __restrict__ leafBox,
                                                       const float* __restrict__ etaLeaf, float4* box, float* eta) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n2 = flags[3];
    if (j >= n2 || n2 < 2u) return;
    const uint32_t leafOffset2 = n2 - 1u;
    uint32_t me = leafOffset2 + j;
    uint32_t node = parent[me];
    float4 lo, hi; float e;
    tt_node_box(me, leafOffset2, vals, leafBox, etaLeaf, box, eta, lo, hi, e);
    while (true) {
        const uint2 ch = child[node];
        const uint32_t up = node ? parent[node] : 0u;
        __threadfence();                                                    // my subtree's box (stored below) before my arrival
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;                // the sibling subtree is not finished yet
        __threadfence();
        float4 l1, h1; float e1;
        tt_node_box(ch.x == me ? ch.y : ch.x, leafOffset2, vals, leafBox, etaLeaf, box, eta, l1, h1, e1);
        lo = make_float4(fminf(lo.x, l1.x), fminf(lo.y, l1.y), fminf(lo.z, l1.z), 0.f);
        hi = make_float4(fmaxf(hi.x, h1.x), fmaxf(hi.y, h1.y), fmaxf(hi.z, h1.z), 0.f);
        e = fmaxf(e, e1);
        __stcg(&box[2ull * node], lo);
        __stcg(&box[2ull * node + 1], hi);
        __stcg(&eta[node], e);
        if (node == 0u) return;
        me = node; node = up;
    }
}

// 4-ary records of the tree (greedy surface-area cut, as pack_wide_kernel); a leaf entry carries the reference's leaf index N - 1 + primitive
__global__ void __launch_bounds__(128) tt_pack_kernel(const unsigned int* __restrict__ flags, const uint2* __restrict__ child, const uint32_t* __restrict__ vals,
                                                      const float4* __restrict__ leafBox, const float* __restrict__ etaLeaf, const float4* __restrict__ box,
                                                      const float* __restrict__ eta, const float* __restrict__ etaRootAll, uint32_t N, uint4* wide) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n2 = flags[3];
    if (n2 < 2u || i >= n2 - 1u) return;
    // the slack of the reference tree's root (the maximum over ALL primitives, big ones included) decides whether the records are grown
    const float er = *etaRootAll;
    const bool slackOk = er >= 0.0f && er < 3.0e38f;
    const uint32_t leafOffset2 = n2 - 1u;
    uint32_t entry[4]; int cnt = 2;
    { const uint2 c = child[i]; entry[0] = c.y; entry[1] = c.x; }
    while (cnt < 4) {
        int pick = -1; float best = -1.0f;
        for (int e = 0; e < cnt; e++) {
            if (entry[e] >= leafOffset2) continue;
            const float4 l = box[2ull * entry[e]], h = box[2ull * entry[e] + 1];
            const float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
            const float area = dx * dy + dy * dz + dz * dx;
            if (pick < 0 || area > best) { pick = e; best = area; }
        }
        if (pick < 0) break;
        const uint2 c = child[entry[pick]];
        for (int e = cnt; e > pick + 1; e--) entry[e] = entry[e - 1];
        entry[pick] = c.y; entry[pick + 1] = c.x;
        cnt++;
    }
    float lo[4][3], hi[4][3];
    uint32_t ids[4], leafMask = 0;
    for (int e = 0; e < cnt; e++) {
        float4 l, h; float sl;
        tt_node_box(entry[e], leafOffset2, vals, leafBox, etaLeaf, box, eta, l, h, sl);
        if (!slackOk) sl = 0.f;
        lo[e][0] = __fsub_rd(l.x, sl); lo[e][1] = __fsub_rd(l.y, sl); lo[e][2] = __fsub_rd(l.z, sl);
        hi[e][0] = __fadd_ru(h.x, sl); hi[e][1] = __fadd_ru(h.y, sl); hi[e][2] = __fadd_ru(h.z, sl);
        if (entry[e] >= leafOffset2) { leafMask |= 1u << e; ids[e] = (N - 1u) + vals[entry[e] - leafOffset2]; }
        else ids[e] = entry[e];
    }
    store_wide_record(wide + 4ull * i, cnt, lo, hi, ids, leafMask);
}

// flags[0] (may the walk cull by t), flags[1] (the record it starts at) and the records that present the big leaves: record k, at index
// N - 1 + k, holds big leaves 3k .. 3k+2 and a link to record k + 1, the last one to the tree's root record 0.  One thread: <= 15 records.
__global__ void tt_top_kernel(const unsigned int* __restrict__ flagsIn, unsigned int* flags, const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ primBounds,
                              const float4* __restrict__ leafBox,
                              const float* __restrict__ etaLeaf, const float* __restrict__ etaRootAll, const float4* __restrict__ box, const float* __restrict__ eta,
                              uint32_t N, uint4* wide) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    const float er = *etaRootAll;
    const bool slackOk = er >= 0.0f && er < 3.0e38f;
    flags[0] = slackOk ? 1u : 0u;
    const uint32_t n2 = flagsIn[3], nb = N - n2;                 // (n2 < N only when the big leaves are really left out)
    if (nb == 0u) { flags[1] = 0u; return; }
    const uint32_t K = (nb + 2u) / 3u;
    for (uint32_t k = 0; k < K; k++) {
        float lo[4][3], hi[4][3]; uint32_t ids[4]; int cnt = 0; uint32_t leafMask = 0;
        for (uint32_t j = 3u * k; j < nb && j < 3u * k + 3u; j++) {
            const uint32_t g = bigList[j];
            const float4 l = leafBox[2ull * g], h = leafBox[2ull * g + 1];
            const float sl = slackOk ? etaLeaf[g] : 0.0f;
            lo[cnt][0] = __fsub_rd(l.x, sl); lo[cnt][1] = __fsub_rd(l.y, sl); lo[cnt][2] = __fsub_rd(l.z, sl);
            hi[cnt][0] = __fadd_ru(h.x, sl); hi[cnt][1] = __fadd_ru(h.y, sl); hi[cnt][2] = __fadd_ru(h.z, sl);
            ids[cnt] = (N - 1u) + g; leafMask |= 1u << cnt; cnt++;
        }
        {   // the link: to the tree's root record (everything below it lies inside the tree's root box) or to the next of these records
            // (which holds more big leaves: the box of ALL primitives, padded like a leaf box can be, grown by the largest slack)
            const bool last = k + 1u == K;
            const float4 l = last ? box[0] : make_float4(ord2f(primBounds[0]) - 0.0005f, ord2f(primBounds[1]) - 0.0005f, ord2f(primBounds[2]) - 0.0005f, 0.f);
            const float4 h = last ? box[1] : make_float4(ord2f(primBounds[3]) + 0.0005f, ord2f(primBounds[4]) + 0.0005f, ord2f(primBounds[5]) + 0.0005f, 0.f);
            const float sl = slackOk ? (last ? eta[0] : er) : 0.0f;
            lo[cnt][0] = __fsub_rd(l.x, sl); lo[cnt][1] = __fsub_rd(l.y, sl); lo[cnt][2] = __fsub_rd(l.z, sl);
            hi[cnt][0] = __fadd_ru(h.x, sl); hi[cnt][1] = __fadd_ru(h.y, sl); hi[cnt][2] = __fadd_ru(h.z, sl);
            ids[cnt] = k + 1u < K ? (N - 1u) + k + 1u : 0u; cnt++;
        }
        store_wide_record(wide + 4ull * ((N - 1u) + k), cnt, lo, hi, ids, leafMask);
    }
    flags[1] = N - 1u;
}

unsigned blocks_of(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

// etaNode: the reference tree's per-node slack (leaf values at [N-1 + g], the maximum over all primitives at [0]); wide: room for
// N - 1 + 16 records.  etaRootReady: recorded behind the reference tree's climb (which may run on another stream); the pack kernels --
// the first readers of etaNode[0] -- wait for it.  Returns #launches.
int launch_traversal_tree(cudaStream_t st, uint32_t N, const void* leafBox, const float* etaNode, const TraversalTreeBuffers& b, void* wide,
                          cudaEvent_t etaRootReady) {
    if (N < 2) { if (etaRootReady) cudaStreamWaitEvent(st, etaRootReady, 0); return 0; }
    const float* etaLeaf = etaNode + (N - 1);
    cudaMemsetAsync(b.arrivals, 0, sizeof(unsigned int) * (N - 1), st);
    tt_keys_kernel<<<blocks_of(N), 256, 0, st>>>((const float4*)leafBox, N, b.primBounds, b.smallBounds, b.flags, b.keys0, b.vals0);
    const int sortLaunches = launch_radix_sort(st, b.keys0, b.vals0, b.keys1, b.vals1, N, b.sortCounts);
    tt_topology_kernel<<<blocks_of(N), 256, 0, st>>>(b.keys0, b.flags, b.child, b.parent);
    tt_refit_kernel<<<blocks_of(N), 256, 0, st>>>(b.flags, b.child, b.parent, b.arrivals, b.vals0, (const float4*)leafBox, etaLeaf, b.box, b.eta);
    if (etaRootReady) cudaStreamWaitEvent(st, etaRootReady, 0);
    tt_pack_kernel<<<(unsigned)(((uint64_t)N + 127) / 128), 128, 0, st>>>(      // (56 registers: CTAs of 128 fit beside a resident trace launch)
        b.flags, b.child, b.vals0, (const float4*)leafBox, etaLeaf, b.box, b.eta, etaNode, N, (uint4*)wide);
    tt_top_kernel<<<1, 32, 0, st>>>(b.flags, b.flags, b.bigList, b.primBounds, (const float4*)leafBox, etaLeaf, etaNode, b.box, b.eta, N, (uint4*)wide);
    return sortLaunches + 5;
}

}  // namespace rtb
