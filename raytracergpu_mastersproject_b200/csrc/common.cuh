// common.cuh -- device-side arithmetic conventions and internal record layouts of librtb200.
//
// Arithmetic contract (must hold for bit-exact parity with the reference shaders as pinned in DESIGN.md):
//   * this library is compiled with --fmad=false: every + - * below is an individually rounded IEEE-754
//     binary32 op; '/' and sqrtf are the correctly rounded CUDA defaults (-prec-div/-prec-sqrt=true, no ftz)
//   * the only fused ops are the explicit fmaf() calls of pin_sincos()
//   * GLSL min/max are spelled as their defining ternaries (raytraceBVH.comp:188-191 semantics for NaN / -0)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rtb200.h"

namespace rtb {

struct f3 { float x, y, z; };

__host__ __device__ __forceinline__ f3 F3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ float gmin(float x, float y) { return y < x ? y : x; }  // GLSL min
__host__ __device__ __forceinline__ float gmax(float x, float y) { return x < y ? y : x; }  // GLSL max
__host__ __device__ __forceinline__ f3 operator+(f3 a, f3 b) { return F3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 operator-(f3 a, f3 b) { return F3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 operator*(f3 a, f3 b) { return F3(a.x * b.x, a.y * b.y, a.z * b.z); }
__host__ __device__ __forceinline__ f3 operator*(float s, f3 a) { return F3(s * a.x, s * a.y, s * a.z); }
__host__ __device__ __forceinline__ f3 operator/(f3 a, float s) { return F3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ f3 operator-(f3 a) { return F3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ float dot(f3 a, f3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ f3 cross(f3 a, f3 b) {
    return F3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ f3 normalize(f3 a) { return a / sqrtf(dot(a, a)); }
__device__ __forceinline__ f3 xyz(float4 v) { return F3(v.x, v.y, v.z); }

// ---- RNG: shaders/include/random.glsl:10-27 (PCG RXS-M-XS 32, inc = 1) ---------------------------------
__host__ __device__ __forceinline__ uint32_t pcg_step(uint32_t s) { return s * 747796405u + 1u; }
__host__ __device__ __forceinline__ uint32_t pcg_word(uint32_t s) {
    uint32_t w = ((s >> ((s >> 28) + 4u)) ^ s) * 277803737u;
    return (w >> 22) ^ w;
}
// stepAndOutputRNGFloat: float(word) / 4294967295.0f, whose literal is 2^32 in binary32 -> exact scaling
__device__ __forceinline__ float pcg_float(uint32_t& s) {
    s = pcg_step(s);
    return __uint2float_rn(pcg_word(s)) * 2.3283064365386963e-10f;
}
// uint(alpha * 4294967294.0f) (raytraceBVH.comp:350): the literal is 2^32; F2I.U32 saturates (pin U1)
__device__ __forceinline__ uint32_t alpha_to_u32(float a) { return __float2uint_rz(a * 4294967296.0f); }

// pinned sin/cos on [0, 2*pi] (DESIGN.md "pins", U11) -- explicit fused steps in a fixed order
__device__ __forceinline__ void pin_sincos(float x, float& s, float& c) {
    const float q = rintf(x * 0x1.45f306p-1f);
    float r = fmaf(q, -0x1.921fb4p+0f, x);
    r = fmaf(q, -0x1.4442d2p-24f, r);
    const float z = r * r;
    float ps = fmaf(-0x1.9943f2p-13f, z, 0x1.11073cp-7f);
    ps = fmaf(ps, z, -0x1.555546p-3f);
    const float sr = fmaf(ps * z, r, r);
    float pc = fmaf(0x1.99eb9cp-16f, z, -0x1.6c0c34p-10f);
    pc = fmaf(pc, z, 0x1.55554ap-5f);
    const float cr = fmaf(pc, z * z, fmaf(-0.5f, z, 1.0f));
    const int k = ((int)q) & 3;
    s = (k & 1) ? cr : sr;
    c = (k & 1) ? sr : cr;
    if (k & 2) s = -s;
    if ((k + 1) & 2) c = -c;
}

// 256-bit read-only global load (sm_100 LDG.E.256): one instruction per 32-byte sector instead of two LDG.128 -- halves the
// L1TEX wavefronts of the divergent node fetch, which is the unit the trace kernel saturates (profiles/r01_wave_kernel_ncu.txt)
struct __align__(32) f8 { float4 lo, hi; };
__device__ __forceinline__ f8 ldg256(const void* p) {
    f8 r;
#ifdef RTB_SIMT_EMU   // tests/emu: the kernel sources compiled for the CPU (SIMT emulation, test infrastructure) -- no PTX there
    r.lo = ((const float4*)p)[0]; r.hi = ((const float4*)p)[1];
#else
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
                 : "l"(p));
#endif
    return r;
}

// ---- traversal records (what the trace kernel fetches; derived from the reference-layout arrays) --------
// One 64-byte, 64-byte-aligned record per INTERNAL node i in [0, N-2]: both child boxes + both child indices,
// so one visit = four LDG.128 from two adjacent 32-byte sectors.  Leaves need no record: child index
// c >= N-1 is primitive g = c-(N-1) (triangle g if g < T else sphere g-T; ConstructHLBVH.comp:152-169).
struct __align__(16) PairNode {
    float4 lLo;   // left  child box min.xyz , w = bits(left index)
    float4 lHi;   // left  child box max.xyz , w = bits(right index)
    float4 rLo;   // right child box min.xyz , w unused
    float4 rHi;   // right child box max.xyz , w unused
};
// Compressed child-pair record, 32 B = ONE 256-bit load (trace_wave_shared.cuh::wave_step_c): origin.xyz (the node's own min
// corner), one power-of-two scale per axis, both child boxes as 8-bit offsets rounded OUTWARD (conservative), Karras split
// position + two leaf flags instead of two child indices.  Exactness is restored at the leaves (exact leaf box + exact
// primitive test): a leaf is tested iff its own exact box passes -- see DESIGN.md "compressed nodes".
// packed triangle: 4 x float4 = 64 B : (v0.xyz, bits(materialIndex)) (n.xyz, u.x) (u.yz, v.xy) (v.z, w.xyz) with u = v1-v0,
//                  v = v2-v0, n = normalize(cross(u,v)), w = cross(u,v)/dot(cross,cross): triangleHit's ray-independent prologue
// packed sphere  : 1 x float4        : (center.xyz, radius) + u32 materialIndex in a side array
// packed material: 1 x float4        : (albedo.xyz, bits(materialType))

struct TraceScene {
    const float4* __restrict__ pairs;    // [N-1][4]
    const float4* __restrict__ tris;     // [T][4]
    const float4* __restrict__ sphs;     // [S]
    const uint32_t* __restrict__ sphMat; // [S]
    const float4* __restrict__ mats;     // [M]
    const float4* __restrict__ rootBox;  // [4]: (min.xyz,0) (max.xyz,0) of node 0; [2..3] the origin region of the t-culled walk (eta_leaf_kernel)
    const uint4* __restrict__ cnodes;    // [N-1][2] 32-byte compressed child-pair records (conservative 8-bit boxes), or null
    const float4* __restrict__ leafBox;  // [N][2]   exact leaf boxes (min.xyz,0) (max.xyz,0), used with cnodes / wide
    const uint4* __restrict__ wide;      // [N-1][4] 64-byte 4-ary records (conservative 8-bit grandchild boxes), or null
    const uint4* __restrict__ top;       // A/B build RTB_SMEM_TOP: [k][4] the first k records in breadth-first order, entry ids inside the table tagged
    const uint32_t* __restrict__ topGlobal;   // [k] node index of each table record
    uint32_t T, S, N;
};

struct Camera {            // raytraceBVH.comp:50-81, hoisted to the host (computed once per submission)
    f3 origin, pixel00, deltaU, deltaV;
};

// streaming kernel: pool of in-flight paths (SoA) + the list of rays to trace in the current iteration
struct StreamPool {
    uint32_t* slot;        // [P] (sample, active pixel) slot index of the path, 0xFFFFFFFF = empty
    float4* colorRng;      // [P] (colour.xyz, bits(rng))
    float4* attDepth;      // [P] (attenuation.xyz, bits(depth))
    float4* org;           // [P] (ray origin.xyz, hit t)
    float4* dir;           // [P] (ray direction.xyz, bits(hit primitive) or 0xFFFFFFFF)
    float4* nrm;           // [P] (hit normal.xyz, bits(material | back << 31))
    uint32_t* rayList;     // [P] pool indices of the rays of this iteration
    unsigned int* cnt;     // [4] ray count (parity 0/1), fetch cursor (parity 0/1)
    uint32_t capacity;
};

struct TraceParams {
    TraceScene sc;
    Camera cam;
    float4* image;
    uint32_t W, H, localRows, bandRows, bandFirst, bandStep;
    uint32_t sampleSkip, sampleCount, maxDepth, randomState;
    uint32_t* hitPrim; float* hitT; uint32_t* rngOut;
    unsigned long long* counters;      // rtb_counters (reference-equivalent work), or null
    unsigned long long* walkCounters;  // rtb_walk_counters (what the kernels fetch), or null
    unsigned int* workCounter;         // persistent-thread tile counter
    unsigned int* errFlag;             // bit0: traversal stack overflow
    const unsigned int* cullAllowed;   // nearest-first walk: 1 = the records were grown by the hit-point slack, t-culling is sound (pack_wide_kernel)
    uint32_t tilesX, tilesY;
    uint32_t sortedPush;               // nearest-first kernel: pick the variant that stacks waiting entries farthest-first
    uint32_t qGate;                    // nearest-first kernel: a lane keeps stepping while its FIFO holds <= qGate candidates
    uint32_t tMin;                     // wave kernel: minimum stepping lanes to stay in the traverse phase (0 = default)
    uint32_t mainCtas, tailThreads;    // wave kernel: cap on resident CTAs per SM (0 = none) / CTA size of the tail launch (host-side launch shape)
    uint32_t sMin;                     // wave kernel: lanes that must wait for the shade / generate phase before it is entered (1 = every time)
    // tail hand-over (trace_wave.cu): once the work queue is drained, a warp with <= coopMax live lanes parks the paths whose next
    // ray is about to start; trace_tail_kernel finishes them one ray per WARP (0 = off)
    uint32_t coopMax;
    // long-ray hand-over: a ray still being walked after coopTurns turns of its lane (a ray tangent to a finely tessellated
    // surface crosses thousands of leaf boxes) is parked with its pending stack and finished by the tail kernel (0 = off)
    uint32_t coopTurns;
    float4* parkBuf;                   // [parkCapacity][PARK_STRIDE]: see trace_wave.cu
    unsigned int* parkCount;           // paths parked by the main launch
    unsigned int* parkCursor;          // fetch cursor of the tail launch
    uint32_t parkCapacity;
    uint32_t parkEpoch;                // ready-flag value of this launch (last word of a park record); changes with every launch that parks, so the flags never need clearing
    unsigned int* doneWarps;           // [0] warps of the main launch that have finished (and park nothing any more), [1] its CTAs that have started
    uint32_t tailConcurrent;           // trace_tail_kernel: 1 = runs beside the main launch and polls, 0 = final drain in stream order
    uint32_t tailSpinUs;               // concurrent tail launch: upper bound of a warp's polling time
    uint32_t mainWarps;                // warps of the main launch: the concurrent tail launch leaves when doneWarps reaches it
    // wave kernel, per pass: (pixel, sample) work items
    unsigned long long* workCounter64; // [0] work-item counter, [1] (as unsigned*) active-pixel count
    unsigned int* activeCount;         // = (unsigned*)(workCounter64 + 1)
    uint32_t* activePix;               // [pixels] local pixel index of every pixel whose primary ray enters the root box
    uint32_t* activeXY;                // [pixels] x | (global row << 16) of the same pixels (saves the item fetch four integer divisions)
    float4* sampleBuf;                 // [samplesPerPass][slotCapacity]: (colour.xyz, incoming alpha) per (sample, active pixel)
    uint32_t slotCapacity;
    uint32_t firstPass, lastPass;
    // primary-hit sharing (trace_wave.cu): 0 off | 1 this launch traces ONE primary ray per active pixel and stores its hit |
    // 2 every (pixel, sample) item starts from the stored hit
    uint32_t primaryMode;
    float4* primaryHits;               // [pixels][3]: (t, normal.xyz), (prim, mat, hit, backFace), (primary direction.xyz, -)
    StreamPool pool;
    f3 background;                     // _BACKGROUND_COLOR (0 for the BVH program, (0.1,0.1,0.3) for the non-BVH program)
};

}  // namespace rtb
