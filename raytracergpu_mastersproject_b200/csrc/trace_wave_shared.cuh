// trace_wave_shared.cuh -- pieces shared by the warp-coherent trace kernels (trace_wave.cu, trace_stream.cu): the filtered
// box test, the per-lane traversal step, and the pre-pass / accumulate kernels of the (pixel, sample) work-item scheme.
#pragma once

#include "trace_common.cuh"

namespace rtb {

constexpr int SSTACK = 32;                // stack entries kept in shared memory; deeper levels spill to local memory
constexpr int QCAP = 8;                   // pending-leaf FIFO entries per lane
constexpr int T_MIN_DEFAULT = 20;
#ifndef RTB_WIDE_POPS
#define RTB_WIDE_POPS 1
#endif
constexpr int WIDE_STACK_DEPTH = 128;      // 4-ary records push up to three entries per level: a deeper spill area than the reference's 64         // leave the traverse phase when fewer lanes than this can step

template <int THREADS>
struct __align__(16) WaveSmem {
    uint32_t stack[SSTACK][THREADS];   // [level][thread]: conflict-free
    uint32_t queue[QCAP][THREADS];     // pending-leaf FIFO
};

// two-sided filter on the reciprocal-multiply slab test; returns 1 = pass, 0 = fail, -1 = undecided
__device__ __forceinline__ int box_filter(const f3 o, const f3 rinv, const float lox, const float loy, const float loz, const float hix,
                                          const float hiy, const float hiz) {
    const float ax = (lox - o.x) * rinv.x, ay = (loy - o.y) * rinv.y, az = (loz - o.z) * rinv.z;
    const float bx = (hix - o.x) * rinv.x, by = (hiy - o.y) * rinv.y, bz = (hiz - o.z) * rinv.z;
    const float tNear = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tFar = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    // max / min are monotone, so tNear and tFar inherit the 3-ulp relative error of the products with respect to
    // THEMSELVES; with the rounding of the subtraction: |diff - (tFar - tNear)_reference| <= 4 ulp * (|tNear| + |tFar|)
    // = 2.384e-7 * (...) (second-order terms ~1e-14).  The filter uses 3e-7 plus an absolute term for the subnormal range.
    const float e = fmaf(fabsf(tNear) + fabsf(tFar), 3.0e-7f, 1.0e-36f);
    const float diff = tFar - tNear;
    if (diff > e) return 1;
    if (diff < -e) return 0;
    return -1;                            // too close to call (or inf / NaN): ask the exact test
}

__device__ __forceinline__ bool box_test(const f3 o, const f3 d, const f3 rinv, const bool exactOnly, const float lox, const float loy,
                                         const float loz, const float hix, const float hiy, const float hiz) {
    if (!exactOnly) {
        const int r = box_filter(o, rinv, lox, loy, loz, hix, hiy, hiz);
        if (r >= 0) return r != 0;
    }
    return box_hit(o, d, lox, loy, loz, hix, hiy, hiz);
}

// global image row of local row j (rtb_trace_args: interleaved bands)
__device__ __forceinline__ uint32_t global_row(const TraceParams& p, uint32_t j) {
    return ((j / p.bandRows) * p.bandStep + p.bandFirst) * p.bandRows + (j % p.bandRows);
}
// getRay :329-342 (no jitter: the same for every sample of a pixel) + rayColor's own normalize :280
__device__ __forceinline__ f3 primary_direction(const TraceParams& p, uint32_t x, uint32_t y) {
    const f3 pixelSample = (p.cam.pixel00 + (float)x * p.cam.deltaU) + (float)y * p.cam.deltaV;
    return normalize(normalize(pixelSample - p.cam.origin));
}

// ---------------------------------------------------------------------------------------------------------------------
// Pre-pass, one thread per local pixel: walk the alpha seed chain for the `passCount` samples of this pass.
//  * primary ray misses the root box (or maxDepth == 0): every sample is "seed, one random(), colour 0" -> finish the pixel here
//  * otherwise: append the pixel to the active list and park sample s's incoming alpha in slot (s, pixel).w
// ---------------------------------------------------------------------------------------------------------------------
template <bool COUNT>
__global__ void __launch_bounds__(256) wave_prepass_kernel(const TraceParams p) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t x = idx % p.W, j = idx / p.W;
    unsigned long long nRays = 0, nSamples = 0;
    bool activePixel = false;
    uint32_t y = 0, base = 0;
    float alpha = 0.f;
    if (j < p.localRows) {
        y = global_row(p, j);
        if (y < p.H) {
            const float4 c = p.image[idx];                                                  // imageLoad :349
            base = (600u * x + y) * (p.randomState + 1u);                                   // random.glsl:10
            alpha = c.w;
            for (uint32_t s = 0; s < p.sampleSkip; s++) { uint32_t t = base + alpha_to_u32(alpha); alpha = pcg_float(t); }
            bool rootPass = false;
            if (p.maxDepth != 0) {
                const f3 dir = primary_direction(p, x, y);
                const f3 ri = F3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
                const bool ex = !(fabsf(ri.x) < 3.0e38f && fabsf(ri.y) < 3.0e38f && fabsf(ri.z) < 3.0e38f);
                const float4 lo = __ldg(p.sc.rootBox), hi = __ldg(p.sc.rootBox + 1);
                rootPass = box_test(p.cam.origin, dir, ri, ex, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
            }
            nSamples = p.sampleCount;
            if (!rootPass) {
                f3 rgb = F3(c.x, c.y, c.z);
                uint32_t rng = 0;
                for (uint32_t s = 0; s < p.sampleCount; s++) {
                    rng = base + alpha_to_u32(alpha);                                       // :350
                    alpha = pcg_float(rng);                                                 // nextRandom :352 -> alpha :372
                    const f3 col = F3(0.f, 0.f, 0.f) + F3(0.f, 0.f, 0.f) * F3(1.f, 1.f, 1.f);   // :278-279,284
                    rgb = col + rgb;
                }
                if (p.maxDepth != 0) nRays = p.sampleCount;
                p.image[idx] = make_float4(rgb.x, rgb.y, rgb.z, alpha);                     // imageStore :374
                if (p.rngOut && p.lastPass) p.rngOut[idx] = rng;
                if (p.hitPrim && p.firstPass) { p.hitPrim[idx] = 0xFFFFFFFFu; if (p.hitT) p.hitT[idx] = 0.0f; }
            } else {
                activePixel = true;
            }
        }
    }
    // warp-aggregated append to the active list
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, activePixel);
    if (bal) {
        const unsigned lane = threadIdx.x & 31;
        uint32_t slot = 0;
        if (lane == (unsigned)(__ffs(bal) - 1)) slot = atomicAdd(p.activeCount, (unsigned)__popc(bal));
        slot = __shfl_sync(0xFFFFFFFFu, slot, __ffs(bal) - 1) + __popc(bal & ((1u << lane) - 1u));
        if (activePixel) {
            p.activePix[slot] = idx;
            p.activeXY[slot] = x | (y << 16);
            for (uint32_t s = 0; s < p.sampleCount; s++) {
                p.sampleBuf[(size_t)s * p.slotCapacity + slot] = make_float4(0.f, 0.f, 0.f, alpha);
                uint32_t t = base + alpha_to_u32(alpha);
                alpha = pcg_float(t);
            }
        }
    }
    if (COUNT) {
        unsigned long long v[2] = { nRays, nSamples };
#pragma unroll
        for (int i = 0; i < 2; i++) {
            unsigned long long t = v[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, off);
            v[i] = t;
        }
        if ((threadIdx.x & 31) == 0 && p.counters) {
            if (v[0]) { atomicAdd(p.counters + 0, v[0]); atomicAdd(p.counters + 1, v[0]); }   // rays, node visits (root only)
            if (v[1]) atomicAdd(p.counters + 5, v[1]);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Accumulate, one thread per active pixel: rgb = colour_s + rgb for s in order (main() :372), alpha = end of the chain.
// ---------------------------------------------------------------------------------------------------------------------
template <int UNUSED = 0>
__global__ void __launch_bounds__(256) wave_accumulate_kernel(const TraceParams p) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= *p.activeCount) return;
    const uint32_t idx = p.activePix[slot];
    const uint32_t x = idx % p.W, y = global_row(p, idx / p.W);
    const float4 c = p.image[idx];
    f3 rgb = F3(c.x, c.y, c.z);
    float alphaIn = 0.f;
    for (uint32_t s = 0; s < p.sampleCount; s++) {
        const float4 e = p.sampleBuf[(size_t)s * p.slotCapacity + slot];
        rgb = F3(e.x, e.y, e.z) + rgb;                                                       // pixelColor + currentColor.xyz
        alphaIn = e.w;
    }
    uint32_t t = (600u * x + y) * (p.randomState + 1u) + alpha_to_u32(alphaIn);
    const float alphaOut = pcg_float(t);                                                     // nextRandom of the last sample
    p.image[idx] = make_float4(rgb.x, rgb.y, rgb.z, alphaOut);                               // imageStore :374
}


// One traverse-phase turn of one lane: expand the child-pair record of `cur` (two 256-bit loads, two filtered box tests,
// straight-line queue / stack bookkeeping in the reference's order: right subtree completely, then left,
// raytraceBVH.comp:241-244), then resume from the stack with at most one predicated pop.
template <bool COUNT, bool CULL, int THREADS>
__device__ __forceinline__ void wave_step(const TraceScene& sc, WaveSmem<THREADS>& sm, const unsigned tid, const uint32_t leafOffset,
                                          const f3 o, const f3 d, const f3 rinv, const bool exactOnly, uint32_t& cur, int& sp,
                                          const uint32_t qHead, uint32_t& qCount, bool& travDone, uint32_t* lstack, Tally& tl, unsigned& err,
                                          const f3 segLo, const f3 segHi) {
    if (cur != 0xFFFFFFFFu) {
        const float4* pr = sc.pairs + 4ull * cur;
        const f8 nl = ldg256(pr), nr = ldg256(pr + 2);          // 64-byte record = two 256-bit loads
        const float4 lLo = nl.lo, lHi = nl.hi, rLo = nr.lo, rHi = nr.hi;
        if (COUNT) tl.visits += 2;
        const uint32_t li = __float_as_uint(lLo.w), ri = __float_as_uint(lHi.w);
        int fR = box_filter(o, rinv, rLo.x, rLo.y, rLo.z, rHi.x, rHi.y, rHi.z);
        int fL = box_filter(o, rinv, lLo.x, lLo.y, lLo.z, lHi.x, lHi.y, lHi.z);
        if (exactOnly || (fR | fL) < 0) {                      // rare: some comparison is too close to call -> the reference's divisions
            if (exactOnly || fR < 0) fR = box_hit(o, d, rLo.x, rLo.y, rLo.z, rHi.x, rHi.y, rHi.z) ? 1 : 0;
            if (exactOnly || fL < 0) fL = box_hit(o, d, lLo.x, lLo.y, lLo.z, lHi.x, lHi.y, lHi.z) ? 1 : 0;
        }
        bool passR = fR != 0, passL = fL != 0;
        if (CULL) {
            passR = passR && !(rLo.x > segHi.x || rHi.x < segLo.x || rLo.y > segHi.y || rHi.y < segLo.y || rLo.z > segHi.z || rHi.z < segLo.z);
            passL = passL && !(lLo.x > segHi.x || lHi.x < segLo.x || lLo.y > segHi.y || lHi.y < segLo.y || lLo.z > segHi.z || lHi.z < segLo.z);
        }
        const bool leafR = ri >= leafOffset, leafL = li >= leafOffset;
        const bool goR = passR && !leafR;                      // descend right now
        const bool enqR = passR && leafR;                      // right child is a leaf: test it first
        const bool pushL = passL && goR;                       // left waits on the stack until right is done
        const bool enqL = passL && !goR && leafL;
        const bool goL = passL && !goR && !leafL;
        uint32_t tail = (qHead + qCount) & (QCAP - 1);
        if (enqR) sm.queue[tail][tid] = ri - leafOffset;
        tail = (tail + (enqR ? 1u : 0u)) & (QCAP - 1);
        if (enqL) sm.queue[tail][tid] = li - leafOffset;
        qCount += (enqR ? 1u : 0u) + (enqL ? 1u : 0u);
        // push: predicated shared-memory store; the local-memory levels (sp >= SSTACK) are a cold branch
        if (pushL && sp < SSTACK) sm.stack[sp][tid] = li;
        if (pushL && sp >= SSTACK) {
            if (sp < STACK_DEPTH) lstack[sp - SSTACK] = li; else err |= 1u;
        }
        sp += (pushL && sp < STACK_DEPTH) ? 1 : 0;
        cur = goR ? ri : (goL ? li : 0xFFFFFFFFu);
    }
    // resume from the stack: one predicated pop per turn (a popped leaf is queued and the lane pops again next turn)
    const bool needPop = cur == 0xFFFFFFFFu && qCount < QCAP;
    if (needPop && sp == 0) travDone = true;
    const bool doPop = needPop && sp > 0;
    uint32_t e = 0;
    if (doPop && sp <= SSTACK) e = sm.stack[sp - 1][tid];
    if (doPop && sp > SSTACK) e = lstack[sp - 1 - SSTACK];
    sp -= doPop ? 1 : 0;
    const bool popLeaf = doPop && e >= leafOffset;
    if (popLeaf) sm.queue[(qHead + qCount) & (QCAP - 1)][tid] = e - leafOffset;
    qCount += popLeaf ? 1u : 0u;
    if (doPop && !popLeaf) cur = e;
}

#ifdef RTB_AB_KERNELS   // measured and not adopted: compiled only into A/B builds (make EXTRA=-DRTB_AB_KERNELS AB=1)
// Traverse-phase turn over the 32-byte COMPRESSED record (one 256-bit load).  Child planes are conservative (outward
// rounded) 8-bit offsets from the node origin; slab parameters are evaluated as t = q * (scale / dir) + (origin - o) / dir
// with one FFMA per plane.  A child is skipped only when it is a definite miss even after granting every rounding error
// (|error| <= 4 ulp * (255 |scale/dir| + |(origin - o)/dir|) per value); everything else is visited.  Leaves
// that survive are queued as CANDIDATES: the L phase first runs the reference's exact box test on the exact leaf box.
// Because the reference's box test is monotone under box inclusion, a leaf's exact box passing implies all its ancestors'
// exact (hence conservative) boxes pass, so the set and order of primitive tests is unchanged.
template <bool CULL, int THREADS>
__device__ __forceinline__ void wave_step_c(const TraceScene& sc, WaveSmem<THREADS>& sm, const unsigned tid, const uint32_t leafOffset,
                                            const f3 o, const f3 rinv, uint32_t& cur, int& sp, const uint32_t qHead, uint32_t& qCount,
                                            bool& travDone, uint32_t* lstack, unsigned& err, const f3 segLo, const f3 segHi) {
    if (cur != 0xFFFFFFFFu) {
        const f8 rec = ldg256(sc.cnodes + 2ull * cur);
        const uint32_t w3 = __float_as_uint(rec.lo.w), q0 = __float_as_uint(rec.hi.x), q1 = __float_as_uint(rec.hi.y),
                       q2 = __float_as_uint(rec.hi.z), split = __float_as_uint(rec.hi.w);
        const float sx = __uint_as_float((w3 & 0xFFu) << 23), sy = __uint_as_float(((w3 >> 8) & 0xFFu) << 23),
                    sz = __uint_as_float(((w3 >> 16) & 0xFFu) << 23);
        const float ax = sx * rinv.x, ay = sy * rinv.y, az = sz * rinv.z;                       // exact (power-of-two scaling)
        const float bx = (rec.lo.x - o.x) * rinv.x, by = (rec.lo.y - o.y) * rinv.y, bz = (rec.lo.z - o.z) * rinv.z;
        const float m = fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
        const float tol = -2.0e-6f * m;                                                         // NaN / inf -> nothing is skipped
#define RTB_Q(w, j) __uint2float_rn(((w) >> (8 * (j))) & 0xFFu)
        // left child: q0 = lo.x lo.y lo.z hi.x ; q1 = hi.y hi.z | right child: q1 = lo.x lo.y ; q2 = lo.z hi.x hi.y hi.z
        const float lLx = fmaf(RTB_Q(q0, 0), ax, bx), lLy = fmaf(RTB_Q(q0, 1), ay, by), lLz = fmaf(RTB_Q(q0, 2), az, bz);
        const float lHx = fmaf(RTB_Q(q0, 3), ax, bx), lHy = fmaf(RTB_Q(q1, 0), ay, by), lHz = fmaf(RTB_Q(q1, 1), az, bz);
        const float rLx = fmaf(RTB_Q(q1, 2), ax, bx), rLy = fmaf(RTB_Q(q1, 3), ay, by), rLz = fmaf(RTB_Q(q2, 0), az, bz);
        const float rHx = fmaf(RTB_Q(q2, 1), ax, bx), rHy = fmaf(RTB_Q(q2, 2), ay, by), rHz = fmaf(RTB_Q(q2, 3), az, bz);
#undef RTB_Q
        const float lNear = fmaxf(fmaxf(fminf(lLx, lHx), fminf(lLy, lHy)), fminf(lLz, lHz));
        const float lFar = fminf(fminf(fmaxf(lLx, lHx), fmaxf(lLy, lHy)), fmaxf(lLz, lHz));
        const float rNear = fmaxf(fmaxf(fminf(rLx, rHx), fminf(rLy, rHy)), fminf(rLz, rHz));
        const float rFar = fminf(fminf(fmaxf(rLx, rHx), fmaxf(rLy, rHy)), fmaxf(rLz, rHz));
        bool passL = !((lFar - lNear) < tol), passR = !((rFar - rNear) < tol);
        if (CULL) {   // conservative planes in world space (a few ulp of slack is covered by the segment margin)
            const float Lx0 = fmaf(__uint2float_rn(q0 & 0xFFu), sx, rec.lo.x), Lx1 = fmaf(__uint2float_rn(q0 >> 24), sx, rec.lo.x);
            const float Ly0 = fmaf(__uint2float_rn((q0 >> 8) & 0xFFu), sy, rec.lo.y), Ly1 = fmaf(__uint2float_rn(q1 & 0xFFu), sy, rec.lo.y);
            const float Lz0 = fmaf(__uint2float_rn((q0 >> 16) & 0xFFu), sz, rec.lo.z), Lz1 = fmaf(__uint2float_rn((q1 >> 8) & 0xFFu), sz, rec.lo.z);
            const float Rx0 = fmaf(__uint2float_rn((q1 >> 16) & 0xFFu), sx, rec.lo.x), Rx1 = fmaf(__uint2float_rn((q2 >> 8) & 0xFFu), sx, rec.lo.x);
            const float Ry0 = fmaf(__uint2float_rn(q1 >> 24), sy, rec.lo.y), Ry1 = fmaf(__uint2float_rn((q2 >> 16) & 0xFFu), sy, rec.lo.y);
            const float Rz0 = fmaf(__uint2float_rn(q2 & 0xFFu), sz, rec.lo.z), Rz1 = fmaf(__uint2float_rn(q2 >> 24), sz, rec.lo.z);
            passL = passL && !(Lx0 > segHi.x || Lx1 < segLo.x || Ly0 > segHi.y || Ly1 < segLo.y || Lz0 > segHi.z || Lz1 < segLo.z);
            passR = passR && !(Rx0 > segHi.x || Rx1 < segLo.x || Ry0 > segHi.y || Ry1 < segLo.y || Rz0 > segHi.z || Rz1 < segLo.z);
        }
        const bool leafL = (w3 >> 24) & 1u, leafR = (w3 >> 25) & 1u;
        const uint32_t li = leafL ? leafOffset + split : split, ri = leafR ? leafOffset + split + 1u : split + 1u;
        const bool goR = passR && !leafR;
        const bool enqR = passR && leafR;
        const bool pushL = passL && goR;
        const bool enqL = passL && !goR && leafL;
        const bool goL = passL && !goR && !leafL;
        uint32_t tail = (qHead + qCount) & (QCAP - 1);
        if (enqR) sm.queue[tail][tid] = split + 1u;                       // primitive id of a leaf = its index - leafOffset
        tail = (tail + (enqR ? 1u : 0u)) & (QCAP - 1);
        if (enqL) sm.queue[tail][tid] = split;
        qCount += (enqR ? 1u : 0u) + (enqL ? 1u : 0u);
        if (pushL && sp < SSTACK) sm.stack[sp][tid] = li;
        if (pushL && sp >= SSTACK) {
            if (sp < STACK_DEPTH) lstack[sp - SSTACK] = li; else err |= 1u;
        }
        sp += (pushL && sp < STACK_DEPTH) ? 1 : 0;
        cur = goR ? ri : (goL ? li : 0xFFFFFFFFu);
    }
    const bool needPop = cur == 0xFFFFFFFFu && qCount < QCAP;
    if (needPop && sp == 0) travDone = true;
    const bool doPop = needPop && sp > 0;
    uint32_t e = 0;
    if (doPop && sp <= SSTACK) e = sm.stack[sp - 1][tid];
    if (doPop && sp > SSTACK) e = lstack[sp - 1 - SSTACK];
    sp -= doPop ? 1 : 0;
    const bool popLeaf = doPop && e >= leafOffset;
    if (popLeaf) sm.queue[(qHead + qCount) & (QCAP - 1)][tid] = e - leafOffset;
    qCount += popLeaf ? 1u : 0u;
    if (doPop && !popLeaf) cur = e;
}

#endif   // RTB_AB_KERNELS

// Traverse-phase turn over the 64-byte WIDE record (two 256-bit loads): up to four grandchild entries of binary node `cur` in
// the reference's visiting order, boxes conservative (8-bit, outward rounded) exactly as in wave_step_c, so one turn covers
// two levels of the binary tree.  The first surviving internal entry is descended into; surviving entries before it are
// leaves and are queued in order; surviving entries after it wait on the stack (pushed last-first so that they pop in order).
// Skipping the intermediate child's own box test is legal because its box contains its children's boxes (conservative).
// The plane bytes are stored one word per plane (byte e = entry e): the ray's direction signs pick the near / far word per
// axis once, so an entry costs six byte conversions, six FFMA and one 3-input max + one 3-input min.  (fma is monotone in the
// byte value, so near <= far holds exactly as with min / max of both planes; NaN planes -- a zero direction component never
// gets here, see exactOnly -- would only drop a constraint, which is conservative.)
template <bool CULL, int THREADS>
__device__ __forceinline__ void wave_step_w(const TraceScene& sc, WaveSmem<THREADS>& sm, const unsigned tid, const uint32_t leafOffset,
                                            const f3 o, const f3 rinv, uint32_t& cur, int& sp, const uint32_t qHead, uint32_t& qCount,
                                            bool& travDone, uint32_t* lstack, unsigned& err, const f3 segLo, const f3 segHi) {
    if (cur != 0xFFFFFFFFu) {
        const uint4* rp = sc.wide + 4ull * cur;
        const f8 h0 = ldg256(rp), h1 = ldg256(rp + 2);
        const uint32_t w3 = __float_as_uint(h0.lo.w);
        const uint32_t lox = __float_as_uint(h0.hi.x), loy = __float_as_uint(h0.hi.y), loz = __float_as_uint(h0.hi.z),
                       hix = __float_as_uint(h0.hi.w), hiy = __float_as_uint(h1.lo.x), hiz = __float_as_uint(h1.lo.y);
        const uint32_t id0 = __float_as_uint(h1.lo.z), id1 = __float_as_uint(h1.lo.w), id2 = __float_as_uint(h1.hi.x),
                       id3 = __float_as_uint(h1.hi.y);
        const float sx = __uint_as_float((w3 & 0xFFu) << 23), sy = __uint_as_float(((w3 >> 8) & 0xFFu) << 23),
                    sz = __uint_as_float(((w3 >> 16) & 0xFFu) << 23);
        const float ax = sx * rinv.x, ay = sy * rinv.y, az = sz * rinv.z;
        const float bx = (h0.lo.x - o.x) * rinv.x, by = (h0.lo.y - o.y) * rinv.y, bz = (h0.lo.z - o.z) * rinv.z;
        const float m = fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
        const float tol = -2.0e-6f * m;                                   // NaN / inf -> nothing is skipped
        const bool ngx = rinv.x < 0.0f, ngy = rinv.y < 0.0f, ngz = rinv.z < 0.0f;
        const uint32_t nX = ngx ? hix : lox, fX = ngx ? lox : hix;
        const uint32_t nY = ngy ? hiy : loy, fY = ngy ? loy : hiy;
        const uint32_t nZ = ngz ? hiz : loz, fZ = ngz ? loz : hiz;
        const uint32_t meta = w3 >> 24;
        uint32_t passMask = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
#define RTB_B(w) __uint2float_rn(((w) >> (8 * e)) & 0xFFu)                 // one I2F.U8 with a byte selector
            const float tn = fmaxf(fmaxf(fmaf(RTB_B(nX), ax, bx), fmaf(RTB_B(nY), ay, by)), fmaf(RTB_B(nZ), az, bz));
            const float tf = fminf(fminf(fmaf(RTB_B(fX), ax, bx), fmaf(RTB_B(fY), ay, by)), fmaf(RTB_B(fZ), az, bz));
            bool pass = !((tf - tn) < tol);
            if (CULL) {
                const float X0 = fmaf(RTB_B(lox), sx, h0.lo.x), Y0 = fmaf(RTB_B(loy), sy, h0.lo.y), Z0 = fmaf(RTB_B(loz), sz, h0.lo.z);
                const float X1 = fmaf(RTB_B(hix), sx, h0.lo.x), Y1 = fmaf(RTB_B(hiy), sy, h0.lo.y), Z1 = fmaf(RTB_B(hiz), sz, h0.lo.z);
                pass = pass && !(X0 > segHi.x || X1 < segLo.x || Y0 > segHi.y || Y1 < segLo.y || Z0 > segHi.z || Z1 < segLo.z);
            }
#undef RTB_B
            passMask |= pass ? (1u << e) : 0u;
        }
        passMask &= meta >> 4;                                            // entries that exist
        const uint32_t leafMask = meta & 0xFu;
        const uint32_t intMask = passMask & ~leafMask;
        const uint32_t first = intMask & (0u - intMask);                  // first surviving internal entry (one-hot, or 0)
        const uint32_t before = (first - 1u) & 0xFu;                      // entries ahead of it (all four when there is none)
        const uint32_t enqMask = passMask & leafMask & before;            // leaves ahead of it: test them first, in order
        const uint32_t pushMask = passMask & ~before & ~first;            // everything behind it waits on the stack
        // FIFO append.  trace_wave.cu drains every lane's FIFO in the L phase and rewinds it, so during T the head is slot 0
        // and the tail is simply qCount (the turn's precondition qCount <= QCAP - 4 leaves room for four appends).
        { const bool en = enqMask & 1u; if (en) sm.queue[qCount][tid] = id0 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 2u; if (en) sm.queue[qCount][tid] = id1 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 4u; if (en) sm.queue[qCount][tid] = id2 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 8u; if (en) sm.queue[qCount][tid] = id3 - leafOffset; qCount += en ? 1u : 0u; }
        if (sp <= SSTACK - 3) {                                           // the usual case: three predicated shared-memory stores
            { const bool pu = pushMask & 8u; if (pu) sm.stack[sp][tid] = id3; sp += pu ? 1 : 0; }
            { const bool pu = pushMask & 4u; if (pu) sm.stack[sp][tid] = id2; sp += pu ? 1 : 0; }
            { const bool pu = pushMask & 2u; if (pu) sm.stack[sp][tid] = id1; sp += pu ? 1 : 0; }
        } else {                                                          // deep: the levels beyond SSTACK live in local memory
            const uint32_t ids[3] = { id3, id2, id1 };
#pragma unroll 1
            for (int k = 0; k < 3; k++) {
                if (!((pushMask >> (3 - k)) & 1u)) continue;
                if (sp < SSTACK) sm.stack[sp][tid] = ids[k];
                else if (sp < STACK_DEPTH) lstack[sp - SSTACK] = ids[k];
                else { err |= 1u; continue; }
                sp++;
            }
        }
        uint32_t next = 0xFFFFFFFFu;
        next = (first & 1u) ? id0 : next;
        next = (first & 2u) ? id1 : next;
        next = (first & 4u) ? id2 : next;
        next = (first & 8u) ? id3 : next;
        cur = next;
    }
    // resume from the stack: a popped leaf only moves to the FIFO, so a lane gets up to RTB_WIDE_POPS predicated pops per turn
#pragma unroll
    for (int r = 0; r < RTB_WIDE_POPS; r++) {
        const bool needPop = cur == 0xFFFFFFFFu && qCount < QCAP;
        if (needPop && sp == 0) travDone = true;
        const bool doPop = needPop && sp > 0;
        uint32_t e = 0;
        if (doPop && sp <= SSTACK) e = sm.stack[sp - 1][tid];
        if (doPop && sp > SSTACK) e = lstack[sp - 1 - SSTACK];
        sp -= doPop ? 1 : 0;
        const bool popLeaf = doPop && e >= leafOffset;
        if (popLeaf) sm.queue[qCount][tid] = e - leafOffset;
        qCount += popLeaf ? 1u : 0u;
        if (doPop && !popLeaf) cur = e;
    }
}

// Nearest-first, t-culled turn over the same 64-byte records (NODES == 3).  Which leaves the reference tests does not depend on
// the order or on hits (no t-interval in its box test), but which of them can still CHANGE the result does: a primitive whose
// accepted hit would lie beyond the closest hit so far, or before tMin, cannot.  The record boxes are grown by the hit-point
// slack eta (pack_wide_kernel), so an accepted hit at parameter t lies inside every grown ancestor box and t is inside that
// box's slab interval; an entry is dropped when its interval, after granting the evaluation error |tol|, lies entirely beyond
// `closest` or before tMin.  Entries are taken nearest-first (the nearest surviving internal entry is descended, the others
// wait on the stack, all surviving leaves are queued); the L phase resolves equal-t ties like the reference's order would
// (leaf_test_unordered).  closest only shrinks, so a dropped entry stays irrelevant.
template <bool SORTED_PUSH, int THREADS>
__device__ __forceinline__ void wave_step_u(const TraceScene& sc, WaveSmem<THREADS>& sm, const unsigned tid, const f3 o, const f3 rinv,
                                            const float closest, const float tMinRay, uint32_t& cur, int& sp, uint32_t& qCount,
                                            bool& travDone, uint32_t* lstack, unsigned& err, const uint32_t leafOffset, const uint4* smTop = nullptr) {
    if (cur != 0xFFFFFFFFu) {
#ifdef RTB_SMEM_TOP   // A/B variant: records of the top levels come from the CTA's shared-memory copy (ids tagged with bit 31)
        f8 h0, h1;
        if (cur & 0x80000000u) {
            const float4* r = reinterpret_cast<const float4*>(smTop) + 4u * (cur & 0x7FFFFFFFu);
            h0.lo = r[0]; h0.hi = r[1]; h1.lo = r[2]; h1.hi = r[3];
        } else {
            const uint4* rp = sc.wide + 4ull * cur;
            h0 = ldg256(rp); h1 = ldg256(rp + 2);
        }
#else
        const uint4* rp = sc.wide + 4ull * cur;
        const f8 h0 = ldg256(rp), h1 = ldg256(rp + 2);
#endif
        const uint32_t w3 = __float_as_uint(h0.lo.w);
        const uint32_t lox = __float_as_uint(h0.hi.x), loy = __float_as_uint(h0.hi.y), loz = __float_as_uint(h0.hi.z),
                       hix = __float_as_uint(h0.hi.w), hiy = __float_as_uint(h1.lo.x), hiz = __float_as_uint(h1.lo.y);
        const uint32_t id0 = __float_as_uint(h1.lo.z), id1 = __float_as_uint(h1.lo.w), id2 = __float_as_uint(h1.hi.x),
                       id3 = __float_as_uint(h1.hi.y);
        const float sx = __uint_as_float((w3 & 0xFFu) << 23), sy = __uint_as_float(((w3 >> 8) & 0xFFu) << 23),
                    sz = __uint_as_float(((w3 >> 16) & 0xFFu) << 23);
        const float ax = sx * rinv.x, ay = sy * rinv.y, az = sz * rinv.z;
        const float bx = (h0.lo.x - o.x) * rinv.x, by = (h0.lo.y - o.y) * rinv.y, bz = (h0.lo.z - o.z) * rinv.z;
        const float m = fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
        const float tol = -2.0e-6f * m;                                   // NaN / inf -> nothing is dropped
        const float farLimit = closest - tol, nearLimit = tMinRay + tol;  // drop when tn > closest + |tol| or tf < tMin - |tol|
        const bool ngx = rinv.x < 0.0f, ngy = rinv.y < 0.0f, ngz = rinv.z < 0.0f;
        const uint32_t nX = ngx ? hix : lox, fX = ngx ? lox : hix;
        const uint32_t nY = ngy ? hiy : loy, fY = ngy ? loy : hiy;
        const uint32_t nZ = ngz ? hiz : loz, fZ = ngz ? loz : hiz;
        const uint32_t meta = w3 >> 24;
        uint32_t passMask = 0;
        float tnE[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
#define RTB_B(w) __uint2float_rn(((w) >> (8 * e)) & 0xFFu)
            const float tn = fmaxf(fmaxf(fmaf(RTB_B(nX), ax, bx), fmaf(RTB_B(nY), ay, by)), fmaf(RTB_B(nZ), az, bz));
            const float tf = fminf(fminf(fmaf(RTB_B(fX), ax, bx), fmaf(RTB_B(fY), ay, by)), fmaf(RTB_B(fZ), az, bz));
#undef RTB_B
            const bool pass = !((tf - tn) < tol) && !(tn > farLimit) && !(tf < nearLimit);
            passMask |= pass ? (1u << e) : 0u;
            tnE[e] = tn;
        }
        passMask &= meta >> 4;
        const uint32_t leafMask = meta & 0xFu;
        const uint32_t intMask = passMask & ~leafMask;
        // nearest surviving internal entry (one-hot; a NaN tn never compares equal -> falls back to the lowest bit)
        const float INF = __int_as_float(0x7f800000);
        const float k0 = (intMask & 1u) ? tnE[0] : INF, k1 = (intMask & 2u) ? tnE[1] : INF, k2 = (intMask & 4u) ? tnE[2] : INF,
                    k3 = (intMask & 8u) ? tnE[3] : INF;
        const float kmin = fminf(fminf(k0, k1), fminf(k2, k3));
        uint32_t first = ((intMask & 1u) && k0 == kmin) ? 1u : ((intMask & 2u) && k1 == kmin) ? 2u : ((intMask & 4u) && k2 == kmin) ? 4u
                       : ((intMask & 8u) && k3 == kmin) ? 8u : 0u;
        if (first == 0u) first = intMask & (0u - intMask);
        const uint32_t enqMask = passMask & leafMask;                     // every surviving leaf is a candidate now
        const uint32_t pushMask = intMask & ~first;
        { const bool en = enqMask & 1u; if (en) sm.queue[qCount][tid] = id0 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 2u; if (en) sm.queue[qCount][tid] = id1 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 4u; if (en) sm.queue[qCount][tid] = id2 - leafOffset; qCount += en ? 1u : 0u; }
        { const bool en = enqMask & 8u; if (en) sm.queue[qCount][tid] = id3 - leafOffset; qCount += en ? 1u : 0u; }
        if (SORTED_PUSH && sp <= SSTACK - 3) {
            // farthest first, so that the nearest waiting entry is popped first: slot of entry e = number of waiting entries
            // that are farther (ties: higher index counts as farther)
            const uint32_t m0 = pushMask & 1u, m1 = (pushMask >> 1) & 1u, m2 = (pushMask >> 2) & 1u, m3 = (pushMask >> 3) & 1u;
            const uint32_t c01 = tnE[1] >= tnE[0], c02 = tnE[2] >= tnE[0], c03 = tnE[3] >= tnE[0], c12 = tnE[2] >= tnE[1],
                           c13 = tnE[3] >= tnE[1], c23 = tnE[3] >= tnE[2];
            // every present pair (e < f) adds one to exactly one of the two slots, so the slots are a permutation even with NaNs
            const uint32_t r0 = (m1 & c01) + (m2 & c02) + (m3 & c03);
            const uint32_t r1 = (m0 & (c01 ^ 1u)) + (m2 & c12) + (m3 & c13);
            const uint32_t r2 = (m0 & (c02 ^ 1u)) + (m1 & (c12 ^ 1u)) + (m3 & c23);
            const uint32_t r3 = (m0 & (c03 ^ 1u)) + (m1 & (c13 ^ 1u)) + (m2 & (c23 ^ 1u));
            if (pushMask & 1u) sm.stack[sp + r0][tid] = id0;
            if (pushMask & 2u) sm.stack[sp + r1][tid] = id1;
            if (pushMask & 4u) sm.stack[sp + r2][tid] = id2;
            if (pushMask & 8u) sm.stack[sp + r3][tid] = id3;
            sp += __popc(pushMask);
        } else if (sp <= SSTACK - 3) {
            { const bool pu = pushMask & 8u; if (pu) sm.stack[sp][tid] = id3; sp += pu ? 1 : 0; }
            { const bool pu = pushMask & 4u; if (pu) sm.stack[sp][tid] = id2; sp += pu ? 1 : 0; }
            { const bool pu = pushMask & 2u; if (pu) sm.stack[sp][tid] = id1; sp += pu ? 1 : 0; }
            { const bool pu = pushMask & 1u; if (pu) sm.stack[sp][tid] = id0; sp += pu ? 1 : 0; }
        } else {
#pragma unroll 1
            for (int k = 3; k >= 0; k--) {
                if (!((pushMask >> k) & 1u)) continue;
                const uint32_t v = k == 3 ? id3 : (k == 2 ? id2 : (k == 1 ? id1 : id0));
                if (sp < SSTACK) sm.stack[sp][tid] = v;
                else if (sp < WIDE_STACK_DEPTH) lstack[sp - SSTACK] = v;
                else { err |= 1u; continue; }
                sp++;
            }
        }
        uint32_t next = 0xFFFFFFFFu;
        next = (first & 1u) ? id0 : next;
        next = (first & 2u) ? id1 : next;
        next = (first & 4u) ? id2 : next;
        next = (first & 8u) ? id3 : next;
        cur = next;
    }
    const bool needPop = cur == 0xFFFFFFFFu;
    if (needPop && sp == 0) travDone = true;
    const bool doPop = needPop && sp > 0;
    uint32_t e = 0xFFFFFFFFu;
    if (doPop && sp <= SSTACK) e = sm.stack[sp - 1][tid];
    if (doPop && sp > SSTACK) e = lstack[sp - 1 - SSTACK];
    sp -= doPop ? 1 : 0;
    if (doPop) cur = e;
}

// the reference's box test on the EXACT box of leaf candidate g (compressed / wide traversal only)
__device__ __forceinline__ bool leaf_box_passes(const TraceScene& sc, const uint32_t g, const f3 o, const f3 d, const f3 rinv, const bool exactOnly) {
    const f8 b = ldg256(sc.leafBox + 2ull * g);
    return box_test(o, d, rinv, exactOnly, b.lo.x, b.lo.y, b.lo.z, b.hi.x, b.hi.y, b.hi.z);
}

}  // namespace rtb
