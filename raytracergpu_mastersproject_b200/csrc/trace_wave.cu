// trace_wave.cu -- the production S2 kernel: raytraceBVH.comp as a persistent, warp-coherent state machine.
//
// Why (profiles/r01_v1_*): the straightforward kernel (trace.cu: one lane = one pixel, nested sample / bounce /
// traversal loops) ran with 4.9 of 32 lanes active per issued instruction -- after the first diffuse bounce every lane
// of a warp is at a different depth of a different tree walk.  It was issue-bound at 15 % SIMT efficiency, not memory-bound.
//
// Design.  Every lane owns one pixel at a time and carries its whole path state in registers.  The warp alternates
// between three phases, each of which is one uniform instruction stream:
//   T  traverse : every lane that has a node to expand does ONE child-pair step (four LDG.128, two box tests, stack
//                 push / pop).  Leaves whose box passed are NOT tested here; their primitive ids go to a per-lane FIFO in
//                 shared memory.  This is legal because the reference's box test has no t-interval (raytraceBVH.comp:
//                 184-193): which leaves a ray visits, and in which order, does not depend on the hits found so far.
//   L  leaves   : lanes pop their FIFO in order and run the triangle / sphere test with the running closest-so-far, so
//                 the sequence of primitive tests per ray -- and every tie-break (pin U9) -- is exactly the reference's.
//   S  shade    : lanes whose ray is finished shade it (emit / diffuse scatter), start the next bounce, the next sample of
//                 the pixel, or pull the next pixel from a global counter (lane-granular persistent threads).
// Warp votes (__ballot_sync) decide the phase switches: T runs while at least T_MIN lanes can step; a warp leaves T
// early only when other lanes are waiting for L or S.
//
// Box test: t = (bound - origin) * (1/dir) with the per-ray reciprocal, plus an error filter.  Each product is within
// 3 ulp of the reference's correctly rounded quotient (bound - origin) / dir, so when |tFar - tNear| exceeds
// 5e-7 * (|tNear| + |tFar|) the comparison tNear < tFar provably has the reference's outcome; otherwise (and for rays
// with a zero / denormal direction component, where the reference produces inf / NaN) the exact division path of
// trace_common.cuh runs.
// Results are bit-identical either way (tests/test_gpu_parity.py compares images, hit ids, RNG states and visit counters).
#include "kernels.h"
#include "trace_common.cuh"

namespace rtb {

constexpr int WAVE_THREADS = 128;
constexpr int WAVE_MIN_BLOCKS = 6;
constexpr int SSTACK = 32;                // stack entries kept in shared memory; deeper levels spill to local memory
constexpr int QCAP = 8;                   // pending-leaf FIFO entries per lane
constexpr int T_MIN_DEFAULT = 20;         // leave the traverse phase when fewer lanes than this can step

struct __align__(16) WaveSmem {
    uint32_t stack[SSTACK][WAVE_THREADS];
    uint32_t queue[QCAP][WAVE_THREADS];
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
    // = 2.4e-7 * (...).  The filter uses 5e-7 (2x margin) plus an absolute term for the subnormal range.
    const float e = fmaf(fabsf(tNear) + fabsf(tFar), 5.0e-7f, 1.0e-36f);
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

template <bool COUNT, bool EXT>
__global__ void __launch_bounds__(WAVE_THREADS, WAVE_MIN_BLOCKS) trace_wave_kernel(const TraceParams p) {
    __shared__ WaveSmem sm;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const TraceScene& sc = p.sc;
    const uint32_t leafOffset = sc.N - 1;
    const uint32_t totalWork = p.tilesX * p.tilesY * (TILE_W * TILE_H);
    const float T_MIN_RAY = 0.001f, T_MAX_RAY = 10000000.0f;                    // sceneHit :268-269

    // ---- per-lane state -------------------------------------------------------------------------------------------
    bool dead = false, havePixel = false, rayActive = false, travDone = true, exactOnly = false, hit = false;
    uint32_t x = 0, y = 0, base = 0, k = 0, depth = 0, rng = 0;
    size_t px = 0;
    float alpha = 0.f, nextRandom = 0.f;
    f3 rgb = F3(0, 0, 0), color = F3(0, 0, 0), att = F3(1, 1, 1), primDir = F3(0, 0, 1);
    f3 o = F3(0, 0, 0), d = F3(0, 0, 1), rinv = F3(0, 0, 0);
    float closest = T_MAX_RAY;
    Hit rec; rec.t = 0.f; rec.normal = F3(0, 0, 0); rec.mat = 0; rec.prim = 0; rec.back = 0;
    uint32_t cur = 0xFFFFFFFFu;           // internal node to expand next, or NONE
    int sp = 0;
    uint32_t qHead = 0, qCount = 0;
    uint32_t lstack[STACK_DEPTH - SSTACK];
    unsigned err = 0;
    Tally tl = { 0, 0, 0, 0, 0 };
    unsigned long long samplesDone = 0;

    auto push = [&](uint32_t v) {
        if (sp < SSTACK) sm.stack[sp][tid] = v;
        else if (sp < STACK_DEPTH) lstack[sp - SSTACK] = v;
        else { err |= 1u; return; }
        sp++;
    };
    auto pop = [&]() -> uint32_t {
        sp--;
        return sp < SSTACK ? sm.stack[sp][tid] : lstack[sp - SSTACK];
    };
    auto enqueue = [&](uint32_t g) {
        sm.queue[(qHead + qCount) & (QCAP - 1)][tid] = g;
        qCount++;
    };

    while (true) {
        // =========================================== S: shade / generate =========================================
        while (!dead && (!rayActive || (travDone && qCount == 0))) {
            if (rayActive) {                                              // the ray is finished: rayColor loop body :283-307
                rayActive = false;
                if (k == 0 && depth == 0 && p.hitPrim) {
                    p.hitPrim[px] = hit ? rec.prim : 0xFFFFFFFFu;
                    if (p.hitT) p.hitT[px] = hit ? rec.t : 0.0f;
                }
                bool pathEnd = true;
                if (!hit) {
                    color = color + F3(0.f, 0.f, 0.f) * att;              // _BACKGROUND_COLOR * globalAttenuation :284
                } else {
                    const float4 m = __ldg(sc.mats + rec.mat);
                    if (COUNT) tl.mat++;
                    const uint32_t type = __float_as_uint(m.w);
                    const f3 albedo = xyz(m);
                    const f3 emitted = (type == RTB_LIGHT) ? albedo : F3(0.f, 0.f, 0.f);      // emitted :94-99
                    color = color + emitted * att;                                             // :293
                    if (type == RTB_DIFFUSE) {                                                 // scatter :100-115
                        const f3 P = o + rec.t * d;
                        const f3 nd = normalize(rec.normal + random_unit_vector(rng));
                        o = P; d = nd;
                        att = att * albedo;
                        pathEnd = false;
                    } else if (EXT && type != RTB_LIGHT) {
                        f3 a2, nd;
                        const f3 P = o + rec.t * d;
                        if (scatter_extension(type, albedo, d, rec, rng, a2, nd)) { o = P; d = nd; att = att * a2; pathEnd = false; }
                    }
                    depth++;
                    if (depth >= p.maxDepth) pathEnd = true;               // for (i < maxRayTraceDepth) :282
                }
                if (pathEnd) {                                            // main() :372-374 for this sample
                    rgb = color + rgb;
                    alpha = nextRandom;
                    k++;
                    depth = 0xFFFFFFFFu;                                  // marks "start a new sample"
                    if (k == p.sampleCount) {
                        p.image[px] = make_float4(rgb.x, rgb.y, rgb.z, alpha);
                        if (p.rngOut) p.rngOut[px] = rng;
                        if (COUNT) samplesDone += p.sampleCount;
                        havePixel = false;
                    }
                }
            } else {
                depth = 0xFFFFFFFFu;
            }
            if (depth == 0xFFFFFFFFu) {                                   // need a primary ray
                if (!havePixel) {                                         // lane-granular persistent threads: next pixel
                    const unsigned act = __activemask();
                    const int leader = __ffs(act) - 1;
                    uint32_t idx = 0;
                    if ((int)lane == leader) idx = atomicAdd(p.workCounter, (unsigned)__popc(act));
                    idx = __shfl_sync(act, idx, leader) + __popc(act & ((1u << lane) - 1u));
                    if (idx >= totalWork) { dead = true; break; }
                    const uint32_t tile = idx / (TILE_W * TILE_H), within = idx % (TILE_W * TILE_H);
                    const uint32_t tx = tile % p.tilesX, ty = tile / p.tilesX;
                    x = tx * TILE_W + (within % TILE_W);
                    const uint32_t j = ty * TILE_H + (within / TILE_W);
                    y = ((j / p.bandRows) * p.bandStep + p.bandFirst) * p.bandRows + (j % p.bandRows);
                    if (x >= p.W || j >= p.localRows || y >= p.H) continue;               // padding of the tile grid
                    px = (size_t)j * p.W + x;
                    const float4 c = p.image[px];                                          // imageLoad :349
                    base = (600u * x + y) * (p.randomState + 1u);                          // random.glsl:10
                    alpha = c.w;
                    for (uint32_t s = 0; s < p.sampleSkip; s++) {                          // fast-forward the seed chain
                        uint32_t t = base + alpha_to_u32(alpha);
                        alpha = pcg_float(t);
                    }
                    // getRay :329-342 (no jitter: the same for every sample) + rayColor's own normalize :280
                    const f3 pixelSample = (p.cam.pixel00 + (float)x * p.cam.deltaU) + (float)y * p.cam.deltaV;
                    primDir = normalize(normalize(pixelSample - p.cam.origin));
                    rgb = F3(c.x, c.y, c.z);
                    k = 0;
                    // The primary ray is the same for every sample of the pixel, so its test against the root box is loop
                    // invariant.  If it fails (or the bounce loop is empty), every sample is: seed, one random(), colour 0.
                    bool rootPass = false;
                    if (p.maxDepth != 0) {
                        const f3 ri = F3(1.0f / primDir.x, 1.0f / primDir.y, 1.0f / primDir.z);
                        const bool ex = !(fabsf(ri.x) < 3.0e38f && fabsf(ri.y) < 3.0e38f && fabsf(ri.z) < 3.0e38f);
                        const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
                        rootPass = box_test(p.cam.origin, primDir, ri, ex, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z);
                    }
                    if (!rootPass) {
                        for (uint32_t s = 0; s < p.sampleCount; s++) {
                            rng = base + alpha_to_u32(alpha);                              // :350
                            alpha = pcg_float(rng);                                        // nextRandom :352 -> alpha :372
                            const f3 col = F3(0.f, 0.f, 0.f) + F3(0.f, 0.f, 0.f) * F3(1.f, 1.f, 1.f);   // :278-279,284
                            rgb = col + rgb;
                        }
                        if (COUNT) { samplesDone += p.sampleCount; if (p.maxDepth != 0) { tl.rays += p.sampleCount; tl.visits += p.sampleCount; } }
                        p.image[px] = make_float4(rgb.x, rgb.y, rgb.z, alpha);
                        if (p.rngOut) p.rngOut[px] = rng;
                        if (p.hitPrim) { p.hitPrim[px] = 0xFFFFFFFFu; if (p.hitT) p.hitT[px] = 0.0f; }
                        continue;                                                          // next pixel
                    }
                    havePixel = true;
                }
                rng = base + alpha_to_u32(alpha);                                          // :350 (stepRNG :351 is a no-op)
                nextRandom = pcg_float(rng);                                               // :352
                color = F3(0.f, 0.f, 0.f); att = F3(1.f, 1.f, 1.f);
                depth = 0;
                o = p.cam.origin; d = primDir;
            }
            // ---- start the ray (hitBVH prologue :196-201 + the root's own box test) ----
            rayActive = true; hit = false; closest = T_MAX_RAY;
            sp = 0; qHead = 0; qCount = 0; cur = 0xFFFFFFFFu; travDone = true;
            rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            exactOnly = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
            if (COUNT) { tl.rays++; tl.visits++; }
            const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
            if (box_test(o, d, rinv, exactOnly, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) {
                if (sc.N == 1) enqueue(0u);                                                // the root is the only leaf
                else { cur = 0; travDone = false; }
            }
        }
        if (__all_sync(FULL, dead)) break;

        // =========================================== T: traverse ================================================
        while (true) {
            const bool can = rayActive && !travDone && qCount <= QCAP - 2;
            const unsigned bal = __ballot_sync(FULL, can);
            if (bal == 0) break;
            if (__popc(bal) < (int)p.tMin) {
                const bool waiting = !dead && !can;                       // lanes that L or S could put back to work
                if (__any_sync(FULL, waiting)) break;
            }
            if (can) {
                if (cur != 0xFFFFFFFFu) {
                    const float4* pr = sc.pairs + 4ull * cur;
                    const float4 lLo = __ldg(pr), lHi = __ldg(pr + 1), rLo = __ldg(pr + 2), rHi = __ldg(pr + 3);
                    if (COUNT) tl.visits += 2;
                    const uint32_t li = __float_as_uint(lLo.w), ri = __float_as_uint(lHi.w);
                    int fR = box_filter(o, rinv, rLo.x, rLo.y, rLo.z, rHi.x, rHi.y, rHi.z);
                    int fL = box_filter(o, rinv, lLo.x, lLo.y, lLo.z, lHi.x, lHi.y, lHi.z);
                    if (exactOnly || (fR | fL) < 0) {                      // rare: some comparison is too close to call
                        fR = box_hit(o, d, rLo.x, rLo.y, rLo.z, rHi.x, rHi.y, rHi.z) ? 1 : 0;
                        fL = box_hit(o, d, lLo.x, lLo.y, lLo.z, lHi.x, lHi.y, lHi.z) ? 1 : 0;
                    }
                    // Reference order (:241-244): the right subtree completely, then the left.  Straight-line bookkeeping:
                    const bool passR = fR != 0, passL = fL != 0;
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
                    if (pushL) push(li);
                    cur = goR ? ri : (goL ? li : 0xFFFFFFFFu);
                }
                while (cur == 0xFFFFFFFFu && qCount < QCAP) {              // resume from the stack (usually 0 or 1 turns)
                    if (sp == 0) { travDone = true; break; }
                    const uint32_t e = pop();
                    if (e >= leafOffset) enqueue(e - leafOffset); else cur = e;
                }
            }
        }

        // =========================================== L: leaf tests ==============================================
        while (true) {
            const bool has = qCount > 0;
            if (!__any_sync(FULL, has)) break;
            if (has) {
                const uint32_t g = sm.queue[qHead][tid];
                qHead = (qHead + 1) & (QCAP - 1);
                qCount--;
                leaf_test<COUNT>(sc, g, o, d, T_MIN_RAY, closest, hit, rec, tl);
            }
        }
    }

    if (err) atomicOr(p.errFlag, err);
    if (COUNT) {
        unsigned long long v[6] = { tl.rays, tl.visits, tl.tri, tl.sph, tl.mat, samplesDone };
#pragma unroll
        for (int i = 0; i < 6; i++) {
            unsigned long long s = v[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
            if (lane == 0 && s) atomicAdd(p.counters + i, s);
        }
    }
}

static int wave_blocks_per_sm(bool count, bool ext) {
    int nb = 0;
    if (count) {
        if (ext) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<true, true>, WAVE_THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<true, false>, WAVE_THREADS, 0);
    } else {
        if (ext) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<false, true>, WAVE_THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<false, false>, WAVE_THREADS, 0);
    }
    return nb > 0 ? nb : 1;
}

void launch_trace_wave(cudaStream_t st, TraceParams p, bool count, bool ext, int smCount) {
    p.tilesX = (p.W + TILE_W - 1) / TILE_W;
    p.tilesY = (p.localRows + TILE_H - 1) / TILE_H;
    const uint64_t numWarps = (uint64_t)p.tilesX * p.tilesY;
    uint64_t grid = (uint64_t)smCount * wave_blocks_per_sm(count, ext);     // persistent: resident CTAs per SM x SMs
    const uint64_t need = (numWarps + WAVE_THREADS / 32 - 1) / (WAVE_THREADS / 32);
    if (grid > need) grid = need;
    if (grid == 0) return;
    if (p.tMin == 0) p.tMin = T_MIN_DEFAULT;
    cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), st);
    if (count) {
        if (ext) trace_wave_kernel<true, true><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
        else trace_wave_kernel<true, false><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
    } else {
        if (ext) trace_wave_kernel<false, true><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
        else trace_wave_kernel<false, false><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
    }
}

}  // namespace rtb
