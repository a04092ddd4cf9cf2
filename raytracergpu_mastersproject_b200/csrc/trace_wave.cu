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
//   S  shade    : lanes whose ray is finished shade it (emit / diffuse scatter), start the next bounce, or pull the next
//                 (pixel, sample) work item from a global counter (lane-granular persistent threads).
//
// Work items are (pixel, sample), not pixels: a pixel's samples are sequential only through two scalars -- the alpha seed
// chain (raytraceBVH.comp:349-352,372), which needs no tracing to advance, and the fp32 running sum.  A pre-pass walks the
// chain of every pixel and parks each sample's seed in a per-sample slot; the trace kernel fills the slot with that
// sample's colour; an accumulate pass then adds the slots in sample order (colour_k + sum, exactly the shader's
// association), so the image is bit-identical to sequential dispatches while no lane is ever tied to a pixel for more
// than one sample (a hit pixel's 64 samples used to keep one lane busy for ~30 ms: a 70 ms kernel tail at 1080p).
// Pixels whose primary ray misses the root box (the same ray for every sample: no jitter) never reach the trace kernel.
// Warp votes (__ballot_sync) decide the phase switches: T runs while at least T_MIN lanes can step; a warp leaves T
// early only when other lanes are waiting for L or S.
//
// Box test: t = (bound - origin) * (1/dir) with the per-ray reciprocal, plus an error filter.  Each product is within
// 3 ulp of the reference's correctly rounded quotient (bound - origin) / dir, so when |tFar - tNear| exceeds
// 3e-7 * (|tNear| + |tFar|) the comparison tNear < tFar provably has the reference's outcome; otherwise (and for rays
// with a zero / denormal direction component, where the reference produces inf / NaN) the exact division path of
// trace_common.cuh runs.
// Results are bit-identical either way (tests/test_gpu_parity.py compares images, hit ids, RNG states and visit counters).
#include <stdio.h>
#include <stdlib.h>

#include "kernels.h"
#include "trace_wave_shared.cuh"

namespace rtb {

constexpr int WAVE_THREADS = 128;
// parked path / ray record, float4 units: 0 (o, rng) 1 (d, depth) 2 (colour, pixel) 3 (attenuation, sample)
// 4 (slot lo, slot hi, cur, flags: bit 0 ray in flight | bit 1 hit | sp << 8) 5 (rec.t, rec.normal) 6 (mat, prim, back, closest) 7..14 stack[32]
// 15 (.x = ready flag: the epoch of the launch that parked it, release-stored after the rest of the record)
constexpr unsigned PARK_STRIDE = 16;
#ifdef RTB_TAIL_PROBE   // debug build only (make EXTRA=-DRTB_TAIL_PROBE): per-warp timeline of the main launch, dumped to $RTB_TAIL_PROBE_FILE
__device__ uint4 g_probeLane[8192 * 32];             // per lane: its longest ray (T steps, pixel, sample | depth << 16 | exact << 31, L tests)
__device__ unsigned long long g_probe[3][8192];     // [0] first item pulled, [1] first lane retired (queue drained), [2] warp exit
__device__ __forceinline__ unsigned long long probe_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif
#ifndef RTB_WIDE_MINB
#define RTB_WIDE_MINB 7
#endif
constexpr int WAVE_MIN_BLOCKS = 6;
constexpr int WIDE_MIN_BLOCKS = RTB_WIDE_MINB;   // resident CTAs per SM for the 4-ary variant

constexpr uint32_t COOP_CAP = 512;
// the pending-set stack of a tail warp: a plain shared array (trace_tail_kernel) or the warp's own slice of the main kernel's
// [level][thread] traversal stack, free once all its lanes are dead (32 levels x 32 lanes = 1024 entries)
struct LinearStack {
    uint32_t* s;
    __device__ __forceinline__ uint32_t& operator[](uint32_t i) const { return s[i]; }
};
template <bool EXT, bool COUNT, class STK>
__device__ __forceinline__ void tail_item(const TraceParams& p, const uint32_t item, const STK stk, const unsigned lane, Tally& tl, unsigned& err);
// A/B build RTB_SMEM_TOP: stack entries of the main kernel may be table indices; the tail kernel walks global records only
__device__ __forceinline__ uint32_t untag_top(const TraceScene& sc, const uint32_t id) {
#ifdef RTB_SMEM_TOP
    if (id != 0xFFFFFFFFu && (id & 0x80000000u)) return __ldg(sc.topGlobal + (id & 0x7FFFFFFFu));
#endif
    return id;
}
// release-store of a park slot's ready flag: the record written before it is visible to whoever reads the flag (acquire side: the
// tail launch reads the flag, fences, then loads the record past L1)
__device__ __forceinline__ void publish_parked(unsigned int* flag, const unsigned int epoch) {
#ifdef RTB_SIMT_EMU
    *flag = epoch;
#else
    __threadfence();
    *flag = epoch;
#endif
}

// CULL (RTB_TRACE_CULLED, default OFF, NOT the reference's traversal): additionally skips children whose box lies outside
// the axis-aligned box of the ray SEGMENT [tMin, closest-so-far] grown by a safety margin.  The reference visits every
// box the ray's LINE crosses (raytraceBVH.comp:184-193 has no t-interval); a primitive it would accept has its hit point
// on that segment and (up to rounding) inside its own leaf box, so skipped subtrees cannot contain an accepted hit as long
// as the margin exceeds the rounding slack of the intersection tests.  That is an argument, not a proof (sliver triangles
// stretch the slack), hence a flag: tests/test_gpu_parity.py checks images and hit ids stay bit-identical on the test
// scenes and bench.py --mode culled reports it as a separate, labelled line.
// NODES: 0 exact 64-B pairs | 1 compressed 32-B pairs | 2 4-ary 64-B records, reference order | 3 4-ary records, nearest-first + t-culled
//        | 4 = 3 with the waiting entries stacked farthest-first (pays off on overlapping sphere fields: +16 % C3, -2..7 % C2 / C4 / C5)
template <bool COUNT, bool EXT, bool CULL, int NODES>
__global__ void __launch_bounds__(WAVE_THREADS, NODES >= 2 ? WIDE_MIN_BLOCKS : WAVE_MIN_BLOCKS) trace_wave_kernel(const TraceParams p) {
    __shared__ WaveSmem<WAVE_THREADS> sm;
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned tid = threadIdx.x;
#ifdef RTB_SMEM_TOP
    __shared__ uint4 smTop[NODES >= 3 ? RTB_SMEM_TOP * 4 : 1];
    if (NODES >= 3) { for (unsigned i = tid; i < RTB_SMEM_TOP * 4; i += WAVE_THREADS) smTop[i] = p.sc.top[i]; __syncthreads(); }
#define RTB_ROOT_RECORD() (NODES >= 3 ? 0x80000000u : 0u)          /* entry 0 of the table = the record the walk starts at */
#else
    const uint4* smTop = nullptr;
    // the record the walk starts at: the root's, or the first of the records that hold the hoisted big leaves (tt_top_kernel);
    // re-read per ray (an L1 hit) rather than held in a register: the kernel sits exactly at its 72-register budget
#define RTB_ROOT_RECORD() (NODES >= 3 ? __ldg(p.cullAllowed + 1) : 0u)
#endif
    const unsigned lane = tid & 31;
    const TraceScene& sc = p.sc;
    const uint32_t leafOffset = sc.N - 1;
    if (NODES >= 3 && p.doneWarps != nullptr && tid == 0) atomicAdd(p.doneWarps + 1, 1u);   // this CTA is resident (see trace_tail_kernel)
    const uint32_t activeCount = *p.activeCount;
    // item i -> group of 32 active pixels g = i / (32 * S), sample s = (i % (32 * S)) / 32, pixel slot g * 32 + i % 32:
    // the 32 items a warp pulls together are the same sample of 32 neighbouring pixels (coherent primary rays)
    const uint32_t groupItems = 32u * (p.primaryMode == 1u ? 1u : p.sampleCount);
    const uint64_t totalWork = (uint64_t)((activeCount + 31u) / 32u) * groupItems;
    const float T_MIN_RAY = 0.001f, T_MAX_RAY = 10000000.0f;                    // sceneHit :268-269
    constexpr bool CN = NODES != 0;                       // conservative internal boxes: leaves are re-checked exactly
    constexpr uint32_t Q_ROOM = NODES >= 2 ? 4u : 2u;     // free FIFO entries a turn may need
    const uint32_t qGate = NODES >= 3 ? min(p.qGate, QCAP - Q_ROOM) : QCAP - Q_ROOM;   // a lane steps while its FIFO holds <= qGate candidates
    const bool cullAllowed = NODES >= 3 ? (*p.cullAllowed != 0u) : false;   // records grown by a finite hit-point slack (pack_wide_kernel)

    // ---- per-lane state -------------------------------------------------------------------------------------------
    bool dead = false, rayActive = false, travDone = true, exactOnly = false, hit = false, cullOk = false;
    uint32_t depth = 0, rng = 0, pix = 0, smp = 0, turns = 0;
    size_t slotIndex = 0;
    f3 color = F3(0, 0, 0), att = F3(1, 1, 1);
    f3 o = F3(0, 0, 0), d = F3(0, 0, 1), rinv = F3(0, 0, 0);
    float closest = T_MAX_RAY;
    f3 segLo = F3(0, 0, 0), segHi = F3(0, 0, 0);      // CULL: box of the ray segment [tMin, closest] + margin
    auto update_segment = [&]() {
        const f3 a = o + T_MIN_RAY * d, b = o + closest * d;
        const float m = 0.05f + 1.0e-4f * fmaxf(fmaxf(fmaxf(fabsf(a.x), fabsf(b.x)), fmaxf(fabsf(a.y), fabsf(b.y))), fmaxf(fabsf(a.z), fabsf(b.z)));
        segLo = F3(fminf(a.x, b.x) - m, fminf(a.y, b.y) - m, fminf(a.z, b.z) - m);
        segHi = F3(fmaxf(a.x, b.x) + m, fmaxf(a.y, b.y) + m, fmaxf(a.z, b.z) + m);
    };
    Hit rec; rec.t = 0.f; rec.normal = F3(0, 0, 0); rec.mat = 0; rec.prim = 0; rec.back = 0;
    uint32_t cur = 0xFFFFFFFFu;           // internal node to expand next, or NONE
    int sp = 0;
    uint32_t qHead = 0, qCount = 0;
    uint32_t lstack[(NODES >= 2 ? WIDE_STACK_DEPTH : STACK_DEPTH) - SSTACK];
    unsigned err = 0;
    Tally tl = { 0, 0, 0, 0, 0 };

#ifdef RTB_TAIL_PROBE
    const unsigned pw = (blockIdx.x * WAVE_THREADS + tid) >> 5;
    bool probeDrained = false;
    uint32_t prSteps = 0, prLeaf = 0; uint4 prMax = make_uint4(0, 0, 0, 0);
    if (lane == 0 && pw < 8192) { g_probe[0][pw] = probe_now(); g_probe[1][pw] = 0; }
#endif
    auto enqueue = [&](uint32_t g) {
        sm.queue[(qHead + qCount) & (QCAP - 1)][tid] = g;
        qCount++;
    };

    while (true) {
        // Tail hand-over (NODES >= 3): lanes only die once the work queue is drained.  From then on a warp with few live lanes
        // stops starting rays: a path whose next ray is about to start is parked for trace_tail_kernel (one ray per WARP) and
        // the lane retires; rays already in flight are finished here.
        bool tailMode = false;
        if (NODES >= 3 && p.coopMax != 0u && p.primaryMode != 1u) {
            const unsigned deadBal = __ballot_sync(FULL, dead);
            tailMode = deadBal != 0u && 32u - (unsigned)__popc(deadBal) <= p.coopMax;
        }
        // Long-ray hand-over: every FIFO is empty here (the L phase drains them), so a ray's whole pending set is cur + its stack.
        if (NODES >= 3 && p.coopTurns != 0u && rayActive && !travDone) {
            turns++;
            if (turns > p.coopTurns && !exactOnly && sp <= SSTACK) {
                const uint32_t k = atomicAdd(p.parkCount, 1u);
                if (k < p.parkCapacity) {
                    float4* e = p.parkBuf + (size_t)PARK_STRIDE * k;
                    e[0] = make_float4(o.x, o.y, o.z, __uint_as_float(rng));
                    e[1] = make_float4(d.x, d.y, d.z, __uint_as_float(depth));
                    e[2] = make_float4(color.x, color.y, color.z, __uint_as_float(pix));
                    e[3] = make_float4(att.x, att.y, att.z, __uint_as_float(smp));
                    e[4] = make_float4(__uint_as_float((uint32_t)slotIndex), __uint_as_float((uint32_t)((unsigned long long)slotIndex >> 32)),
                                       __uint_as_float(cur), __uint_as_float(1u | (hit ? 2u : 0u) | ((uint32_t)sp << 8)));
                    e[5] = make_float4(rec.t, rec.normal.x, rec.normal.y, rec.normal.z);
                    e[6] = make_float4(__uint_as_float(rec.mat), __uint_as_float(rec.prim), __uint_as_float((uint32_t)rec.back), closest);
#pragma unroll
                    for (int j = 0; j < SSTACK / 4; j++)
                        e[7 + j] = make_float4(__uint_as_float(sm.stack[4 * j][tid]), __uint_as_float(sm.stack[4 * j + 1][tid]),
                                               __uint_as_float(sm.stack[4 * j + 2][tid]), __uint_as_float(sm.stack[4 * j + 3][tid]));
                    rayActive = false; travDone = true; sp = 0; cur = 0xFFFFFFFFu;      // the lane is free for the next item
                    publish_parked((unsigned int*)(e + 15), p.parkEpoch);                  // the concurrent tail launch may take it from here
                    if (COUNT) tl.parked++;
                }
            }
        }
        // =========================================== S: shade / generate =========================================
        // The S phase costs the warp ~450 instructions however few lanes need it (5-8 on average when every finished ray is served at
        // once: a quarter of all instructions once the traverse phase got short).  It is therefore entered only when at least sMin lanes
        // wait for it -- or nothing else can run (no lane can step; the T loop below keeps stepping meanwhile).
        bool runS = true;
        if (p.sMin > 1u) {
            const unsigned sBal = __ballot_sync(FULL, !dead && (!rayActive || (travDone && qCount == 0)));
            runS = (unsigned)__popc(sBal) >= p.sMin || !__any_sync(FULL, rayActive && !travDone && qCount <= qGate);
        }
        while (runS && !dead && (!rayActive || (travDone && qCount == 0))) {
            bool needItem = !rayActive;
            if (rayActive && p.primaryMode == 1u) {                       // primary-hit launch: keep the hit record, no shading
                rayActive = false; needItem = true;
                float4* h = p.primaryHits + 3ull * pix;
                h[0] = make_float4(rec.t, rec.normal.x, rec.normal.y, rec.normal.z);
                h[1] = make_float4(__uint_as_float(rec.prim), __uint_as_float(rec.mat), __uint_as_float(hit ? 1u : 0u),
                                   __uint_as_float((uint32_t)rec.back));
                h[2] = make_float4(d.x, d.y, d.z, 0.f);
            } else if (rayActive) {                                       // the ray is finished: rayColor loop body :283-307
                rayActive = false;
#ifdef RTB_TAIL_PROBE
                if (prSteps > prMax.x) prMax = make_uint4(prSteps, pix, smp | (depth << 16) | (exactOnly ? 0x80000000u : 0u), prLeaf);
                prSteps = 0; prLeaf = 0;
#endif
                if (depth == 0 && smp == 0 && p.firstPass && p.hitPrim) {
                    p.hitPrim[pix] = hit ? rec.prim : 0xFFFFFFFFu;
                    if (p.hitT) p.hitT[pix] = hit ? rec.t : 0.0f;
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
                if (pathEnd) {                                            // this sample's pixelColor is final
                    float4* e = p.sampleBuf + slotIndex;
                    e->x = color.x; e->y = color.y; e->z = color.z;       // .w keeps the sample's incoming alpha
                    if (p.rngOut && p.lastPass && smp + 1 == p.sampleCount) p.rngOut[pix] = rng;
                    needItem = true;
                    if (COUNT) tl.paths++;
                }
            }
            if (needItem) {                                               // lane-granular persistent threads: next (pixel, sample)
                const unsigned act = __activemask();
                const int leader = __ffs(act) - 1;
                unsigned long long i = 0;
                if ((int)lane == leader) i = atomicAdd(p.workCounter64, (unsigned long long)__popc(act));
                i = __shfl_sync(act, i, leader) + (unsigned long long)__popc(act & ((1u << lane) - 1u));
                if (i >= totalWork) { dead = true; break; }
                uint32_t g, r;
                if (totalWork <= 0xFFFFFFFFull) { g = (uint32_t)i / groupItems; r = (uint32_t)i - g * groupItems; }   // warp-uniform branch
                else { g = (uint32_t)(i / groupItems); r = (uint32_t)(i % groupItems); }
                const uint32_t slot = g * 32u + (r & 31u);
                if (slot >= activeCount) continue;                        // padding of the last group
                if (COUNT) tl.items++;
                smp = r >> 5;
                pix = p.activePix[slot];
                slotIndex = (size_t)smp * p.slotCapacity + slot;
                const uint32_t xy = p.activeXY[slot];
                const uint32_t x = xy & 0xFFFFu, y = xy >> 16;
                const float alphaIn = p.sampleBuf[slotIndex].w;
                rng = (600u * x + y) * (p.randomState + 1u) + alpha_to_u32(alphaIn);       // random.glsl:10 + :350 (:351 is a no-op)
                (void)pcg_float(rng);                                                      // nextRandom :352 (kept by the pre-pass)
                color = F3(0.f, 0.f, 0.f); att = F3(1.f, 1.f, 1.f);
                depth = 0;
                o = p.cam.origin;
                if (p.primaryMode == 2u) {                                // the primary ray's hitBVH result, traced once per pixel
                    const float4 h0 = __ldg(p.primaryHits + 3ull * pix), h1 = __ldg(p.primaryHits + 3ull * pix + 1), h2 = __ldg(p.primaryHits + 3ull * pix + 2);
                    d = xyz(h2);                                          // primary_direction(p, x, y), as the primary launch evaluated it
                    rec.t = h0.x; rec.normal = F3(h0.y, h0.z, h0.w);
                    rec.prim = __float_as_uint(h1.x); rec.mat = __float_as_uint(h1.y);
                    hit = __float_as_uint(h1.z) != 0u; rec.back = (int)__float_as_uint(h1.w);
                    rayActive = true; travDone = true; qCount = 0; qHead = 0; sp = 0; cur = 0xFFFFFFFFu;
                    continue;                                             // straight to shading
                }
                d = primary_direction(p, x, y);
            }
            if (NODES >= 3 && tailMode) {                                 // park the path (state at a ray boundary) and retire
                const uint32_t k = atomicAdd(p.parkCount, 1u);
                if (k < p.parkCapacity) {
                    float4* e = p.parkBuf + (size_t)PARK_STRIDE * k;
                    e[0] = make_float4(o.x, o.y, o.z, __uint_as_float(rng));
                    e[1] = make_float4(d.x, d.y, d.z, __uint_as_float(depth));
                    e[2] = make_float4(color.x, color.y, color.z, __uint_as_float(pix));
                    e[3] = make_float4(att.x, att.y, att.z, __uint_as_float(smp));
                    e[4] = make_float4(__uint_as_float((uint32_t)slotIndex), __uint_as_float((uint32_t)((unsigned long long)slotIndex >> 32)), 0.f, 0.f);
                    rayActive = false;
                    publish_parked((unsigned int*)(e + 15), p.parkEpoch);
                    if (COUNT) tl.parked++;
                    continue;                                             // pulls from the drained queue -> dead
                }
            }
            // ---- start the ray (hitBVH prologue :196-201 + the root's own box test) ----
            rayActive = true; hit = false; closest = T_MAX_RAY; turns = 0;
            sp = 0; qHead = 0; qCount = 0; cur = 0xFFFFFFFFu; travDone = true;
            rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            exactOnly = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
            if (COUNT) { tl.rays++; tl.visits++; }
            if (CULL) update_segment();                                                    // closest = tMax: nothing is culled yet
            const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
            if (NODES >= 3) {   // the slack the records were grown by covers rays that start inside the origin region (bvh_build.cu): others are not culled
                const float4 rl = __ldg(sc.rootBox + 2), rh = __ldg(sc.rootBox + 3);
                cullOk = cullAllowed && o.x >= rl.x && o.x <= rh.x && o.y >= rl.y && o.y <= rh.y && o.z >= rl.z && o.z <= rh.z;
            }
            if (box_test(o, d, rinv, exactOnly, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) {
                if (sc.N == 1) enqueue(0u);                                                // the root is the only leaf
                else { cur = exactOnly ? 0u : RTB_ROOT_RECORD(); travDone = false; }
            }
        }
#ifdef RTB_TAIL_PROBE
        if (!probeDrained && __any_sync(FULL, dead)) { probeDrained = true; if (lane == 0 && pw < 8192) g_probe[1][pw] = probe_now(); }
#endif
        if (__all_sync(FULL, dead)) break;

        // =========================================== T: traverse ================================================
        while (true) {
            const bool can = rayActive && !travDone && qCount <= qGate;
            const unsigned bal = __ballot_sync(FULL, can);
            if (bal == 0) break;
            if (__popc(bal) < (int)p.tMin) {
                // lanes that L could put back to work (candidates queued), or enough lanes waiting for S to make it worth entering
                const bool needS = !dead && (!rayActive || (travDone && qCount == 0));
                const bool needL = !dead && !can && !needS;
                if (__any_sync(FULL, needL) || (unsigned)__popc(__ballot_sync(FULL, needS)) >= p.sMin) break;
            }
#ifdef RTB_TAIL_PROBE
            if (can) prSteps++;
#endif
            if (COUNT) {
                if (can) {
                    tl.lsteps++;
                    if (cur != 0xFFFFFFFFu) {      // lanes of a warp that expand the SAME record share one fetch (L1 coalesces them): count it once
                        tl.rec++;
                        const unsigned same = __match_any_sync(__activemask(), cur);
                        if ((unsigned)(__ffs(same) - 1) == lane) tl.recUnique++;
                    }
                }
                if (lane == 0) tl.wsteps++;
            }
            if (can) {
                // CN: 32-byte compressed records; rays with a zero / denormal direction component (reference yields inf / NaN)
                // keep to the exact 64-byte records
                if (NODES >= 3 && !exactOnly) wave_step_u<NODES == 4>(sc, sm, tid, o, rinv, cullOk ? closest : __int_as_float(0x7f800000), cullOk ? T_MIN_RAY : __int_as_float(0xff800000), cur, sp, qCount, travDone, lstack, err, leafOffset, smTop);
                else if (NODES == 2 && !exactOnly) wave_step_w<CULL>(sc, sm, tid, leafOffset, o, rinv, cur, sp, qHead, qCount, travDone, lstack, err, segLo, segHi);
#ifdef RTB_AB_KERNELS
                else if (NODES == 1 && !exactOnly) wave_step_c<CULL>(sc, sm, tid, leafOffset, o, rinv, cur, sp, qHead, qCount, travDone, lstack, err, segLo, segHi);
#endif
                else wave_step<COUNT, CULL>(sc, sm, tid, leafOffset, o, d, rinv, exactOnly, cur, sp, qHead, qCount, travDone, lstack, tl, err, segLo, segHi);
            }
        }

        // =========================================== L: leaf tests ==============================================
        while (true) {
            const bool has = qCount > 0;
            if (!__any_sync(FULL, has)) break;
#ifdef RTB_TAIL_PROBE
            if (has) prLeaf++;
#endif
            if (has) {
                const uint32_t g = sm.queue[qHead][tid];
                qHead = (qHead + 1) & (QCAP - 1);
                qCount--;
                const float before = closest;
                if (COUNT && CN && !exactOnly) tl.lbox++;
                if (NODES >= 3 && !exactOnly) {
                    bool poison = false;
                    if (leaf_box_passes(sc, g, o, d, rinv, false)) leaf_test_unordered<COUNT>(sc, g, o, d, T_MIN_RAY, closest, hit, rec, tl, poison);
                    if (poison) {                                         // NaN hit: the outcome depends on the reference's order -> re-trace in it
                        exactOnly = true; hit = false; closest = T_MAX_RAY;
                        sp = 0; qHead = 0; qCount = 0; cur = 0; travDone = false;
                    }
                } else
                if (!CN || exactOnly || leaf_box_passes(sc, g, o, d, rinv, exactOnly)) leaf_test<COUNT>(sc, g, o, d, T_MIN_RAY, closest, hit, rec, tl);
                if (CULL && closest != before) update_segment();
            }
        }
        qHead = 0;                                                        // every FIFO is empty: rewind (wave_step_w relies on it)
    }

    // tell the concurrently running trace_tail_kernel that this warp parks nothing any more
    if (NODES >= 3 && p.doneWarps != nullptr) { __threadfence(); if (lane == 0) atomicAdd(p.doneWarps, 1u); }
#ifdef RTB_TAIL_PROBE
    if (lane == 0 && pw < 8192) g_probe[2][pw] = probe_now();
    if (pw < 8192) g_probeLane[pw * 32 + lane] = prMax;
#endif
    if (err) atomicOr(p.errFlag, err);
    if (COUNT) flush_tally(tl, p.counters, p.walkCounters, lane);
}

// ---------------------------------------------------------------------------------------------------------------------
// trace_tail_kernel: finishes the paths the main launch parked, ONE RAY PER WARP.  At the end of the main launch every
// remaining path is a chain of dependent rays, each ray a chain of dependent node fetches; walked by one lane, the chain's
// latency is the launch's tail (~2.5 ms of a C2 frame, and it does not shrink when the frame is split over more GPUs).  Here
// all 32 lanes hold the same path state and expand up to 32 pending entries of the current ray per turn: the pending set lives
// in a per-warp shared-memory stack, each lane decodes one 4-ary record (same records, same conservative test and t-culling
// as wave_step_u), tests the surviving leaf candidates itself -- exact leaf box, then the reference's primitive test with the
// order-free acceptance of leaf_test_unordered against its lane-local closest hit -- and appends the surviving internal
// entries with ballot-ranked stores; the culling bound is the warp-wide minimum of the lane-local closest hits.  The ray's result is the
// lexicographic minimum (t, primitive id) over the lanes: what the per-lane walk produces, because that walk's outcome does
// not depend on the order of the tests.  A NaN hit (poison), a zero / denormal direction component or an overfull pending
// set send the ray to hit_bvh(): the reference's own walk on the exact records.  Shading is the main kernel's S phase,
// evaluated redundantly by every lane (uniform state, no divergence).
// ---------------------------------------------------------------------------------------------------------------------

// Finishes ONE parked path / ray, one ray per warp (see above).  Every lane of the warp calls it with the same item.
template <bool EXT, bool COUNT, class STK>
__device__ __forceinline__ void tail_item(const TraceParams& p, const uint32_t item, const STK stk, const unsigned lane, Tally& tl, unsigned& err) {
    const unsigned FULL = 0xFFFFFFFFu;
    const TraceScene& sc = p.sc;
    const uint32_t leafOffset = sc.N - 1;
    const float T_MIN_RAY = 0.001f, T_MAX_RAY = 10000000.0f;
    const float INF = __int_as_float(0x7f800000);
    const float4* e = p.parkBuf + (size_t)PARK_STRIDE * item;
    const float4 q0 = __ldcg(e), q1 = __ldcg(e + 1), q2 = __ldcg(e + 2), q3 = __ldcg(e + 3), q4 = __ldcg(e + 4);
    uint32_t resume = __float_as_uint(q4.w);                          // bit 0: the first ray is in flight (pending set + closest hit parked)
    f3 o = xyz(q0), d = xyz(q1), color = xyz(q2), att = xyz(q3);
    uint32_t rng = __float_as_uint(q0.w), depth = __float_as_uint(q1.w);
    const uint32_t pix = __float_as_uint(q2.w), smp = __float_as_uint(q3.w);
    const size_t slotIndex = (size_t)__float_as_uint(q4.x) | ((size_t)__float_as_uint(q4.y) << 32);
    while (true) {                                                    // one ray of the path per turn
        bool hit = false;
        Hit rec; rec.t = 0.f; rec.normal = F3(0, 0, 0); rec.mat = 0; rec.prim = 0; rec.back = 0;
        const f3 rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
        bool needExact = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
        const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
        if (COUNT && lane == 0) { tl.tailRays++; if (!(resume & 1u)) tl.rays++; }   // a resumed ray was counted by the launch that started it
        if (!needExact && ((resume & 1u) || box_test(o, d, rinv, false, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z))) {
            const float4 rl = __ldg(sc.rootBox + 2), rh = __ldg(sc.rootBox + 3);      // origin region of the t-culled walk (bvh_build.cu)
            const bool cullOk = *p.cullAllowed != 0u && o.x >= rl.x && o.x <= rh.x && o.y >= rl.y && o.y <= rh.y && o.z >= rl.z && o.z <= rh.z;
            float closest = T_MAX_RAY, best = T_MAX_RAY;              // lane-local closest hit; warp-wide culling bound
            bool poison = false, fit = true;
            uint32_t n = 1;
            if (resume & 1u) {                                        // continue a parked ray: its stack + cur, lane 0 keeps its closest hit
                const uint32_t psp = resume >> 8, pcur = __float_as_uint(q4.z);
                const float4 q5 = __ldcg(e + 5), q6 = __ldcg(e + 6);
                if (lane < psp) stk[lane] = untag_top(sc, __float_as_uint(__ldcg((const float*)(e + 7) + lane)));
                if (lane == 0 && pcur != 0xFFFFFFFFu) stk[psp] = untag_top(sc, pcur);
                n = psp + (pcur != 0xFFFFFFFFu ? 1u : 0u);
                closest = best = q6.w;                                // the others may only accept t <= the parked closest; ties: see the reduce
                if (lane == 0 && (resume & 2u)) {
                    hit = true; rec.t = q5.x; rec.normal = F3(q5.y, q5.z, q5.w);
                    rec.mat = __float_as_uint(q6.x); rec.prim = __float_as_uint(q6.y); rec.back = (int)__float_as_uint(q6.z);
                }
                resume = 0;
            } else if (lane == 0) stk[0] = p.cullAllowed[1];              // the record the walk starts at (tt_top_kernel)
            __syncwarp();
            while (n > 0) {
                if (COUNT && lane == 0) tl.tailTurns++;
                const uint32_t take = n < 32u ? n : 32u;
                const uint32_t my = lane < take ? stk[n - 1u - lane] : 0xFFFFFFFFu;
                n -= take;
                __syncwarp();
                uint32_t intMask = 0, enqMask = 0, id0 = 0, id1 = 0, id2 = 0, id3 = 0;
                if (my != 0xFFFFFFFFu) {
                    if (COUNT) { tl.rec++; tl.recUnique++; }      // the lanes of a tail warp expand distinct entries of one ray
                    const uint4* rp = sc.wide + 4ull * my;
                    const f8 h0 = ldg256(rp), h1 = ldg256(rp + 2);
                    const uint32_t w3 = __float_as_uint(h0.lo.w);
                    const uint32_t lox = __float_as_uint(h0.hi.x), loy = __float_as_uint(h0.hi.y), loz = __float_as_uint(h0.hi.z),
                                   hix = __float_as_uint(h0.hi.w), hiy = __float_as_uint(h1.lo.x), hiz = __float_as_uint(h1.lo.y);
                    id0 = __float_as_uint(h1.lo.z); id1 = __float_as_uint(h1.lo.w); id2 = __float_as_uint(h1.hi.x); id3 = __float_as_uint(h1.hi.y);
                    const float sx = __uint_as_float((w3 & 0xFFu) << 23), sy = __uint_as_float(((w3 >> 8) & 0xFFu) << 23),
                                sz = __uint_as_float(((w3 >> 16) & 0xFFu) << 23);
                    const float ax = sx * rinv.x, ay = sy * rinv.y, az = sz * rinv.z;
                    const float bx = (h0.lo.x - o.x) * rinv.x, by = (h0.lo.y - o.y) * rinv.y, bz = (h0.lo.z - o.z) * rinv.z;
                    const float m = fmaxf(fmaxf(fmaf(255.0f, fabsf(ax), fabsf(bx)), fmaf(255.0f, fabsf(ay), fabsf(by))), fmaf(255.0f, fabsf(az), fabsf(bz)));
                    const float tol = -2.0e-6f * m;                   // the error budget of wave_step_u; NaN / inf -> nothing is dropped
                    const float farLimit = (cullOk ? best : INF) - tol, nearLimit = (cullOk ? T_MIN_RAY : -INF) + tol;
                    const bool ngx = rinv.x < 0.0f, ngy = rinv.y < 0.0f, ngz = rinv.z < 0.0f;
                    const uint32_t nX = ngx ? hix : lox, fX = ngx ? lox : hix;
                    const uint32_t nY = ngy ? hiy : loy, fY = ngy ? loy : hiy;
                    const uint32_t nZ = ngz ? hiz : loz, fZ = ngz ? loz : hiz;
                    const uint32_t meta = w3 >> 24;
                    uint32_t passMask = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
#define RTB_B(w) __uint2float_rn(((w) >> (8 * k)) & 0xFFu)
                        const float tn = fmaxf(fmaxf(fmaf(RTB_B(nX), ax, bx), fmaf(RTB_B(nY), ay, by)), fmaf(RTB_B(nZ), az, bz));
                        const float tf = fminf(fminf(fmaf(RTB_B(fX), ax, bx), fmaf(RTB_B(fY), ay, by)), fmaf(RTB_B(fZ), az, bz));
#undef RTB_B
                        const bool pass = !((tf - tn) < tol) && !(tn > farLimit) && !(tf < nearLimit);
                        passMask |= pass ? (1u << k) : 0u;
                    }
                    passMask &= meta >> 4;
                    const uint32_t leafMask = meta & 0xFu;
                    intMask = passMask & ~leafMask;
                    enqMask = passMask & leafMask;
                }
                // the leaf candidates of all lanes are dealt out evenly (ballot-ranked list behind the pending set): a turn's critical path is
                // one record fetch + ceil(candidates / 32) leaf tests, not the largest per-lane candidate count (up to 4) of them
                uint32_t nLeaf = 0;
                { const bool b = enqMask & 1u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[COOP_CAP + nLeaf + __popc(bal & ((1u << lane) - 1u))] = id0 - leafOffset; nLeaf += __popc(bal); }
                { const bool b = enqMask & 2u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[COOP_CAP + nLeaf + __popc(bal & ((1u << lane) - 1u))] = id1 - leafOffset; nLeaf += __popc(bal); }
                { const bool b = enqMask & 4u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[COOP_CAP + nLeaf + __popc(bal & ((1u << lane) - 1u))] = id2 - leafOffset; nLeaf += __popc(bal); }
                { const bool b = enqMask & 8u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[COOP_CAP + nLeaf + __popc(bal & ((1u << lane) - 1u))] = id3 - leafOffset; nLeaf += __popc(bal); }
                __syncwarp();
#pragma unroll 1
                for (uint32_t base = lane; base < nLeaf; base += 32u) {
                    const uint32_t g = stk[COOP_CAP + base];
                    if (COUNT) tl.lbox++;
                    if (leaf_box_passes(sc, g, o, d, rinv, false)) leaf_test_unordered<COUNT>(sc, g, o, d, T_MIN_RAY, closest, hit, rec, tl, poison);
                }
                if (n + 128u > COOP_CAP) { fit = false; break; }      // warp-uniform
                { const bool b = intMask & 1u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[n + __popc(bal & ((1u << lane) - 1u))] = id0; n += __popc(bal); }
                { const bool b = intMask & 2u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[n + __popc(bal & ((1u << lane) - 1u))] = id1; n += __popc(bal); }
                { const bool b = intMask & 4u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[n + __popc(bal & ((1u << lane) - 1u))] = id2; n += __popc(bal); }
                { const bool b = intMask & 8u; const unsigned bal = __ballot_sync(FULL, b); if (b) stk[n + __popc(bal & ((1u << lane) - 1u))] = id3; n += __popc(bal); }
                __syncwarp();
                float mn = closest;                                   // never NaN: a NaN hit only raises poison
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, off));
                best = mn;
            }
            __syncwarp();
#ifdef RTB_TAIL_TEST_FALLBACK   // test builds only (tests/test_emulated_kernels.py): every third ray takes the exact fallback below
            if ((item + depth) % 3u == 0u) fit = false;
#endif
            if (!fit || __any_sync(FULL, poison)) {
                needExact = true;
            } else {
                // lexicographic minimum (t, primitive id) over the lanes that hold a hit, broadcast to every lane
                float bt = hit ? rec.t : INF;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) bt = fminf(bt, __shfl_xor_sync(FULL, bt, off));
                uint32_t bp = (hit && rec.t == bt) ? rec.prim : 0xFFFFFFFFu;
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) bp = min(bp, __shfl_xor_sync(FULL, bp, off));
                const unsigned winBal = __ballot_sync(FULL, hit && rec.t == bt && rec.prim == bp);
                hit = winBal != 0u;
                const int win = hit ? __ffs(winBal) - 1 : 0;
                rec.t = __shfl_sync(FULL, rec.t, win);
                rec.normal = F3(__shfl_sync(FULL, rec.normal.x, win), __shfl_sync(FULL, rec.normal.y, win), __shfl_sync(FULL, rec.normal.z, win));
                rec.mat = __shfl_sync(FULL, rec.mat, win); rec.prim = __shfl_sync(FULL, rec.prim, win); rec.back = __shfl_sync(FULL, rec.back, win);
            }
        }
        if (needExact) {                                              // the reference's walk, every lane alike (rare: NaN hits, axis-parallel rays)
            Tally scratch = { 0, 0, 0, 0, 0 };
            hit = hit_bvh<COUNT>(sc, o, d, T_MIN_RAY, T_MAX_RAY, rec, scratch, err);
            if (COUNT && lane == 0) { tl.rec += (scratch.visits - 1) / 2; tl.tri += scratch.tri; tl.sph += scratch.sph; }
        }
        if (p.primaryMode == 1u) {                                    // primary-hit launch: keep the hit record, no shading
            if (lane == 0) {
                float4* h = p.primaryHits + 3ull * pix;
                h[0] = make_float4(rec.t, rec.normal.x, rec.normal.y, rec.normal.z);
                h[1] = make_float4(__uint_as_float(rec.prim), __uint_as_float(rec.mat), __uint_as_float(hit ? 1u : 0u), __uint_as_float((uint32_t)rec.back));
                h[2] = make_float4(d.x, d.y, d.z, 0.f);
            }
            break;
        }
        // ---- the main kernel's S phase (rayColor loop body :283-307), uniform over the warp ----
        if (depth == 0 && smp == 0 && p.firstPass && p.hitPrim && lane == 0) {
            p.hitPrim[pix] = hit ? rec.prim : 0xFFFFFFFFu;
            if (p.hitT) p.hitT[pix] = hit ? rec.t : 0.0f;
        }
        bool pathEnd = true;
        if (!hit) {
            color = color + F3(0.f, 0.f, 0.f) * att;
        } else {
            const float4 m = __ldg(sc.mats + rec.mat);
            if (COUNT && lane == 0) tl.mat++;
            const uint32_t type = __float_as_uint(m.w);
            const f3 albedo = xyz(m);
            const f3 emitted = (type == RTB_LIGHT) ? albedo : F3(0.f, 0.f, 0.f);
            color = color + emitted * att;
            if (type == RTB_DIFFUSE) {
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
            if (depth >= p.maxDepth) pathEnd = true;
        }
        if (pathEnd) {
            if (COUNT && lane == 0) tl.paths++;
            if (lane == 0) {
                float4* out = p.sampleBuf + slotIndex;
                out->x = color.x; out->y = color.y; out->z = color.z;   // .w keeps the sample's incoming alpha
                if (p.rngOut && p.lastPass && smp + 1 == p.sampleCount) p.rngOut[pix] = rng;
            }
            break;
        }
    }
}

template <bool EXT, bool COUNT>
__global__ void __launch_bounds__(WAVE_THREADS) trace_tail_kernel(const TraceParams p) {
    __shared__ uint32_t coopStack[WAVE_THREADS / 32][COOP_CAP + 128];     // per warp: pending set + the turn's leaf candidates
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned lane = threadIdx.x & 31;
    const LinearStack stk{ coopStack[threadIdx.x >> 5] };
    unsigned err = 0;
    Tally tl = { 0, 0, 0, 0, 0 };
    if (!p.tailConcurrent) {
        // FINAL launch (in stream order behind the main launch): everything parked and not yet taken is complete and visible
        const uint32_t parked = min(*p.parkCount, p.parkCapacity);
        while (true) {
            uint32_t item = 0;
            if (lane == 0) item = atomicAdd(p.parkCursor, 1u);
            item = __shfl_sync(FULL, item, 0);
            if (item >= parked) break;
            tail_item<EXT, COUNT>(p, item, stk, lane, tl, err);
        }
    } else {
        // CONCURRENT launch (second stream, enqueued right behind the main launch): its CTAs become resident as CTAs of the main launch
        // exit, take what has been published so far (slot below the park count whose ready flag carries this launch's epoch) and leave
        // once every warp of the main launch has signed off and nothing is left -- the long rays that used to be the launch's tail are
        // worked off DURING the launch.  Deadlock freedom: a CTA of this launch that becomes resident before ALL CTAs of the (persistent)
        // main launch have started leaves at once -- it must not hold an SM slot a main CTA is waiting for -- and no warp polls for more
        // than tailSpinUs (20 ms); whatever is left over is finished by the FINAL launch.
        const unsigned mainCtas = p.mainWarps / (WAVE_THREADS / 32);
        if (*(volatile unsigned int*)(p.doneWarps + 1) < mainCtas) return;
        unsigned long long t0 = 0;
#ifndef RTB_SIMT_EMU
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
#endif
        while (true) {
            uint32_t item = 0xFFFFFFFFu;
            if (lane == 0) {
                while (true) {
                    const unsigned done = *(volatile unsigned int*)p.doneWarps;       // read BEFORE the count: once all producers are gone the count is final
                    __threadfence();
                    const unsigned n = min(*(volatile unsigned int*)p.parkCount, p.parkCapacity);
                    const unsigned c = *(volatile unsigned int*)p.parkCursor;
                    if (c < n) {
                        if (*(volatile unsigned int*)(p.parkBuf + (size_t)PARK_STRIDE * c + 15) == p.parkEpoch) {
                            if (atomicCAS(p.parkCursor, c, c + 1u) == c) { item = c; break; }
                            continue;
                        }
                    } else if (done >= p.mainWarps) break;
#ifdef RTB_SIMT_EMU
                    break;                    // (emulated launches are synchronous: nothing can change while this one runs)
#else
                    unsigned long long t1;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                    if (t1 - t0 > (unsigned long long)p.tailSpinUs * 1000ull) break;
                    __nanosleep(400);
#endif
                }
            }
            item = __shfl_sync(FULL, item, 0);
            if (item == 0xFFFFFFFFu) break;
            __threadfence();
            tail_item<EXT, COUNT>(p, item, stk, lane, tl, err);
        }
    }
    if (err) atomicOr(p.errFlag, err);
    if (COUNT) flush_tally(tl, nullptr, p.walkCounters, lane);
}

template <bool COUNT, bool EXT, bool CULL, int NODES>
static void launch_wave_variant(cudaStream_t st, TraceParams& p, int smCount, uint64_t need) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<COUNT, EXT, CULL, NODES>, WAVE_THREADS, 0);
    if (p.mainCtas != 0u && nb > (int)p.mainCtas) nb = (int)p.mainCtas;
    uint64_t grid = (uint64_t)smCount * (nb > 0 ? nb : 1);                 // persistent: resident CTAs per SM x SM count
    if (grid > need) grid = need;
    p.mainWarps = (uint32_t)grid * (WAVE_THREADS / 32);                    // what the concurrent tail launch waits for
    trace_wave_kernel<COUNT, EXT, CULL, NODES><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
}

static void dispatch_wave(cudaStream_t st, TraceParams& p, bool count, bool ext, bool cull, int nodesMode, int smCount, uint64_t need) {
    const int v = (count ? 4 : 0) | (ext ? 2 : 0) | (cull ? 1 : 0);
    if (nodesMode == 3 && p.sortedPush) {
        if (count) { if (ext) launch_wave_variant<true, true, false, 4>(st, p, smCount, need); else launch_wave_variant<true, false, false, 4>(st, p, smCount, need); }
        else if (ext) launch_wave_variant<false, true, false, 4>(st, p, smCount, need);
        else launch_wave_variant<false, false, false, 4>(st, p, smCount, need);
    } else if (nodesMode == 3) {
        if (count) { if (ext) launch_wave_variant<true, true, false, 3>(st, p, smCount, need); else launch_wave_variant<true, false, false, 3>(st, p, smCount, need); }
        else if (ext) launch_wave_variant<false, true, false, 3>(st, p, smCount, need);
        else launch_wave_variant<false, false, false, 3>(st, p, smCount, need);
    } else if (nodesMode == 2) {
        switch (v) {
        case 0: launch_wave_variant<false, false, false, 2>(st, p, smCount, need); break;
        case 1: launch_wave_variant<false, false, true, 2>(st, p, smCount, need); break;
        case 2: launch_wave_variant<false, true, false, 2>(st, p, smCount, need); break;
        default: launch_wave_variant<false, true, true, 2>(st, p, smCount, need); break;
        }
#ifdef RTB_AB_KERNELS
    } else if (nodesMode == 1) {
        switch (v) {
        case 0: launch_wave_variant<false, false, false, 1>(st, p, smCount, need); break;
        case 1: launch_wave_variant<false, false, true, 1>(st, p, smCount, need); break;
        case 2: launch_wave_variant<false, true, false, 1>(st, p, smCount, need); break;
        default: launch_wave_variant<false, true, true, 1>(st, p, smCount, need); break;
        }
#endif
    } else {
        switch (v) {
        case 0: launch_wave_variant<false, false, false, 0>(st, p, smCount, need); break;
        case 1: launch_wave_variant<false, false, true, 0>(st, p, smCount, need); break;
        case 2: launch_wave_variant<false, true, false, 0>(st, p, smCount, need); break;
        case 3: launch_wave_variant<false, true, true, 0>(st, p, smCount, need); break;
        case 4: launch_wave_variant<true, false, false, 0>(st, p, smCount, need); break;
        case 5: launch_wave_variant<true, false, true, 0>(st, p, smCount, need); break;
        case 6: launch_wave_variant<true, true, false, 0>(st, p, smCount, need); break;
        default: launch_wave_variant<true, true, true, 0>(st, p, smCount, need); break;
        }
    }
}

// One S2 submission = ceil(sampleCount / samplesPerPass) passes of { pre-pass, [primary hits,] trace, accumulate }.  Returns #launches.
int launch_trace_wave(cudaStream_t st, TraceParams p, bool count, bool ext, bool cull, int nodesMode, int smCount, uint32_t samplesPerPass,
                      const TailOverlap& ov) {
    const bool walk = p.walkCounters != nullptr;           // RTB_TRACE_WALK_COUNT: the production walk, counting what it fetches
    if ((count && !walk) || p.sc.N < 2) nodesMode = 0;     // RTB_TRACE_COUNT counts the reference's visits: exact records, reference order
    if (walk && nodesMode != 3) nodesMode = 0;             // (the walk counters are built into the two production variants)
#ifdef RTB_AB_KERNELS
    if (nodesMode == 1 && !p.sc.cnodes) nodesMode = 0;
#else
    if (nodesMode == 1) nodesMode = 0;
#endif
    if (nodesMode >= 2 && !p.sc.wide) nodesMode = 0;
    if (nodesMode == 3 && cull) nodesMode = 2;             // the segment-box extension belongs to the reference-order walk
    if (p.tMin == 0) p.tMin = nodesMode >= 2 ? 20 : T_MIN_DEFAULT;   // swept on C2 (RTB_WAVE_TMIN)
    // Primary-hit sharing: the reference's getRay has no jitter (raytraceBVH.comp:329-342), so the samples of a pixel all start
    // with the same primary ray and hitBVH returns the same record for each of them.  It is traced once per pixel per submission
    // (nothing survives the call); the instrumented variant keeps tracing it per sample because it counts the reference's work.
    const bool share = (!count || walk) && p.primaryHits != nullptr && p.sampleCount > 1;
    const uint32_t pixels = p.W * p.localRows;
    const uint32_t totalSamples = p.sampleCount, skip0 = p.sampleSkip;
    int launches = 0;
    for (uint32_t first = 0; first < totalSamples; first += samplesPerPass) {
        p.sampleCount = (totalSamples - first < samplesPerPass) ? totalSamples - first : samplesPerPass;
        p.sampleSkip = first == 0 ? skip0 : 0;                              // later passes continue from the image's alpha
        p.firstPass = first == 0;
        p.lastPass = first + p.sampleCount >= totalSamples;
        cudaMemsetAsync(p.workCounter64, 0, 32, st);                        // work counter + active-pixel count + park count / cursor
        if (count) wave_prepass_kernel<true><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        else wave_prepass_kernel<false><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        const bool tail = nodesMode == 3 && (p.coopMax != 0u || p.coopTurns != 0u);
        auto launch_tail = [&]() {                                          // finish the parked paths / rays, one ray per warp
            int nb = 0;
            if (ext) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_tail_kernel<true, false>, WAVE_THREADS, 0);
            else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_tail_kernel<false, false>, WAVE_THREADS, 0);
            const unsigned tt = p.tailThreads == 64u ? 64u : (unsigned)WAVE_THREADS;       // same warps in total, in CTAs half the size
            const unsigned grid = (unsigned)smCount * (unsigned)(nb > 0 ? nb : 1) * ((unsigned)WAVE_THREADS / tt);
            auto go = [&](cudaStream_t ts) {
                if (walk) { if (ext) trace_tail_kernel<true, true><<<grid, tt, 0, ts>>>(p); else trace_tail_kernel<false, true><<<grid, tt, 0, ts>>>(p); }
                else if (ext) trace_tail_kernel<true, false><<<grid, tt, 0, ts>>>(p);
                else trace_tail_kernel<false, false><<<grid, tt, 0, ts>>>(p);
                launches++;
            };
            if (ov.aux) {      // concurrent launch on the second stream (may start while the main launch, just enqueued on `st`, is running) ...
                p.tailConcurrent = 1u;
                cudaStreamWaitEvent(ov.aux, ov.fork, 0);
                go(ov.aux);
                cudaEventRecord(ov.join, ov.aux);
                cudaStreamWaitEvent(st, ov.join, 0);
            }
            p.tailConcurrent = 0u;                                          // ... and the final one, in stream order: whatever is left
            go(st);
        };
        if (share && first == 0) {                                          // one primary ray per active pixel -> primaryHits
            p.primaryMode = 1;
            p.parkEpoch++;
            if (tail && ov.aux) cudaEventRecord(ov.fork, st);               // everything the tail launch depends on precedes this point
            dispatch_wave(st, p, count, ext, cull, nodesMode, smCount, ((uint64_t)pixels + WAVE_THREADS - 1) / WAVE_THREADS);
            if (tail && p.coopTurns != 0u) launch_tail();
            cudaMemsetAsync(p.workCounter64, 0, 8, st);                     // rewind the work counter, keep the active-pixel count
            if (tail) cudaMemsetAsync((unsigned int*)p.workCounter64 + 4, 0, 16, st);   // park count / cursor
            launches++;
        }
        p.primaryMode = share ? 2u : 0u;
        p.parkEpoch++;
        const uint64_t need = ((uint64_t)pixels * p.sampleCount + WAVE_THREADS - 1) / WAVE_THREADS;   // never more lanes than items
        if (tail && ov.aux) cudaEventRecord(ov.fork, st);
        dispatch_wave(st, p, count, ext, cull, nodesMode, smCount, need);
        if (tail) launch_tail();
        wave_accumulate_kernel<0><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        launches += 3;
#ifdef RTB_TAIL_PROBE
        if (const char* pf = getenv("RTB_TAIL_PROBE_FILE")) {
            static unsigned long long host[3][8192];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(host, g_probe, sizeof(host));
            static uint4 hostLane[8192 * 32];
            cudaMemcpyFromSymbol(hostLane, g_probeLane, sizeof(hostLane));
            if (FILE* f = fopen(pf, "wb")) { fwrite(host, 1, sizeof(host), f); fwrite(hostLane, 1, sizeof(hostLane), f); fclose(f); }
        }
#endif
    }
    return launches;
}

}  // namespace rtb
