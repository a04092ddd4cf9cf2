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
#include "kernels.h"
#include "trace_wave_shared.cuh"

namespace rtb {

constexpr int WAVE_THREADS = 128;
#ifndef RTB_WIDE_MINB
#define RTB_WIDE_MINB 7
#endif
constexpr int WAVE_MIN_BLOCKS = 6;
constexpr int WIDE_MIN_BLOCKS = RTB_WIDE_MINB;   // resident CTAs per SM for the 4-ary variant

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
    const unsigned lane = tid & 31;
    const TraceScene& sc = p.sc;
    const uint32_t leafOffset = sc.N - 1;
    const uint32_t activeCount = *p.activeCount;
    // item i -> group of 32 active pixels g = i / (32 * S), sample s = (i % (32 * S)) / 32, pixel slot g * 32 + i % 32:
    // the 32 items a warp pulls together are the same sample of 32 neighbouring pixels (coherent primary rays)
    const uint32_t groupItems = 32u * (p.primaryMode == 1u ? 1u : p.sampleCount);
    const uint64_t totalWork = (uint64_t)((activeCount + 31u) / 32u) * groupItems;
    const float T_MIN_RAY = 0.001f, T_MAX_RAY = 10000000.0f;                    // sceneHit :268-269
    constexpr bool CN = NODES != 0;                       // conservative internal boxes: leaves are re-checked exactly
    constexpr uint32_t Q_ROOM = NODES >= 2 ? 4u : 2u;     // free FIFO entries a turn may need
    const uint32_t qGate = NODES >= 3 ? min(p.qGate, QCAP - Q_ROOM) : QCAP - Q_ROOM;   // a lane steps while its FIFO holds <= qGate candidates

    // ---- per-lane state -------------------------------------------------------------------------------------------
    bool dead = false, rayActive = false, travDone = true, exactOnly = false, hit = false, cullOk = false;
    uint32_t depth = 0, rng = 0, pix = 0, smp = 0;
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

    auto enqueue = [&](uint32_t g) {
        sm.queue[(qHead + qCount) & (QCAP - 1)][tid] = g;
        qCount++;
    };

    while (true) {
        // =========================================== S: shade / generate =========================================
        while (!dead && (!rayActive || (travDone && qCount == 0))) {
            bool needItem = !rayActive;
            if (rayActive && p.primaryMode == 1u) {                       // primary-hit launch: keep the hit record, no shading
                rayActive = false; needItem = true;
                float4* h = p.primaryHits + 2ull * pix;
                h[0] = make_float4(rec.t, rec.normal.x, rec.normal.y, rec.normal.z);
                h[1] = make_float4(__uint_as_float(rec.prim), __uint_as_float(rec.mat), __uint_as_float(hit ? 1u : 0u),
                                   __uint_as_float((uint32_t)rec.back));
            } else if (rayActive) {                                       // the ray is finished: rayColor loop body :283-307
                rayActive = false;
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
                }
            }
            if (needItem) {                                               // lane-granular persistent threads: next (pixel, sample)
                const unsigned act = __activemask();
                const int leader = __ffs(act) - 1;
                unsigned long long i = 0;
                if ((int)lane == leader) i = atomicAdd(p.workCounter64, (unsigned long long)__popc(act));
                i = __shfl_sync(act, i, leader) + (unsigned long long)__popc(act & ((1u << lane) - 1u));
                if (i >= totalWork) { dead = true; break; }
                const uint32_t g = (uint32_t)(i / groupItems), r = (uint32_t)(i % groupItems);
                const uint32_t slot = g * 32u + (r & 31u);
                if (slot >= activeCount) continue;                        // padding of the last group
                smp = r >> 5;
                pix = p.activePix[slot];
                slotIndex = (size_t)smp * p.slotCapacity + slot;
                const uint32_t x = pix % p.W, y = global_row(p, pix / p.W);
                const float alphaIn = p.sampleBuf[slotIndex].w;
                rng = (600u * x + y) * (p.randomState + 1u) + alpha_to_u32(alphaIn);       // random.glsl:10 + :350 (:351 is a no-op)
                (void)pcg_float(rng);                                                      // nextRandom :352 (kept by the pre-pass)
                color = F3(0.f, 0.f, 0.f); att = F3(1.f, 1.f, 1.f);
                depth = 0;
                o = p.cam.origin; d = primary_direction(p, x, y);
                if (p.primaryMode == 2u) {                                // the primary ray's hitBVH result, traced once per pixel
                    const float4 h0 = __ldg(p.primaryHits + 2ull * pix), h1 = __ldg(p.primaryHits + 2ull * pix + 1);
                    rec.t = h0.x; rec.normal = F3(h0.y, h0.z, h0.w);
                    rec.prim = __float_as_uint(h1.x); rec.mat = __float_as_uint(h1.y);
                    hit = __float_as_uint(h1.z) != 0u; rec.back = (int)__float_as_uint(h1.w);
                    rayActive = true; travDone = true; qCount = 0; qHead = 0; sp = 0; cur = 0xFFFFFFFFu;
                    continue;                                             // straight to shading
                }
            }
            // ---- start the ray (hitBVH prologue :196-201 + the root's own box test) ----
            rayActive = true; hit = false; closest = T_MAX_RAY;
            sp = 0; qHead = 0; qCount = 0; cur = 0xFFFFFFFFu; travDone = true;
            rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
            exactOnly = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
            if (COUNT) { tl.rays++; tl.visits++; }
            if (CULL) update_segment();                                                    // closest = tMax: nothing is culled yet
            const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
            if (NODES >= 3) {   // the slack the records were grown by covers rays that start inside the scene's box (+ 0.1 %): others are not culled
                const float grow = 1.0e-3f * fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
                cullOk = o.x >= lo.x - grow && o.x <= hi.x + grow && o.y >= lo.y - grow && o.y <= hi.y + grow && o.z >= lo.z - grow && o.z <= hi.z + grow;
            }
            if (box_test(o, d, rinv, exactOnly, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) {
                if (sc.N == 1) enqueue(0u);                                                // the root is the only leaf
                else { cur = 0; travDone = false; }
            }
        }
        if (__all_sync(FULL, dead)) break;

        // =========================================== T: traverse ================================================
        while (true) {
            const bool can = rayActive && !travDone && qCount <= qGate;
            const unsigned bal = __ballot_sync(FULL, can);
            if (bal == 0) break;
            if (__popc(bal) < (int)p.tMin) {
                const bool waiting = !dead && !can;                       // lanes that L or S could put back to work
                if (__any_sync(FULL, waiting)) break;
            }
            if (can) {
                // CN: 32-byte compressed records; rays with a zero / denormal direction component (reference yields inf / NaN)
                // keep to the exact 64-byte records
                if (NODES >= 3 && !exactOnly) wave_step_u<NODES == 4>(sc, sm, tid, o, rinv, cullOk ? closest : __int_as_float(0x7f800000), cullOk ? T_MIN_RAY : __int_as_float(0xff800000), cur, sp, qCount, travDone, lstack, err, leafOffset);
                else if (NODES == 2 && !exactOnly) wave_step_w<CULL>(sc, sm, tid, leafOffset, o, rinv, cur, sp, qHead, qCount, travDone, lstack, err, segLo, segHi);
                else if (NODES == 1 && !exactOnly) wave_step_c<CULL>(sc, sm, tid, leafOffset, o, rinv, cur, sp, qHead, qCount, travDone, lstack, err, segLo, segHi);
                else wave_step<COUNT, CULL>(sc, sm, tid, leafOffset, o, d, rinv, exactOnly, cur, sp, qHead, qCount, travDone, lstack, tl, err, segLo, segHi);
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
                const float before = closest;
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

    if (err) atomicOr(p.errFlag, err);
    if (COUNT) {
        unsigned long long v[5] = { tl.rays, tl.visits, tl.tri, tl.sph, tl.mat };
#pragma unroll
        for (int i = 0; i < 5; i++) {
            unsigned long long s = v[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(FULL, s, off);
            if (lane == 0 && s) atomicAdd(p.counters + i, s);
        }
    }
}

template <bool COUNT, bool EXT, bool CULL, int NODES>
static void launch_wave_variant(cudaStream_t st, const TraceParams& p, int smCount, uint64_t need) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_wave_kernel<COUNT, EXT, CULL, NODES>, WAVE_THREADS, 0);
    uint64_t grid = (uint64_t)smCount * (nb > 0 ? nb : 1);                 // persistent: resident CTAs per SM x SM count
    if (grid > need) grid = need;
    trace_wave_kernel<COUNT, EXT, CULL, NODES><<<(unsigned)grid, WAVE_THREADS, 0, st>>>(p);
}

static void dispatch_wave(cudaStream_t st, const TraceParams& p, bool count, bool ext, bool cull, int nodesMode, int smCount, uint64_t need) {
    const int v = (count ? 4 : 0) | (ext ? 2 : 0) | (cull ? 1 : 0);
    if (nodesMode == 3 && p.sortedPush) {
        if (ext) launch_wave_variant<false, true, false, 4>(st, p, smCount, need);
        else launch_wave_variant<false, false, false, 4>(st, p, smCount, need);
    } else if (nodesMode == 3) {
        if (ext) launch_wave_variant<false, true, false, 3>(st, p, smCount, need);
        else launch_wave_variant<false, false, false, 3>(st, p, smCount, need);
    } else if (nodesMode == 2) {
        switch (v) {
        case 0: launch_wave_variant<false, false, false, 2>(st, p, smCount, need); break;
        case 1: launch_wave_variant<false, false, true, 2>(st, p, smCount, need); break;
        case 2: launch_wave_variant<false, true, false, 2>(st, p, smCount, need); break;
        default: launch_wave_variant<false, true, true, 2>(st, p, smCount, need); break;
        }
    } else if (nodesMode == 1) {
        switch (v) {
        case 0: launch_wave_variant<false, false, false, 1>(st, p, smCount, need); break;
        case 1: launch_wave_variant<false, false, true, 1>(st, p, smCount, need); break;
        case 2: launch_wave_variant<false, true, false, 1>(st, p, smCount, need); break;
        default: launch_wave_variant<false, true, true, 1>(st, p, smCount, need); break;
        }
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
int launch_trace_wave(cudaStream_t st, TraceParams p, bool count, bool ext, bool cull, int nodesMode, int smCount, uint32_t samplesPerPass) {
    if (count || p.sc.N < 2) nodesMode = 0;                // the instrumented variant counts the reference's visits: exact records
    if (nodesMode == 1 && !p.sc.cnodes) nodesMode = 0;
    if (nodesMode >= 2 && !p.sc.wide) nodesMode = 0;
    if (nodesMode == 3 && cull) nodesMode = 2;             // the segment-box extension belongs to the reference-order walk
    if (p.tMin == 0) p.tMin = nodesMode >= 2 ? 20 : T_MIN_DEFAULT;   // swept on C2 (RTB_WAVE_TMIN)
    // Primary-hit sharing: the reference's getRay has no jitter (raytraceBVH.comp:329-342), so the samples of a pixel all start
    // with the same primary ray and hitBVH returns the same record for each of them.  It is traced once per pixel per submission
    // (nothing survives the call); the instrumented variant keeps tracing it per sample because it counts the reference's work.
    const bool share = !count && p.primaryHits != nullptr && p.sampleCount > 1;
    const uint32_t pixels = p.W * p.localRows;
    const uint32_t totalSamples = p.sampleCount, skip0 = p.sampleSkip;
    int launches = 0;
    for (uint32_t first = 0; first < totalSamples; first += samplesPerPass) {
        p.sampleCount = (totalSamples - first < samplesPerPass) ? totalSamples - first : samplesPerPass;
        p.sampleSkip = first == 0 ? skip0 : 0;                              // later passes continue from the image's alpha
        p.firstPass = first == 0;
        p.lastPass = first + p.sampleCount >= totalSamples;
        cudaMemsetAsync(p.workCounter64, 0, 16, st);                        // work counter + active-pixel count
        if (count) wave_prepass_kernel<true><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        else wave_prepass_kernel<false><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        if (share && first == 0) {                                          // one primary ray per active pixel -> primaryHits
            p.primaryMode = 1;
            dispatch_wave(st, p, count, ext, cull, nodesMode, smCount, ((uint64_t)pixels + WAVE_THREADS - 1) / WAVE_THREADS);
            cudaMemsetAsync(p.workCounter64, 0, 8, st);                     // rewind the work counter, keep the active-pixel count
            launches++;
        }
        p.primaryMode = share ? 2u : 0u;
        const uint64_t need = ((uint64_t)pixels * p.sampleCount + WAVE_THREADS - 1) / WAVE_THREADS;   // never more lanes than items
        dispatch_wave(st, p, count, ext, cull, nodesMode, smCount, need);
        wave_accumulate_kernel<0><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        launches += 3;
    }
    return launches;
}

}  // namespace rtb
