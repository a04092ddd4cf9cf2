// trace_stream.cu -- S2 as a streaming (wavefront) path tracer inside ONE persistent cooperative kernel.
//
// trace_wave.cu keeps a path in the registers of one lane from its first ray to its last; when a lane's ray ends, the
// whole warp has to leave the traverse phase to shade it, so traversal runs at ~23 of 32 lanes and shading at ~9
// (profiles/r01_wave_kernel_c2_ncu.txt).  Here paths live in a pool in global memory and the grid alternates, separated
// by grid-wide barriers, between two phases that are each uniform across the whole machine:
//   shade / generate : one thread per pool path: shade the finished ray (emit, diffuse scatter), start the next bounce,
//                      or retire the path (write the sample's colour to its slot) and pull the next (pixel, sample) item;
//                      rays that enter the root box are appended to the ray list (warp-aggregated atomics)
//   trace            : lanes pull rays from the list one at a time; each warp alternates the T (child-pair steps) and
//                      L (queued leaf tests) phases of trace_wave.cu; a finished lane only stores its hit record and
//                      pulls the next ray, so the traverse phase stays (nearly) full.
// Per ray this adds ~170 bytes of pool traffic to ~4 KB of node fetches.  All arithmetic -- seeds, rays, traversal
// order, intersection, scatter, accumulation order (per-sample slots + ordered accumulate pass) -- is identical to
// trace_wave.cu, so results are bit-identical (same tests).
#include <cooperative_groups.h>

#include "kernels.h"
#include "trace_wave_shared.cuh"

namespace cg = cooperative_groups;

namespace rtb {

constexpr int STREAM_THREADS = 128;
constexpr int STREAM_MIN_BLOCKS = 7;
constexpr int STREAM_TMIN_DEFAULT = 26;
constexpr uint32_t PATH_EMPTY = 0xFFFFFFFFu;

template <bool COUNT, bool EXT>
__global__ void __launch_bounds__(STREAM_THREADS, STREAM_MIN_BLOCKS) trace_stream_kernel(const TraceParams p) {
    __shared__ WaveSmem<STREAM_THREADS> sm;
    cg::grid_group grid = cg::this_grid();
    const unsigned FULL = 0xFFFFFFFFu;
    const unsigned tid = threadIdx.x;
    const unsigned lane = tid & 31;
    const TraceScene& sc = p.sc;
    const StreamPool& pool = p.pool;
    const uint32_t leafOffset = sc.N - 1;
    const uint32_t activeCount = *p.activeCount;
    const uint32_t groupItems = 32u * p.sampleCount;
    const uint64_t totalWork = (uint64_t)((activeCount + 31u) / 32u) * groupItems;
    const float T_MIN_RAY = 0.001f, T_MAX_RAY = 10000000.0f;
    const uint32_t P = pool.capacity;
    const uint32_t gthreads = gridDim.x * blockDim.x;
    const uint32_t gtid = blockIdx.x * blockDim.x + tid;
    const float4 rootLo = __ldg(sc.rootBox), rootHi = __ldg(sc.rootBox + 1);

    Tally tl = { 0, 0, 0, 0, 0 };
    unsigned err = 0;

    for (uint32_t it = 0;; ++it) {
        const uint32_t par = it & 1u;
        // =================================== shade / generate: one thread per pool path ===================================
        for (uint32_t base = gtid - lane; base < P; base += gthreads) {
            const uint32_t i = base + lane;
            bool produced = false;
            if (i < P) {
                uint32_t slotIndex = __ldcg(&pool.slot[i]);
                f3 color = F3(0, 0, 0), att = F3(1, 1, 1), o = F3(0, 0, 0), d = F3(0, 0, 1);
                uint32_t rng = 0, depth = 0;
                bool haveRay = false;
                if (slotIndex != PATH_EMPTY) {                               // the path's ray is finished: rayColor body :283-307
                    const float4 c4 = __ldcg(&pool.colorRng[i]), a4 = __ldcg(&pool.attDepth[i]), o4 = __ldcg(&pool.org[i]), d4 = __ldcg(&pool.dir[i]),
                                 n4 = __ldcg(&pool.nrm[i]);   // L2-scoped: written by other SMs in the trace phase
                    color = xyz(c4); rng = __float_as_uint(c4.w); att = xyz(a4); depth = __float_as_uint(a4.w);
                    o = xyz(o4); d = xyz(d4);
                    const uint32_t prim = __float_as_uint(d4.w);
                    const bool hit = prim != 0xFFFFFFFFu;
                    const uint32_t smp = slotIndex / p.slotCapacity;
                    if (depth == 0 && smp == 0 && p.firstPass && p.hitPrim) {
                        const uint32_t pix = p.activePix[slotIndex - smp * p.slotCapacity];
                        p.hitPrim[pix] = prim;
                        if (p.hitT) p.hitT[pix] = hit ? o4.w : 0.0f;
                    }
                    bool pathEnd = true;
                    if (!hit) {
                        color = color + F3(0.f, 0.f, 0.f) * att;             // _BACKGROUND_COLOR * globalAttenuation :284
                    } else {
                        const uint32_t matBits = __float_as_uint(n4.w);
                        Hit rec; rec.t = o4.w; rec.normal = xyz(n4); rec.mat = matBits & 0x7FFFFFFFu; rec.back = (int)(matBits >> 31); rec.prim = prim;
                        const float4 m = __ldg(sc.mats + rec.mat);
                        if (COUNT) tl.mat++;
                        const uint32_t type = __float_as_uint(m.w);
                        const f3 albedo = xyz(m);
                        const f3 emitted = (type == RTB_LIGHT) ? albedo : F3(0.f, 0.f, 0.f);   // emitted :94-99
                        color = color + emitted * att;                                        // :293
                        if (type == RTB_DIFFUSE) {                                            // scatter :100-115
                            const f3 Pt = o + rec.t * d;
                            const f3 nd = normalize(rec.normal + random_unit_vector(rng));
                            o = Pt; d = nd;
                            att = att * albedo;
                            pathEnd = false;
                        } else if (EXT && type != RTB_LIGHT) {
                            f3 a2, nd;
                            const f3 Pt = o + rec.t * d;
                            if (scatter_extension(type, albedo, d, rec, rng, a2, nd)) { o = Pt; d = nd; att = att * a2; pathEnd = false; }
                        }
                        depth++;
                        if (depth >= p.maxDepth) pathEnd = true;                              // for (i < maxRayTraceDepth) :282
                    }
                    haveRay = !pathEnd;
                }
                // find a ray that enters the root box: the bounce just made, or the primary ray of the next item(s)
                while (true) {
                    if (!haveRay) {
                        if (slotIndex != PATH_EMPTY) {                       // retire the path: this sample's pixelColor is final
                            float4* e = p.sampleBuf + slotIndex;
                            e->x = color.x; e->y = color.y; e->z = color.z;  // .w keeps the sample's incoming alpha
                            const uint32_t smp = slotIndex / p.slotCapacity;
                            if (p.rngOut && p.lastPass && smp + 1 == p.sampleCount) p.rngOut[p.activePix[slotIndex - smp * p.slotCapacity]] = rng;
                            slotIndex = PATH_EMPTY;
                        }
                        const unsigned act = __activemask();                 // next (pixel, sample) item
                        const int leader = __ffs(act) - 1;
                        unsigned long long w = 0;
                        if ((int)lane == leader) w = atomicAdd(p.workCounter64, (unsigned long long)__popc(act));
                        w = __shfl_sync(act, w, leader) + (unsigned long long)__popc(act & ((1u << lane) - 1u));
                        if (w >= totalWork) break;                           // no work left: the path stays empty
                        const uint32_t g = (uint32_t)(w / groupItems), r = (uint32_t)(w % groupItems);
                        const uint32_t slot = g * 32u + (r & 31u);
                        if (slot >= activeCount) continue;                   // padding of the last group
                        const uint32_t smp = r >> 5;
                        const uint32_t pix = p.activePix[slot];
                        slotIndex = smp * p.slotCapacity + slot;
                        const uint32_t x = pix % p.W, y = global_row(p, pix / p.W);
                        const float alphaIn = p.sampleBuf[slotIndex].w;
                        rng = (600u * x + y) * (p.randomState + 1u) + alpha_to_u32(alphaIn);   // random.glsl:10 + :350
                        (void)pcg_float(rng);                                // nextRandom :352 (kept by the pre-pass)
                        color = F3(0.f, 0.f, 0.f); att = F3(1.f, 1.f, 1.f);
                        depth = 0;
                        o = p.cam.origin; d = primary_direction(p, x, y);
                        haveRay = true;
                    }
                    // hitBVH prologue: the root's own box test (:208-211)
                    if (COUNT) { tl.rays++; tl.visits++; }
                    const f3 rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    const bool exactOnly = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
                    if (box_test(o, d, rinv, exactOnly, rootLo.x, rootLo.y, rootLo.z, rootHi.x, rootHi.y, rootHi.z)) break;
                    // the ray misses the whole scene: shade the miss right here (:283-286) and retire the path
                    if (depth == 0 && p.firstPass && p.hitPrim && slotIndex / p.slotCapacity == 0) {
                        const uint32_t pix = p.activePix[slotIndex % p.slotCapacity];
                        p.hitPrim[pix] = 0xFFFFFFFFu;
                        if (p.hitT) p.hitT[pix] = 0.0f;
                    }
                    color = color + F3(0.f, 0.f, 0.f) * att;
                    haveRay = false;
                }
                pool.slot[i] = slotIndex;
                if (slotIndex != PATH_EMPTY) {
                    pool.colorRng[i] = make_float4(color.x, color.y, color.z, __uint_as_float(rng));
                    pool.attDepth[i] = make_float4(att.x, att.y, att.z, __uint_as_float(depth));
                    pool.org[i] = make_float4(o.x, o.y, o.z, 0.f);
                    pool.dir[i] = make_float4(d.x, d.y, d.z, __uint_as_float(0xFFFFFFFFu));
                    produced = true;
                }
            }
            const unsigned bal = __ballot_sync(FULL, produced);              // append to the ray list
            if (bal) {
                uint32_t at = 0;
                if (lane == (unsigned)(__ffs(bal) - 1)) at = atomicAdd(&pool.cnt[par], (unsigned)__popc(bal));
                at = __shfl_sync(FULL, at, __ffs(bal) - 1) + __popc(bal & ((1u << lane) - 1u));
                if (produced) pool.rayList[at] = i;
            }
        }
        if (gtid == 0) { pool.cnt[par ^ 1u] = 0; pool.cnt[2 + (par ^ 1u)] = 0; }   // counters of the next iteration
        grid.sync();
        const uint32_t numRays = __ldcg(&pool.cnt[par]);
        if (numRays == 0) break;

        // =================================== trace: lanes pull rays from the list ===================================
        {
            bool haveRay = false, listDone = false, travDone = true, exactOnly = false, hit = false;
            uint32_t pathIdx = 0;
            f3 o = F3(0, 0, 0), d = F3(0, 0, 1), rinv = F3(0, 0, 0);
            float closest = T_MAX_RAY;
            Hit rec; rec.t = 0.f; rec.normal = F3(0, 0, 0); rec.mat = 0; rec.prim = 0; rec.back = 0;
            uint32_t cur = 0xFFFFFFFFu;
            int sp = 0;
            uint32_t qHead = 0, qCount = 0;
            uint32_t lstack[STACK_DEPTH - SSTACK];
            while (true) {
                // ---- finish / fetch ----
                if (haveRay && travDone && qCount == 0) {                    // store the hit record for the shade phase
                    pool.org[pathIdx].w = hit ? rec.t : 0.f;
                    pool.dir[pathIdx].w = __uint_as_float(hit ? rec.prim : 0xFFFFFFFFu);
                    if (hit) pool.nrm[pathIdx] = make_float4(rec.normal.x, rec.normal.y, rec.normal.z, __uint_as_float(rec.mat | ((uint32_t)rec.back << 31)));
                    haveRay = false;
                }
                if (!haveRay && !listDone) {
                    const unsigned act = __activemask();
                    const int leader = __ffs(act) - 1;
                    uint32_t k = 0;
                    if ((int)lane == leader) k = atomicAdd(&pool.cnt[2 + par], (unsigned)__popc(act));
                    k = __shfl_sync(act, k, leader) + __popc(act & ((1u << lane) - 1u));
                    if (k >= numRays) listDone = true;
                    else {
                        pathIdx = __ldcg(&pool.rayList[k]);
                        const float4 o4 = __ldcg(&pool.org[pathIdx]), d4 = __ldcg(&pool.dir[pathIdx]);
                        o = xyz(o4); d = xyz(d4);
                        rinv = F3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                        exactOnly = !(fabsf(rinv.x) < 3.0e38f && fabsf(rinv.y) < 3.0e38f && fabsf(rinv.z) < 3.0e38f);
                        haveRay = true; hit = false; closest = T_MAX_RAY;
                        sp = 0; qHead = 0; qCount = 0;
                        if (sc.N == 1) { sm.queue[0][tid] = 0u; qCount = 1; cur = 0xFFFFFFFFu; travDone = true; }   // the root is the only leaf
                        else { cur = 0; travDone = false; }                  // the root box was tested by the shade phase
                    }
                }
                if (!__any_sync(FULL, haveRay)) break;                       // (lanes without a ray have listDone set by now)

                // ---- T: child-pair steps ----
                while (true) {
                    const bool can = haveRay && !travDone && qCount <= QCAP - 2;
                    const unsigned bal = __ballot_sync(FULL, can);
                    if (bal == 0) break;
                    if (__popc(bal) < (int)p.tMin) {
                        const bool waiting = (haveRay && !can) || (!haveRay && !listDone);
                        if (__any_sync(FULL, waiting)) break;
                    }
                    if (can) wave_step<COUNT, false>(sc, sm, tid, leafOffset, o, d, rinv, exactOnly, cur, sp, qHead, qCount, travDone, lstack, tl, err, F3(0, 0, 0), F3(0, 0, 0));
                }
                // ---- L: queued leaf tests, in order ----
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
        }
        grid.sync();
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

template <bool COUNT, bool EXT>
static int launch_stream_variant(cudaStream_t st, TraceParams& p, int smCount) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_stream_kernel<COUNT, EXT>, STREAM_THREADS, 0);
    if (nb < 1) nb = 1;
    const unsigned grid = (unsigned)smCount * (unsigned)nb;                 // cooperative launch: every CTA resident
    void* args[] = { (void*)&p };
    cudaLaunchCooperativeKernel((const void*)trace_stream_kernel<COUNT, EXT>, dim3(grid), dim3(STREAM_THREADS), args, 0, st);
    return 1;
}

// One S2 submission = passes of { pre-pass, streaming trace (cooperative), accumulate }.  Returns #launches.
int launch_trace_stream(cudaStream_t st, TraceParams p, bool count, bool ext, int smCount, uint32_t samplesPerPass) {
    if (p.tMin == 0) p.tMin = STREAM_TMIN_DEFAULT;
    const uint32_t pixels = p.W * p.localRows;
    const uint32_t totalSamples = p.sampleCount, skip0 = p.sampleSkip;
    int launches = 0;
    for (uint32_t first = 0; first < totalSamples; first += samplesPerPass) {
        p.sampleCount = (totalSamples - first < samplesPerPass) ? totalSamples - first : samplesPerPass;
        p.sampleSkip = first == 0 ? skip0 : 0;
        p.firstPass = first == 0;
        p.lastPass = first + p.sampleCount >= totalSamples;
        cudaMemsetAsync(p.workCounter64, 0, 16, st);                        // work counter + active-pixel count
        cudaMemsetAsync(p.pool.cnt, 0, 16, st);                             // ray-list counters
        cudaMemsetAsync(p.pool.slot, 0xFF, sizeof(uint32_t) * (size_t)p.pool.capacity, st);   // every pool path empty
        if (count) wave_prepass_kernel<true><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        else wave_prepass_kernel<false><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        if (count) { if (ext) launch_stream_variant<true, true>(st, p, smCount); else launch_stream_variant<true, false>(st, p, smCount); }
        else { if (ext) launch_stream_variant<false, true>(st, p, smCount); else launch_stream_variant<false, false>(st, p, smCount); }
        wave_accumulate_kernel<0><<<(pixels + 255) / 256, 256, 0, st>>>(p);
        launches += 3;
    }
    return launches;
}

}  // namespace rtb
