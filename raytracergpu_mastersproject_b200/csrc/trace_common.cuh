// trace_common.cuh -- device functions shared by the trace kernels: the reference's intersection routines and scatter,
// restated with the shader's operation order (raytraceBVH.comp:84-193, random.glsl:33-62).
#pragma once

#include "common.cuh"

namespace rtb {

constexpr int TILE_W = 8, TILE_H = 4;     // work is handed out in 8x4 pixel tiles (coherent primary rays)
constexpr int STACK_DEPTH = 64;           // Karras tree depth <= 62 (30 code bits + 32 index bits); reference: 128

struct Hit {
    float t;
    f3 normal;
    uint32_t mat;
    uint32_t prim;      // global primitive id g (triangle g < T, sphere T + idx)
    int back;           // backFaceInt
};

// rays .. mat: the reference-equivalent work counters (rtb_counters); rec .. tailTurns: what the kernels really fetch
// (rtb_walk_counters, RTB_TRACE_WALK_COUNT).  Only the COUNT instantiations touch any of it.
struct Tally {
    unsigned long long rays, visits, tri, sph, mat;
    unsigned long long rec, lbox, items, paths, parked, lsteps, wsteps, tailRays, tailTurns, recUnique;
};
constexpr int WALK_COUNTER_WORDS = 14;
// warp-reduce the per-lane tallies and add them to the rtb_counters / rtb_walk_counters buffers (either may be null)
__device__ __forceinline__ void flush_tally(const Tally& tl, unsigned long long* counters, unsigned long long* walk, const unsigned lane) {
    const unsigned long long a[5] = { tl.rays, tl.visits, tl.tri, tl.sph, tl.mat };
    const unsigned long long b[WALK_COUNTER_WORDS] = { tl.rays, tl.rec, tl.lbox, tl.tri, tl.sph, tl.mat, tl.items, tl.paths, tl.parked,
                                                       tl.lsteps, tl.wsteps, tl.tailRays, tl.tailTurns, tl.recUnique };
    if (counters) {
#pragma unroll
        for (int i = 0; i < 5; i++) {
            unsigned long long s = a[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
            if (lane == 0 && s) atomicAdd(counters + i, s);
        }
    }
    if (walk) {
#pragma unroll
        for (int i = 0; i < WALK_COUNTER_WORDS; i++) {
            unsigned long long s = b[i];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, off);
            if (lane == 0 && s) atomicAdd(walk + i, s);
        }
    }
}

// AABBhitCheck, raytraceBVH.comp:184-193 : exact IEEE divisions, GLSL min/max, no t-interval
__device__ __forceinline__ bool box_hit(const f3 o, const f3 d, const float lox, const float loy, const float loz, const float hix,
                                        const float hiy, const float hiz) {
    const float ax = (lox - o.x) / d.x, ay = (loy - o.y) / d.y, az = (loz - o.z) / d.z;
    const float bx = (hix - o.x) / d.x, by = (hiy - o.y) / d.y, bz = (hiz - o.z) / d.z;
    const float t1x = gmin(ax, bx), t1y = gmin(ay, by), t1z = gmin(az, bz);
    const float t2x = gmax(ax, bx), t2y = gmax(ay, by), t2z = gmax(az, bz);
    const float tNear = gmax(gmax(t1x, t1y), t1z);
    const float tFar = gmin(gmin(t2x, t2y), t2z);
    return tNear < tFar;
}

// triangleHit, raytraceBVH.comp:118-149.  The ray-independent part of the shader's function -- u, v, the normalised normal
// and w = N / dot(N, N) (:120-125) -- is evaluated once per triangle by pack_prims_kernel with the shader's operation
// order and fetched here as one 64-byte record (two 256-bit loads); the ray-dependent part is the shader's, verbatim.
__device__ __forceinline__ bool triangle_hit_rec(const f8 r0, const f8 r1, const f3 o, const f3 d, const float tMin, const float tMax,
                                                 Hit& rec) {
    const f3 v0 = xyz(r0.lo);
    const f3 n = xyz(r0.hi);
    const f3 u = F3(r0.hi.w, r1.lo.x, r1.lo.y);
    const f3 v = F3(r1.lo.z, r1.lo.w, r1.hi.x);
    const f3 w = F3(r1.hi.y, r1.hi.z, r1.hi.w);
    const float D = dot(n, v0);
    const float denom = dot(n, d);
    if (fabsf(denom) < 0.0001f) return false;
    const float t = (D - dot(n, o)) / denom;
    if (t < tMin || t > tMax) return false;
    const f3 P = o + t * d;
    const f3 pp = P - v0;
    const float aa = dot(w, cross(pp, v));
    const float bb = dot(w, cross(u, pp));
    if (aa < 0 || bb < 0 || aa + bb > 1) return false;
    rec.t = t;
    const int back = dot(d, n) > 0 ? 1 : 0;
    rec.normal = (float)(1 - 2 * back) * n;
    rec.back = back;
    rec.mat = __float_as_uint(r0.lo.w);
    return true;
}
__device__ __forceinline__ bool triangle_hit(const TraceScene& sc, const uint32_t idx, const f3 o, const f3 d, const float tMin,
                                             const float tMax, Hit& rec) {
    const float4* tp = sc.tris + 4ull * idx;
    const f8 r0 = ldg256(tp), r1 = ldg256(tp + 2);
    return triangle_hit_rec(r0, r1, o, d, tMin, tMax, rec);
}

// sphereHit, raytraceBVH.comp:152-181 (rec.u / rec.v are dead values)
__device__ __forceinline__ bool sphere_hit(const TraceScene& sc, const uint32_t idx, const f3 o, const f3 d, const float tMin,
                                           const float tMax, Hit& rec) {
    const float4 s = __ldg(sc.sphs + idx);
    const f3 ctr = xyz(s);
    const f3 oc = o - ctr;
    const float a = dot(d, d);
    const float halfB = dot(oc, d);
    const float c = dot(oc, oc) - (s.w * s.w);
    const float underRadical = (halfB * halfB) - (a * c);
    if (underRadical < 0) return false;
    const float radical = sqrtf(underRadical);
    float root = (-halfB - radical) / a;
    if (root < tMin || root > tMax) {
        root = (-halfB + radical) / a;
        if (root < tMin || root > tMax) return false;
    }
    rec.t = root;
    const f3 P = o + root * d;
    f3 n = (P - ctr) / s.w;
    const int back = dot(d, n) > 0 ? 1 : 0;
    rec.normal = (float)(1 - 2 * back) * n;
    rec.back = back;
    rec.mat = __ldg(sc.sphMat + idx);
    return true;
}

template <bool COUNT>
__device__ __forceinline__ void leaf_test(const TraceScene& sc, const uint32_t g, const f3 o, const f3 d, const float tMin, float& closest,
                                          bool& hit, Hit& rec, Tally& tl) {
    if (g < sc.T) {
        if (COUNT) tl.tri++;
        if (triangle_hit(sc, g, o, d, tMin, closest, rec)) { hit = true; closest = rec.t; rec.prim = g; }
    } else {
        if (COUNT) tl.sph++;
        if (sphere_hit(sc, g - sc.T, o, d, tMin, closest, rec)) { hit = true; closest = rec.t; rec.prim = g; }
    }
}

// Order-free acceptance for the nearest-first traversal (trace_wave.cu, NODES == 3).  The reference tests a ray's leaves in
// descending primitive id (right-first DFS, leaves kept in input order: ConstructHLBVH.comp:152-169) and accepts
// tMin <= t <= closest-so-far, so its result is the smallest t and, among equal t, the SMALLEST id -- provided no accepted t
// is NaN (a zero-area triangle's NaN normal passes every comparison of triangleHit; from then on the outcome depends on the
// order).  Tested in any order, the same result follows from "t <= closest, but an equal t only replaces a larger id";
// a NaN t raises `poison` and the caller re-traces the ray in the reference's order.
template <bool COUNT>
__device__ __forceinline__ void leaf_test_unordered(const TraceScene& sc, const uint32_t g, const f3 o, const f3 d, const float tMin,
                                                    float& closest, bool& hit, Hit& rec, Tally& tl, bool& poison) {
    Hit tmp;
    bool ok;
    if (g < sc.T) { if (COUNT) tl.tri++; ok = triangle_hit(sc, g, o, d, tMin, closest, tmp); }
    else { if (COUNT) tl.sph++; ok = sphere_hit(sc, g - sc.T, o, d, tMin, closest, tmp); }
    if (ok) {
        if (tmp.t != tmp.t) poison = true;
        else if (!(hit && tmp.t == closest && g > rec.prim)) {
            rec.t = tmp.t; rec.normal = tmp.normal; rec.mat = tmp.mat; rec.back = tmp.back; rec.prim = g;
            closest = tmp.t; hit = true;
        }
    }
}

// hitBVH, raytraceBVH.comp:195-265, over child-pair records.  Every stack entry is a node whose own box test has
// already passed (tested when its parent was expanded); entries >= leafOffset are leaves awaiting their
// primitive test.  Reference visit count: 1 (root) + 2 per expanded internal node.
template <bool COUNT>
__device__ __forceinline__ bool hit_bvh(const TraceScene& sc, const f3 o, const f3 d, const float tMin, const float tMax, Hit& rec,
                                        Tally& tl, unsigned& err) {
    bool hit = false;
    float closest = tMax;
    if (COUNT) { tl.rays++; tl.visits++; }
    {
        const float4 lo = __ldg(sc.rootBox), hi = __ldg(sc.rootBox + 1);
        if (!box_hit(o, d, lo.x, lo.y, lo.z, hi.x, hi.y, hi.z)) return false;
    }
    if (sc.N == 1) {           // the root is the only leaf
        leaf_test<COUNT>(sc, 0u, o, d, tMin, closest, hit, rec, tl);
        return hit;
    }
    const uint32_t leafOffset = sc.N - 1;
    uint32_t stack[STACK_DEPTH];
    int sp = 0;
    uint32_t cur = 0;
    while (true) {
        const float4* p = sc.pairs + 4ull * cur;
        const float4 lLo = __ldg(p), lHi = __ldg(p + 1), rLo = __ldg(p + 2), rHi = __ldg(p + 3);
        if (COUNT) tl.visits += 2;
        const uint32_t li = __float_as_uint(lLo.w), ri = __float_as_uint(lHi.w);
        const bool passR = box_hit(o, d, rLo.x, rLo.y, rLo.z, rHi.x, rHi.y, rHi.z);
        const bool passL = box_hit(o, d, lLo.x, lLo.y, lLo.z, lHi.x, lHi.y, lHi.z);
        uint32_t next = 0xFFFFFFFFu;
        if (passR) {           // right subtree first (reference: descend right, left waits on the stack)
            if (ri >= leafOffset) leaf_test<COUNT>(sc, ri - leafOffset, o, d, tMin, closest, hit, rec, tl);
            else next = ri;
        }
        if (passL) {
            if (next != 0xFFFFFFFFu) {
                if (sp >= STACK_DEPTH) { err |= 1u; break; }
                stack[sp++] = li;                      // must wait until the whole right subtree is done
            } else if (li >= leafOffset) {
                leaf_test<COUNT>(sc, li - leafOffset, o, d, tMin, closest, hit, rec, tl);
            } else {
                next = li;
            }
        }
        while (next == 0xFFFFFFFFu) {
            if (sp == 0) return hit;
            const uint32_t e = stack[--sp];
            if (e >= leafOffset) leaf_test<COUNT>(sc, e - leafOffset, o, d, tMin, closest, hit, rec, tl);
            else next = e;
        }
        cur = next;
    }
    return hit;
}

// randomUnitVector, random.glsl:33-38,60-62 with the pinned sin/cos
__device__ __forceinline__ f3 random_unit_vector(uint32_t& rng) {
    const float PI = 3.1415926535897932385f;
    const float rho = pcg_float(rng);
    const float theta = 0.0f + ((2.0f * PI) - 0.0f) * pcg_float(rng);
    const float phi = 0.0f + (PI - 0.0f) * pcg_float(rng);
    float sp, cp, st, ct;
    pin_sincos(phi, sp, cp);
    pin_sincos(theta, st, ct);
    return normalize(F3(rho * sp * ct, rho * sp * st, rho * cp));
}

// Extension N1 (NOT reference behaviour, DESIGN.md "extensions"): mirror metal, Schlick dielectric with IOR 1.5.
// Same operation order as the oracle's scatter_extension so the two stay bit-identical.
__device__ __forceinline__ bool scatter_extension(const uint32_t type, const f3 albedo, const f3 d, const Hit& rec, uint32_t& rng,
                                                  f3& attenuation, f3& outDir) {
    if (type == RTB_METALLIC) {
        const float dn = dot(d, rec.normal);
        const f3 refl = d - (2.0f * dn) * rec.normal;
        attenuation = albedo;
        outDir = normalize(refl);
        return dot(outDir, rec.normal) > 0;
    }
    if (type == RTB_DIELECTRIC) {
        const float ior = 1.5f;
        const float ri = rec.back ? ior : 1.0f / ior;
        const float cosT = gmin(dot(-d, rec.normal), 1.0f);
        const float sinT = sqrtf(1.0f - cosT * cosT);
        float r0 = (1.0f - ri) / (1.0f + ri);
        r0 = r0 * r0;
        const float om = 1.0f - cosT;
        const float refl = r0 + (1.0f - r0) * ((om * om) * (om * om) * om);
        const float u = pcg_float(rng);
        f3 dir;
        if (ri * sinT > 1.0f || refl > u) {
            dir = d - (2.0f * dot(d, rec.normal)) * rec.normal;
        } else {
            const f3 perp = ri * (d + cosT * rec.normal);
            const float k = 1.0f - dot(perp, perp);
            const f3 par = (-sqrtf(fabsf(k))) * rec.normal;
            dir = perp + par;
        }
        attenuation = albedo;
        outDir = normalize(dir);
        return true;
    }
    return false;
}

}  // namespace rtb
