// trace.cu -- S2 of the frame: raytraceBVH.comp as one persistent sm_100a kernel.
//
// The reference dispatches raytraceBVH.comp once per sample with an image barrier in between
// (RaytracerBVH.cpp:1025-1050); every dispatch re-reads and re-writes the RGBA32F pixel and re-derives the
// camera.  Here one launch renders all samples: each pixel's accumulator and its alpha seed chain
// (raytraceBVH.comp:349-352,372) live in registers, the pixel is read once and written once, and the per-sample
// arithmetic -- seed, ray, traversal order, intersection formulas, scatter, fp32 accumulation order -- is the
// shader's, so the image is bit-identical to `sampleCount` dispatches.
//
// Traversal: explicit-stack DFS over 64-byte child-pair records (common.cuh).  Visiting internal node i fetches
// BOTH child boxes with four LDG.128; the reference's order "push left, descend right" (raytraceBVH.comp:241-244)
// and its `tNear < tFar` line test without t-interval (:184-193) are kept exactly, so the sequence of primitive
// tests -- and therefore every tie-break on equal t (pin U9) -- is the reference's.
#include "common.cuh"
#include "kernels.h"

namespace rtb {

constexpr int TRACE_THREADS = 128;
constexpr int TILE_W = 8, TILE_H = 4;     // one warp = one 8x4 pixel tile (coherent primary rays)
constexpr int STACK_DEPTH = 64;           // Karras tree depth <= 62 (30 code bits + 32 index bits); reference: 128

struct Hit {
    float t;
    f3 normal;
    uint32_t mat;
    uint32_t prim;      // global primitive id g (triangle g < T, sphere T + idx)
    int back;           // backFaceInt
};

struct Tally { unsigned long long rays, visits, tri, sph, mat; };

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

// triangleHit, raytraceBVH.comp:118-149
__device__ __forceinline__ bool triangle_hit(const TraceScene& sc, const uint32_t idx, const f3 o, const f3 d, const float tMin,
                                             const float tMax, Hit& rec) {
    const float4 a = __ldg(sc.tris + 3ull * idx), b = __ldg(sc.tris + 3ull * idx + 1), c = __ldg(sc.tris + 3ull * idx + 2);
    const f3 v0 = xyz(a);
    const f3 u = xyz(b) - v0;
    const f3 v = xyz(c) - v0;
    const f3 nU = cross(u, v);
    const f3 n = normalize(nU);
    const float D = dot(n, v0);
    const f3 w = nU / dot(nU, nU);
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
    rec.mat = __float_as_uint(a.w);
    return true;
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

template <bool COUNT, bool EXT>
__device__ __forceinline__ f3 ray_color(const TraceParams& p, const f3 origin, const f3 dirIn, uint32_t& rng, Tally& tl, unsigned& err,
                                        uint32_t* firstPrim, float* firstT) {
    const TraceScene& sc = p.sc;
    f3 color = F3(0.f, 0.f, 0.f);
    f3 att = F3(1.f, 1.f, 1.f);
    f3 o = origin;
    f3 d = normalize(dirIn);                                     // rayColor: unitDir = normalize(r.direction) (:280)
    for (uint32_t depth = 0; depth < p.maxDepth; depth++) {
        Hit rec;
        const bool hit = hit_bvh<COUNT>(sc, o, d, 0.001f, 10000000.0f, rec, tl, err);     // sceneHit :267-274
        if (depth == 0 && firstPrim) { *firstPrim = hit ? rec.prim : 0xFFFFFFFFu; *firstT = hit ? rec.t : 0.0f; }
        if (!hit) {
            color = color + F3(0.f, 0.f, 0.f) * att;             // _BACKGROUND_COLOR * globalAttenuation (:284)
            break;
        }
        const float4 m = __ldg(sc.mats + rec.mat);
        if (COUNT) tl.mat++;
        const uint32_t type = __float_as_uint(m.w);
        const f3 albedo = xyz(m);
        const f3 emitted = (type == RTB_LIGHT) ? albedo : F3(0.f, 0.f, 0.f);              // emitted :94-99
        color = color + emitted * att;                                                     // :293
        if (type == RTB_DIFFUSE) {                                                         // scatter :100-115
            const f3 P = o + rec.t * d;
            const f3 nd = normalize(rec.normal + random_unit_vector(rng));
            o = P; d = nd;
            att = att * albedo;
        } else if (EXT && type != RTB_LIGHT) {
            f3 a2, nd;
            const f3 P = o + rec.t * d;
            if (!scatter_extension(type, albedo, d, rec, rng, a2, nd)) break;
            o = P; d = nd;
            att = att * a2;
        } else {
            break;                                               // LIGHT / METALLIC / DIELECTRIC absorb (D4, U8)
        }
    }
    return color;
}

template <bool COUNT, bool EXT>
__global__ void __launch_bounds__(TRACE_THREADS) trace_kernel(const TraceParams p) {
    const unsigned lane = threadIdx.x & 31;
    Tally tl = { 0, 0, 0, 0, 0 };
    unsigned long long samplesDone = 0;
    unsigned err = 0;
    const uint32_t numTiles = p.tilesX * p.tilesY;
    while (true) {
        // persistent threads: every warp pulls the next 8x4 pixel tile from a global counter
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(p.workCounter, 1u);
        tile = __shfl_sync(0xFFFFFFFFu, tile, 0);
        if (tile >= numTiles) break;
        const uint32_t tx = tile % p.tilesX, ty = tile / p.tilesX;
        const uint32_t x = tx * TILE_W + (lane & (TILE_W - 1));
        const uint32_t j = ty * TILE_H + (lane / TILE_W);                                   // local row
        const uint32_t y = ((j / p.bandRows) * p.bandStep + p.bandFirst) * p.bandRows + (j % p.bandRows);
        if (x >= p.W || j >= p.localRows || y >= p.H) continue;
        const size_t px = (size_t)j * p.W + x;
        const float4 cur = p.image[px];                                                     // imageLoad :349
        const uint32_t base = (600u * x + y) * (p.randomState + 1u);                        // random.glsl:10
        float alpha = cur.w;
        for (uint32_t k = 0; k < p.sampleSkip; k++) {                                       // fast-forward the seed chain
            uint32_t s = base + alpha_to_u32(alpha);
            alpha = pcg_float(s);
        }
        // getRay :329-342 -- no jitter, pinhole: identical for every sample of the pixel
        const f3 pixelSample = (p.cam.pixel00 + (float)x * p.cam.deltaU) + (float)y * p.cam.deltaV;
        const f3 rayDir = normalize(pixelSample - p.cam.origin);
        f3 rgb = F3(cur.x, cur.y, cur.z);
        uint32_t rng = 0;
        for (uint32_t k = 0; k < p.sampleCount; k++) {                                      // one reference dispatch each
            rng = base + alpha_to_u32(alpha);                                               // :350 (stepRNG :351 is a no-op, U2)
            const float nextRandom = pcg_float(rng);                                        // :352
            uint32_t fp = 0xFFFFFFFFu; float ft = 0.f;
            const bool wantFirst = (k == 0) && (p.hitPrim != nullptr);
            const f3 color = ray_color<COUNT, EXT>(p, p.cam.origin, rayDir, rng, tl, err, wantFirst ? &fp : nullptr, &ft);
            rgb = color + rgb;                                                              // pixelColor + currentColor.xyz :372
            alpha = nextRandom;
            if (wantFirst) { p.hitPrim[px] = fp; if (p.hitT) p.hitT[px] = ft; }
        }
        if (COUNT) samplesDone += p.sampleCount;
        p.image[px] = make_float4(rgb.x, rgb.y, rgb.z, alpha);                              // imageStore :374
        if (p.rngOut) p.rngOut[px] = rng;
    }
    if (err) atomicOr(p.errFlag, err);
    if (COUNT) {
        unsigned long long v[6] = { tl.rays, tl.visits, tl.tri, tl.sph, tl.mat, samplesDone };
#pragma unroll
        for (int i = 0; i < 6; i++) {
            unsigned long long x = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xFFFFFFFFu, x, o);
            if (lane == 0 && x) atomicAdd(p.counters + i, x);
        }
    }
}

// clear to (0,0,0,1): vkCmdClearColorImage, RaytracerBVH.cpp:772-776
__global__ void __launch_bounds__(256) clear_image_kernel(float4* img, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        img[i] = make_float4(0.f, 0.f, 0.f, 1.f);
}

// SingleTriangleFullScreen.frag:13-21: clamp(sqrt(rgb / raysPerPixel), 0, 1) -> UNORM8 (NaN pinned to 0)
__global__ void __launch_bounds__(256) resolve_kernel(const float4* __restrict__ img, size_t n, float rpp, uchar4* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float4 c = img[i];
        float r = sqrtf(c.x / rpp), g = sqrtf(c.y / rpp), b = sqrtf(c.z / rpp);
        r = (r > 0.f) ? r : 0.f; r = (r < 1.f) ? r : 1.f;
        g = (g > 0.f) ? g : 0.f; g = (g < 1.f) ? g : 1.f;
        b = (b > 0.f) ? b : 0.f; b = (b < 1.f) ? b : 1.f;
        out[i] = make_uchar4((unsigned char)(r * 255.0f + 0.5f), (unsigned char)(g * 255.0f + 0.5f), (unsigned char)(b * 255.0f + 0.5f), 255);
    }
}

int trace_blocks_per_sm(bool count, bool ext) {
    int nb = 0;
    if (count) {
        if (ext) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<true, true>, TRACE_THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<true, false>, TRACE_THREADS, 0);
    } else {
        if (ext) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<false, true>, TRACE_THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<false, false>, TRACE_THREADS, 0);
    }
    return nb > 0 ? nb : 1;
}

void launch_trace(cudaStream_t st, TraceParams p, bool count, bool ext, int smCount) {
    p.tilesX = (p.W + TILE_W - 1) / TILE_W;
    p.tilesY = (p.localRows + TILE_H - 1) / TILE_H;
    const uint64_t numWarps = (uint64_t)p.tilesX * p.tilesY;
    // persistent grid: resident CTAs per SM x SM count (never more warps than tiles)
    uint64_t grid = (uint64_t)smCount * trace_blocks_per_sm(count, ext);
    const uint64_t need = (numWarps + TRACE_THREADS / 32 - 1) / (TRACE_THREADS / 32);
    if (grid > need) grid = need;
    if (grid == 0) return;
    cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), st);
    if (count) {
        if (ext) trace_kernel<true, true><<<(unsigned)grid, TRACE_THREADS, 0, st>>>(p);
        else trace_kernel<true, false><<<(unsigned)grid, TRACE_THREADS, 0, st>>>(p);
    } else {
        if (ext) trace_kernel<false, true><<<(unsigned)grid, TRACE_THREADS, 0, st>>>(p);
        else trace_kernel<false, false><<<(unsigned)grid, TRACE_THREADS, 0, st>>>(p);
    }
}

void launch_clear_image(cudaStream_t st, void* img, size_t pixels, int smCount) {
    if (!pixels) return;
    size_t grid = (pixels + 255) / 256;
    const size_t cap = (size_t)smCount * 16;
    if (grid > cap) grid = cap;
    clear_image_kernel<<<(unsigned)grid, 256, 0, st>>>((float4*)img, pixels);
}
void launch_resolve(cudaStream_t st, const void* img, size_t pixels, uint32_t rpp, void* out, int smCount) {
    if (!pixels) return;
    size_t grid = (pixels + 255) / 256;
    const size_t cap = (size_t)smCount * 16;
    if (grid > cap) grid = cap;
    resolve_kernel<<<(unsigned)grid, 256, 0, st>>>((const float4*)img, pixels, (float)rpp, (uchar4*)out);
}

}  // namespace rtb
