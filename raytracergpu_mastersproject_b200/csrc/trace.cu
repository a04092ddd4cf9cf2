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
#include "kernels.h"
#include "trace_common.cuh"

namespace rtb {

// sceneHit of the NON-BVH program (raytrace.comp:167-190, the Config::Programs::Raytracer host): every triangle in index
// order, then every sphere, with the running closest-so-far.  O(N) per ray -- only meaningful for small scenes, and an
// independent check of the BVH walk's closest hits.
template <bool COUNT>
__device__ __forceinline__ bool hit_linear(const TraceScene& sc, const f3 o, const f3 d, const float tMin, const float tMax, Hit& rec, Tally& tl) {
    bool hit = false;
    float closest = tMax;
    if (COUNT) tl.rays++;
    for (uint32_t g = 0; g < sc.N; g++) leaf_test<COUNT>(sc, g, o, d, tMin, closest, hit, rec, tl);
    return hit;
}

constexpr int TRACE_THREADS = 128;

template <bool COUNT, bool EXT, bool LINEAR>
__device__ __forceinline__ f3 ray_color(const TraceParams& p, const f3 origin, const f3 dirIn, uint32_t& rng, Tally& tl, unsigned& err,
                                        uint32_t* firstPrim, float* firstT) {
    const TraceScene& sc = p.sc;
    f3 color = F3(0.f, 0.f, 0.f);
    f3 att = F3(1.f, 1.f, 1.f);
    f3 o = origin;
    f3 d = normalize(dirIn);                                     // rayColor: unitDir = normalize(r.direction) (:280)
    for (uint32_t depth = 0; depth < p.maxDepth; depth++) {
        Hit rec;
        const bool hit = LINEAR ? hit_linear<COUNT>(sc, o, d, 0.001f, 10000000.0f, rec, tl)
                                : hit_bvh<COUNT>(sc, o, d, 0.001f, 10000000.0f, rec, tl, err);     // sceneHit :267-274
        if (depth == 0 && firstPrim) { *firstPrim = hit ? rec.prim : 0xFFFFFFFFu; *firstT = hit ? rec.t : 0.0f; }
        if (!hit) {
            color = color + p.background * att;                  // _BACKGROUND_COLOR * globalAttenuation (:284)
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

template <bool COUNT, bool EXT, bool LINEAR>
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
            const f3 color = ray_color<COUNT, EXT, LINEAR>(p, p.cam.origin, rayDir, rng, tl, err, wantFirst ? &fp : nullptr, &ft);
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

// logistic.comp:23-35 -- the LogisticMap demo program: x' = x * r * (1 - x), then plot (r / 4 * W, (1 - x') * H)
__global__ void __launch_bounds__(256) logistic_kernel(float2* points, uint32_t count, uchar4* image, uint32_t W, uint32_t H, uchar4 color) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 c = points[i];
    const float nx = c.x * c.y * (1.0f - c.x);
    points[i] = make_float2(nx, c.y);
    const int xc = __float2int_rz((c.y / 4.0f) * (float)W);          // int(float): truncation, F2I saturates, NaN -> 0
    const int yc = __float2int_rz(((1 - nx) / 1.0f) * (float)H);
    if (xc >= 0 && yc >= 0 && (uint32_t)xc < W && (uint32_t)yc < H) image[(size_t)yc * W + xc] = color;
}
void launch_logistic(cudaStream_t st, void* points, uint32_t count, void* image, uint32_t W, uint32_t H, const float* pixelColor) {
    if (!count) return;
    unsigned char c[4];
    for (int k = 0; k < 4; k++) {
        float v = pixelColor[k];
        v = (v > 0.0f) ? v : 0.0f; v = (v < 1.0f) ? v : 1.0f;
        c[k] = (unsigned char)(v * 255.0f + 0.5f);
    }
    logistic_kernel<<<(count + 255) / 256, 256, 0, st>>>((float2*)points, count, (uchar4*)image, W, H, make_uchar4(c[0], c[1], c[2], c[3]));
}

template <bool COUNT, bool EXT, bool LINEAR>
static void launch_trace_variant(cudaStream_t st, const TraceParams& p, int smCount, uint64_t need) {
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, trace_kernel<COUNT, EXT, LINEAR>, TRACE_THREADS, 0);
    uint64_t grid = (uint64_t)smCount * (nb > 0 ? nb : 1);   // persistent grid: resident CTAs per SM x SM count
    if (grid > need) grid = need;
    trace_kernel<COUNT, EXT, LINEAR><<<(unsigned)grid, TRACE_THREADS, 0, st>>>(p);
}

void launch_trace(cudaStream_t st, TraceParams p, bool count, bool ext, bool linear, int smCount) {
    p.tilesX = (p.W + TILE_W - 1) / TILE_W;
    p.tilesY = (p.localRows + TILE_H - 1) / TILE_H;
    const uint64_t numWarps = (uint64_t)p.tilesX * p.tilesY;
    const uint64_t need = (numWarps + TRACE_THREADS / 32 - 1) / (TRACE_THREADS / 32);
    if (need == 0) return;
    p.background = linear ? F3(0.1f, 0.1f, 0.3f) : F3(0.f, 0.f, 0.f);   // raytrace.comp:43 / raytraceBVH.comp:52
    cudaMemsetAsync(p.workCounter, 0, sizeof(unsigned int), st);
    const int v = (count ? 4 : 0) | (ext ? 2 : 0) | (linear ? 1 : 0);
    switch (v) {
    case 0: launch_trace_variant<false, false, false>(st, p, smCount, need); break;
    case 1: launch_trace_variant<false, false, true>(st, p, smCount, need); break;
    case 2: launch_trace_variant<false, true, false>(st, p, smCount, need); break;
    case 3: launch_trace_variant<false, true, true>(st, p, smCount, need); break;
    case 4: launch_trace_variant<true, false, false>(st, p, smCount, need); break;
    case 5: launch_trace_variant<true, false, true>(st, p, smCount, need); break;
    case 6: launch_trace_variant<true, true, false>(st, p, smCount, need); break;
    default: launch_trace_variant<true, true, true>(st, p, smCount, need); break;
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
