// bvh_build.cu -- S1 of the frame: the reference's five BVH-build dispatches + ModelToWorld as sm_100a kernels
// (K1..K3, K5, K6 of DESIGN.md; K4, the sort, is radix_sort.cu), plus the kernels that derive the 16-byte
// aligned traversal records from the reference-layout arrays.
//
// Output contract: the HLBVHNode[2N-1] / MortonPrimitive[N] / enclosing-AABB buffers are bit-identical to what
// the reference's shaders leave in their SSBOs (as pinned in DESIGN.md); every kernel cites the shader it replaces.
#include "build_shared.cuh"
#include "common.cuh"
#include "kernels.h"

namespace rtb {

// ---------------------------------------------------------------------------------------------------------
// K1  ModelSpaceToWorldSpace.comp:32-48 -- in place; one thread per primitive; grid covers T+S
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 mat_mul_point(const float4 c0, const float4 c1, const float4 c2, const float4 c3, float4 p) {
    // mat4 * vec4(p.xyz, 1.0), per component ((c0*x + c1*y) + c2*z) + c3*w
    const float x = p.x, y = p.y, z = p.z, w = 1.0f;
    float4 r;
    r.x = ((c0.x * x + c1.x * y) + c2.x * z) + c3.x * w;
    r.y = ((c0.y * x + c1.y * y) + c2.y * z) + c3.y * w;
    r.z = ((c0.z * x + c1.z * y) + c2.z * z) + c3.z * w;
    r.w = ((c0.w * x + c1.w * y) + c2.w * z) + c3.w * w;
    return r;
}

__global__ void __launch_bounds__(256) model_to_world_kernel(const float4* __restrict__ models, float4* tris, uint32_t T,
                                                             float4* sphs, uint32_t S) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) {
        float4* t = tris + 4ull * i;                       // 64-byte record: v0 v1 v2 (mat, model, pad, pad)
        const uint4 idx = *reinterpret_cast<const uint4*>(t + 3);
        const float4* m = models + 4ull * idx.y;
        const float4 c0 = __ldg(m), c1 = __ldg(m + 1), c2 = __ldg(m + 2), c3 = __ldg(m + 3);
        t[0] = mat_mul_point(c0, c1, c2, c3, t[0]);
        t[1] = mat_mul_point(c0, c1, c2, c3, t[1]);
        t[2] = mat_mul_point(c0, c1, c2, c3, t[2]);
    } else if (i < T + S) {
        float4* s = sphs + 2ull * (i - T);                 // 32-byte record: center (radius, mat, model, pad)
        const uint4 idx = *reinterpret_cast<const uint4*>(s + 1);
        const float4* m = models + 4ull * idx.z;
        s[0] = mat_mul_point(__ldg(m), __ldg(m + 1), __ldg(m + 2), __ldg(m + 3), s[0]);
    }
}

// ---------------------------------------------------------------------------------------------------------
// K2  GetEnclosingAABB.comp:64-110 -- the reference reduces in ONE workgroup with a spin semaphore; here a
// grid-wide reduction: registers -> warp redux -> global atomics on order-preserving integer images of the
// floats (total order, -0 < +0, so the result is independent of the reduction order).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ f3 prim_center(const float4* __restrict__ tris, uint32_t T, const float4* __restrict__ sphs, uint32_t i) {
    if (i < T) {  // getTriangleCenter: ((v0 + v1 + v2) / 3).xyz  (GetEnclosingAABB.comp:40-42)
        const float4 a = tris[4ull * i], b = tris[4ull * i + 1], c = tris[4ull * i + 2];
        return F3(((a.x + b.x) + c.x) / 3.0f, ((a.y + b.y) + c.y) / 3.0f, ((a.z + b.z) + c.z) / 3.0f);
    }
    return xyz(sphs[2ull * (i - T)]);
}

__global__ void enclosing_init_kernel(uint32_t* red, int initInf) {
    // pin U4: the shader's localMin/localMax are never initialised (:69-70) -> read as 0.0
    const float lo = initInf ? __int_as_float(0x7f800000) : 0.0f;
    const float hi = initInf ? __int_as_float(0xff800000) : 0.0f;
    if (threadIdx.x < 3) red[threadIdx.x] = f2ord(lo);
    else if (threadIdx.x < 6) red[threadIdx.x] = f2ord(hi);
    else if (threadIdx.x < 9) red[threadIdx.x] = 0xFFFFFFFFu;       // [6..11]: bounds of the primitives (fused build only)
    else if (threadIdx.x < 12) red[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) enclosing_reduce_kernel(const float4* __restrict__ tris, uint32_t T,
                                                               const float4* __restrict__ sphs, uint32_t S, uint32_t* red) {
    uint32_t lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u };
    const uint32_t n = T + S;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const f3 c = prim_center(tris, T, sphs, i);
        const uint32_t ox = f2ord(c.x), oy = f2ord(c.y), oz = f2ord(c.z);
        lo[0] = min(lo[0], ox); lo[1] = min(lo[1], oy); lo[2] = min(lo[2], oz);
        hi[0] = max(hi[0], ox); hi[1] = max(hi[1], oy); hi[2] = max(hi[2], oz);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        lo[k] = __reduce_min_sync(0xFFFFFFFFu, lo[k]);
        hi[k] = __reduce_max_sync(0xFFFFFFFFu, hi[k]);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (lo[k] != 0xFFFFFFFFu) atomicMin(&red[k], lo[k]);
            if (hi[k] != 0u) atomicMax(&red[3 + k], hi[k]);
        }
    }
}

// Fused S1 front end of rtb_build_bvh: K1 (transform in place) and K2's reduction in ONE pass over the primitives -- the centroid
// is taken from the values just stored, so the reduction sees exactly what enclosing_reduce_kernel would read back.  Also reduces
// the bounds of the primitives themselves into red[6..11] (the origin region / R of the hit-point slack, eta_leaf below, without
// waiting for the refit to produce the root box).  Block-level reduction first: 12 atomics per block.
__global__ void __launch_bounds__(256) model_to_world_enclosing_kernel(const float4* __restrict__ models, float4* tris, uint32_t T,
                                                                       float4* sphs, uint32_t S, uint32_t* red) {
    __shared__ uint32_t sm[12][8];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v[12] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };
    if (i < T + S) {
        f3 c, bl, bh;
        if (i < T) {
            float4* t = tris + 4ull * i;
            const uint4 idx = *reinterpret_cast<const uint4*>(t + 3);
            const float4* m = models + 4ull * idx.y;
            const float4 c0 = __ldg(m), c1 = __ldg(m + 1), c2 = __ldg(m + 2), c3 = __ldg(m + 3);
            const float4 a = mat_mul_point(c0, c1, c2, c3, t[0]), b = mat_mul_point(c0, c1, c2, c3, t[1]), d = mat_mul_point(c0, c1, c2, c3, t[2]);
            t[0] = a; t[1] = b; t[2] = d;
            c = F3(((a.x + b.x) + d.x) / 3.0f, ((a.y + b.y) + d.y) / 3.0f, ((a.z + b.z) + d.z) / 3.0f);      // getTriangleCenter :40-42
            bl = F3(fminf(a.x, fminf(b.x, d.x)), fminf(a.y, fminf(b.y, d.y)), fminf(a.z, fminf(b.z, d.z)));
            bh = F3(fmaxf(a.x, fmaxf(b.x, d.x)), fmaxf(a.y, fmaxf(b.y, d.y)), fmaxf(a.z, fmaxf(b.z, d.z)));
        } else {
            float4* s = sphs + 2ull * (i - T);
            const uint4 idx = *reinterpret_cast<const uint4*>(s + 1);
            const float4* m = models + 4ull * idx.z;
            const float4 a = mat_mul_point(__ldg(m), __ldg(m + 1), __ldg(m + 2), __ldg(m + 3), s[0]);
            s[0] = a;
            c = xyz(a);
            const float r = fabsf(__uint_as_float(idx.x));
            bl = F3(a.x - r, a.y - r, a.z - r); bh = F3(a.x + r, a.y + r, a.z + r);
        }
        v[0] = f2ord(c.x); v[1] = f2ord(c.y); v[2] = f2ord(c.z); v[3] = v[0]; v[4] = v[1]; v[5] = v[2];
        v[6] = f2ord(bl.x); v[7] = f2ord(bl.y); v[8] = f2ord(bl.z); v[9] = f2ord(bh.x); v[10] = f2ord(bh.y); v[11] = f2ord(bh.z);
    }
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const bool isMin = (k % 6) < 3;
        const uint32_t r = isMin ? __reduce_min_sync(0xFFFFFFFFu, v[k]) : __reduce_max_sync(0xFFFFFFFFu, v[k]);
        if (lane == 0) sm[k][warp] = r;
    }
    __syncthreads();
    if (threadIdx.x < 12) {
        const int k = threadIdx.x;
        const bool isMin = (k % 6) < 3;
        uint32_t r = sm[k][0];
        for (int w = 1; w < 8; w++) r = isMin ? min(r, sm[k][w]) : max(r, sm[k][w]);
        if (isMin) { if (r != 0xFFFFFFFFu) atomicMin(&red[k], r); }
        else if (r != 0u) atomicMax(&red[k], r);
    }
}

__global__ void enclosing_finalize_kernel(const uint32_t* __restrict__ red, float4* enclosing) {
    if (threadIdx.x != 0) return;
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
        // the shader presets eMin/eMax to +-1e9 (:73-74) and folds the subgroup results in with min/max
        float a = ord2f(red[k]), b = ord2f(red[3 + k]);
        lo[k] = (a < 1000000000.0f) ? a : 1000000000.0f;
        hi[k] = (b > -1000000000.0f) ? b : -1000000000.0f;
        if (hi[k] - lo[k] < 0.001f) { lo[k] -= 0.0005f; hi[k] += 0.0005f; }   // padAABB :49-62
    }
    enclosing[0] = make_float4(lo[0], lo[1], lo[2], 0.0f);   // w lanes: min(1e9, 0) / max(-1e9, 0)
    enclosing[1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// K3  GenerateMortonCodesOfPrimitives.comp:41-94
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t separate_bits_by3(uint32_t v) {   // :41-50
    if (v == 1024u) v--;
    v = (v | (v << 16)) & 50331903u;
    v = (v | (v << 8)) & 50393103u;
    v = (v | (v << 4)) & 51130563u;
    v = (v | (v << 2)) & 153391689u;
    return v;
}
__device__ __forceinline__ uint32_t morton_of(f3 c, float4 eMin, float4 eMax) {
    // quantizeForMorton :60-65 divides the coordinate itself (not coord - eMin) by the span: follow the code (U5)
    const float sx = eMax.x - eMin.x, sy = eMax.y - eMin.y, sz = eMax.z - eMin.z;
    const uint32_t qx = __float2uint_rz((c.x / sx) * 1024.0f);   // F2I.U32 saturates, NaN -> 0
    const uint32_t qy = __float2uint_rz((c.y / sy) * 1024.0f);
    const uint32_t qz = __float2uint_rz((c.z / sz) * 1024.0f);
    return separate_bits_by3(qz) << 2 | separate_bits_by3(qy) << 1 | separate_bits_by3(qx);
}

// writes either the 12-byte MortonPrimitive records (morton != nullptr) or the SoA key/value arrays the sort uses
__global__ void __launch_bounds__(256) morton_kernel(const float4* __restrict__ tris, uint32_t T, const float4* __restrict__ sphs,
                                                     uint32_t S, const float4* __restrict__ enclosing, uint32_t* morton,
                                                     uint32_t* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T + S) return;
    const uint32_t code = morton_of(prim_center(tris, T, sphs, i), __ldg(enclosing), __ldg(enclosing + 1));
    if (morton) {
        morton[3ull * i] = code;
        morton[3ull * i + 1] = i < T ? i : i - T;
        morton[3ull * i + 2] = i < T ? RTB_TRIANGLE_PRIMITIVE : RTB_SPHERE_PRIMITIVE;
    }
    if (keys) { keys[i] = code; vals[i] = i; }
}

// MortonPrimitive records <-> SoA (code, global primitive id)
__global__ void __launch_bounds__(256) morton_unpack_kernel(const uint32_t* __restrict__ morton, uint32_t n, uint32_t T,
                                                            uint32_t* keys, uint32_t* vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    keys[i] = morton[3ull * i];
    vals[i] = morton[3ull * i + 1] + (morton[3ull * i + 2] == RTB_SPHERE_PRIMITIVE ? T : 0u);
}
__global__ void __launch_bounds__(256) morton_repack_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                            uint32_t n, uint32_t T, uint32_t* morton) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t g = vals[i];
    morton[3ull * i] = keys[i];
    morton[3ull * i + 1] = g < T ? g : g - T;
    morton[3ull * i + 2] = g < T ? RTB_TRIANGLE_PRIMITIVE : RTB_SPHERE_PRIMITIVE;
}

// ---------------------------------------------------------------------------------------------------------
// K5  ConstructHLBVH.comp -- Karras topology over the sorted codes + leaf boxes in ORIGINAL primitive order
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pad_axis(float& lo, float& hi) {          // padAABB :41-54
    if (hi - lo < 0.001f) { lo -= 0.0005f; hi += 0.0005f; }
}

// Optional by-products of the K5 leaf pass in the fused build (rtb_build_bvh): the thread that builds leaf g already holds the
// primitive, so it also writes the exact leaf box, the packed primitive record (pack_prims_kernel's) and the primitive's hit-point
// slack (eta_leaf_kernel's, with the origin region taken from the primitive bounds model_to_world_enclosing_kernel reduced).
struct LeafExtras {
    float4* leafBox;            // [N][2]
    float4* ptris; float4* psphs; uint32_t* sphMat;
    float* etaNode;             // [2N-1]: leaf values at [N-1+g]
    const uint32_t* primBounds; // ordered-int min.xyz, max.xyz of all primitives (red[6..11])
    float4* originRegion;       // [2]
    f3 cam;
    unsigned int* bigCount;     // traversal hierarchy (traversal_tree.cu): number of big leaves, the first MAX_BIG of them,
    uint32_t* bigList;
    uint32_t* smallBounds;      // and the ordered-int bounds of the centres of all the others
};
__device__ __forceinline__ float eta_triangle(const f3 u, const f3 v, const f3 w, const float R);
__device__ __forceinline__ float eta_sphere(const float4 sp, const float4 lo, const float4 hi, const float R);
__device__ __forceinline__ void origin_region_of(float4 lo, float4 hi, const f3 cam, const bool mayExtend, float4& rlo, float4& rhi);

// A triangle's slack grows linearly with the region's size (1e-5 R: negligible against a triangle), a sphere's with the SQUARE of its
// diameter divided by the radius: extending the region to a distant camera would inflate every sphere box (measured on C3, 100 k spheres
// of radius 1..4: 43.0 -> 48.7 record fetches per ray, 491 -> 552 ms).  So the region is extended only for scenes that are mostly triangles.
__device__ __forceinline__ bool region_may_extend(const uint32_t T, const uint32_t S) { return S <= T / 16u; }

template <bool EXTRAS>
__global__ void __launch_bounds__(256) hlbvh_kernel(const float4* __restrict__ tris, uint32_t T, const float4* __restrict__ sphs,
                                                    uint32_t S, Codes codes, uint32_t* nodes /*10 words each*/, uint2* cinfo, const LeafExtras ex) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = (int)(T + S);
    const int leafOffset = n - 1;
    uint32_t sb[6] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u, 0u };   // EXTRAS: ordered-int centre of this leaf if it is not big
    if (g < (uint32_t)n) {  // leaf :152-169
        float mnx, mxx, mny, mxy, mnz, mxz;
        uint32_t type, prim;
        float4 rlo, rhi; float R = 0.f;
        if (EXTRAS) {
            // bounds of all primitives, padded like a leaf box can be (padAABB), -> origin region and R
            const float4 pl = make_float4(ord2f(ex.primBounds[0]) - 0.0005f, ord2f(ex.primBounds[1]) - 0.0005f, ord2f(ex.primBounds[2]) - 0.0005f, 0.f);
            const float4 ph = make_float4(ord2f(ex.primBounds[3]) + 0.0005f, ord2f(ex.primBounds[4]) + 0.0005f, ord2f(ex.primBounds[5]) + 0.0005f, 0.f);
            origin_region_of(pl, ph, ex.cam, region_may_extend(T, S), rlo, rhi);
            if (g == 0) { ex.originRegion[0] = rlo; ex.originRegion[1] = rhi; }
            R = fmaxf(fmaxf(fmaxf(fabsf(rlo.x), fabsf(rhi.x)), fmaxf(fabsf(rlo.y), fabsf(rhi.y))), fmaxf(fabsf(rlo.z), fabsf(rhi.z)));
        }
        if (g < T) {        // getTriangleAABB :132-143 : min(v0, min(v1, v2))
            const float4 a = tris[4ull * g], b = tris[4ull * g + 1], c = tris[4ull * g + 2];
            mnx = gmin(a.x, gmin(b.x, c.x)); mxx = gmax(a.x, gmax(b.x, c.x));
            mny = gmin(a.y, gmin(b.y, c.y)); mxy = gmax(a.y, gmax(b.y, c.y));
            mnz = gmin(a.z, gmin(b.z, c.z)); mxz = gmax(a.z, gmax(b.z, c.z));
            type = RTB_TRIANGLE_PRIMITIVE; prim = g;
            if (EXTRAS) {   // pack_prims_kernel's triangle record: triangleHit's ray-independent prologue, same operation order
                const uint4 idx = *reinterpret_cast<const uint4*>(tris + 4ull * g + 3);
                const f3 v0 = xyz(a), u = xyz(b) - v0, v = xyz(c) - v0;
                const f3 nU = cross(u, v);
                const f3 nn = normalize(nU);
                const f3 w = nU / dot(nU, nU);
                ex.ptris[4ull * g] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(idx.x));
                ex.ptris[4ull * g + 1] = make_float4(nn.x, nn.y, nn.z, u.x);
                ex.ptris[4ull * g + 2] = make_float4(u.y, u.z, v.x, v.y);
                ex.ptris[4ull * g + 3] = make_float4(v.z, w.x, w.y, w.z);
                ex.etaNode[leafOffset + (int)g] = eta_triangle(u, v, w, R);
            }
        } else {            // getSphereAABB :117-130
            const float4 c = sphs[2ull * (g - T)];
            const float4 rr = sphs[2ull * (g - T) + 1];
            const float r = rr.x;
            const float lx = c.x - r, ly = c.y - r, lz = c.z - r, rx = c.x + r, ry = c.y + r, rz = c.z + r;
            mnx = gmin(lx, rx); mxx = gmax(lx, rx);
            mny = gmin(ly, ry); mxy = gmax(ly, ry);
            mnz = gmin(lz, rz); mxz = gmax(lz, rz);
            type = RTB_SPHERE_PRIMITIVE; prim = g - T;
            if (EXTRAS) {
                ex.psphs[g - T] = make_float4(c.x, c.y, c.z, r);
                ex.sphMat[g - T] = __float_as_uint(rr.y);
                ex.etaNode[leafOffset + (int)g] = eta_sphere(make_float4(c.x, c.y, c.z, r), rlo, rhi, R);
            }
        }
        pad_axis(mnx, mxx); pad_axis(mny, mxy); pad_axis(mnz, mxz);
        if (EXTRAS) {
            ex.leafBox[2ull * g] = make_float4(mnx, mny, mnz, 0.f);
            ex.leafBox[2ull * g + 1] = make_float4(mxx, mxy, mxz, 0.f);
            // traversal hierarchy (traversal_tree.cu): the BIG leaves are listed, the centres of the others are bounded (block reduction below)
            if (box_is_big(mnx, mxx, mny, mxy, mnz, mxz, hoist_threshold(ex.primBounds))) {
                const unsigned int k = atomicAdd(ex.bigCount, 1u);
                if (k < MAX_BIG) ex.bigList[k] = g;
            } else {
                const float cx = 0.5f * (mnx + mxx), cy = 0.5f * (mny + mxy), cz = 0.5f * (mnz + mxz);
                sb[0] = sb[3] = f2ord(cx); sb[1] = sb[4] = f2ord(cy); sb[2] = sb[5] = f2ord(cz);
            }
        }
        uint32_t* nd = nodes + 10ull * (uint32_t)(leafOffset + (int)g);   // 40-byte records: 8-byte aligned
        reinterpret_cast<float2*>(nd)[0] = make_float2(mnx, mxx);
        reinterpret_cast<float2*>(nd)[1] = make_float2(mny, mxy);
        reinterpret_cast<float2*>(nd)[2] = make_float2(mnz, mxz);
        reinterpret_cast<uint2*>(nd)[3] = make_uint2(0u, 0u);
        reinterpret_cast<uint2*>(nd)[4] = make_uint2(prim, type);
    }
    if ((int)g < n - 1) {   // internal :172-209
        int leftChild, rightChild;
        karras_children(codes, (int)g, leafOffset, leftChild, rightChild);   // determineRange :72-96, findSplit :98-115
        uint32_t* nd = nodes + 10ull * g;
        reinterpret_cast<float2*>(nd)[0] = make_float2(0.f, 0.f);
        reinterpret_cast<float2*>(nd)[1] = make_float2(0.f, 0.f);
        reinterpret_cast<float2*>(nd)[2] = make_float2(0.f, 0.f);
        reinterpret_cast<uint2*>(nd)[3] = make_uint2((uint32_t)leftChild, (uint32_t)rightChild);
        reinterpret_cast<uint2*>(nd)[4] = make_uint2(0u, 0u);
        cinfo[leftChild] = make_uint2(g, 0u);
        cinfo[rightChild] = make_uint2(g, 0u);
    }
    if (g == 0) cinfo[0] = make_uint2(0u, 0u);   // :212-214
    if (EXTRAS) {                                // bounds of the non-big leaves' centres -> ex.smallBounds (ordered ints), 6 atomics per block
        __shared__ uint32_t sm[6][8];
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < 6; k++) {
            const uint32_t r = k < 3 ? __reduce_min_sync(0xFFFFFFFFu, sb[k]) : __reduce_max_sync(0xFFFFFFFFu, sb[k]);
            if (lane == 0) sm[k][warp] = r;
        }
        __syncthreads();
        if (threadIdx.x < 6) {
            const int k = threadIdx.x;
            uint32_t r = sm[k][0];
            for (int w = 1; w < 8; w++) r = k < 3 ? min(r, sm[k][w]) : max(r, sm[k][w]);
            if (k < 3) { if (r != 0xFFFFFFFFu) atomicMin(&ex.smallBounds[k], r); }
            else if (r != 0u) atomicMax(&ex.smallBounds[k], r);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// K6  ConstructAABBsOfInternalNodes.comp:38-65 -- bottom-up refit; the second arrival at a node unions its
// children.  The shader relies on `coherent`; here: L2-scoped loads (__ldcg) + __threadfence() before the
// counter atomic.  fp min/max of fixed operands (left, right) -> deterministic whatever the arrival order.
// ---------------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256) refit_kernel(uint32_t* nodes, uint2* cinfo, uint32_t n, float4* pairs, float4* rootBox, float* etaNode) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n) return;
    const uint32_t leafOffset = n - 1;
    // The climbing thread keeps the box (and slack) of the subtree it comes from in registers; a node's static words -- its children and
    // its parent -- are fetched beside the arrival atomic: two dependent round trips per level (atomic, sibling box) instead of four.
    uint32_t me = leafOffset + g;
    uint32_t nodeId = __ldcg(&cinfo[me]).x;
    float2 mx, my, mz;
    { const float2* B = reinterpret_cast<const float2*>(nodes + 10ull * me); mx = __ldcg(B); my = __ldcg(B + 1); mz = __ldcg(B + 2); }
    float me_eta = etaNode ? __ldcg(&etaNode[me]) : 0.f;
    while (true) {
        uint32_t* nd = nodes + 10ull * nodeId;
        const uint2 ch = __ldcg(reinterpret_cast<const uint2*>(nd) + 3);
        const uint32_t up = __ldcg(&cinfo[nodeId]).x;
        __threadfence();
        const int visitations = atomicAdd(reinterpret_cast<int*>(&cinfo[nodeId]) + 1, 1);
        if (visitations < 1) return;
        __threadfence();
        const bool iAmLeft = ch.x == me;
        const uint32_t sib = iAmLeft ? ch.y : ch.x;
        const float2* S = reinterpret_cast<const float2*>(nodes + 10ull * sib);
        const float2 sx = __ldcg(S), sy = __ldcg(S + 1), sz = __ldcg(S + 2);
        const float2 lx = iAmLeft ? mx : sx, ly = iAmLeft ? my : sy, lz = iAmLeft ? mz : sz;
        const float2 rx = iAmLeft ? sx : mx, ry = iAmLeft ? sy : my, rz = iAmLeft ? sz : mz;
        // combineAABB(left, right) :27-36
        mx = make_float2(gmin(lx.x, rx.x), gmax(lx.y, rx.y));
        my = make_float2(gmin(ly.x, ry.x), gmax(ly.y, ry.y));
        mz = make_float2(gmin(lz.x, rz.x), gmax(lz.y, rz.y));
        __stcg(reinterpret_cast<float2*>(nd), mx);
        __stcg(reinterpret_cast<float2*>(nd) + 1, my);
        __stcg(reinterpret_cast<float2*>(nd) + 2, mz);
        if (etaNode) { me_eta = fmaxf(me_eta, __ldcg(&etaNode[sib])); __stcg(&etaNode[nodeId], me_eta); }   // largest hit-point slack of the subtree
        if (pairs) {    // the thread that unions a node holds both child boxes: emit the node's 64-byte traversal record here
            float4* out = pairs + 4ull * nodeId;
            out[0] = make_float4(lx.x, ly.x, lz.x, __uint_as_float(ch.x));
            out[1] = make_float4(lx.y, ly.y, lz.y, __uint_as_float(ch.y));
            out[2] = make_float4(rx.x, ry.x, rz.x, 0.f);
            out[3] = make_float4(rx.y, ry.y, rz.y, 0.f);
            if (nodeId == 0) {
                rootBox[0] = make_float4(mx.x, my.x, mz.x, 0.f);
                rootBox[1] = make_float4(mx.y, my.y, mz.y, 0.f);
            }
        }
        if (nodeId == 0) return;
        me = nodeId; nodeId = up;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Traversal-record derivation (rtb_bind_trace_buffers): reference-layout arrays -> 16-byte aligned records
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_pairs_kernel(const uint32_t* __restrict__ nodes, uint32_t n, float4* pairs, float4* rootBox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) {
        const float* b = reinterpret_cast<const float*>(nodes);
        rootBox[0] = make_float4(b[0], b[2], b[4], 0.f);
        rootBox[1] = make_float4(b[1], b[3], b[5], 0.f);
    }
    if (n < 2 || i >= n - 1) return;
    const uint32_t* nd = nodes + 10ull * i;
    const uint32_t li = nd[6], ri = nd[7];
    const float* L = reinterpret_cast<const float*>(nodes + 10ull * li);
    const float* R = reinterpret_cast<const float*>(nodes + 10ull * ri);
    float4* out = pairs + 4ull * i;
    out[0] = make_float4(L[0], L[2], L[4], __uint_as_float(li));
    out[1] = make_float4(L[1], L[3], L[5], __uint_as_float(ri));
    out[2] = make_float4(R[0], R[2], R[4], 0.f);
    out[3] = make_float4(R[1], R[3], R[5], 0.f);
}


// Compressed traversal records (common.cuh): one thread per internal node.  Quantisation is done with directed rounding so
// that, as REAL numbers, origin + qlo * scale <= lo and origin + qhi * scale >= hi for every plane of both children.
__global__ void __launch_bounds__(256) pack_cnodes_kernel(const uint32_t* __restrict__ nodes, uint32_t n, uint4* cnodes, float4* leafBox) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {                                               // exact box of leaf i (input order)
        const float* b = reinterpret_cast<const float*>(nodes + 10ull * (n - 1 + i));
        leafBox[2ull * i] = make_float4(b[0], b[2], b[4], 0.f);
        leafBox[2ull * i + 1] = make_float4(b[1], b[3], b[5], 0.f);
    }
    if (!cnodes || n < 2 || i >= n - 1) return;
    const uint32_t* nd = nodes + 10ull * i;
    const uint32_t li = nd[6], ri = nd[7];
    const float* L = reinterpret_cast<const float*>(nodes + 10ull * li);
    const float* R = reinterpret_cast<const float*>(nodes + 10ull * ri);
    const uint32_t leafOffset = n - 1;
    const bool leafL = li >= leafOffset, leafR = ri >= leafOffset;
    const uint32_t split = leafL ? li - leafOffset : li;       // ConstructHLBVH.comp:180-196: left = split or leafOffset + split
    float org[3], inv[3];
    uint32_t E[3], q[12];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float llo = L[2 * k], lhi = L[2 * k + 1], rlo = R[2 * k], rhi = R[2 * k + 1];
        const float lo = fminf(llo, rlo), hi = fmaxf(lhi, rhi);
        org[k] = lo;
        const float ext = __fsub_ru(hi, lo);                   // >= the real extent
        uint32_t e = 1;                                        // scale = 2^(e - 127); need 255 * scale >= ext
        if (ext > 0.f) {
            const uint32_t bits = __float_as_uint(__fdiv_ru(ext, 255.0f));
            e = (bits >> 23) + ((bits & 0x7FFFFFu) ? 1u : 0u);
            if (e < 1) e = 1;
            if (e > 253) e = 253;                              // (only reachable with non-finite boxes)
        }
        E[k] = e;
        inv[k] = __uint_as_float((254u - e) << 23);            // 1 / scale, exact
        const float dl[2] = { __fsub_rd(llo, lo), __fsub_rd(rlo, lo) };   // <= real (lo_child - origin)
        const float dh[2] = { __fsub_ru(lhi, lo), __fsub_ru(rhi, lo) };   // >= real (hi_child - origin)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            float ql = floorf(dl[c] * inv[k]); ql = fminf(fmaxf(ql, 0.f), 255.f);
            float qh = ceilf(dh[c] * inv[k]); qh = fminf(fmaxf(qh, 0.f), 255.f);
            q[c * 6 + k] = (uint32_t)ql;                       // child c: planes lo.x lo.y lo.z hi.x hi.y hi.z
            q[c * 6 + 3 + k] = (uint32_t)qh;
        }
    }
    uint4 a, bq;
    a.x = __float_as_uint(org[0]); a.y = __float_as_uint(org[1]); a.z = __float_as_uint(org[2]);
    a.w = E[0] | (E[1] << 8) | (E[2] << 16) | ((leafL ? 1u : 0u) << 24) | ((leafR ? 1u : 0u) << 25);
    bq.x = q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24);
    bq.y = q[4] | (q[5] << 8) | (q[6] << 16) | (q[7] << 24);
    bq.z = q[8] | (q[9] << 8) | (q[10] << 16) | (q[11] << 24);
    bq.w = split;
    cnodes[2ull * i] = a;
    cnodes[2ull * i + 1] = bq;
}

// Hit-point slack, for the nearest-first t-culled traversal: a point P the shader's primitive tests ACCEPT as a hit
// of a ray whose origin lies inside the (slightly grown) root box is within eta of the primitive, hence inside the primitive's
// leaf box grown by eta.  eps = 2^-24; R = largest coordinate magnitude of the grown root box.
//  * triangle (raytraceBVH.comp:118-149).  Off the shader's plane {n.x = n.v0}: the three dot products, the subtraction and
//    the division leave |n.(o + t d) - n.v0| <= 8 eps (|o| + |v0| + |t d|) <= 3e-6 R; rounding of P, of u = v1 - v0 and
//    v = v2 - v0: < 1e-6 R.  Together < 1e-5 R.  In the plane: the tests aa >= 0, bb >= 0, aa + bb <= 1 are evaluated with
//    rounding (|error| <= 7 eps |w| |pp| |v|) and with the rounded w, whose relative error is <= 12 eps / sin(theta) + 5 eps
//    because cross(u, v) cancels (theta = angle between u and v, 1 / sin(theta) = |u| |v| |w|); the stored normal tilts by the
//    same 4 eps / sin(theta).  As a displacement: <= 24 eps (1 + 1/sin(theta)) |w| |u| |v| (|u| + |v|); the kernel uses 64 eps.
//    Needle angles below 1e-4 rad, zero-area triangles (NaN normal) and microscopic ones (|w| > 1e18) get eta = inf.
//  * sphere (raytraceBVH.comp:152-181).  The returned root t satisfies | |P(t) - c|^2 - r^2 | <= 31 eps (|oc| + r)^2 however
//    small the discriminant (the perturbed quadratic it solves exactly differs by that much), so | |P - c| - r | <=
//    31 eps (|oc| + r)^2 / r; |oc| <= distance from the centre to the farthest corner of the grown root box.  The kernel
//    uses 62 eps = 4e-6 (+ the same 1e-5 R for the rounding of P).
// The per-node value is the maximum over the node's subtree (climbed like the refit: the second arrival at a node continues).
//
// ORIGIN REGION.  Everything above holds for rays whose origin lies in a region Omega with R = the largest coordinate magnitude of
// Omega (the bounds use |o| <= sqrt(3) R, |v0| <= sqrt(3) R, |t d| = |P - o| <= 2 sqrt(3) R).  Omega is the root box grown by 0.1 %
// of its largest extent -- every secondary ray starts on a primitive, i.e. inside it -- extended to contain the CAMERA position
// when that at most quadruples R (a camera far away from a small scene would inflate every record box; its rays then simply stay
// un-culled, as rays from outside Omega always do).  The kernel stores Omega in originRegion[0..1]; the trace kernels cull a ray
// by t only if its origin lies inside it.
__device__ __forceinline__ void origin_region_of(float4 lo, float4 hi, const f3 cam, const bool mayExtend, float4& rlo, float4& rhi) {
    const float grow = 1.0e-3f * fmaxf(fmaxf(hi.x - lo.x, hi.y - lo.y), hi.z - lo.z);
    lo.x -= grow; lo.y -= grow; lo.z -= grow; hi.x += grow; hi.y += grow; hi.z += grow;
    const float R = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
    const float rc = fmaxf(fmaxf(fabsf(cam.x), fabsf(cam.y)), fabsf(cam.z));
    if (mayExtend && rc <= 4.0f * R) {      // (false for a NaN camera)
        lo.x = fminf(lo.x, cam.x); lo.y = fminf(lo.y, cam.y); lo.z = fminf(lo.z, cam.z);
        hi.x = fmaxf(hi.x, cam.x); hi.y = fmaxf(hi.y, cam.y); hi.z = fmaxf(hi.z, cam.z);
    }
    rlo = lo; rhi = hi;
}
__device__ __forceinline__ float eta_triangle(const f3 u, const f3 v, const f3 w, const float R) {
    const float lu = sqrtf(dot(u, u)), lv = sqrtf(dot(v, v)), lw = sqrtf(dot(w, w));
    const float invSin = lu * lv * lw;
    float eta = 64.0f * 5.9604645e-8f * (1.0f + invSin) * lw * lu * lv * (lu + lv) + 1.0e-5f * R;
    if (!(invSin < 1.0e4f) || !(lw < 1.0e18f)) eta = __int_as_float(0x7f800000);
    if (!(eta < 3.0e38f)) eta = __int_as_float(0x7f800000);               // NaN / overflow
    return eta;
}
__device__ __forceinline__ float eta_sphere(const float4 sp, const float4 lo, const float4 hi, const float R) {
    const float dx = fmaxf(fabsf(sp.x - lo.x), fabsf(hi.x - sp.x)), dy = fmaxf(fabsf(sp.y - lo.y), fabsf(hi.y - sp.y)),
                dz = fmaxf(fabsf(sp.z - lo.z), fabsf(hi.z - sp.z));
    const float r = fabsf(sp.w), D = sqrtf(dx * dx + dy * dy + dz * dz) + r;
    float eta = 4.0e-6f * D * D / r + 1.0e-5f * R;
    if (!(eta < 3.0e38f)) eta = __int_as_float(0x7f800000);               // NaN / overflow
    return eta;
}
__global__ void __launch_bounds__(256) eta_leaf_kernel(const float4* __restrict__ ptris, uint32_t T, const float4* __restrict__ psphs,
                                                       uint32_t S, const float4* __restrict__ rootBox, const f3 cam, float* etaLeaf,
                                                       float4* originRegion) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= T + S) return;
    float4 lo, hi;
    origin_region_of(rootBox[0], rootBox[1], cam, region_may_extend(T, S), lo, hi);
    if (g == 0) { originRegion[0] = lo; originRegion[1] = hi; }
    const float R = fmaxf(fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fabsf(lo.y), fabsf(hi.y))), fmaxf(fabsf(lo.z), fabsf(hi.z)));
    float eta;
    if (g < T) {
        const float4 r1 = ptris[4ull * g + 1], r2 = ptris[4ull * g + 2], r3 = ptris[4ull * g + 3];
        eta = eta_triangle(F3(r1.w, r2.x, r2.y), F3(r2.z, r2.w, r3.x), F3(r3.y, r3.z, r3.w), R);
    } else {
        eta = eta_sphere(psphs[g - T], lo, hi, R);
    }
    etaLeaf[g] = eta;
}
__global__ void __launch_bounds__(256) parent_kernel(const uint32_t* __restrict__ nodes, uint32_t n, uint32_t* parent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 2 || i >= n - 1) return;
    parent[nodes[10ull * i + 6]] = i;
    parent[nodes[10ull * i + 7]] = i;
}
__global__ void __launch_bounds__(256) eta_climb_kernel(const uint32_t* __restrict__ nodes, const uint32_t* __restrict__ parent, uint32_t n,
                                                        float* etaNode, unsigned int* arrivals) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 2 || g >= n) return;
    uint32_t node = parent[n - 1 + g];
    while (true) {
        __threadfence();
        if (atomicAdd(&arrivals[node], 1u) == 0u) return;                  // the sibling subtree is not finished yet
        __threadfence();
        const float a = __ldcg(&etaNode[nodes[10ull * node + 6]]), b = __ldcg(&etaNode[nodes[10ull * node + 7]]);
        __stcg(&etaNode[node], fmaxf(a, b));
        if (node == 0) return;
        node = parent[node];
    }
}

// Wide (4-ary) traversal records, 64 B, one per binary internal node X: up to four DESCENDANT entries of X that together
// cover X's subtree, in the reference's visiting order (e.g. [right.right, right.left, left.right, left.left]), each with its
// exact box quantised OUTWARD to 8 bits per plane relative to the record's origin / power-of-two scales.  Layout (16 words):
//   0-2 origin.xyz | 3: Ex, Ey, Ez, meta (bits 0-3 leaf flags, bits 4-7 present flags) | 4-9: one word per plane
//   (lo.x lo.y lo.z hi.x hi.y hi.z), byte e = entry e, so the kernel picks a ray's near / far planes of all four entries
//   with one select per word | 10-13: entry node indices (a leaf is leafOffset + primitive id) | 14-15 unused
// The records are grown by the hit-point slack only if the ROOT's slack (= the scene's largest) is finite; otherwise they stay
// tight and flags[0] = 0 tells the trace kernels to walk them without t-culling (decided on the device: no read-back).
// flags[0] = 1: the records were grown by a finite hit-point slack, the walk may cull by t | flags[1] = record the walk starts at (0: the root's)
// These are the records over the REFERENCE's tree (every leaf in its place, entries in the reference's visiting order): what the reference-order
// walk needs, and what scenes outside the fused build get.  rtb_build_bvh derives the records of its own traversal hierarchy instead (traversal_tree.cu).
__global__ void __launch_bounds__(256) pack_wide_kernel(const uint32_t* __restrict__ nodes, uint32_t n, uint4* wide, const float* __restrict__ etaNode,
                                                        unsigned int* flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < 2 || i >= n - 1) return;
    bool slackOk = false;
    if (etaNode) { const float etaRoot = etaNode[0]; slackOk = etaRoot >= 0.0f && etaRoot < 3.0e38f; }
    if (i == 0 && flags) { flags[0] = slackOk ? 1u : 0u; flags[1] = 0u; }
    const uint32_t leafOffset = n - 1;
    // The entries are a cut through X's subtree, kept in visiting order (right before left, raytraceBVH.comp:241-244): start from
    // X's two children and keep replacing the internal entry with the largest surface area by its own two children while a slot
    // is free (the entry a ray is most likely to enter is the one worth resolving inside this record).
    uint32_t entry[4]; int cnt = 2;
    entry[0] = nodes[10ull * i + 7]; entry[1] = nodes[10ull * i + 6];
    while (cnt < 4) {
        int pick = -1; float best = -1.0f;
        for (int e = 0; e < cnt; e++) {
            if (entry[e] >= leafOffset) continue;
            const float* b = reinterpret_cast<const float*>(nodes + 10ull * entry[e]);
            const float dx = b[1] - b[0], dy = b[3] - b[2], dz = b[5] - b[4];
            const float area = dx * dy + dy * dz + dz * dx;
            if (pick < 0 || area > best) { pick = e; best = area; }
        }
        if (pick < 0) break;
        const uint32_t k = entry[pick];
        for (int e = cnt; e > pick + 1; e--) entry[e] = entry[e - 1];
        entry[pick] = nodes[10ull * k + 7]; entry[pick + 1] = nodes[10ull * k + 6];
        cnt++;
    }
    float lo[4][3], hi[4][3];
    uint32_t leafMask = 0;
    for (int e = 0; e < cnt; e++) {
        const float* b = reinterpret_cast<const float*>(nodes + 10ull * entry[e]);
        const float eta = slackOk ? etaNode[entry[e]] : 0.0f;                          // largest hit-point slack in the entry's subtree
        for (int k = 0; k < 3; k++) { lo[e][k] = __fsub_rd(b[2 * k], eta); hi[e][k] = __fadd_ru(b[2 * k + 1], eta); }
        if (entry[e] >= leafOffset) leafMask |= 1u << e;
    }
    store_wide_record(wide + 4ull * i, cnt, lo, hi, entry, leafMask);
}

#ifdef RTB_SMEM_TOP
// A/B variant "top tree levels staged in shared memory": the first RTB_SMEM_TOP records of the 4-ary hierarchy in breadth-first
// order, copied into a table whose internal entry ids are replaced by TOP_FLAG | table index when the entry is in the table too.
// One thread (sequential breadth-first queue): measurement aid, not tuned.
__global__ void build_top_table_kernel(const uint4* __restrict__ wide, uint32_t n, uint4* top, uint32_t* topGlobal, const unsigned int* flags) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    uint32_t count = 1;
    topGlobal[0] = flags[1];                                     // the record the walk starts at (tt_top_kernel)
    for (uint32_t i = 0; i < RTB_SMEM_TOP; i++) {
        if (i >= count) { for (int j = 0; j < 4; j++) top[4 * i + j] = make_uint4(0, 0, 0, 0); topGlobal[i] = 0; continue; }
        uint4 r[4];
        for (int j = 0; j < 4; j++) r[j] = wide[4ull * topGlobal[i] + j];
        const uint32_t meta = r[0].w >> 24;
        uint32_t ids[4] = { r[2].z, r[2].w, r[3].x, r[3].y };
        for (int e = 0; e < 4; e++) {
            const bool present = (meta >> (4 + e)) & 1u, leaf = (meta >> e) & 1u;
            if (!present || leaf || count >= RTB_SMEM_TOP) continue;
            topGlobal[count] = ids[e];
            ids[e] = 0x80000000u | count;
            count++;
        }
        r[2].z = ids[0]; r[2].w = ids[1]; r[3].x = ids[2]; r[3].y = ids[3];
        for (int j = 0; j < 4; j++) top[4 * i + j] = r[j];
    }
}
void launch_build_top_table(cudaStream_t st, const void* wide, uint32_t n, void* top, void* topGlobal, const unsigned int* flags) {
    build_top_table_kernel<<<1, 32, 0, st>>>((const uint4*)wide, n, (uint4*)top, (uint32_t*)topGlobal, flags);
}
#endif

__global__ void __launch_bounds__(256) pack_prims_kernel(const float4* __restrict__ tris, uint32_t T, const float4* __restrict__ sphs,
                                                         uint32_t S, const float4* __restrict__ mats, uint32_t M, float4* ptris,
                                                         float4* psphs, uint32_t* sphMat, float4* pmats) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) {
        const float4 a = tris[4ull * i], b = tris[4ull * i + 1], c = tris[4ull * i + 2];
        const uint4 idx = *reinterpret_cast<const uint4*>(tris + 4ull * i + 3);
        // ray-independent prologue of triangleHit (raytraceBVH.comp:119-125), same operation order
        const f3 v0 = xyz(a);
        const f3 u = xyz(b) - v0;
        const f3 v = xyz(c) - v0;
        const f3 nU = cross(u, v);
        const f3 n = normalize(nU);
        const f3 w = nU / dot(nU, nU);
        ptris[4ull * i] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(idx.x));
        ptris[4ull * i + 1] = make_float4(n.x, n.y, n.z, u.x);
        ptris[4ull * i + 2] = make_float4(u.y, u.z, v.x, v.y);
        ptris[4ull * i + 3] = make_float4(v.z, w.x, w.y, w.z);
    }
    if (i < S) {
        const float4 c = sphs[2ull * i];
        const float4 r = sphs[2ull * i + 1];
        psphs[i] = make_float4(c.x, c.y, c.z, r.x);
        sphMat[i] = __float_as_uint(r.y);
    }
    if (i < M) {
        const float4 a = mats[2ull * i];
        const float4 t = mats[2ull * i + 1];
        pmats[i] = make_float4(a.x, a.y, a.z, t.x);   // t.x carries the materialType bits
    }
}

// ---------------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

void launch_model_to_world(cudaStream_t st, const void* models, void* tris, uint32_t T, void* sphs, uint32_t S) {
    if (T + S == 0) return;
    model_to_world_kernel<<<blocks_for(T + S, 256), 256, 0, st>>>((const float4*)models, (float4*)tris, T, (float4*)sphs, S);
}
int launch_enclosing(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, uint32_t* red, void* enclosing,
                     int initInf, int smCount) {
    enclosing_init_kernel<<<1, 32, 0, st>>>(red, initInf);
    const uint64_t n = (uint64_t)T + S;
    unsigned grid = blocks_for(n, 256);
    const unsigned cap = (unsigned)smCount * 8u;       // persistent-style grid: a multiple of the SM count
    if (grid > cap) grid = cap;
    if (grid) enclosing_reduce_kernel<<<grid, 256, 0, st>>>((const float4*)tris, T, (const float4*)sphs, S, red);
    enclosing_finalize_kernel<<<1, 32, 0, st>>>(red, (float4*)enclosing);
    return grid ? 3 : 2;
}
void launch_morton(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const void* enclosing, void* morton,
                   uint32_t* keys, uint32_t* vals) {
    morton_kernel<<<blocks_for((uint64_t)T + S, 256), 256, 0, st>>>((const float4*)tris, T, (const float4*)sphs, S,
                                                                    (const float4*)enclosing, (uint32_t*)morton, keys, vals);
}
void launch_morton_unpack(cudaStream_t st, const void* morton, uint32_t n, uint32_t T, uint32_t* keys, uint32_t* vals) {
    morton_unpack_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t*)morton, n, T, keys, vals);
}
void launch_morton_repack(cudaStream_t st, const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t T, void* morton) {
    morton_repack_kernel<<<blocks_for(n, 256), 256, 0, st>>>(keys, vals, n, T, (uint32_t*)morton);
}
void launch_hlbvh(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const uint32_t* codes,
                  uint32_t codeStrideWords, void* nodes, void* cinfo) {
    Codes c{ codes, codeStrideWords, (int)(T + S) };
    LeafExtras none{};
    hlbvh_kernel<false><<<blocks_for((uint64_t)T + S, 256), 256, 0, st>>>((const float4*)tris, T, (const float4*)sphs, S, c,
                                                                          (uint32_t*)nodes, (uint2*)cinfo, none);
}
// K5 of the fused build: also writes the exact leaf boxes, the packed primitive records and the per-primitive hit-point slack
void launch_hlbvh_fused(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const uint32_t* codes, void* nodes,
                        void* cinfo, void* leafBox, void* ptris, void* psphs, void* sphMat, float* etaNode, const uint32_t* primBounds,
                        void* originRegion, const float* camPos, unsigned int* bigCount, uint32_t* bigList, uint32_t* smallBounds) {
    Codes c{ codes, 1, (int)(T + S) };
    cudaMemsetAsync(bigCount, 0, sizeof(unsigned int), st);
    cudaMemsetAsync(smallBounds, 0xFF, 3 * sizeof(uint32_t), st);           // ordered-int minima start at the top, maxima at 0
    cudaMemsetAsync(smallBounds + 3, 0, 3 * sizeof(uint32_t), st);
    LeafExtras ex{ (float4*)leafBox, (float4*)ptris, (float4*)psphs, (uint32_t*)sphMat, etaNode, primBounds, (float4*)originRegion,
                   F3(camPos[0], camPos[1], camPos[2]), bigCount, bigList, smallBounds };
    hlbvh_kernel<true><<<blocks_for((uint64_t)T + S, 256), 256, 0, st>>>((const float4*)tris, T, (const float4*)sphs, S, c,
                                                                         (uint32_t*)nodes, (uint2*)cinfo, ex);
}
void launch_refit(cudaStream_t st, void* nodes, void* cinfo, uint32_t n, void* pairs, void* rootBox, float* etaNode) {
    refit_kernel<<<blocks_for(n, 256), 256, 0, st>>>((uint32_t*)nodes, (uint2*)cinfo, n, (float4*)pairs, (float4*)rootBox, etaNode);
}
// K1 + K2 in one pass (fused build).  Returns #launches.
int launch_model_to_world_enclosing(cudaStream_t st, const void* models, void* tris, uint32_t T, void* sphs, uint32_t S, uint32_t* red,
                                    void* enclosing, int initInf) {
    enclosing_init_kernel<<<1, 32, 0, st>>>(red, initInf);
    model_to_world_enclosing_kernel<<<blocks_for((uint64_t)T + S, 256), 256, 0, st>>>((const float4*)models, (float4*)tris, T, (float4*)sphs, S, red);
    enclosing_finalize_kernel<<<1, 32, 0, st>>>(red, (float4*)enclosing);
    return 3;
}
void launch_pack_pairs(cudaStream_t st, const void* nodes, uint32_t n, void* pairs, void* rootBox) {
    const uint32_t work = n > 1 ? n - 1 : 1;
    pack_pairs_kernel<<<blocks_for(work, 256), 256, 0, st>>>((const uint32_t*)nodes, n, (float4*)pairs, (float4*)rootBox);
}
void launch_pack_cnodes(cudaStream_t st, const void* nodes, uint32_t n, void* cnodes, void* leafBox) {
    pack_cnodes_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t*)nodes, n, (uint4*)cnodes, (float4*)leafBox);
}
void launch_pack_wide(cudaStream_t st, const void* nodes, uint32_t n, void* wide, const float* etaNode, unsigned int* flags) {
    if (n < 2) return;
    pack_wide_kernel<<<blocks_for(n - 1, 256), 256, 0, st>>>((const uint32_t*)nodes, n, (uint4*)wide, etaNode, flags);
}
// etaNode[2n-1] <- per-node hit-point slack; parent[2n-1], arrivals[n-1] are scratch.  Returns #launches.
int launch_eta(cudaStream_t st, const void* nodes, uint32_t n, const void* ptris, uint32_t T, const void* psphs, uint32_t S,
               void* rootBox, const float* camPos, float* etaNode, uint32_t* parent, unsigned int* arrivals) {
    if (n < 2) return 0;
    cudaMemsetAsync(arrivals, 0, sizeof(unsigned int) * (n - 1), st);
    eta_leaf_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const float4*)ptris, T, (const float4*)psphs, S, (const float4*)rootBox,
                                                        F3(camPos[0], camPos[1], camPos[2]), etaNode + (n - 1), (float4*)rootBox + 2);
    parent_kernel<<<blocks_for(n - 1, 256), 256, 0, st>>>((const uint32_t*)nodes, n, parent);
    eta_climb_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const uint32_t*)nodes, parent, n, etaNode, arrivals);
    return 3;
}
void launch_pack_prims(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const void* mats, uint32_t M,
                       void* ptris, void* psphs, void* sphMat, void* pmats) {
    uint32_t m = T > S ? T : S;
    if (M > m) m = M;
    if (!m) return;
    pack_prims_kernel<<<blocks_for(m, 256), 256, 0, st>>>((const float4*)tris, T, (const float4*)sphs, S, (const float4*)mats, M,
                                                          (float4*)ptris, (float4*)psphs, (uint32_t*)sphMat, (float4*)pmats);
}

}  // namespace rtb
