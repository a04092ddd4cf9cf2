// build_shared.cuh -- device helpers shared by the reference-layout build (bvh_build.cu) and the traversal hierarchy (traversal_tree.cu)
#pragma once

#include "common.cuh"

namespace rtb {

// order-preserving integer image of a float (total order, -0 < +0): reductions with integer atomics are order-independent
__device__ __forceinline__ uint32_t f2ord(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// sorted Morton codes + the reference's tie-break on the index (ConstructHLBVH.comp:58-70)
struct Codes {
    const uint32_t* __restrict__ p; uint32_t stride; int n;
    __device__ __forceinline__ uint32_t operator[](int i) const { return __ldg(p + (size_t)i * stride); }
};
// countLeadingZeroesFromDifference :58-70 ; 31 - findMSB(x) == clz(x) for x != 0
__device__ __forceinline__ int delta(const Codes& c, int i, uint32_t codeI, int j) {
    if (j < 0 || j > c.n - 1) return -1;
    const uint32_t codeJ = c[j];
    if (codeI == codeJ) return 32 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clz(codeI ^ codeJ);
}
// Karras range + split of internal node `id` (determineRange :72-96, findSplit :98-115): children as node indices, a leaf is leafOffset + position
__device__ __forceinline__ void karras_children(const Codes& codes, const int id, const int leafOffset, int& leftChild, int& rightChild) {
    const uint32_t codeI = codes[id];
    const int deltaL = delta(codes, id, codeI, id - 1);
    const int deltaR = delta(codes, id, codeI, id + 1);
    const int dir = (deltaR >= deltaL) ? 1 : -1;
    const int deltaMin = min(deltaL, deltaR);
    int lMax = 2;
    while (delta(codes, id, codeI, id + lMax * dir) > deltaMin) lMax <<= 1;
    int l = 0;
    for (int t = lMax >> 1; t > 0; t >>= 1)
        if (delta(codes, id, codeI, id + (l + t) * dir) > deltaMin) l += t;
    const int endId = id + l * dir;
    const int first = min(id, endId), last = max(id, endId);
    const uint32_t codeF = codes[first];
    const int commonPrefix = delta(codes, first, codeF, last);
    int split = first, stride = last - first;
    do {
        stride = (stride + 1) >> 1;
        const int newSplit = split + stride;
        if (newSplit < last && delta(codes, first, codeF, newSplit) > commonPrefix) split = newSplit;
    } while (stride > 1);
    leftChild = (split == first) ? leafOffset + split : split;
    rightChild = (split + 1 == last) ? leafOffset + split + 1 : split + 1;
}

// BIG primitives (traversal_tree.cu): leaf box above 1/256 of the surface area of the box of all primitives
constexpr uint32_t MAX_BIG = 45;
__device__ __forceinline__ float hoist_threshold(const uint32_t* pb) {
    const float dx = ord2f(pb[3]) - ord2f(pb[0]), dy = ord2f(pb[4]) - ord2f(pb[1]), dz = ord2f(pb[5]) - ord2f(pb[2]);
    return (dx * dy + dy * dz + dz * dx) * (1.0f / 256.0f);
}
__device__ __forceinline__ bool box_is_big(const float mnx, const float mxx, const float mny, const float mxy, const float mnz, const float mxz, const float thr) {
    const float dx = mxx - mnx, dy = mxy - mny, dz = mxz - mnz;
    return dx * dy + dy * dz + dz * dx > thr;      // false for a NaN threshold or box
}

// quantise up to four entry boxes OUTWARD to 8 bits per plane relative to the record's origin / power-of-two scales and store the 64-byte
// 4-ary record (layout: common.cuh / DESIGN.md "data layout")
__device__ __forceinline__ void store_wide_record(uint4* out, const int cnt, const float (*lo)[3], const float (*hi)[3], const uint32_t* ids, const uint32_t leafMask) {
    float org[3], top[3];
    for (int k = 0; k < 3; k++) { org[k] = __int_as_float(0x7f800000); top[k] = __int_as_float(0xff800000); }
    for (int e = 0; e < cnt; e++)
        for (int k = 0; k < 3; k++) { org[k] = fminf(org[k], lo[e][k]); top[k] = fmaxf(top[k], hi[e][k]); }
    if (cnt == 0) { org[0] = org[1] = org[2] = 0.f; top[0] = top[1] = top[2] = 0.f; }
    uint32_t E[3], q[24];
    for (int j = 0; j < 24; j++) q[j] = 0;
    for (int k = 0; k < 3; k++) {
        const float ext = __fsub_ru(top[k], org[k]);
        uint32_t ex = 1;
        if (ext > 0.f) {
            const uint32_t bits = __float_as_uint(__fdiv_ru(ext, 255.0f));
            ex = (bits >> 23) + ((bits & 0x7FFFFFu) ? 1u : 0u);
            if (ex < 1) ex = 1;
            if (ex > 253) ex = 253;
        }
        E[k] = ex;
        const float inv = __uint_as_float((254u - ex) << 23);
        for (int e = 0; e < cnt; e++) {
            float ql = floorf(__fsub_rd(lo[e][k], org[k]) * inv); ql = fminf(fmaxf(ql, 0.f), 255.f);
            float qh = ceilf(__fsub_ru(hi[e][k], org[k]) * inv); qh = fminf(fmaxf(qh, 0.f), 255.f);
            q[4 * k + e] = (uint32_t)ql;
            q[4 * (3 + k) + e] = (uint32_t)qh;
        }
    }
    uint32_t meta = leafMask & 0xFu;
    for (int e = 0; e < cnt; e++) meta |= 1u << (4 + e);
    uint32_t w[16];
    w[0] = __float_as_uint(org[0]); w[1] = __float_as_uint(org[1]); w[2] = __float_as_uint(org[2]);
    w[3] = E[0] | (E[1] << 8) | (E[2] << 16) | (meta << 24);
    for (int j = 0; j < 6; j++) w[4 + j] = q[4 * j] | (q[4 * j + 1] << 8) | (q[4 * j + 2] << 16) | (q[4 * j + 3] << 24);
    for (int e = 0; e < 4; e++) w[10 + e] = e < cnt ? ids[e] : 0u;
    w[14] = 0; w[15] = 0;
    for (int j = 0; j < 4; j++) out[j] = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}

}  // namespace rtb
