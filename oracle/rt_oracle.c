/*
 * rt_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See rt_oracle.h for the contract.
 *
 * Pinned bit for bit against the reference's own compiled shaders (shaders/compiled/ *.spv run by oracle/spirv_interp.py ->
 * tests/golden/spirv_*.npz, tests/test_spirv_golden.py); driver-defined built-in arithmetic pinned by convention (U1..U14).
 *
 * Every function restates one reference shader and cites it.  The code follows the shader text literally
 * (same stack traversal, same 40-byte node records, one "dispatch" per sample) -- it is deliberately NOT
 * structured like the CUDA kernels so that the two implementations are independent restatements.
 *
 * Build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC (oracle/Makefile)
 */
#include "rt_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------ */
/* GLSL built-ins, pinned (SURVEY.md Appendix B U10/U11, Appendix E conventions)                     */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;

static inline float gmin(float x, float y) { return y < x ? y : x; }   /* GLSL min(x,y): y<x ? y : x */
static inline float gmax(float x, float y) { return x < y ? y : x; }   /* GLSL max(x,y): x<y ? y : x */
static inline v3 V3(float x, float y, float z) { v3 r = { x, y, z }; return r; }
static inline v3 vadd(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vsub(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 vmul(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 vscale(float s, v3 a) { return V3(s * a.x, s * a.y, s * a.z); }
static inline v3 vdivs(v3 a, float s) { return V3(a.x / s, a.y / s, a.z / s); }
static inline v3 vneg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline float vdot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 vcross(v3 a, v3 b) {
    return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 vnormalize(v3 a) { return vdivs(a, sqrtf(vdot(a, a))); }  /* v / sqrt(dot(v,v)) */
static inline v3 v3of(const float* p) { return V3(p[0], p[1], p[2]); }

/* uint(float): out-of-range is undefined in GLSL; pin U1/U5 = NVIDIA F2I.U32 saturation, NaN -> 0 */
uint32_t orc_float_to_u32_sat(float f) {
    if (!(f > 0.0f)) return 0u;                 /* negatives, -0, NaN */
    if (f >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)f;                         /* truncation toward zero */
}

/* pinned sin/cos for x in [0, 2*pi] (U11): quadrant reduction with two fused steps, Cephes-style minimax
 * polynomials evaluated with explicit fmaf in a fixed order.  The CUDA kernels implement the same recipe. */
void orc_pin_sincos(float x, float* s, float* c) {
    const float TWO_OVER_PI = 0x1.45f306p-1f;
    const float PIO2_HI = 0x1.921fb4p+0f, PIO2_LO = 0x1.4442d2p-24f;
    float q = rintf(x * TWO_OVER_PI);
    float r = fmaf(q, -PIO2_HI, x);
    r = fmaf(q, -PIO2_LO, r);
    float z = r * r;
    float ps = fmaf(-0x1.9943f2p-13f, z, 0x1.11073cp-7f);
    ps = fmaf(ps, z, -0x1.555546p-3f);
    float sr = fmaf(ps * z, r, r);
    float pc = fmaf(0x1.99eb9cp-16f, z, -0x1.6c0c34p-10f);
    pc = fmaf(pc, z, 0x1.55554ap-5f);
    float cr = fmaf(pc, z * z, fmaf(-0.5f, z, 1.0f));
    int k = ((int)q) & 3;
    switch (k) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* RNG: shaders/include/random.glsl                                                                  */
/* ------------------------------------------------------------------------------------------------ */
uint32_t orc_pcg_step(uint32_t state) { return state * 747796405u + 1u; }          /* random.glsl:13-15 */
uint32_t orc_pcg_word(uint32_t st) {                                               /* random.glsl:20-21 */
    uint32_t word = ((st >> ((st >> 28) + 4u)) ^ st) * 277803737u;
    return (word >> 22) ^ word;
}
float orc_pcg_float(uint32_t* state) {                                             /* random.glsl:18-23 */
    *state = orc_pcg_step(*state);
    uint32_t word = orc_pcg_word(*state);
    /* float(word) / 4294967295.0f : the literal rounds to 2^32 in binary32, so this is an exact scaling */
    return (float)word / 4294967296.0f;
}
uint32_t orc_seed_base(uint32_t x, uint32_t y, uint32_t randomState) {             /* random.glsl:10 */
    return (600u * x + y) * (randomState + 1u);
}
uint32_t orc_alpha_to_u32(float alpha) {                                           /* raytraceBVH.comp:350 */
    /* 4294967294.0f is 2^32 in binary32; alpha == 1.0 (the clear value) saturates: pin U1 */
    return orc_float_to_u32_sat(alpha * 4294967296.0f);
}

static const float ORC_PI = 3.1415926535897932385f;                                /* random.glsl:8 */

typedef struct { uint32_t rng; } RngCtx;
static inline float rnd(RngCtx* r) { return orc_pcg_float(&r->rng); }              /* random.glsl:25-27 */
static inline float rnd_range(RngCtx* r, float lo, float hi) { return lo + (hi - lo) * rnd(r); } /* :29-31 */
static v3 random_in_unit_sphere(RngCtx* r) {                                       /* random.glsl:33-38 */
    float rho = rnd(r);
    float theta = rnd_range(r, 0.0f, 2.0f * ORC_PI);
    float phi = rnd_range(r, 0.0f, ORC_PI);
    float sp, cp, st, ct;
    orc_pin_sincos(phi, &sp, &cp);
    orc_pin_sincos(theta, &st, &ct);
    return V3(rho * sp * ct, rho * sp * st, rho * cp);
}
static v3 random_unit_vector(RngCtx* r) { return vnormalize(random_in_unit_sphere(r)); } /* :60-62 */

/* ------------------------------------------------------------------------------------------------ */
/* K1  ModelSpaceToWorldSpace.comp:32-48                                                             */
/* ------------------------------------------------------------------------------------------------ */
static void mat_mul_point(const float* m, const float* p, float* out) {
    /* mat4 * vec4(p.xyz, 1.0) = ((c0*x + c1*y) + c2*z) + c3*w, per component */
    float x = p[0], y = p[1], z = p[2], w = 1.0f;
    float r[4];
    for (int i = 0; i < 4; i++) r[i] = ((m[0 + i] * x + m[4 + i] * y) + m[8 + i] * z) + m[12 + i] * w;
    for (int i = 0; i < 4; i++) out[i] = r[i];
}
void orc_model_to_world(const OrcModel* models, OrcTriangle* tris, uint32_t T, OrcSphere* sphs, uint32_t S) {
    for (uint32_t i = 0; i < T; i++) {
        const float* m = models[tris[i].modelIndex].m;
        mat_mul_point(m, tris[i].v0, tris[i].v0);
        mat_mul_point(m, tris[i].v1, tris[i].v1);
        mat_mul_point(m, tris[i].v2, tris[i].v2);
    }
    for (uint32_t i = 0; i < S; i++) mat_mul_point(models[sphs[i].modelIndex].m, sphs[i].center, sphs[i].center);
}

/* ------------------------------------------------------------------------------------------------ */
/* K2  GetEnclosingAABB.comp:64-110                                                                  */
/* ------------------------------------------------------------------------------------------------ */
static v3 triangle_center(const OrcTriangle* t) {       /* GetEnclosingAABB.comp:40-42: ((v0+v1+v2)/3).xyz */
    return V3(((t->v0[0] + t->v1[0]) + t->v2[0]) / 3.0f, ((t->v0[1] + t->v1[1]) + t->v2[1]) / 3.0f,
              ((t->v0[2] + t->v1[2]) + t->v2[2]) / 3.0f);
}
static v3 prim_center(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t i) {
    return i < T ? triangle_center(&tris[i]) : v3of(sphs[i - T].center);
}
/* total-order min / max: like fminf/fmaxf but -0 < +0, so the reduction is order independent */
static float tmin(float a, float b) { if (a == b) return signbit(a) ? a : b; return a < b ? a : b; }
static float tmax(float a, float b) { if (a == b) return signbit(a) ? b : a; return a > b ? a : b; }
static const float ORC_DELTA = 0.001f;
static void pad_axis(float* lo, float* hi) {            /* GetEnclosingAABB.comp:49-62, ConstructHLBVH.comp:41-54 */
    const float PADDING = ORC_DELTA / 2;
    if (*hi - *lo < ORC_DELTA) { *lo -= PADDING; *hi += PADDING; }
}
void orc_enclosing_aabb(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                        const OrcOptions* opt, OrcEnclosing* out) {
    /* Pin U4: localMin/localMax are never initialised in the shader (:69-70); they are read as 0.0, so
     * eMin = min(0, min centroid), eMax = max(0, max centroid).  The shader's own presets +-1e9 (:73-74)
     * never survive because 0 is always folded in.  With enclosingInitInf the locals start at +-inf and the
     * +-1e9 presets apply.  The w lanes end as min(1e9, 0) = 0 and max(-1e9, 0) = 0 (:100-101).
     * Zero signs: the reduction uses a total order (-0 < +0) so the result does not depend on lane order. */
    int inf = opt && opt->enclosingInitInf;
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) { lo[k] = inf ? INFINITY : 0.0f; hi[k] = inf ? -INFINITY : 0.0f; }
    for (uint32_t i = 0; i < T + S; i++) {
        v3 c = prim_center(tris, T, sphs, i);
        lo[0] = tmin(lo[0], c.x); lo[1] = tmin(lo[1], c.y); lo[2] = tmin(lo[2], c.z);
        hi[0] = tmax(hi[0], c.x); hi[1] = tmax(hi[1], c.y); hi[2] = tmax(hi[2], c.z);
    }
    for (int k = 0; k < 3; k++) {
        out->eMin[k] = tmin(1000000000.0f, lo[k]);
        out->eMax[k] = tmax(-1000000000.0f, hi[k]);
        pad_axis(&out->eMin[k], &out->eMax[k]);
    }
    out->eMin[3] = 0.0f; out->eMax[3] = 0.0f;
}

/* ------------------------------------------------------------------------------------------------ */
/* K3  GenerateMortonCodesOfPrimitives.comp:41-94                                                    */
/* ------------------------------------------------------------------------------------------------ */
uint32_t orc_separate_bits(uint32_t val) {              /* :41-50 */
    if (val == 1024u) val--;
    val = (val | (val << 16)) & 50331903u;
    val = (val | (val << 8)) & 50393103u;
    val = (val | (val << 4)) & 51130563u;
    val = (val | (val << 2)) & 153391689u;
    return val;
}
uint32_t orc_morton3(uint32_t qx, uint32_t qy, uint32_t qz) {   /* :52-58 */
    return orc_separate_bits(qz) << 2 | orc_separate_bits(qy) << 1 | orc_separate_bits(qx);
}
void orc_morton_codes(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                      const OrcEnclosing* enc, OrcMorton* out) {
    for (uint32_t i = 0; i < T + S; i++) {
        v3 c = prim_center(tris, T, sphs, i);
        /* quantizeForMorton :60-65 -- uses coord/span, NOT (coord-eMin)/span (pin U5: follow the code) */
        float sx = enc->eMax[0] - enc->eMin[0], sy = enc->eMax[1] - enc->eMin[1], sz = enc->eMax[2] - enc->eMin[2];
        uint32_t qx = orc_float_to_u32_sat((c.x / sx) * 1024.0f);
        uint32_t qy = orc_float_to_u32_sat((c.y / sy) * 1024.0f);
        uint32_t qz = orc_float_to_u32_sat((c.z / sz) * 1024.0f);
        out[i].code = orc_morton3(qx, qy, qz);
        out[i].primitiveIndex = i < T ? i : i - T;
        out[i].primitiveType = i < T ? ORC_TRIANGLE : ORC_SPHERE;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K4  RadixSortSimple.comp:52-158 -- 4 LSD passes of 8 bits, stable, ping-pong m1<->m2, result in m1 */
/* ------------------------------------------------------------------------------------------------ */
void orc_radix_sort(OrcMorton* m1, OrcMorton* m2, uint32_t n) {
    for (uint32_t iteration = 0; iteration < 4; iteration++) {
        uint32_t shift = 8u * iteration;
        const OrcMorton* in = (iteration % 2 == 0) ? m1 : m2;
        OrcMorton* outp = (iteration % 2 == 0) ? m2 : m1;
        uint32_t histogram[256];
        memset(histogram, 0, sizeof(histogram));
        for (uint32_t e = 0; e < n; e++) histogram[(in[e].code >> shift) & 255u]++;         /* :70-73 */
        uint32_t offsets[256], sum = 0;
        for (int b = 0; b < 256; b++) { offsets[b] = sum; sum += histogram[b]; }             /* :78-97 */
        /* :106-156: blocks of 256 in order; inside a block an element's slot is the count of lower-indexed
         * elements of the same bin -> stable */
        for (uint32_t e = 0; e < n; e++) outp[offsets[(in[e].code >> shift) & 255u]++] = in[e];
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K5  ConstructHLBVH.comp                                                                           */
/* ------------------------------------------------------------------------------------------------ */
static int find_msb(uint32_t v) { return v == 0 ? -1 : 31 - __builtin_clz(v); }   /* GLSL findMSB */
int orc_delta(const OrcMorton* mp, int n, int i, int j) {        /* countLeadingZeroesFromDifference :58-70 */
    if (j < 0 || j > n - 1) return -1;
    uint32_t codeI = mp[i].code, codeJ = mp[j].code;
    if (codeI == codeJ) return 32 + 31 - find_msb((uint32_t)i ^ (uint32_t)j);
    return 31 - find_msb(codeI ^ codeJ);
}
static void determine_range(const OrcMorton* mp, int n, int id, int* lower, int* upper) {  /* :72-96 */
    const int deltaL = orc_delta(mp, n, id, id - 1);
    const int deltaR = orc_delta(mp, n, id, id + 1);
    const int dir = (deltaR >= deltaL) ? 1 : -1;
    const int deltaMin = deltaL < deltaR ? deltaL : deltaR;
    int lMax = 2;
    while (orc_delta(mp, n, id, id + lMax * dir) > deltaMin) lMax <<= 1;
    int l = 0;
    for (int t = lMax >> 1; t > 0; t >>= 1)
        if (orc_delta(mp, n, id, id + (l + t) * dir) > deltaMin) l += t;
    int endId = id + l * dir;
    *lower = id < endId ? id : endId;
    *upper = id > endId ? id : endId;
}
static int find_split(const OrcMorton* mp, int n, int first, int last) {                   /* :98-115 */
    int commonPrefix = orc_delta(mp, n, first, last);
    int split = first;
    int stride = last - first;
    do {
        stride = (stride + 1) >> 1;
        int newSplit = split + stride;
        if (newSplit < last) {
            int splitPrefix = orc_delta(mp, n, first, newSplit);
            if (splitPrefix > commonPrefix) split = newSplit;
        }
    } while (stride > 1);
    return split;
}
static void pad_aabb(OrcAABB* b) { pad_axis(&b->minX, &b->maxX); pad_axis(&b->minY, &b->maxY); pad_axis(&b->minZ, &b->maxZ); }
static OrcAABB sphere_aabb(const OrcSphere* s) {                                           /* :117-130 */
    OrcAABB box;
    float lx = s->center[0] - s->radius, ly = s->center[1] - s->radius, lz = s->center[2] - s->radius;
    float rx = s->center[0] + s->radius, ry = s->center[1] + s->radius, rz = s->center[2] + s->radius;
    box.minX = gmin(lx, rx); box.maxX = gmax(lx, rx);
    box.minY = gmin(ly, ry); box.maxY = gmax(ly, ry);
    box.minZ = gmin(lz, rz); box.maxZ = gmax(lz, rz);
    return box;
}
static OrcAABB triangle_aabb(const OrcTriangle* t) {                                       /* :132-143 */
    OrcAABB box;
    box.minX = gmin(t->v0[0], gmin(t->v1[0], t->v2[0])); box.maxX = gmax(t->v0[0], gmax(t->v1[0], t->v2[0]));
    box.minY = gmin(t->v0[1], gmin(t->v1[1], t->v2[1])); box.maxY = gmax(t->v0[1], gmax(t->v1[1], t->v2[1]));
    box.minZ = gmin(t->v0[2], gmin(t->v1[2], t->v2[2])); box.maxZ = gmax(t->v0[2], gmax(t->v1[2], t->v2[2]));
    return box;
}
void orc_construct_hlbvh(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                         const OrcMorton* sorted, OrcNode* nodes, OrcCInfo* cinfo) {       /* main :145-215 */
    const int primitiveCount = (int)(T + S);
    const int leafOffset = primitiveCount - 1;
    for (int g = 0; g < primitiveCount; g++) {          /* leaves, in ORIGINAL primitive order (D8) */
        OrcNode nd;
        if ((uint32_t)g < T) { nd.primitiveType = ORC_TRIANGLE; nd.primitiveIndex = (uint32_t)g; nd.aabb = triangle_aabb(&tris[g]); }
        else { nd.primitiveType = ORC_SPHERE; nd.primitiveIndex = (uint32_t)g - T; nd.aabb = sphere_aabb(&sphs[(uint32_t)g - T]); }
        pad_aabb(&nd.aabb);
        nd.leftIndex = 0; nd.rightIndex = 0;
        nodes[leafOffset + g] = nd;
    }
    for (int g = 0; g < primitiveCount - 1; g++) {      /* internal nodes over the SORTED codes */
        int first, last;
        determine_range(sorted, primitiveCount, g, &first, &last);
        int split = find_split(sorted, primitiveCount, first, last);
        int leftChild = (split == first) ? leafOffset + split : split;
        int rightChild = (split + 1 == last) ? leafOffset + split + 1 : split + 1;
        OrcNode nd;
        memset(&nd, 0, sizeof(nd));
        nd.leftIndex = (uint32_t)leftChild; nd.rightIndex = (uint32_t)rightChild;
        nodes[g] = nd;
        cinfo[leftChild].parent = (uint32_t)g; cinfo[leftChild].visitationCount = 0;
        cinfo[rightChild].parent = (uint32_t)g; cinfo[rightChild].visitationCount = 0;
    }
    cinfo[0].parent = 0; cinfo[0].visitationCount = 0;  /* :212-214 (written last in program order here; for
                                                          N == 1 it is the only record) */
}

/* ------------------------------------------------------------------------------------------------ */
/* K6  ConstructAABBsOfInternalNodes.comp:27-65                                                      */
/* ------------------------------------------------------------------------------------------------ */
static OrcAABB combine_aabb(const OrcAABB* a, const OrcAABB* b) {                          /* :27-36 */
    OrcAABB c;
    c.minX = gmin(a->minX, b->minX); c.maxX = gmax(a->maxX, b->maxX);
    c.minY = gmin(a->minY, b->minY); c.maxY = gmax(a->maxY, b->maxY);
    c.minZ = gmin(a->minZ, b->minZ); c.maxZ = gmax(a->maxZ, b->maxZ);
    return c;
}
void orc_refit_aabbs(OrcNode* nodes, OrcCInfo* cinfo, uint32_t N) {                        /* main :38-65 */
    const uint32_t leafOffset = N - 1;
    for (uint32_t g = 0; g < N; g++) {                  /* one "invocation" per leaf, run to completion */
        uint32_t nodeId = cinfo[leafOffset + g].parent;
        for (;;) {
            int visitations = cinfo[nodeId].visitationCount++;     /* atomicAdd */
            if (visitations < 1) break;
            OrcNode node = nodes[nodeId];
            node.aabb = combine_aabb(&nodes[node.leftIndex].aabb, &nodes[node.rightIndex].aabb);
            nodes[nodeId] = node;
            if (nodeId == 0) break;
            nodeId = cinfo[nodeId].parent;
        }
    }
}

void orc_build_bvh(const OrcModel* models, OrcTriangle* tris, uint32_t T, OrcSphere* sphs, uint32_t S,
                   const OrcOptions* opt, OrcEnclosing* enc, OrcMorton* m1, OrcMorton* m2,
                   OrcNode* nodes, OrcCInfo* cinfo) {   /* RaytracerBVH.cpp:778-991 dispatch order */
    orc_model_to_world(models, tris, T, sphs, S);
    orc_enclosing_aabb(tris, T, sphs, S, opt, enc);
    orc_morton_codes(tris, T, sphs, S, enc, m1);
    orc_radix_sort(m1, m2, T + S);
    orc_construct_hlbvh(tris, T, sphs, S, m1, nodes, cinfo);
    orc_refit_aabbs(nodes, cinfo, T + S);
}

/* ------------------------------------------------------------------------------------------------ */
/* K7  raytraceBVH.comp                                                                              */
/* ------------------------------------------------------------------------------------------------ */
typedef struct { v3 origin, direction; } Ray;
typedef struct { v3 p, normal; uint32_t materialIndex; float t; int backFaceInt; } HitRecord;

typedef struct {
    const OrcUBO* ubo;
    const OrcTriangle* tris; const OrcSphere* sphs; const OrcMaterial* mats; const OrcNode* nodes;
    v3 camPos, pixel00, pixelDeltaU, pixelDeltaV;
    int extMaterials;
    int linearScan;         /* 1: raytrace.comp (the non-BVH program): loop over all primitives */
    v3 background;          /* _BACKGROUND_COLOR: raytraceBVH.comp:52 = 0 ; raytrace.comp:43 = (0.1, 0.1, 0.3) */
    uint32_t leafOffset;
} TraceCtx;

typedef struct { uint64_t rays, nodeVisits, triTests, sphTests, matReads; int stackOverflow; uint32_t lastPrim; } TraceStats;

static v3 point_on_ray(const Ray* r, float t) { return vadd(r->origin, vscale(t, r->direction)); }   /* :84-86 */

static int triangle_hit(const TraceCtx* c, uint32_t idx, const Ray* r, float tMin, float tMax, HitRecord* rec) { /* :118-149 */
    const OrcTriangle* tri = &c->tris[idx];
    v3 v0 = v3of(tri->v0);
    v3 u = vsub(v3of(tri->v1), v0);
    v3 v = vsub(v3of(tri->v2), v0);
    v3 nU = vcross(u, v);
    v3 n = vnormalize(nU);
    float D = vdot(n, v0);
    v3 w = vdivs(nU, vdot(nU, nU));
    float denom = vdot(n, r->direction);
    if (fabsf(denom) < 0.0001f) return 0;
    float t = (D - vdot(n, r->origin)) / denom;
    if (t < tMin || t > tMax) return 0;
    v3 P = point_on_ray(r, t);
    v3 pp = vsub(P, v0);
    float a = vdot(w, vcross(pp, v));
    float b = vdot(w, vcross(u, pp));
    if (a < 0 || b < 0 || a + b > 1) return 0;
    rec->t = t; rec->p = P;
    rec->normal = n;
    rec->backFaceInt = vdot(r->direction, rec->normal) > 0 ? 1 : 0;
    rec->normal = vscale((float)(1 - 2 * rec->backFaceInt), rec->normal);
    rec->materialIndex = tri->materialIndex;
    return 1;
}
static int sphere_hit(const TraceCtx* c, uint32_t idx, const Ray* r, float tMin, float tMax, HitRecord* rec) { /* :152-181 */
    const OrcSphere* s = &c->sphs[idx];
    v3 ctr = v3of(s->center);
    v3 oc = vsub(r->origin, ctr);
    float a = vdot(r->direction, r->direction);
    float halfB = vdot(oc, r->direction);
    float cc = vdot(oc, oc) - (s->radius * s->radius);
    float underRadical = (halfB * halfB) - (a * cc);
    if (underRadical < 0) return 0;
    float radical = sqrtf(underRadical);
    float root = (-halfB - radical) / a;
    if (root < tMin || root > tMax) {
        root = (-halfB + radical) / a;
        if (root < tMin || root > tMax) return 0;
    }
    rec->t = root;
    rec->p = point_on_ray(r, rec->t);
    /* rec.u / rec.v (:172-173) are dead values -- never read */
    rec->normal = vdivs(vsub(rec->p, ctr), s->radius);
    rec->backFaceInt = vdot(r->direction, rec->normal) > 0 ? 1 : 0;
    rec->normal = vscale((float)(1 - 2 * rec->backFaceInt), rec->normal);
    rec->materialIndex = s->materialIndex;
    return 1;
}
static int aabb_hit_check(const Ray* r, v3 boxMin, v3 boxMax) {                            /* :184-193 */
    v3 tMin = V3((boxMin.x - r->origin.x) / r->direction.x, (boxMin.y - r->origin.y) / r->direction.y, (boxMin.z - r->origin.z) / r->direction.z);
    v3 tMax = V3((boxMax.x - r->origin.x) / r->direction.x, (boxMax.y - r->origin.y) / r->direction.y, (boxMax.z - r->origin.z) / r->direction.z);
    v3 t1 = V3(gmin(tMin.x, tMax.x), gmin(tMin.y, tMax.y), gmin(tMin.z, tMax.z));
    v3 t2 = V3(gmax(tMin.x, tMax.x), gmax(tMin.y, tMax.y), gmax(tMin.z, tMax.z));
    float tNear = gmax(gmax(t1.x, t1.y), t1.z);
    float tFar = gmin(gmin(t2.x, t2.y), t2.z);
    return tNear < tFar;
}
#define ORC_MAX_STACK_DEPTH 128
static int hit_bvh(const TraceCtx* c, const Ray* r, float tMin, float tMax, HitRecord* rec, TraceStats* st) { /* :195-265 */
    int hit = 0;
    float closestSoFar = tMax;
    uint32_t stack[ORC_MAX_STACK_DEPTH];
    uint32_t toVisitOffset = 0;
    uint32_t currentNodeIndex = 0;
    st->rays++;
    for (;;) {
        OrcNode node = c->nodes[currentNodeIndex];
        st->nodeVisits++;
        if (aabb_hit_check(r, V3(node.aabb.minX, node.aabb.minY, node.aabb.minZ), V3(node.aabb.maxX, node.aabb.maxY, node.aabb.maxZ))) {
            if (node.leftIndex == 0 && node.rightIndex == 0) {       /* leaf */
                if (node.primitiveType == ORC_SPHERE) {
                    st->sphTests++;
                    if (sphere_hit(c, node.primitiveIndex, r, tMin, closestSoFar, rec)) { hit = 1; closestSoFar = rec->t; st->lastPrim = currentNodeIndex - c->leafOffset; }
                } else if (node.primitiveType == ORC_TRIANGLE) {
                    st->triTests++;
                    if (triangle_hit(c, node.primitiveIndex, r, tMin, closestSoFar, rec)) { hit = 1; closestSoFar = rec->t; st->lastPrim = currentNodeIndex - c->leafOffset; }
                }
                if (toVisitOffset == 0) break;
                currentNodeIndex = stack[--toVisitOffset];
            } else {                                                   /* internal: push left, go right */
                if (toVisitOffset >= ORC_MAX_STACK_DEPTH) { st->stackOverflow = 1; break; }
                stack[toVisitOffset++] = node.leftIndex;
                currentNodeIndex = node.rightIndex;
            }
        } else {
            if (toVisitOffset == 0) break;
            currentNodeIndex = stack[--toVisitOffset];
        }
    }
    return hit;
}
/* sceneHit of the non-BVH program, raytrace.comp:167-190: all triangles in index order, then all spheres */
static int scene_hit_linear(const TraceCtx* c, const Ray* r, HitRecord* rec, TraceStats* st) {
    const float tMin = 0.001f, tMax = 10000000.0f;
    HitRecord temp;
    memset(&temp, 0, sizeof(temp));
    int hitAny = 0;
    float closestSoFar = tMax;
    st->rays++;
    for (uint32_t i = 0; i < c->ubo->numTriangles; i++) {
        st->triTests++;
        if (triangle_hit(c, i, r, tMin, closestSoFar, &temp)) { hitAny = 1; closestSoFar = temp.t; *rec = temp; st->lastPrim = i; }
    }
    for (uint32_t i = 0; i < c->ubo->numSpheres; i++) {
        st->sphTests++;
        if (sphere_hit(c, i, r, tMin, closestSoFar, &temp)) { hitAny = 1; closestSoFar = temp.t; *rec = temp; st->lastPrim = c->ubo->numTriangles + i; }
    }
    return hitAny;
}
static int scene_hit(const TraceCtx* c, const Ray* r, HitRecord* rec, TraceStats* st) {    /* raytraceBVH.comp:267-274 */
    if (c->linearScan) return scene_hit_linear(c, r, rec, st);
    return hit_bvh(c, r, 0.001f, 10000000.0f, rec, st);
}

/* Extension N1 (NOT reference behaviour): mirror metal, Schlick dielectric with IOR 1.5. */
static int scatter_extension(const TraceCtx* c, const Ray* rIn, const HitRecord* rec, const OrcMaterial* m,
                             RngCtx* rng, v3* attenuation, Ray* scattered) {
    v3 d = rIn->direction;
    if (m->materialType == ORC_METALLIC) {
        float dn = vdot(d, rec->normal);
        v3 refl = vsub(d, vscale(2.0f * dn, rec->normal));
        *attenuation = v3of(m->albedo);
        scattered->origin = rec->p; scattered->direction = vnormalize(refl);
        return vdot(scattered->direction, rec->normal) > 0;
    }
    if (m->materialType == ORC_DIELECTRIC) {
        const float ior = 1.5f;
        float ri = rec->backFaceInt ? ior : 1.0f / ior;
        float cosT = gmin(vdot(vneg(d), rec->normal), 1.0f);
        float sinT = sqrtf(1.0f - cosT * cosT);
        float r0 = (1.0f - ri) / (1.0f + ri); r0 = r0 * r0;
        float om = 1.0f - cosT;
        float refl = r0 + (1.0f - r0) * ((om * om) * (om * om) * om);
        v3 dir;
        float u = rnd(rng);
        if (ri * sinT > 1.0f || refl > u) {
            dir = vsub(d, vscale(2.0f * vdot(d, rec->normal), rec->normal));
        } else {
            v3 perp = vscale(ri, vadd(d, vscale(cosT, rec->normal)));
            float k = 1.0f - vdot(perp, perp);
            v3 par = vscale(-sqrtf(fabsf(k)), rec->normal);
            dir = vadd(perp, par);
        }
        *attenuation = v3of(m->albedo);
        scattered->origin = rec->p; scattered->direction = vnormalize(dir);
        return 1;
    }
    (void)c;
    return 0;
}

static v3 ray_color(const TraceCtx* c, const Ray* rIn, RngCtx* rng, TraceStats* st, uint32_t* firstLeaf, float* firstT) { /* :276-315 */
    HitRecord rec;
    memset(&rec, 0, sizeof(rec));
    v3 color = V3(0, 0, 0);
    v3 globalAttenuation = V3(1, 1, 1);
    Ray curr; curr.origin = rIn->origin; curr.direction = vnormalize(rIn->direction);
    for (uint32_t i = 0; i < c->ubo->maxRayTraceDepth; i++) {
        int hit = scene_hit(c, &curr, &rec, st);
        if (i == 0 && firstLeaf) { *firstLeaf = hit ? st->lastPrim : 0xFFFFFFFFu; if (firstT) *firstT = hit ? rec.t : 0.0f; }
        if (!hit) {
            color = vadd(color, vmul(c->background, globalAttenuation));        /* _BACKGROUND_COLOR * globalAttenuation (:284) */
            break;
        }
        const OrcMaterial* m = &c->mats[rec.materialIndex];
        st->matReads++;
        v3 emittedColor = (m->materialType == ORC_LIGHT) ? v3of(m->albedo) : V3(0, 0, 0); /* emitted :94-99 */
        color = vadd(color, vmul(emittedColor, globalAttenuation));
        v3 attenuation; int scattered;
        if (m->materialType == ORC_DIFFUSE) {                                   /* scatter :100-115 */
            attenuation = v3of(m->albedo);
            Ray sc; sc.origin = rec.p; sc.direction = vnormalize(vadd(rec.normal, random_unit_vector(rng)));
            curr = sc; scattered = 1;
        } else if (c->extMaterials && m->materialType != ORC_LIGHT) {
            Ray sc; scattered = scatter_extension(c, &curr, &rec, m, rng, &attenuation, &sc);
            if (scattered) curr = sc;
        } else {
            scattered = 0;              /* LIGHT, METALLIC, DIELECTRIC: absorbed (D4 / U8) */
        }
        if (!scattered) break;          /* the shader multiplies an uninitialised attenuation first; unobservable */
        globalAttenuation = vmul(globalAttenuation, attenuation);
    }
    return color;
}

static void setup_camera(TraceCtx* c, uint32_t W, uint32_t H) {                            /* :50-81 */
    const OrcUBO* ubo = c->ubo;
    const float FOCAL = 10.0f;
    float aspect = (float)W / (float)H;
    float theta = ubo->verticalFOV * 0.017453292519943295f;      /* radians() */
    float h = tanf(theta / 2);
    float viewportHeight = 2.0f * h * FOCAL;
    float viewportWidth = viewportHeight * aspect;
    v3 camPos = v3of(ubo->camPos);
    v3 camW = vnormalize(vsub(camPos, v3of(ubo->camLookAt)));
    v3 camU = vnormalize(vcross(v3of(ubo->camUpDir), camW));
    v3 camV = vcross(camW, camU);
    v3 viewportU = vscale(viewportWidth, camU);
    v3 viewportV = vscale(viewportHeight, vneg(camV));
    v3 dU = vdivs(viewportU, (float)W);
    v3 dV = vdivs(viewportV, (float)H);
    v3 upperLeft = vsub(vsub(vsub(camPos, vscale(FOCAL, camW)), vdivs(viewportU, 2)), vdivs(viewportV, 2));
    c->camPos = camPos;
    c->pixelDeltaU = dU; c->pixelDeltaV = dV;
    c->pixel00 = vadd(upperLeft, vscale(0.5f, vadd(dU, dV)));
}

void orc_clear_image(float* rgba, uint32_t W, uint32_t H) {
    for (uint64_t i = 0; i < (uint64_t)W * H; i++) { rgba[4 * i] = 0; rgba[4 * i + 1] = 0; rgba[4 * i + 2] = 0; rgba[4 * i + 3] = 1.0f; }
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int orc_raytrace(const OrcUBO* ubo, float* rgba, uint32_t W, uint32_t H, uint32_t y0, uint32_t y1,
                 const OrcTriangle* tris, const OrcSphere* sphs, const OrcMaterial* mats, const OrcNode* nodes,
                 uint32_t spp, const OrcOptions* opt,
                 uint32_t* hitPrim, float* hitT, uint32_t* rngOut, OrcCounters* counters) {
    TraceCtx c;
    memset(&c, 0, sizeof(c));
    c.ubo = ubo; c.tris = tris; c.sphs = sphs; c.mats = mats; c.nodes = nodes;
    c.extMaterials = opt ? opt->extMaterials : 0;
    c.linearScan = opt ? opt->linearScan : 0;
    c.background = c.linearScan ? V3(0.1f, 0.1f, 0.3f) : V3(0, 0, 0);
    setup_camera(&c, W, H);
    const uint32_t N = ubo->numTriangles + ubo->numSpheres;
    c.leafOffset = N - 1;
    uint64_t rays = 0, visits = 0, tt = 0, stt = 0, mr = 0;
    int overflow = 0;
    int threads = (opt && opt->threads > 0) ? opt->threads : orc_max_threads();
    (void)threads;
    /* one dispatch per sample with a full barrier between dispatches: RaytracerBVH.cpp:1025-1050 */
    for (uint32_t s = 0; s < spp; s++) {
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : rays, visits, tt, stt, mr) reduction(| : overflow)
        for (int64_t yy = (int64_t)y0; yy < (int64_t)y1; yy++) {
            uint32_t y = (uint32_t)yy;
            for (uint32_t x = 0; x < W; x++) {                                  /* main() :345-375 */
                TraceStats st; memset(&st, 0, sizeof(st));
                float* px = rgba + 4 * ((uint64_t)y * W + x);
                float cur[4] = { px[0], px[1], px[2], px[3] };
                RngCtx rng;
                rng.rng = orc_seed_base(x, y, ubo->randomState);
                rng.rng += orc_alpha_to_u32(cur[3]);
                /* stepRNG(rngState); -- result discarded, no state change (pin U2) */
                float nextRandom = rnd(&rng);
                Ray r;                                                           /* getRay :329-342 */
                r.origin = c.camPos;
                v3 pixelSample = vadd(vadd(c.pixel00, vscale((float)x, c.pixelDeltaU)), vscale((float)y, c.pixelDeltaV));
                r.direction = vnormalize(vsub(pixelSample, r.origin));
                uint32_t leaf = 0xFFFFFFFFu; float ft = 0.0f;
                v3 color = ray_color(&c, &r, &rng, &st, (s == 0) ? &leaf : NULL, &ft);
                px[0] = color.x + cur[0]; px[1] = color.y + cur[1]; px[2] = color.z + cur[2]; px[3] = nextRandom;
                if (s == 0 && hitPrim) hitPrim[(uint64_t)y * W + x] = leaf;
                if (s == 0 && hitT) hitT[(uint64_t)y * W + x] = ft;
                if (rngOut && s + 1 == spp) rngOut[(uint64_t)y * W + x] = rng.rng;
                rays += st.rays; visits += st.nodeVisits; tt += st.triTests; stt += st.sphTests; mr += st.matReads;
                overflow |= st.stackOverflow;
            }
        }
    }
    if (counters) {
        counters->rays += rays; counters->nodeVisits += visits; counters->triTests += tt;
        counters->sphTests += stt; counters->matReads += mr; counters->samples += (uint64_t)W * (y1 - y0) * spp;
    }
    return overflow ? -1 : 0;
}

/* raytrace.comp:167-190 (the non-BVH program): loop all triangles, then all spheres, closest so far */
void orc_primary_hits_bruteforce(const OrcUBO* ubo, uint32_t W, uint32_t H,
                                 const OrcTriangle* tris, const OrcSphere* sphs, uint32_t* hitPrim, float* hitT) {
    TraceCtx c;
    memset(&c, 0, sizeof(c));
    c.ubo = ubo; c.tris = tris; c.sphs = sphs;
    setup_camera(&c, W, H);
    const uint32_t T = ubo->numTriangles, S = ubo->numSpheres;
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t yy = 0; yy < (int64_t)H; yy++) {
        for (uint32_t x = 0; x < W; x++) {
            uint32_t y = (uint32_t)yy;
            Ray r; r.origin = c.camPos;
            v3 pixelSample = vadd(vadd(c.pixel00, vscale((float)x, c.pixelDeltaU)), vscale((float)y, c.pixelDeltaV));
            r.direction = vnormalize(vsub(pixelSample, r.origin));
            r.direction = vnormalize(r.direction);
            HitRecord rec; memset(&rec, 0, sizeof(rec));
            float closest = 10000000.0f; uint32_t best = 0xFFFFFFFFu;
            for (uint32_t i = 0; i < T; i++) if (triangle_hit(&c, i, &r, 0.001f, closest, &rec)) { closest = rec.t; best = i; }
            for (uint32_t i = 0; i < S; i++) if (sphere_hit(&c, i, &r, 0.001f, closest, &rec)) { closest = rec.t; best = T + i; }
            hitPrim[(uint64_t)y * W + x] = best;
            if (hitT) hitT[(uint64_t)y * W + x] = best == 0xFFFFFFFFu ? 0.0f : closest;
        }
    }
}

/* SingleTriangleFullScreen.frag:13-21 then B8G8R8A8_UNORM store (SwapChain.cpp:393): round(c*255) */
void orc_resolve_rgba8(const float* rgba, uint32_t W, uint32_t H, uint32_t raysPerPixel, uint8_t* out) {
    for (uint64_t i = 0; i < (uint64_t)W * H; i++) {
        for (int k = 0; k < 3; k++) {
            float v = sqrtf(rgba[4 * i + k] / (float)raysPerPixel);
            v = (v > 0.0f) ? v : 0.0f;                  /* clamp(x,0,1); NaN (undefined in GLSL) pinned to 0 */
            v = (v < 1.0f) ? v : 1.0f;
            out[4 * i + k] = (uint8_t)(v * 255.0f + 0.5f);
        }
        out[4 * i + 3] = 255;
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* K9  logistic.comp:23-35 -- the LogisticMap demo program (Config::Programs::LogisticMap): one step of     */
/* x' = x * r * (1 - x) per point, then plot the point into an rgba8 image (never cleared between steps).   */
/* ------------------------------------------------------------------------------------------------ */
static int f2i_trunc_sat(float f) {            /* GLSL int(float): truncation; out of range / NaN pinned like CUDA F2I */
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)(-2147483647 - 1);
    return (int)f;
}
void orc_logistic_step(float* points /* (x, r) pairs */, uint32_t count, uint8_t* rgba8, uint32_t W, uint32_t H, const float* pixelColor) {
    uint8_t c[4];
    for (int k = 0; k < 4; k++) {              /* rgba8 unorm store of ubo.pixelColor */
        float v = pixelColor[k];
        v = (v > 0.0f) ? v : 0.0f; v = (v < 1.0f) ? v : 1.0f;
        c[k] = (uint8_t)(v * 255.0f + 0.5f);
    }
    for (uint32_t i = 0; i < count; i++) {
        const float x = points[2 * i], r = points[2 * i + 1];
        const float nx = x * r * (1.0f - x);   /* :29, left to right */
        points[2 * i] = nx;
        const int xc = f2i_trunc_sat((r / 4.0f) * (float)W);            /* :31 */
        const int yc = f2i_trunc_sat(((1 - nx) / 1.0f) * (float)H);     /* :32 */
        if (xc >= 0 && yc >= 0 && (uint32_t)xc < W && (uint32_t)yc < H)  /* imageStore outside the image is discarded */
            memcpy(rgba8 + 4 * ((size_t)yc * W + (size_t)xc), c, 4);
    }
}
