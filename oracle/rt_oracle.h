/*
 * rt_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the compute shaders of silvercorked/RaytracerGPU_MastersProject
 * (paths below are relative to RaytracerGPU_MastersProject/ in the reference tree).
 *
 * HOW IT IS PINNED: the reference ships no tests, golden vectors or fixtures for this path and cannot run on a device
 * in this image (no Vulkan loader/ICD, no glslc) -- but it ships every shader COMPILED (shaders/compiled/ *.spv).
 * oracle/spirv_interp.py executes those binaries on the CPU; tests/golden/make_spirv_golden.py committed their outputs
 * (every buffer after every dispatch, nine scenes) as tests/golden/spirv_*.npz, and tests/test_spirv_golden.py checks
 * that every function below reproduces them BIT FOR BIT, stage by stage.  What the binaries do not define -- the
 * arithmetic of the driver's built-ins (normalize, dot association, sin / cos / tan, min / max of NaN, out-of-range
 * float->uint) -- follows the pins of SURVEY.md Appendix B (U1..U14), shared with the interpreter and repeated at the
 * place each is applied: that part is pinned by convention, not by the reference.  Additional checks: hand-derived
 * known-answer vectors (tests/test_oracle_kat.py) and structural properties (tests/).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library.  The product (raytracergpu_mastersproject_b200/) never includes, links or calls it.
 *
 * Arithmetic conventions (pin U11): every float op is an IEEE-754 binary32 round-to-nearest-even
 * + - * / sqrt with NO fused contraction (compile with -ffp-contract=off); the only fused ops are the
 * explicit fmaf() calls inside orc_pin_sincos() which define the pinned sin/cos.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- record layouts: shaders/include/definitions.glsl:6-77 == VulkanWrapper/SceneTypes.hpp:32-123 ---- */
typedef struct { float m[16]; } OrcModel;                                   /* mat4, column-major, 64 B */
typedef struct { float v0[4], v1[4], v2[4]; uint32_t materialIndex, modelIndex, _pad[2]; } OrcTriangle; /* 64 B */
typedef struct { float center[4]; float radius; uint32_t materialIndex, modelIndex, _pad; } OrcSphere;  /* 32 B */
typedef struct { float albedo[4]; uint32_t materialType, _pad[3]; } OrcMaterial;                        /* 32 B */
typedef struct { float minX, maxX, minY, maxY, minZ, maxZ; } OrcAABB;                                    /* 24 B */
typedef struct { OrcAABB aabb; uint32_t leftIndex, rightIndex, primitiveIndex, primitiveType; } OrcNode; /* 40 B */
typedef struct { uint32_t code, primitiveIndex, primitiveType; } OrcMorton;                              /* 12 B */
typedef struct { uint32_t parent; int32_t visitationCount; } OrcCInfo;                                   /*  8 B */
typedef struct { float eMin[4], eMax[4]; } OrcEnclosing;                                                 /* 32 B */

/* ParameterUBO, std140: raytraceBVH.comp:7-18 == RaytracerBVH.hpp:31-42 (80 bytes) */
typedef struct {
    float camPos[4], camLookAt[4], camUpDir[4];
    float verticalFOV;
    uint32_t numTriangles, numSpheres, numMaterials, numLights, maxRayTraceDepth, randomState;
    uint32_t _pad;
} OrcUBO;

enum { ORC_LIGHT = 0, ORC_DIFFUSE = 1, ORC_METALLIC = 2, ORC_DIELECTRIC = 3 };   /* definitions.glsl:79-82 */
enum { ORC_SPHERE = 0, ORC_TRIANGLE = 1 };                                       /* definitions.glsl:84-85 */

/* work counters for the roofline figure: rays = hitBVH calls, V = node visits,
 * Tt/St = triangle / sphere tests, H = hits that read a material. */
typedef struct { uint64_t rays, nodeVisits, triTests, sphTests, matReads, samples; } OrcCounters;

/* options for the pinned / extension behaviour */
typedef struct {
    int enclosingInitInf;   /* 0: pin U4 (uninitialised localMin/localMax read as 0.0); 1: +-inf */
    int extMaterials;       /* 0: reference behaviour (metal/dielectric absorb, D4); 1: extension N1 */
    int threads;            /* OpenMP threads for orc_raytrace (0 = all) */
    int linearScan;         /* 0: BVH program (raytraceBVH.comp); 1: non-BVH program (raytrace.comp, nodes unused) */
} OrcOptions;

/* ---- RNG (shaders/include/random.glsl:10-27) ---- */
uint32_t orc_pcg_step(uint32_t state);                       /* stepRNG */
uint32_t orc_pcg_word(uint32_t steppedState);                /* output permutation of an already stepped state */
float    orc_pcg_float(uint32_t* state);                     /* stepAndOutputRNGFloat */
uint32_t orc_seed_base(uint32_t x, uint32_t y, uint32_t randomState);   /* random.glsl:10 */
uint32_t orc_alpha_to_u32(float alpha);                      /* raytraceBVH.comp:350, pin U1 */
void     orc_pin_sincos(float x, float* s, float* c);        /* pinned sin/cos (U11) */
uint32_t orc_morton3(uint32_t qx, uint32_t qy, uint32_t qz); /* GenerateMortonCodesOfPrimitives.comp:41-58 */
uint32_t orc_separate_bits(uint32_t v);
uint32_t orc_float_to_u32_sat(float f);
int      orc_delta(const OrcMorton* sorted, int n, int i, int j); /* ConstructHLBVH.comp:58-70 */

/* ---- S1: BVH build, one function per dispatch (RaytracerBVH.cpp:778-991) ---- */
void orc_model_to_world(const OrcModel* models, OrcTriangle* tris, uint32_t T, OrcSphere* sphs, uint32_t S);
void orc_enclosing_aabb(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                        const OrcOptions* opt, OrcEnclosing* out);
void orc_morton_codes(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                      const OrcEnclosing* enc, OrcMorton* out);
void orc_radix_sort(OrcMorton* m1, OrcMorton* m2, uint32_t n);   /* result in m1 */
void orc_construct_hlbvh(const OrcTriangle* tris, uint32_t T, const OrcSphere* sphs, uint32_t S,
                         const OrcMorton* sorted, OrcNode* nodes, OrcCInfo* cinfo);
void orc_refit_aabbs(OrcNode* nodes, OrcCInfo* cinfo, uint32_t N);
/* all of the above in dispatch order; tris/sphs are transformed in place (K1) */
void orc_build_bvh(const OrcModel* models, OrcTriangle* tris, uint32_t T, OrcSphere* sphs, uint32_t S,
                   const OrcOptions* opt, OrcEnclosing* enc, OrcMorton* m1, OrcMorton* m2,
                   OrcNode* nodes, OrcCInfo* cinfo);

/* ---- S2: trace (RaytracerBVH.cpp:998-1050 + raytraceBVH.comp) ---- */
void orc_clear_image(float* rgba, uint32_t W, uint32_t H);       /* (0,0,0,1): RaytracerBVH.cpp:772 */
/* `spp` dispatches of raytraceBVH.comp over rows [y0,y1) of a W x H RGBA32F image (row 0 = top).
 * Optional outputs (may be NULL): hitPrim[W*H] = leaf primitive g of the primary ray of the first dispatch
 * (triangle g<T, sphere T+idx, 0xFFFFFFFF = miss), hitT[W*H] its t; rngOut[W*H] = rngState after the last
 * dispatch; counters. */
int orc_raytrace(const OrcUBO* ubo, float* rgba, uint32_t W, uint32_t H, uint32_t y0, uint32_t y1,
                 const OrcTriangle* tris, const OrcSphere* sphs, const OrcMaterial* mats, const OrcNode* nodes,
                 uint32_t spp, const OrcOptions* opt,
                 uint32_t* hitPrim, float* hitT, uint32_t* rngOut, OrcCounters* counters);
/* brute force closest hit of the primary rays (raytrace.comp:167-190 order: triangles then spheres) */
void orc_primary_hits_bruteforce(const OrcUBO* ubo, uint32_t W, uint32_t H,
                                 const OrcTriangle* tris, const OrcSphere* sphs, uint32_t* hitPrim, float* hitT);
/* fragment resolve: SingleTriangleFullScreen.frag:13-21 -> RGBA8 */
void orc_resolve_rgba8(const float* rgba, uint32_t W, uint32_t H, uint32_t raysPerPixel, uint8_t* out);
int  orc_max_threads(void);
/* LogisticMap demo program, one dispatch of logistic.comp (LogisticMap.cpp:384): points = (x, r) float pairs */
void orc_logistic_step(float* points, uint32_t count, uint8_t* rgba8, uint32_t W, uint32_t H, const float* pixelColor);

#ifdef __cplusplus
}
#endif
#endif
