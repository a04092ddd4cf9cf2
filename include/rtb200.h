/*
 * rtb200.h -- C-ABI of librtb200.so: the B200 (sm_100a) compute path of the RaytracerGPU path tracer.
 *
 * This is the drop-in boundary.  The reference (silvercorked/RaytracerGPU_MastersProject) has no FFI: its
 * renderer hosts call C++ wrapper classes over raw Vulkan.  Every entry point below replaces one of those
 * call sites; the reference file:line it stands in for is cited per function (paths relative to
 * RaytracerGPU_MastersProject/).  Plain pointers and sizes only; no C++ / torch types.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; rtb_last_error() describes the last failure
 *    of the calling thread (the reference throws std::runtime_error at the same places, e.g.
 *    RaytracerBVH.cpp:263-264, RaytracerBVH.hpp:398-404; the C++ shims in host/ re-throw).
 *  - one CUDA stream per context; all kernel entry points are asynchronous on that stream and ordered,
 *    rtb_sync() is the fence wait (vkQueueSubmit + vkWaitForFences, RaytracerBVH.hpp:398-404,474-478).
 *  - a context is not thread-safe (neither is the reference: single host thread, one frame in flight).
 *  - "device pointer" arguments may come from rtb_alloc or from any other allocator of the same CUDA
 *    primary context (e.g. a torch tensor's data_ptr()).
 *  - there is NO CPU fallback: without a CUDA device rtb_ctx_create fails.
 *
 * Record layouts are the reference's std430/std140 layouts byte for byte
 * (shaders/include/definitions.glsl:6-77 == VulkanWrapper/SceneTypes.hpp:32-123).
 */
#ifndef RTB200_H
#define RTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_VERSION 200

/* ---- record layouts ------------------------------------------------------------------------------- */
typedef struct { float m[16]; } rtb_model;                         /* mat4 column-major, 64 B (definitions.glsl:6-8) */
typedef struct { float v0[4], v1[4], v2[4]; uint32_t materialIndex, modelIndex, _pad[2]; } rtb_triangle; /* 64 B (:15-21) */
typedef struct { float center[4]; float radius; uint32_t materialIndex, modelIndex, _pad; } rtb_sphere;   /* 32 B (:23-28) */
typedef struct { float albedo[4]; uint32_t materialType, _pad[3]; } rtb_material;                         /* 32 B (:10-13) */
typedef struct { float minX, maxX, minY, maxY, minZ, maxZ; } rtb_aabb;                                     /* 24 B (:52-56) */
typedef struct { rtb_aabb aabb; uint32_t leftIndex, rightIndex, primitiveIndex, primitiveType; } rtb_bvh_node; /* 40 B (:66-72) */
typedef struct { uint32_t code, primitiveIndex, primitiveType; } rtb_morton_primitive;                    /* 12 B (:58-62) */
typedef struct { uint32_t parent; int32_t visitationCount; } rtb_construction_info;                       /*  8 B (:74-77) */
typedef struct { float eMin[4], eMax[4]; } rtb_enclosing_box;                  /* 32 B (GetEnclosingAABB.comp:25-28) */

/* RaytracingUniformBufferObject, 80 B std140 (RaytracerBVH.hpp:31-42 == raytraceBVH.comp:7-18) */
typedef struct {
    float camPos[4], camLookAt[4], camUpDir[4];
    float verticalFOV;
    uint32_t numTriangles, numSpheres, numMaterials, numLights, maxRayTraceDepth, randomState;
    uint32_t _pad;
} rtb_ubo;

enum { RTB_LIGHT = 0, RTB_DIFFUSE = 1, RTB_METALLIC = 2, RTB_DIELECTRIC = 3 };  /* definitions.glsl:79-82 */
enum { RTB_SPHERE_PRIMITIVE = 0, RTB_TRIANGLE_PRIMITIVE = 1 };                  /* definitions.glsl:84-85 */

/* flags for rtb_trace_args.flags */
enum {
    RTB_TRACE_COUNT = 1u << 0,          /* instrumented variant: fill `counters` (same traversal, same results) */
    RTB_TRACE_EXT_MATERIALS = 1u << 1,  /* extension N1: metal / dielectric scatter (NOT reference behaviour) */
    RTB_TRACE_ENCLOSING_INF = 1u << 2,  /* build option: enclosing-AABB locals start at +-inf instead of pin U4 (0.0) */
    RTB_TRACE_SIMPLE_KERNEL = 1u << 3,  /* run the straightforward one-lane-one-pixel kernel (trace.cu) instead of the
                                           warp-coherent one (trace_wave.cu); same results, kept for A/B measurements */
    RTB_TRACE_LINEAR_SCAN = 1u << 4,    /* the NON-BVH program (Config::Programs::Raytracer): raytrace.comp's sceneHit loops
                                           over all triangles then all spheres (raytrace.comp:167-190), background
                                           (0.1,0.1,0.3) (:43).  Needs no BVH: bind with nodes = NULL. */
    RTB_TRACE_STREAM_KERNEL = 1u << 6,  /* run the streaming (wavefront) kernel trace_stream.cu instead of trace_wave.cu; same
                                           results (A/B switch while both exist) */
    /* Traversal records.  Default (no flag): 64-byte 4-ary records with conservative 8-bit boxes for scenes of >= 512
     * primitives (two binary levels per step; exactness is restored at the leaves), the exact 64-byte child
     * pairs below that (the derivation of the 4-ary records costs more than it saves on tiny scenes).  The flags force
     * one representation (A/B measurements); results are identical in all cases.  The instrumented variant
     * (RTB_TRACE_COUNT) always walks the exact child pairs because it counts the reference's node visits. */
    RTB_TRACE_COMPRESSED_NODES = 1u << 7, /* 32-byte compressed child pairs (+6 % on C3, -2..4 % on C2 / C4) */
    RTB_TRACE_WIDE_NODES = 1u << 8,     /* 64-byte 4-ary records (+6..21 % on C2..C5) */
    RTB_TRACE_EXACT_NODES = 1u << 9,    /* exact 64-byte child pairs */
    /* The reference's camera rays have no jitter (raytraceBVH.comp:329-342): the sampleCount samples of a pixel start with the
     * same primary ray, whose hitBVH result is therefore traced once per pixel per rtb_raytrace call and shared by the samples
     * (identical results; nothing is kept between calls).  This flag traces it once per sample like the shader does. */
    RTB_TRACE_NO_PRIMARY_SHARING = 1u << 10,
    /* Default for >= 512 primitives when the scene's hit-point slack is small (DESIGN.md "nearest-first traversal"): the
     * 4-ary records are walked nearest-first and entries that cannot change the result (beyond the closest hit so far, or
     * before tMin) are dropped; equal-t ties are resolved as the reference's visiting order would.  This flag keeps the
     * reference's visiting order with no t-interval, like the shader. */
    RTB_TRACE_REFERENCE_ORDER = 1u << 11,
    /* Instrumented variant of the PRODUCTION walk: the same kernels, records, visiting order, hand-over and primary-hit
     * sharing as an un-flagged call (same results), additionally filling `walkCounters` with what those kernels really
     * fetch -- the figures bench.py's roofline is computed from.  (RTB_TRACE_COUNT, by contrast, walks the exact records in
     * the reference's order and counts the REFERENCE's work.) */
    RTB_TRACE_WALK_COUNT = 1u << 12,
    RTB_TRACE_CULLED = 1u << 5          /* extension, default off: also skip subtrees outside the box of the ray segment
                                           [tMin, closest] (+ margin).  NOT the reference's traversal (it has no t-interval);
                                           fewer node visits, results empirically identical (see trace_wave.cu) */
};

/* Device-side work counters (u64 each), see DESIGN.md "roofline": rays = hitBVH calls, nodeVisits = nodes
 * whose box was tested (reference-equivalent count), triTests / sphTests = leaf primitive tests,
 * matReads = hits that read a material, samples = (pixel, sample) pairs. */
typedef struct { uint64_t rays, nodeVisits, triTests, sphTests, matReads, samples; } rtb_counters;

/* What the production trace kernels fetch (RTB_TRACE_WALK_COUNT), u64 each, accumulated into:
 *   rays            rays actually walked (a pixel's shared primary ray counts once)
 *   recordFetches   64-byte traversal records fetched (4-ary records; exact child pairs for small scenes / fallback rays)
 *   leafBoxFetches  32-byte exact leaf boxes fetched (candidate re-check)
 *   triTests        64-byte packed triangle records fetched and tested;  sphTests: 16-byte sphere + 4-byte material index
 *   matReads        16-byte material records read by the shading step
 *   items           (pixel, sample) work items pulled: 8 B pixel index / xy + 4 B seed (+ 48 B shared primary hit)
 *   paths           finished paths: one 12-byte colour store into the sample slot
 *   parked          paths / rays handed over to the tail kernel: 240 B written by the main launch and read back by the tail launch
 *   laneSteps, warpSteps   traverse-phase turns per lane / per warp (their ratio = lanes active per T step)
 *   tailRays        rays finished one-ray-per-warp by the tail kernel;  tailTurns: its warp turns
 *   uniqueRecordFetches  recordFetches of the main launch with the lanes of a warp step that expand the SAME record counted once
 *                   (what reaches L1 as distinct sectors: compare with ncu's l1tex__t_sectors_pipe_lsu_mem_global_op_ld) */
typedef struct {
    uint64_t rays, recordFetches, leafBoxFetches, triTests, sphTests, matReads, items, paths, parked, laneSteps, warpSteps,
             tailRays, tailTurns, uniqueRecordFetches, _reserved[2];
} rtb_walk_counters;

/* What one S2 submission renders (the raysPerPixel dispatch loop of RaytracerBVH.cpp:1025-1050).
 * The image buffer is RGBA32F, `localRows` x imageWidth, row-major, local row j holding global image row
 *     y(j) = ((j / bandRows) * bandStep + bandFirst) * bandRows + j % bandRows          (rows y >= H untouched)
 * Single GPU: bandRows = imageHeight, bandFirst = 0, bandStep = 1, localRows = imageHeight.
 * Tile sharding over n ranks: bandStep = n, bandFirst = rank.
 * sampleSkip = number of links of the per-pixel alpha seed chain (raytraceBVH.comp:349-352,372) to
 * fast-forward, starting from the alpha currently in the image, before the first rendered sample
 * (sample-range sharding: cleared image + sampleSkip = first sample index of this rank). */
typedef struct {
    uint32_t imageWidth, imageHeight;   /* full image: defines the camera and the RNG seeds */
    uint32_t localRows;
    uint32_t bandRows, bandFirst, bandStep;
    uint32_t sampleSkip, sampleCount;
    uint32_t flags;
    uint32_t _pad;
    void* hitPrim;      /* optional device u32[localRows*W]: leaf primitive of the primary ray of the first rendered
                           sample (triangle g < T, sphere T + idx, 0xFFFFFFFF miss) */
    void* hitT;         /* optional device f32[localRows*W] */
    void* rngOut;       /* optional device u32[localRows*W]: rngState after the last rendered sample */
    void* counters;     /* device rtb_counters*, required with RTB_TRACE_COUNT (accumulated into) */
    void* walkCounters; /* device rtb_walk_counters*, required with RTB_TRACE_WALK_COUNT (accumulated into) */
} rtb_trace_args;

typedef struct rtb_ctx rtb_ctx;

/* ---- context == Device + compute queue (VulkanWrapper/Device.hpp:27-113, Device.cpp:115-231) ------ */
const char* rtb_last_error(void);
int rtb_version(void);
int rtb_device_count(int* count);
/* `stream`: a cudaStream_t to launch on (e.g. torch's current stream) or NULL to create a private one. */
int rtb_ctx_create(int device, void* stream, rtb_ctx** out);
int rtb_ctx_destroy(rtb_ctx* ctx);
int rtb_sync(rtb_ctx* ctx);                                   /* vkWaitForFences: RaytracerBVH.hpp:402,476 */
int rtb_device_name(rtb_ctx* ctx, char* buf, size_t len);     /* Device.cpp:137 prints it */
int rtb_sm_count(rtb_ctx* ctx, int* count);

/* ---- Buffer (VulkanWrapper/Buffer.hpp:23-57) + Device::copyBuffer (Device.hpp:102) ---------------- */
int rtb_alloc(rtb_ctx* ctx, size_t bytes, void** dptr);       /* Buffer(device, size, count, STORAGE, DEVICE_LOCAL) */
int rtb_free(rtb_ctx* ctx, void* dptr);                       /* Buffer::~Buffer (Buffer.cpp:36-40) */
int rtb_upload(rtb_ctx* ctx, void* dst, const void* host, size_t bytes);    /* staging map/write + copyBuffer (RaytraceScene.hpp:139-178) */
int rtb_download(rtb_ctx* ctx, void* host, const void* src, size_t bytes);  /* DEBUGgetDeployedBufferAs (RaytracerBVH.hpp:575-617) */
int rtb_memset(rtb_ctx* ctx, void* dst, int byte, size_t bytes);
int rtb_host_alloc(size_t bytes, void** hptr);                /* pinned staging memory (HOST_VISIBLE|HOST_COHERENT) */
int rtb_host_free(void* hptr);
/* CUDA-event timing on the context's stream (the reference uses std::chrono around submit+wait,
 * RaytracerBVH.hpp:394-408,470-478) */
int rtb_timer_start(rtb_ctx* ctx);
int rtb_timer_stop_ms(rtb_ctx* ctx, float* ms);               /* synchronises */

/* ---- S1: BVH build, one entry point per dispatch; argument order = descriptor binding order -------- */
/* K1 ModelSpaceToWorldSpace.comp (bindings RaytracerBVH.cpp:638-643, dispatch :789): in place */
int rtb_model_to_world(rtb_ctx* ctx, const rtb_ubo* ubo, const void* models, void* triangles, void* spheres);
/* K2 GetEnclosingAABB.comp (bindings :644-650, dispatch :842) */
int rtb_enclosing_aabb(rtb_ctx* ctx, const rtb_ubo* ubo, void* enclosing, const void* triangles, const void* spheres, uint32_t flags);
/* K3 GenerateMortonCodesOfPrimitives.comp (bindings :651-658, dispatch :879) */
int rtb_morton_codes(rtb_ctx* ctx, const rtb_ubo* ubo, const void* enclosing, const void* triangles, const void* spheres, void* morton1);
/* K4 RadixSortSimple.comp (bindings :659-663, dispatch :916): stable ascending by code, result in morton1 */
int rtb_sort_morton(rtb_ctx* ctx, const rtb_ubo* ubo, void* morton1, void* morton2);
/* K5 ConstructHLBVH.comp (bindings :664-671, dispatch :954) */
int rtb_build_hlbvh(rtb_ctx* ctx, const rtb_ubo* ubo, const void* triangles, const void* spheres, const void* morton1, void* nodes, void* constructionInfo);
/* K6 ConstructAABBsOfInternalNodes.comp (bindings :672-676, dispatch :991) */
int rtb_refit_aabbs(rtb_ctx* ctx, const rtb_ubo* ubo, void* nodes, void* constructionInfo);
/* Whole S1 command buffer (recordComputeS1CommandBuffer, RaytracerBVH.cpp:734-997) minus the image clear.
 * Any of enclosing / morton1 / morton2 / constructionInfo may be NULL (context scratch is used); `nodes`
 * (reference 40-byte layout, 2N-1 records) may be NULL when the caller does not need the array.
 * Also binds the result for rtb_raytrace (see rtb_bind_trace_buffers). */
int rtb_build_bvh(rtb_ctx* ctx, const rtb_ubo* ubo, const void* models, void* triangles, void* spheres,
                  const void* materials, void* enclosing, void* morton1, void* morton2, void* nodes,
                  void* constructionInfo, uint32_t flags);

/* ---- S2: trace ------------------------------------------------------------------------------------ */
/* vkCmdClearColorImage to (0,0,0,1) (RaytracerBVH.cpp:772-776) */
int rtb_clear_image(rtb_ctx* ctx, void* image, uint32_t width, uint32_t rows);
/* DescriptorWriter.writeBuffer for the raytrace set (RaytracerBVH.cpp:677-685): binds world-space
 * triangles / spheres, materials and the reference-layout node array, and derives the 16-byte-aligned
 * traversal records the kernel actually fetches (DESIGN.md "data layout"). */
int rtb_bind_trace_buffers(rtb_ctx* ctx, const rtb_ubo* ubo, const void* triangles, const void* spheres,
                           const void* materials, const void* nodes /* NULL for the non-BVH program (Raytracer.cpp:394-538) */);
/* recordComputeS2CommandBuffer + submit (RaytracerBVH.cpp:998-1050): all samples of args->sampleCount in
 * one persistent launch; bit-identical to sampleCount dispatches of raytraceBVH.comp. */
int rtb_raytrace(rtb_ctx* ctx, const rtb_ubo* ubo, void* image, const rtb_trace_args* args);
/* SingleTriangleFullScreen.frag:13-21 (+ FragmentUniformBufferObject RaytracerBVH.hpp:47-49) -> RGBA8 */
int rtb_resolve_rgba8(rtb_ctx* ctx, const void* image, uint32_t width, uint32_t rows, uint32_t raysPerPixel, void* outRgba8);
/* ---- the LogisticMap demo program (Config::Programs::LogisticMap; not part of the render path) ------------------- */
/* One dispatch of logistic.comp (LogisticMap.cpp:384, bindings: UBO {pixelColor, iteration, width, height}, the (x, r)
 * point SSBO, an rgba8 storage image that is never cleared): x' = x * r * (1 - x) per point, in place, then
 * imageStore(int(r / 4 * width), int((1 - x') * height)) = pixelColor. */
int rtb_logistic_step(rtb_ctx* ctx, void* points /* float2[count] */, uint32_t count, void* imageRgba8, uint32_t width,
                      uint32_t height, const float* pixelColor /* [4] */);
/* ---- multi-GPU (no reference counterpart: the reference renders on the first device only, VulkanWrapper/Device.cpp:115-138) ----
 * One context per GPU -- one process per GPU (torchrun, MPI, ...) or one host thread per GPU of one process.  The scene is
 * replicated and every rank builds the same BVH; a frame is partitioned either by interleaved 8-row bands (rtb_trace_args
 * bandFirst = rank, bandStep = nRanks: bit-identical to one GPU) or by sample range (sampleSkip / sampleCount: fp32 sum order
 * differs, tolerance).  The calls below are the ONE exchange at the end of a frame.  They run over NCCL (NVLink / NVSwitch),
 * bound at run time (libnccl.so.2; RTB_NCCL_LIBRARY overrides the name), are enqueued on the context's stream like every
 * kernel entry point, and must be called by all ranks of the communicator. */
#define RTB_COMM_ID_BYTES 128
/* ncclGetUniqueId: one rank creates the id, the host distributes it to the others by any means (file, MPI, torch.distributed) */
int rtb_comm_unique_id(void* id128);
int rtb_comm_init_rank(rtb_ctx* ctx, int nRanks, int rank, const void* id128);
int rtb_comm_destroy(rtb_ctx* ctx);                           /* also done by rtb_ctx_destroy */
int rtb_comm_info(rtb_ctx* ctx, int* rank, int* nRanks);      /* (0, 1) without a communicator */
/* In-place all-gather of a device buffer of nRanks * bytesPerRank bytes whose slice [rank * bytesPerRank, +bytesPerRank) this
 * rank has filled: each rank uploads 1/n of the scene arrays over its own PCIe link and the rest arrives over NVLink
 * (replaces n full RaytraceScene::updateScene uploads, RaytraceScene.cpp:78-113). */
int rtb_comm_all_gather(rtb_ctx* ctx, void* buffer, size_t bytesPerRank);
int rtb_comm_broadcast(rtb_ctx* ctx, void* buffer, size_t bytes, int root);
/* Tile mode.  localImage: this rank's RGBA32F bands as rendered with bandRows / bandFirst = rank / bandStep = nRanks
 * (localRows = ceil(ceil(height / bandRows) / nRanks) * bandRows rows of `width` pixels).  On return (stream order) EVERY rank
 * holds the assembled frame: frameRgba32f (height x width RGBA32F, may be NULL) and / or frameRgba8 (the fragment-shader
 * resolve, SingleTriangleFullScreen.frag:13-21, fused into the re-assembly; may be NULL).  With frameRgba32f == NULL the bands
 * are resolved before the exchange and only 4 bytes per pixel cross NVLink. */
int rtb_gather_tiles(rtb_ctx* ctx, const void* localImage, uint32_t width, uint32_t height, uint32_t bandRows,
                     void* frameRgba32f, uint32_t raysPerPixel, void* frameRgba8);
/* Sample-range mode.  image: this rank's full-frame partial sums (rendered from a cleared image with its own sampleSkip /
 * sampleCount); it is overwritten.  On `root` it ends up holding the sum of the rgb planes over all ranks and the LAST rank's
 * alpha (= the end of the per-pixel seed chain); frameRgba8 (root only, may be NULL) receives the resolved frame. */
int rtb_reduce_samples(rtb_ctx* ctx, void* image, uint32_t width, uint32_t height, int root, uint32_t raysPerPixel, void* frameRgba8);

/* Measurement aid: rate (GB/s) at which this GPU delivers divergent 64-byte record fetches -- every lane of a persistent grid of
 * the trace kernels' shape fetching records at random indices of a scratch buffer of `footprintBytes` -- i.e. the node-fetch
 * roofline at a scene's record footprint (bench.py reports the trace kernels' fetch rate against it).  Synchronises. */
int rtb_probe_gather(rtb_ctx* ctx, size_t footprintBytes, float* gbPerSecond);
/* rtb_download without the wait: the copy is complete after the next rtb_sync (or any later synchronising call) of this context.
 * `host` should be pinned (rtb_host_alloc) for the copy to overlap other streams. */
int rtb_download_async(rtb_ctx* ctx, void* host, const void* src, size_t bytes);

/* Diagnostic: the per-primitive hit-point slack eta (bvh_build.cu eta_leaf_kernel) the 4-ary records of the bound scene were grown
 * by, one float per primitive in leaf order (triangles, then spheres), and the origin region (6 floats: min.xyz, max.xyz) inside
 * which rays are t-culled.  Fails if the bound scene has no 4-ary records yet.  Synchronises.  (tests: every accepted hit must lie
 * within eta of its primitive's leaf box) */
int rtb_export_hit_slack(rtb_ctx* ctx, float* hostEta, size_t count, float* hostOriginRegion6);

/* number of kernels / collectives this context has launched so far (bench.py's gpu_launches) */
int rtb_launch_count(rtb_ctx* ctx, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif
