// capi.cu -- the C-ABI of librtb200.so (include/rtb200.h): context, buffers, and one entry point per reference
// dispatch.  Host-side work here is limited to what the reference does on the host for the same call (filling the
// UBO-derived camera once per submission, choosing launch shapes) -- all data-path arithmetic runs in the kernels.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "common.cuh"
#include "ctx.h"
#include "kernels.h"

using namespace rtb;

namespace rtb {
std::string& last_error() { thread_local std::string e; return e; }
}  // namespace rtb

namespace {


// Camera globals of raytraceBVH.comp:50-81, evaluated once on the host in binary32 with the shader's operation
// order (compiled with -ffp-contract=off); tan() is libm's tanf.
Camera make_camera(const rtb_ubo* ubo, uint32_t W, uint32_t H) {
    const float FOCAL = 10.0f;
    const float aspect = (float)W / (float)H;
    const float theta = ubo->verticalFOV * 0.017453292519943295f;
    const float h = tanf(theta / 2);
    const float viewportHeight = 2.0f * h * FOCAL;
    const float viewportWidth = viewportHeight * aspect;
    const f3 camPos = F3(ubo->camPos[0], ubo->camPos[1], ubo->camPos[2]);
    const f3 lookAt = F3(ubo->camLookAt[0], ubo->camLookAt[1], ubo->camLookAt[2]);
    const f3 up = F3(ubo->camUpDir[0], ubo->camUpDir[1], ubo->camUpDir[2]);
    const f3 camW = normalize(camPos - lookAt);
    const f3 camU = normalize(cross(up, camW));
    const f3 camV = cross(camW, camU);
    const f3 viewportU = viewportWidth * camU;
    const f3 viewportV = viewportHeight * (-camV);
    Camera cam;
    cam.deltaU = viewportU / (float)W;
    cam.deltaV = viewportV / (float)H;
    const f3 upperLeft = ((camPos - FOCAL * camW) - viewportU / 2.0f) - viewportV / 2.0f;
    cam.pixel00 = upperLeft + 0.5f * (cam.deltaU + cam.deltaV);
    cam.origin = camPos;
    return cam;
}

int bind_internal(rtb_ctx* c, uint32_t T, uint32_t S, uint32_t M, const void* tris, const void* sphs, const void* mats, const void* nodes,
                  bool pairsDone = false, bool primsDone = false) {
    const uint32_t N = T + S;
    if (ensure(c, c->pairs, sizeof(float4) * 4ull * (N > 1 ? N - 1 : 1))) return 1;
    if (ensure(c, c->ptris, sizeof(float4) * 4ull * T)) return 1;
    if (ensure(c, c->psphs, sizeof(float4) * (size_t)S)) return 1;
    if (ensure(c, c->psphMat, sizeof(uint32_t) * (size_t)S)) return 1;
    if (ensure(c, c->pmats, sizeof(float4) * (size_t)M)) return 1;
    if (ensure(c, c->rootBox, sizeof(float4) * 4)) return 1;
    if (ensure(c, c->workCounter, 32)) return 1;
    if (ensure(c, c->errFlag, 16)) return 1;
    if (nodes && !pairsDone) launch_pack_pairs(c->stream, nodes, N, c->pairs.p, c->rootBox.p);
    if (primsDone) launch_pack_prims(c->stream, tris, 0, sphs, 0, mats, M, c->ptris.p, c->psphs.p, c->psphMat.p, c->pmats.p);   // materials only
    else launch_pack_prims(c->stream, tris, T, sphs, S, mats, M, c->ptris.p, c->psphs.p, c->psphMat.p, c->pmats.p);
    c->boundNodesPtr = nodes;
    c->cnodesReady = false;                 // the compressed / wide records are derived on first use
    c->wideReady = false; c->wideRefReady = false; c->hoisted = false;
    c->leafBoxReady = false;
    if (check_launch(c, 2, "pack traversal records")) return 1;
    c->bound = true; c->boundNodes = nodes != nullptr; c->bT = T; c->bS = S; c->bM = M; c->bN = N;
    return 0;
}

// Derived traversal records (DESIGN.md "data layout"): exact leaf boxes + the 32-byte (mode 1) or 4-ary 64-byte (mode 2) records,
// once per bound node array.  Returns the number of launches, -1 on failure.
int derive_records(rtb_ctx* c, int nodesMode, const float* camPos) {
    int extra = 0;
    if (nodesMode && (!c->leafBoxReady || (nodesMode == 1 && !c->cnodesReady))) {   // exact leaf boxes (+ the 32-byte records)
        if (nodesMode == 1 && ensure(c, c->cnodes, 32ull * (c->bN - 1))) return -1;
        if (ensure(c, c->leafBox, 32ull * c->bN)) return -1;
        launch_pack_cnodes(c->stream, c->boundNodesPtr, c->bN, nodesMode == 1 ? c->cnodes.p : nullptr, c->leafBox.p);
        c->leafBoxReady = true; if (nodesMode == 1) c->cnodesReady = true;
        extra++;
    }
    if (nodesMode == 2 && !c->wideReady) {
        // Hit-point slack (bvh_build.cu eta_leaf_kernel): a point the reference's primitive tests accept, for a
        // ray that starts inside the scene's box, lies within eta of the primitive, so inside every record box grown by
        // the largest eta of its subtree.  If any primitive's slack is not finite (zero-area triangle) the records stay
        // tight and the walk drops nothing by t; pack_wide_kernel decides that on the device (walkFlag), no read-back.
        const size_t nn = 2ull * c->bN - 1;
        if (ensure(c, c->wide, 64ull * (c->bN - 1)) || ensure(c, c->etaNode, 4 * nn) || ensure(c, c->etaParent, 4 * nn) ||
            ensure(c, c->etaArrivals, 4ull * c->bN) || ensure(c, c->walkFlag, 64)) return -1;
        extra += launch_eta(c->stream, c->boundNodesPtr, c->bN, c->ptris.p, c->bT, c->psphs.p, c->bS, c->rootBox.p, camPos,
                            (float*)c->etaNode.p, (uint32_t*)c->etaParent.p, (unsigned int*)c->etaArrivals.p);
        launch_pack_wide(c->stream, c->boundNodesPtr, c->bN, c->wide.p, (const float*)c->etaNode.p, (unsigned int*)c->walkFlag.p);
        c->wideReady = true; extra++;
#ifdef RTB_SMEM_TOP
        if (ensure(c, c->topTable, 64ull * RTB_SMEM_TOP) || ensure(c, c->topGlobal, 4ull * RTB_SMEM_TOP)) return -1;
        launch_build_top_table(c->stream, c->wide.p, c->bN, c->topTable.p, c->topGlobal.p, (const unsigned int*)c->walkFlag.p); extra++;
#endif
    }
    return extra;
}

}  // namespace

extern "C" {

const char* rtb_last_error(void) { return rtb::last_error().c_str(); }
int rtb_version(void) { return RTB_VERSION; }

int rtb_device_count(int* count) {
    REQUIRE(count, "rtb_device_count: null argument");
    CK(cudaGetDeviceCount(count));
    return 0;
}

int rtb_ctx_create(int device, void* stream, rtb_ctx** out) {
    REQUIRE(out, "rtb_ctx_create: null out pointer");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) return fail("rtb_ctx_create: no CUDA device (librtb200 has no CPU fallback)", e);
    REQUIRE(device >= 0 && device < n, "rtb_ctx_create: device index out of range");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("rtb_ctx_create: librtb200 is built for sm_100a (Blackwell B200) only");
    rtb_ctx* c = new rtb_ctx();
    c->device = device;
    c->smCount = prop.multiProcessorCount;
    snprintf(c->name, sizeof(c->name), "%s", prop.name);
    if (stream) { c->stream = (cudaStream_t)stream; c->ownStream = false; }
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return fail("cudaStreamCreate", e); }
        c->ownStream = true;
    }
    cudaEventCreate(&c->ev0);
    cudaEventCreate(&c->ev1);
    cudaStreamCreateWithFlags(&c->buildStream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&c->evBuildFork, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->evBuildJoin, cudaEventDisableTiming);
    // RTB_WAVE_TAIL_OVERLAP=1 (experimental, default off): trace_tail_kernel is additionally launched on a second stream beside the
    // main trace launch, so that parked long rays are worked off while it drains.  Measured on B200 (profiles/r02_tail_overlap.txt):
    // correct but slower (36.1 -> 62.4 ms per C2 frame), so the tail launch stays strictly behind the main launch.
    const char* ovl = getenv("RTB_WAVE_TAIL_OVERLAP");
    if (ovl && atoi(ovl) != 0) {
        cudaStreamCreateWithFlags(&c->auxStream, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming);
    }
    if (const char* e = getenv("RTB_WAVE_TMIN")) c->knobs.tMin = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_SORTED_PUSH")) c->knobs.sortedPush = atoi(e);
    if (const char* e = getenv("RTB_WAVE_QGATE")) c->knobs.qGate = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WIDE_MIN")) c->knobs.wideMin = (uint32_t)(atoi(e) > 64 ? atoi(e) : 64);
    if (const char* e = getenv("RTB_WAVE_SMIN")) c->knobs.sMin = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_MAIN_CTAS")) c->knobs.mainCtas = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_TAIL_THREADS")) c->knobs.tailThreads = atoi(e) == 64 ? 64u : 128u;
    if (const char* e = getenv("RTB_WAVE_COOP")) c->knobs.coopMax = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_COOP_TURNS")) c->knobs.coopTurns = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_TAIL_SPIN_US")) c->knobs.tailSpinUs = (uint32_t)atoi(e);
    if (const char* e = getenv("RTB_WAVE_SAMPLE_BUF_MB")) c->knobs.sampleBufBytes = (size_t)atoll(e) << 20;
    if (const char* e = getenv("RTB_STREAM_POOL")) c->knobs.streamPool = (size_t)atoll(e);
    *out = c;
    return 0;
}

int rtb_ctx_destroy(rtb_ctx* c) {
    if (!c) return 0;
    if (c->comm) rtb_comm_destroy(c);
    Activate act(c);
    cudaStreamSynchronize(c->stream);
    for (Scratch* s : { &c->sortKeys[0], &c->sortKeys[1], &c->sortVals[0], &c->sortVals[1], &c->sortCounts, &c->encRed, &c->enclosing,
                        &c->cinfo, &c->nodes, &c->pairs, &c->ptris, &c->psphs, &c->psphMat, &c->pmats, &c->rootBox, &c->workCounter,
                        &c->errFlag, &c->walkFlag, &c->bigList, &c->wideRef, &c->ttChild, &c->ttParent, &c->ttArrivals, &c->ttBox, &c->ttEta, &c->topTable, &c->topGlobal, &c->gatherBuf, &c->bandRgba8, &c->etaNode, &c->etaParent, &c->etaArrivals, &c->cnodes, &c->leafBox, &c->wide, &c->activePix, &c->activeXY, &c->sampleBuf, &c->primaryHits, &c->poolSlot, &c->poolColor,
                        &c->poolAtt, &c->poolOrg, &c->poolDir, &c->poolNrm, &c->poolList, &c->poolCnt, &c->parkBuf })
        release(*s);
    cudaEventDestroy(c->ev0);
    cudaEventDestroy(c->ev1);
    if (c->buildStream) { cudaStreamSynchronize(c->buildStream); cudaStreamDestroy(c->buildStream); cudaEventDestroy(c->evBuildFork); cudaEventDestroy(c->evBuildJoin); }
    if (c->auxStream) { cudaStreamSynchronize(c->auxStream); cudaStreamDestroy(c->auxStream); cudaEventDestroy(c->evFork); cudaEventDestroy(c->evJoin); }
    if (c->ownStream) cudaStreamDestroy(c->stream);
    delete c;
    return 0;
}

// Every entry point that synchronises (rtb_sync, rtb_download, rtb_timer_stop_ms) ends here: wait for the stream, then surface
// (and clear) the device-side error flag, so a frame rendered with an overflowed traversal stack is never returned as good.
static int sync_and_check(rtb_ctx* c) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->errFlag.p && c->traced) {
        unsigned int f = 0;
        CK(cudaMemcpy(&f, c->errFlag.p, sizeof(f), cudaMemcpyDeviceToHost));
        c->traced = false;
        if (f) {
            cudaMemset(c->errFlag.p, 0, sizeof(f));
            return fail("rtb_raytrace: traversal stack overflow (tree deeper than the traversal stack)");
        }
    }
    return 0;
}

int rtb_sync(rtb_ctx* c) {
    REQUIRE(c, "rtb_sync: null context");
    Activate act(c);
    return sync_and_check(c);
}

int rtb_device_name(rtb_ctx* c, char* buf, size_t len) {
    REQUIRE(c && buf && len, "rtb_device_name: bad argument");
    snprintf(buf, len, "%s", c->name);
    return 0;
}
int rtb_sm_count(rtb_ctx* c, int* count) { REQUIRE(c && count, "rtb_sm_count: bad argument"); *count = c->smCount; return 0; }

int rtb_alloc(rtb_ctx* c, size_t bytes, void** dptr) {
    REQUIRE(c && dptr, "rtb_alloc: bad argument");
    Activate act(c);
    *dptr = nullptr;
    CK(cudaMalloc(dptr, bytes ? bytes : 16));
    return 0;
}
int rtb_free(rtb_ctx* c, void* dptr) {     // c may be NULL (a Buffer outliving its Device): cudaFree synchronises the device itself
    if (!dptr) return 0;
    if (c) {
        Activate act(c);
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaFree(dptr));
    } else {
        CK(cudaFree(dptr));
    }
    return 0;
}
int rtb_upload(rtb_ctx* c, void* dst, const void* host, size_t bytes) {
    REQUIRE(c && (bytes == 0 || (dst && host)), "rtb_upload: bad argument");
    Activate act(c);
    if (bytes) CK(cudaMemcpyAsync(dst, host, bytes, cudaMemcpyHostToDevice, c->stream));
    return 0;
}
int rtb_download(rtb_ctx* c, void* host, const void* src, size_t bytes) {
    REQUIRE(c && (bytes == 0 || (src && host)), "rtb_download: bad argument");
    Activate act(c);
    if (bytes) CK(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return sync_and_check(c);
}
int rtb_download_async(rtb_ctx* c, void* host, const void* src, size_t bytes) {
    REQUIRE(c && (bytes == 0 || (src && host)), "rtb_download_async: bad argument");
    Activate act(c);
    if (bytes) CK(cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return 0;
}
int rtb_memset(rtb_ctx* c, void* dst, int byte, size_t bytes) {
    REQUIRE(c && (bytes == 0 || dst), "rtb_memset: bad argument");
    Activate act(c);
    if (bytes) CK(cudaMemsetAsync(dst, byte, bytes, c->stream));
    return 0;
}
int rtb_host_alloc(size_t bytes, void** hptr) {
    REQUIRE(hptr, "rtb_host_alloc: null argument");
    CK(cudaMallocHost(hptr, bytes ? bytes : 16));
    return 0;
}
int rtb_host_free(void* hptr) { if (hptr) CK(cudaFreeHost(hptr)); return 0; }

int rtb_timer_start(rtb_ctx* c) {
    REQUIRE(c, "rtb_timer_start: null context");
    Activate act(c);
    CK(cudaEventRecord(c->ev0, c->stream));
    return 0;
}
int rtb_timer_stop_ms(rtb_ctx* c, float* ms) {
    REQUIRE(c && ms, "rtb_timer_stop_ms: bad argument");
    Activate act(c);
    CK(cudaEventRecord(c->ev1, c->stream));
    if (sync_and_check(c)) return 1;
    CK(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return 0;
}
int rtb_launch_count(rtb_ctx* c, uint64_t* count) { REQUIRE(c && count, "rtb_launch_count: bad argument"); *count = c->launches; return 0; }

// ---- S1 --------------------------------------------------------------------------------------------------------
int rtb_model_to_world(rtb_ctx* c, const rtb_ubo* ubo, const void* models, void* triangles, void* spheres) {
    REQUIRE(c && ubo && models, "rtb_model_to_world: bad argument");
    REQUIRE((ubo->numTriangles == 0 || triangles) && (ubo->numSpheres == 0 || spheres), "rtb_model_to_world: null primitive buffer");
    Activate act(c);
    launch_model_to_world(c->stream, models, triangles, ubo->numTriangles, spheres, ubo->numSpheres);
    return check_launch(c, 1, "model_to_world_kernel");
}

int rtb_enclosing_aabb(rtb_ctx* c, const rtb_ubo* ubo, void* enclosing, const void* triangles, const void* spheres, uint32_t flags) {
    REQUIRE(c && ubo && enclosing, "rtb_enclosing_aabb: bad argument");
    Activate act(c);
    if (ensure(c, c->encRed, 64)) return 1;
    const int n = launch_enclosing(c->stream, triangles, ubo->numTriangles, spheres, ubo->numSpheres, (uint32_t*)c->encRed.p, enclosing,
                                   (flags & RTB_TRACE_ENCLOSING_INF) ? 1 : 0, c->smCount);
    return check_launch(c, n, "enclosing_aabb kernels");
}

int rtb_morton_codes(rtb_ctx* c, const rtb_ubo* ubo, const void* enclosing, const void* triangles, const void* spheres, void* morton1) {
    REQUIRE(c && ubo && enclosing && morton1, "rtb_morton_codes: bad argument");
    REQUIRE(ubo->numTriangles + ubo->numSpheres > 0, "rtb_morton_codes: empty scene");
    Activate act(c);
    launch_morton(c->stream, triangles, ubo->numTriangles, spheres, ubo->numSpheres, enclosing, morton1, nullptr, nullptr);
    return check_launch(c, 1, "morton_kernel");
}

static int sort_scratch(rtb_ctx* c, uint32_t n) {
    for (int k = 0; k < 2; k++) {
        if (ensure(c, c->sortKeys[k], sizeof(uint32_t) * (size_t)n)) return 1;
        if (ensure(c, c->sortVals[k], sizeof(uint32_t) * (size_t)n)) return 1;
    }
    return ensure(c, c->sortCounts, radix_sort_counts_bytes(n));
}

int rtb_sort_morton(rtb_ctx* c, const rtb_ubo* ubo, void* morton1, void* morton2) {
    REQUIRE(c && ubo && morton1, "rtb_sort_morton: bad argument");
    (void)morton2;   // the reference ping-pongs through morton2; its final content is never read (RadixSortSimple.comp:146-151)
    const uint32_t T = ubo->numTriangles, n = T + ubo->numSpheres;
    if (n == 0) return 0;
    Activate act(c);
    if (sort_scratch(c, n)) return 1;
    uint32_t *k0 = (uint32_t*)c->sortKeys[0].p, *k1 = (uint32_t*)c->sortKeys[1].p;
    uint32_t *v0 = (uint32_t*)c->sortVals[0].p, *v1 = (uint32_t*)c->sortVals[1].p;
    launch_morton_unpack(c->stream, morton1, n, T, k0, v0);
    const int ls = launch_radix_sort(c->stream, k0, v0, k1, v1, n, (uint32_t*)c->sortCounts.p);
    launch_morton_repack(c->stream, k0, v0, n, T, morton1);
    return check_launch(c, ls + 2, "radix sort kernels");
}

int rtb_build_hlbvh(rtb_ctx* c, const rtb_ubo* ubo, const void* triangles, const void* spheres, const void* morton1, void* nodes, void* cinfo) {
    REQUIRE(c && ubo && morton1 && nodes && cinfo, "rtb_build_hlbvh: bad argument");
    REQUIRE(ubo->numTriangles + ubo->numSpheres > 0, "rtb_build_hlbvh: empty scene");
    Activate act(c);
    launch_hlbvh(c->stream, triangles, ubo->numTriangles, spheres, ubo->numSpheres, (const uint32_t*)morton1, 3, nodes, cinfo);
    return check_launch(c, 1, "hlbvh_kernel");
}

int rtb_refit_aabbs(rtb_ctx* c, const rtb_ubo* ubo, void* nodes, void* cinfo) {
    REQUIRE(c && ubo && nodes && cinfo, "rtb_refit_aabbs: bad argument");
    Activate act(c);
    launch_refit(c->stream, nodes, cinfo, ubo->numTriangles + ubo->numSpheres, nullptr, nullptr, nullptr);
    return check_launch(c, 1, "refit_kernel");
}

int rtb_build_bvh(rtb_ctx* c, const rtb_ubo* ubo, const void* models, void* triangles, void* spheres, const void* materials, void* enclosing,
                  void* morton1, void* morton2, void* nodes, void* cinfo, uint32_t flags) {
    REQUIRE(c && ubo && models && materials, "rtb_build_bvh: bad argument");
    (void)morton2;
    const uint32_t T = ubo->numTriangles, S = ubo->numSpheres, N = T + S;
    REQUIRE(N > 0, "rtb_build_bvh: empty scene");
    REQUIRE((T == 0 || triangles) && (S == 0 || spheres), "rtb_build_bvh: null primitive buffer");
    Activate act(c);
    if (!enclosing) { if (ensure(c, c->enclosing, 32)) return 1; enclosing = c->enclosing.p; }
    if (!cinfo) { if (ensure(c, c->cinfo, 8ull * (2ull * N - 1))) return 1; cinfo = c->cinfo.p; }
    if (!nodes) { if (ensure(c, c->nodes, 40ull * (2ull * N - 1))) return 1; nodes = c->nodes.p; }
    if (ensure(c, c->encRed, 64)) return 1;
    if (sort_scratch(c, N)) return 1;
    uint32_t *k0 = (uint32_t*)c->sortKeys[0].p, *k1 = (uint32_t*)c->sortKeys[1].p;
    uint32_t *v0 = (uint32_t*)c->sortVals[0].p, *v1 = (uint32_t*)c->sortVals[1].p;
    int launches = 0;
    if (N >= c->knobs.wideMin) {
        // Fused build for scenes that are walked over the 4-ary records: K1 + K2 in one pass, the K5 leaf pass also emits the exact leaf
        // boxes / packed primitives / per-primitive hit-point slack, the K6 climb also carries the slack and the child-pair records.
        // The reference-layout outputs (nodes, morton1, enclosing, cinfo) are the same bits as the stage-by-stage entry points produce.
        const size_t nn = 2ull * N - 1;
        if (ensure(c, c->pairs, sizeof(float4) * 4ull * (N - 1)) || ensure(c, c->rootBox, sizeof(float4) * 4) || ensure(c, c->ptris, sizeof(float4) * 4ull * T) ||
            ensure(c, c->psphs, sizeof(float4) * (size_t)S) || ensure(c, c->psphMat, sizeof(uint32_t) * (size_t)S) || ensure(c, c->leafBox, 32ull * N) ||
            ensure(c, c->wide, 64ull * (N - 1 + 16)) || ensure(c, c->etaNode, 4 * nn) || ensure(c, c->walkFlag, 64) ||
            ensure(c, c->bigList, 4ull * 64) || ensure(c, c->ttChild, 8ull * (N - 1)) || ensure(c, c->ttParent, 4 * nn) ||
            ensure(c, c->ttArrivals, 4ull * (N - 1)) || ensure(c, c->ttBox, 32ull * (N - 1)) || ensure(c, c->ttEta, 4ull * (N - 1))) return 1;
        launches += launch_model_to_world_enclosing(c->stream, models, triangles, T, spheres, S, (uint32_t*)c->encRed.p, enclosing,
                                                    (flags & RTB_TRACE_ENCLOSING_INF) ? 1 : 0);                   // K1 + K2
        launch_morton(c->stream, triangles, T, spheres, S, enclosing, nullptr, k0, v0); launches++;              // K3 (SoA out)
        launches += launch_radix_sort(c->stream, k0, v0, k1, v1, N, (uint32_t*)c->sortCounts.p);                  // K4
        if (morton1) { launch_morton_repack(c->stream, k0, v0, N, T, morton1); launches++; }
        launch_hlbvh_fused(c->stream, triangles, T, spheres, S, k0, nodes, cinfo, c->leafBox.p, c->ptris.p, c->psphs.p, c->psphMat.p,
                           (float*)c->etaNode.p, (const uint32_t*)c->encRed.p + 6, (float4*)c->rootBox.p + 2, ubo->camPos,
                           (unsigned int*)c->walkFlag.p + 2, (uint32_t*)c->bigList.p, (uint32_t*)c->walkFlag.p + 8); launches++;   // K5
        // K6 (+ pair records, slack) on the build stream: the climb is a chain of dependent atomics that leaves the GPU nearly idle, and
        // the first half of the traversal hierarchy's build (keys, sort, topology, its own climb) needs only what K5 wrote
        cudaEventRecord(c->evBuildFork, c->stream);
        cudaStreamWaitEvent(c->buildStream, c->evBuildFork, 0);
        launch_refit(c->buildStream, nodes, cinfo, N, c->pairs.p, c->rootBox.p, (float*)c->etaNode.p); launches++;
        cudaEventRecord(c->evBuildJoin, c->buildStream);
        if (check_launch(c, launches, "BVH build kernels")) return 1;
        if (bind_internal(c, T, S, ubo->numMaterials, triangles, spheres, materials, nodes, /*pairsDone=*/true, /*primsDone=*/true)) return 1;
        // the hierarchy the order-free walk descends (traversal_tree.cu): its own, better tree over the same leaves, big leaves in front of the root
        const TraversalTreeBuffers tb{ k0, v0, k1, v1, (uint32_t*)c->sortCounts.p, (uint2*)c->ttChild.p, (uint32_t*)c->ttParent.p, (unsigned int*)c->ttArrivals.p,
                                       (float4*)c->ttBox.p, (float*)c->ttEta.p, (unsigned int*)c->walkFlag.p, (const uint32_t*)c->bigList.p,
                                       (const uint32_t*)c->encRed.p + 6, (const uint32_t*)c->walkFlag.p + 8 };
        int packs = launch_traversal_tree(c->stream, N, c->leafBox.p, (const float*)c->etaNode.p, tb, c->wide.p, c->evBuildJoin);
        c->leafBoxReady = true; c->wideReady = true; c->hoisted = true;
#ifdef RTB_SMEM_TOP
        if (ensure(c, c->topTable, 64ull * RTB_SMEM_TOP) || ensure(c, c->topGlobal, 4ull * RTB_SMEM_TOP)) return 1;
        launch_build_top_table(c->stream, c->wide.p, N, c->topTable.p, c->topGlobal.p, (const unsigned int*)c->walkFlag.p); packs++;
#endif
        return check_launch(c, packs, "pack_wide_kernel");
    }
    launch_model_to_world(c->stream, models, triangles, T, spheres, S); launches++;                             // K1
    launches += launch_enclosing(c->stream, triangles, T, spheres, S, (uint32_t*)c->encRed.p, enclosing,
                                 (flags & RTB_TRACE_ENCLOSING_INF) ? 1 : 0, c->smCount);                          // K2
    launch_morton(c->stream, triangles, T, spheres, S, enclosing, nullptr, k0, v0); launches++;                  // K3 (SoA out)
    launches += launch_radix_sort(c->stream, k0, v0, k1, v1, N, (uint32_t*)c->sortCounts.p);                      // K4
    if (morton1) { launch_morton_repack(c->stream, k0, v0, N, T, morton1); launches++; }
    launch_hlbvh(c->stream, triangles, T, spheres, S, k0, 1, nodes, cinfo); launches++;                           // K5
    if (ensure(c, c->pairs, sizeof(float4) * 4ull * (N > 1 ? N - 1 : 1))) return 1;
    if (ensure(c, c->rootBox, sizeof(float4) * 4)) return 1;
    launch_refit(c->stream, nodes, cinfo, N, N > 1 ? c->pairs.p : nullptr, c->rootBox.p, nullptr); launches++;   // K6 (+ pair records)
    if (check_launch(c, launches, "BVH build kernels")) return 1;
    return bind_internal(c, T, S, ubo->numMaterials, triangles, spheres, materials, nodes, /*pairsDone=*/N > 1);
}

// ---- S2 --------------------------------------------------------------------------------------------------------
int rtb_clear_image(rtb_ctx* c, void* image, uint32_t width, uint32_t rows) {
    REQUIRE(c && image, "rtb_clear_image: bad argument");
    Activate act(c);
    launch_clear_image(c->stream, image, (size_t)width * rows, c->smCount);
    return check_launch(c, 1, "clear_image_kernel");
}

int rtb_bind_trace_buffers(rtb_ctx* c, const rtb_ubo* ubo, const void* triangles, const void* spheres, const void* materials, const void* nodes) {
    REQUIRE(c && ubo && materials, "rtb_bind_trace_buffers: bad argument");   // nodes == NULL: the non-BVH program's set
    REQUIRE(ubo->numTriangles + ubo->numSpheres > 0, "rtb_bind_trace_buffers: empty scene");
    Activate act(c);
    return bind_internal(c, ubo->numTriangles, ubo->numSpheres, ubo->numMaterials, triangles, spheres, materials, nodes);
}

int rtb_raytrace(rtb_ctx* c, const rtb_ubo* ubo, void* image, const rtb_trace_args* a) {
    REQUIRE(c && ubo && image && a, "rtb_raytrace: bad argument");
    REQUIRE(c->bound, "rtb_raytrace: no buffers bound (call rtb_build_bvh or rtb_bind_trace_buffers first)");
    REQUIRE(c->boundNodes || (a->flags & RTB_TRACE_LINEAR_SCAN), "rtb_raytrace: no BVH nodes bound; only RTB_TRACE_LINEAR_SCAN can run");
    REQUIRE(ubo->numTriangles == c->bT && ubo->numSpheres == c->bS, "rtb_raytrace: UBO primitive counts differ from the bound set");
    REQUIRE(a->imageWidth && a->imageHeight && a->bandRows && a->bandStep, "rtb_raytrace: bad image / band description");
    REQUIRE(!(a->flags & RTB_TRACE_COUNT) || a->counters, "rtb_raytrace: RTB_TRACE_COUNT needs a counters buffer");
    REQUIRE(!(a->flags & RTB_TRACE_WALK_COUNT) || a->walkCounters, "rtb_raytrace: RTB_TRACE_WALK_COUNT needs a walkCounters buffer");
    REQUIRE(!((a->flags & RTB_TRACE_WALK_COUNT) && (a->flags & (RTB_TRACE_COUNT | RTB_TRACE_LINEAR_SCAN | RTB_TRACE_SIMPLE_KERNEL | RTB_TRACE_STREAM_KERNEL | RTB_TRACE_CULLED))),
            "rtb_raytrace: RTB_TRACE_WALK_COUNT instruments the production walk only (no COUNT / LINEAR_SCAN / SIMPLE / STREAM / CULLED)");
    // the wave kernels pack a pixel as x | y << 16 (activeXY)
    REQUIRE(a->imageWidth <= 65536u && a->imageHeight <= 65535u, "rtb_raytrace: image larger than 65536 x 65535");
    if (a->sampleCount == 0 || a->localRows == 0) return 0;
    Activate act(c);
    TraceParams p;
    memset(&p, 0, sizeof(p));
    p.sc.pairs = (const float4*)c->pairs.p;
    p.sc.tris = (const float4*)c->ptris.p;
    p.sc.sphs = (const float4*)c->psphs.p;
    p.sc.sphMat = (const uint32_t*)c->psphMat.p;
    p.sc.mats = (const float4*)c->pmats.p;
    p.sc.rootBox = (const float4*)c->rootBox.p;
    p.sc.cnodes = nullptr;
    p.sc.leafBox = nullptr;
    p.sc.wide = nullptr;
    p.sc.T = c->bT; p.sc.S = c->bS; p.sc.N = c->bN;
    p.cam = make_camera(ubo, a->imageWidth, a->imageHeight);
    p.image = (float4*)image;
    p.W = a->imageWidth; p.H = a->imageHeight; p.localRows = a->localRows;
    p.bandRows = a->bandRows; p.bandFirst = a->bandFirst; p.bandStep = a->bandStep;
    p.sampleSkip = a->sampleSkip; p.sampleCount = a->sampleCount;
    p.maxDepth = ubo->maxRayTraceDepth; p.randomState = ubo->randomState;
    p.hitPrim = (uint32_t*)a->hitPrim; p.hitT = (float*)a->hitT; p.rngOut = (uint32_t*)a->rngOut;
    p.workCounter = (unsigned int*)c->workCounter.p;
    p.errFlag = (unsigned int*)c->errFlag.p;
    p.tMin = c->knobs.tMin;
    // farthest-first stacking of the waiting entries: +16 % on C3 over the reference-tree records in round 1, but -1.6 % there (and -3 % on
    // C2 / C4) once the big leaves were hoisted (profiles/r02_knob_sweep_after_hoisting.txt): off unless RTB_WAVE_SORTED_PUSH=1
    p.sortedPush = c->knobs.sortedPush > 0 ? 1u : 0u;
    p.sMin = c->knobs.sMin ? c->knobs.sMin : 1u;
    p.mainCtas = c->knobs.mainCtas; p.tailThreads = c->knobs.tailThreads;
    p.qGate = c->knobs.qGate; p.coopMax = c->knobs.coopMax; p.coopTurns = c->knobs.coopTurns; p.tailSpinUs = c->knobs.tailSpinUs;
    const bool walk = (a->flags & RTB_TRACE_WALK_COUNT) != 0;
    const bool count = (a->flags & RTB_TRACE_COUNT) != 0 || walk, ext = (a->flags & RTB_TRACE_EXT_MATERIALS) != 0;
    p.counters = walk ? nullptr : (unsigned long long*)a->counters;
    p.walkCounters = walk ? (unsigned long long*)a->walkCounters : nullptr;
    int launches = 1;
    const bool linear = (a->flags & RTB_TRACE_LINEAR_SCAN) != 0;
    if (linear || (a->flags & RTB_TRACE_SIMPLE_KERNEL)) {
        launch_trace(c->stream, p, count, ext, linear, c->smCount);
    } else {
        // per-(sample, pixel) colour slots: as many samples per pass as fit the scratch budget (default 4 GiB)
        const size_t pixels = (size_t)a->imageWidth * a->localRows;
        const size_t budget = c->knobs.sampleBufBytes;
        size_t perPass = budget / (pixels * sizeof(float4));
        if (perPass < 1) perPass = 1;
        if (perPass > a->sampleCount) perPass = a->sampleCount;
        if (ensure(c, c->activePix, pixels * sizeof(uint32_t)) || ensure(c, c->activeXY, pixels * sizeof(uint32_t))) return 1;
        if (ensure(c, c->sampleBuf, pixels * perPass * sizeof(float4))) return 1;
        p.workCounter64 = (unsigned long long*)c->workCounter.p;
        p.activeCount = (unsigned int*)c->workCounter.p + 2;
        p.activePix = (uint32_t*)c->activePix.p;
        p.activeXY = (uint32_t*)c->activeXY.p;
        p.sampleBuf = (float4*)c->sampleBuf.p;
        p.slotCapacity = (uint32_t)pixels;
        const bool cull = (a->flags & RTB_TRACE_CULLED) != 0;
#ifndef RTB_AB_KERNELS
        REQUIRE(!(a->flags & (RTB_TRACE_STREAM_KERNEL | RTB_TRACE_COMPRESSED_NODES)),
                "rtb_raytrace: the streaming kernel and the 32-byte compressed records are A/B variants (measured, not adopted): rebuild with make AB=1");
        {
#else
        if ((a->flags & RTB_TRACE_STREAM_KERNEL) && !cull) {
            // pool of in-flight paths: ~18 paths per resident lane keeps the per-iteration tails short
            size_t cap = (size_t)c->smCount * 128 * 7 * 18;
            if (c->knobs.streamPool) cap = c->knobs.streamPool;
            if (cap > pixels * perPass) cap = pixels * perPass;
            if (cap < 1024) cap = 1024;
            if (ensure(c, c->poolSlot, cap * 4) || ensure(c, c->poolColor, cap * 16) || ensure(c, c->poolAtt, cap * 16) ||
                ensure(c, c->poolOrg, cap * 16) || ensure(c, c->poolDir, cap * 16) || ensure(c, c->poolNrm, cap * 16) ||
                ensure(c, c->poolList, cap * 4) || ensure(c, c->poolCnt, 16)) return 1;
            p.pool.slot = (uint32_t*)c->poolSlot.p; p.pool.colorRng = (float4*)c->poolColor.p; p.pool.attDepth = (float4*)c->poolAtt.p;
            p.pool.org = (float4*)c->poolOrg.p; p.pool.dir = (float4*)c->poolDir.p; p.pool.nrm = (float4*)c->poolNrm.p;
            p.pool.rayList = (uint32_t*)c->poolList.p; p.pool.cnt = (unsigned int*)c->poolCnt.p; p.pool.capacity = (uint32_t)cap;
            launches = launch_trace_stream(c->stream, p, count, ext, c->smCount, (uint32_t)perPass);
        } else {
#endif
            const bool derived = c->boundNodes && c->bN > 1 && (!count || walk);
            int nodesMode = !derived ? 0 : (a->flags & RTB_TRACE_EXACT_NODES) ? 0 : (a->flags & RTB_TRACE_WIDE_NODES) ? 2
                                : (a->flags & RTB_TRACE_COMPRESSED_NODES) ? 1 : (c->bN >= c->knobs.wideMin ? 2 : 0);
            int extra = derive_records(c, nodesMode, ubo->camPos);
            if (extra < 0) return 1;
            if (nodesMode) { p.sc.cnodes = nodesMode == 1 ? (const uint4*)c->cnodes.p : nullptr; p.sc.leafBox = (const float4*)c->leafBox.p; }
            if (nodesMode == 2) p.sc.wide = (const uint4*)c->wide.p;
            const bool orderFree = nodesMode == 2 && !cull && !(a->flags & RTB_TRACE_REFERENCE_ORDER);
            if (nodesMode == 2 && !orderFree && c->hoisted) {
                // the reference-order walk (and the segment-box extension on top of it) needs the records in the reference's visiting
                // order with every leaf in its place: an un-hoisted set, derived on first use
                if (!c->wideRefReady) {
                    if (ensure(c, c->wideRef, 64ull * (c->bN - 1)) || ensure(c, c->walkFlag, 64)) return 1;
                    launch_pack_wide(c->stream, c->boundNodesPtr, c->bN, c->wideRef.p, (const float*)c->etaNode.p, (unsigned int*)c->walkFlag.p + 4);
                    c->wideRefReady = true; extra++;
                }
                p.sc.wide = (const uint4*)c->wideRef.p;
            }
            p.sc.top = (const uint4*)c->topTable.p; p.sc.topGlobal = (const uint32_t*)c->topGlobal.p;
            if (orderFree) nodesMode = 3;
            p.cullAllowed = (const unsigned int*)c->walkFlag.p;
            p.primaryHits = nullptr; p.primaryMode = 0;
            if ((!count || walk) && a->sampleCount > 1 && !(a->flags & RTB_TRACE_NO_PRIMARY_SHARING)) {
                if (ensure(c, c->primaryHits, pixels * 3 * sizeof(float4))) return 1;
                p.primaryHits = (float4*)c->primaryHits.p;
            }
            if (nodesMode == 3 && (p.coopMax || p.coopTurns)) {      // parking is optional for a lane: a full buffer just means "walk it yourself"
                const size_t cap = (size_t)c->smCount * 8 * 128;
                if (ensure(c, c->parkBuf, cap * 16 * sizeof(float4))) return 1;
                p.parkBuf = (float4*)c->parkBuf.p; p.parkCapacity = (uint32_t)cap;
                p.parkEpoch = c->parkEpoch;
                c->parkEpoch += 2u * (uint32_t)((a->sampleCount + perPass - 1) / perPass);      // two parking launches per pass, one epoch each
                p.parkCount = (unsigned int*)c->workCounter.p + 4; p.parkCursor = (unsigned int*)c->workCounter.p + 5;
                p.doneWarps = (unsigned int*)c->workCounter.p + 6;
            } else { p.coopMax = 0; p.coopTurns = 0; p.doneWarps = nullptr; }
            const TailOverlap ov{ c->auxStream, c->evFork, c->evJoin };
            launches = extra + launch_trace_wave(c->stream, p, count, ext, cull, nodesMode, c->smCount, (uint32_t)perPass, ov);
        }
    }
    c->traced = true;
    return check_launch(c, launches, "trace kernel");
}

int rtb_export_hit_slack(rtb_ctx* c, float* hostEta, size_t count, float* hostOriginRegion6) {
    REQUIRE(c && hostEta && hostOriginRegion6, "rtb_export_hit_slack: bad argument");
    REQUIRE(c->bound && c->wideReady && c->bN > 1, "rtb_export_hit_slack: the bound scene has no 4-ary records (fewer than 8192 primitives and never traced with RTB_TRACE_WIDE_NODES)");
    REQUIRE(count == c->bN, "rtb_export_hit_slack: count must equal the number of primitives");
    Activate act(c);
    float4 region[2];
    CK(cudaMemcpyAsync(hostEta, (const float*)c->etaNode.p + (c->bN - 1), sizeof(float) * count, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(region, (const float4*)c->rootBox.p + 2, sizeof(region), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    hostOriginRegion6[0] = region[0].x; hostOriginRegion6[1] = region[0].y; hostOriginRegion6[2] = region[0].z;
    hostOriginRegion6[3] = region[1].x; hostOriginRegion6[4] = region[1].y; hostOriginRegion6[5] = region[1].z;
    return 0;
}

int rtb_logistic_step(rtb_ctx* c, void* points, uint32_t count, void* imageRgba8, uint32_t width, uint32_t height, const float* pixelColor) {
    REQUIRE(c && points && imageRgba8 && pixelColor && width && height, "rtb_logistic_step: bad argument");
    Activate act(c);
    launch_logistic(c->stream, points, count, imageRgba8, width, height, pixelColor);
    return check_launch(c, count ? 1 : 0, "logistic_kernel");
}

int rtb_resolve_rgba8(rtb_ctx* c, const void* image, uint32_t width, uint32_t rows, uint32_t raysPerPixel, void* out) {
    REQUIRE(c && image && out && raysPerPixel, "rtb_resolve_rgba8: bad argument");
    Activate act(c);
    launch_resolve(c->stream, image, (size_t)width * rows, raysPerPixel, out, c->smCount);
    return check_launch(c, 1, "resolve_kernel");
}

}  // extern "C"
