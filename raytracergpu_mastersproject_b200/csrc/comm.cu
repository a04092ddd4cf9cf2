// comm.cu -- the multi-GPU entry points of the C-ABI (include/rtb200.h): one context per GPU, scene
// replicated, the frame partitioned by 8-row bands (tile mode) or by sample range; ONE exchange at the end of a frame, issued on
// the context's stream right behind the kernels that produced the data, so nothing waits on the host.
//
// NCCL is the transport (NVLink 5 / NVSwitch on a B200 box).  It is bound at run time (dlopen of libnccl.so.2 -- the copy a host
// process such as PyTorch has already loaded is reused, so both see the same library) so that librtb200.so itself has no hard
// NCCL dependency: single-GPU hosts, like the reference's own, never touch it.  The few prototypes used are declared here with
// the values of nccl.h (NCCL 2.x ABI: ncclUniqueId = 128 bytes, ncclUint8 = 1, ncclFloat32 = 7, ncclSum = 0).
//
// What is fused around the collective (kernels of this file):
//   tile mode    : the RGBA32F bands of every rank are all-gathered, and ONE kernel re-assembles the interleaved bands into the
//                  frame and applies the fragment-shader resolve (SingleTriangleFullScreen.frag:13-21) in the same pass.  When only
//                  the presented RGBA8 frame is wanted, the resolve runs BEFORE the exchange on the rank's own bands: 4 bytes
//                  per pixel cross NVLink instead of 16.
//   sample ranges: partial sums are reduced to the root (fp32 sum), the alpha channel (end of the seed chain,
//                  raytraceBVH.comp:372) is taken from the last rank, and the root resolves.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "ctx.h"

using namespace rtb;

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;      // NCCL_UNIQUE_ID_BYTES
enum { NCCL_SUCCESS = 0, NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_SUM = 0 };

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = { getenv("RTB_NCCL_LIBRARY"), "libnccl.so.2", "libnccl.so" };
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) { n.why = "libnccl.so.2 not found (set RTB_NCCL_LIBRARY)"; return; }
#define RTB_SYM(field, name) *(void**)(&n.field) = dlsym(n.handle, name); if (!n.field) { n.why = std::string("NCCL symbol missing: ") + name; n.handle = nullptr; return; }
        RTB_SYM(GetUniqueId, "ncclGetUniqueId") RTB_SYM(CommInitRank, "ncclCommInitRank") RTB_SYM(CommDestroy, "ncclCommDestroy")
        RTB_SYM(AllGather, "ncclAllGather") RTB_SYM(Reduce, "ncclReduce") RTB_SYM(Broadcast, "ncclBroadcast")
        RTB_SYM(GetErrorString, "ncclGetErrorString")
#undef RTB_SYM
    });
    return n;
}

int nccl_fail(const char* what, int rc) {
    std::string m = what;
    m += ": ";
    m += nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error";
    return fail(m.c_str());
}
#define NC(call) do { int _r = (call); if (_r != NCCL_SUCCESS) return nccl_fail(#call, _r); } while (0)
#define NEED_NCCL() do { if (!nccl().handle) return fail(("multi-GPU entry point: " + nccl().why).c_str()); } while (0)
#define NEED_COMM(c) do { REQUIRE((c) && (c)->comm, "multi-GPU entry point: the context has no communicator (rtb_comm_init_rank)"); } while (0)

// fragment-shader resolve of one pixel (SingleTriangleFullScreen.frag:13-21), as resolve_kernel in trace.cu
__device__ __forceinline__ uchar4 resolve_px(const float4 c, const float rpp) {
    float r = sqrtf(c.x / rpp), g = sqrtf(c.y / rpp), b = sqrtf(c.z / rpp);
    r = (r > 0.f) ? r : 0.f; r = (r < 1.f) ? r : 1.f;
    g = (g > 0.f) ? g : 0.f; g = (g < 1.f) ? g : 1.f;
    b = (b > 0.f) ? b : 0.f; b = (b < 1.f) ? b : 1.f;
    return make_uchar4((unsigned char)(r * 255.0f + 0.5f), (unsigned char)(g * 255.0f + 0.5f), (unsigned char)(b * 255.0f + 0.5f), 255);
}

// gathered[rank][localRows][W] -> frame[H][W] (+ resolved RGBA8), one pass.  Global row y lives in band y / bandRows, which
// belongs to rank band % n as its local band band / n (rtb_trace_args: bandFirst = rank, bandStep = n).
__global__ void __launch_bounds__(256) assemble_f32_kernel(const float4* __restrict__ gathered, uint32_t W, uint32_t H, uint32_t bandRows,
                                                           uint32_t n, uint32_t localRows, float4* frame, uchar4* rgba8, float rpp) {
    const size_t total = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (size_t)y * W);
        const uint32_t band = y / bandRows, rank = band % n, j = (band / n) * bandRows + y % bandRows;
        const float4 c = gathered[((size_t)rank * localRows + j) * W + x];
        if (frame) frame[i] = c;
        if (rgba8) rgba8[i] = resolve_px(c, rpp);
    }
}
__global__ void __launch_bounds__(256) assemble_u8_kernel(const uchar4* __restrict__ gathered, uint32_t W, uint32_t H, uint32_t bandRows,
                                                          uint32_t n, uint32_t localRows, uchar4* rgba8) {
    const size_t total = (size_t)W * H;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / W), x = (uint32_t)(i - (size_t)y * W);
        const uint32_t band = y / bandRows, rank = band % n, j = (band / n) * bandRows + y % bandRows;
        rgba8[i] = gathered[((size_t)rank * localRows + j) * W + x];
    }
}
// this rank's bands, resolved in place of the exchange buffer slot it owns
__global__ void __launch_bounds__(256) resolve_bands_kernel(const float4* __restrict__ local, size_t pixels, float rpp, uchar4* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x)
        out[i] = resolve_px(local[i], rpp);
}
// sample-range mode: the reduced alpha must be the LAST rank's (the end of the seed chain); every other rank contributes 0
__global__ void __launch_bounds__(256) zero_alpha_kernel(float4* img, size_t pixels) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x) img[i].w = 0.0f;
}
__global__ void __launch_bounds__(256) resolve_frame_kernel(const float4* __restrict__ img, size_t pixels, float rpp, uchar4* out) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (size_t)gridDim.x * blockDim.x)
        out[i] = resolve_px(img[i], rpp);
}

unsigned grid_for(const rtb_ctx* c, size_t n) {
    size_t g = (n + 255) / 256;
    const size_t cap = (size_t)c->smCount * 16;
    if (g > cap) g = cap;
    return (unsigned)(g ? g : 1);
}

}  // namespace

extern "C" {

int rtb_comm_unique_id(void* id128) {
    REQUIRE(id128, "rtb_comm_unique_id: null argument");
    NEED_NCCL();
    ncclUniqueId id;
    NC(nccl().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
    return 0;
}

int rtb_comm_init_rank(rtb_ctx* c, int nRanks, int rank, const void* id128) {
    REQUIRE(c && id128 && nRanks >= 1 && rank >= 0 && rank < nRanks, "rtb_comm_init_rank: bad argument");
    REQUIRE(!c->comm, "rtb_comm_init_rank: the context already has a communicator");
    NEED_NCCL();
    Activate act(c);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    NC(nccl().CommInitRank(&comm, nRanks, id, rank));
    c->comm = comm; c->commRank = rank; c->commSize = nRanks;
    return 0;
}

int rtb_comm_destroy(rtb_ctx* c) {
    REQUIRE(c, "rtb_comm_destroy: null context");
    if (!c->comm) return 0;
    Activate act(c);
    cudaStreamSynchronize(c->stream);
    const int r = nccl().CommDestroy((ncclComm_t)c->comm);
    c->comm = nullptr; c->commRank = 0; c->commSize = 1;
    if (r != NCCL_SUCCESS) return nccl_fail("ncclCommDestroy", r);
    return 0;
}

int rtb_comm_info(rtb_ctx* c, int* rank, int* nRanks) {
    REQUIRE(c && rank && nRanks, "rtb_comm_info: bad argument");
    *rank = c->commRank; *nRanks = c->comm ? c->commSize : 1;
    return 0;
}

int rtb_comm_all_gather(rtb_ctx* c, void* buffer, size_t bytesPerRank) {
    NEED_COMM(c);
    REQUIRE(buffer || bytesPerRank == 0, "rtb_comm_all_gather: null buffer");
    if (!bytesPerRank) return 0;
    Activate act(c);
    NC(nccl().AllGather((const char*)buffer + (size_t)c->commRank * bytesPerRank, buffer, bytesPerRank, NCCL_UINT8, (ncclComm_t)c->comm, c->stream));
    c->launches++;
    return 0;
}

int rtb_comm_broadcast(rtb_ctx* c, void* buffer, size_t bytes, int root) {
    NEED_COMM(c);
    REQUIRE((buffer || bytes == 0) && root >= 0 && root < c->commSize, "rtb_comm_broadcast: bad argument");
    if (!bytes) return 0;
    Activate act(c);
    NC(nccl().Broadcast(buffer, buffer, bytes, NCCL_UINT8, root, (ncclComm_t)c->comm, c->stream));
    c->launches++;
    return 0;
}

int rtb_gather_tiles(rtb_ctx* c, const void* localImage, uint32_t width, uint32_t height, uint32_t bandRows, void* frameRgba32f,
                     uint32_t raysPerPixel, void* frameRgba8) {
    NEED_COMM(c);
    REQUIRE(localImage && width && height && bandRows, "rtb_gather_tiles: bad argument");
    REQUIRE(frameRgba32f || frameRgba8, "rtb_gather_tiles: no output buffer");
    REQUIRE(!frameRgba8 || raysPerPixel, "rtb_gather_tiles: raysPerPixel is 0");
    Activate act(c);
    const uint32_t n = (uint32_t)c->commSize;
    const uint32_t bands = (height + bandRows - 1) / bandRows, localRows = ((bands + n - 1) / n) * bandRows;
    const size_t localPx = (size_t)localRows * width, framePx = (size_t)width * height;
    int launches = 0;
    if (frameRgba32f) {          // the accumulation image itself is wanted: exchange RGBA32F, re-assemble (+ resolve) in one pass
        if (ensure(c, c->gatherBuf, localPx * n * sizeof(float4))) return 1;
        NC(nccl().AllGather(localImage, c->gatherBuf.p, localPx * 4, NCCL_FLOAT32, (ncclComm_t)c->comm, c->stream));
        assemble_f32_kernel<<<grid_for(c, framePx), 256, 0, c->stream>>>((const float4*)c->gatherBuf.p, width, height, bandRows, n, localRows,
                                                                         (float4*)frameRgba32f, (uchar4*)frameRgba8, (float)raysPerPixel);
        launches = 2;
    } else {                     // presented frame only: resolve the own bands first, 4 bytes per pixel cross NVLink
        if (ensure(c, c->gatherBuf, localPx * n * sizeof(uchar4))) return 1;
        uchar4* mine = (uchar4*)c->gatherBuf.p + (size_t)c->commRank * localPx;
        resolve_bands_kernel<<<grid_for(c, localPx), 256, 0, c->stream>>>((const float4*)localImage, localPx, (float)raysPerPixel, mine);
        NC(nccl().AllGather(mine, c->gatherBuf.p, localPx * 4, NCCL_UINT8, (ncclComm_t)c->comm, c->stream));
        assemble_u8_kernel<<<grid_for(c, framePx), 256, 0, c->stream>>>((const uchar4*)c->gatherBuf.p, width, height, bandRows, n, localRows,
                                                                        (uchar4*)frameRgba8);
        launches = 3;
    }
    return check_launch(c, launches, "rtb_gather_tiles kernels");
}

int rtb_reduce_samples(rtb_ctx* c, void* image, uint32_t width, uint32_t height, int root, uint32_t raysPerPixel, void* frameRgba8) {
    NEED_COMM(c);
    REQUIRE(image && width && height && root >= 0 && root < c->commSize, "rtb_reduce_samples: bad argument");
    REQUIRE(!frameRgba8 || raysPerPixel, "rtb_reduce_samples: raysPerPixel is 0");
    Activate act(c);
    const size_t px = (size_t)width * height;
    int launches = 1;
    if (c->commRank != c->commSize - 1) { zero_alpha_kernel<<<grid_for(c, px), 256, 0, c->stream>>>((float4*)image, px); launches++; }
    NC(nccl().Reduce(image, image, px * 4, NCCL_FLOAT32, NCCL_SUM, root, (ncclComm_t)c->comm, c->stream));
    if (c->commRank == root && frameRgba8) {
        resolve_frame_kernel<<<grid_for(c, px), 256, 0, c->stream>>>((const float4*)image, px, (float)raysPerPixel, (uchar4*)frameRgba8);
        launches++;
    }
    return check_launch(c, launches, "rtb_reduce_samples kernels");
}

}  // extern "C"
