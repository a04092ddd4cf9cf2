// tests/emu/cuda_runtime.h -- SIMT EMULATION OF THE CUDA KERNELS ON THE CPU.  TEST INFRASTRUCTURE ONLY.
//
// tests/emu/build_emu.py compiles the UNMODIFIED kernel sources of raytracergpu_mastersproject_b200/csrc/*.cu with g++ against this
// header (found first on the include path, so `#include <cuda_runtime.h>` resolves here; the `<<<...>>>` launches are rewritten
// to emu::launch by the build script).  Every CUDA thread becomes a fiber (ucontext); the fibers of one thread block run
// interleaved on one OS thread and meet at the warp collectives (__ballot_sync, __shfl_sync, ...) and at __syncthreads, so
// the kernels' warp-level control flow -- the thing a plain host port would not exercise -- is executed as written.
// Thread blocks run one after the other (the emulated device has one multiprocessor, so a cooperative launch is one block).  The result is a library with the C-ABI of librtb200.so that executes the same
// kernel code on the CPU, used ONLY by tests/test_emulated_kernels.py (the CPU tier's check of the kernel sources against
// the oracle) and as a development aid when no GPU is at hand.  It is built into a scratch directory, is never loaded by the
// product (raytracergpu_mastersproject_b200/capi.py loads librtb200.so, which has no CPU path), never travels to the GPU box,
// and says "SIMT-EMU" where the product says "NVIDIA B200".
#pragma once
#define RTB_SIMT_EMU 1

#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <functional>

// ---- qualifiers -----------------------------------------------------------------------------------------------------
#define __host__
#define __device__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static        /* one thread block at a time: a function-local static is the block's shared memory */

// ---- vector types ---------------------------------------------------------------------------------------------------
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct uchar4 { unsigned char x, y, z, w; };
static inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return uchar4{ x, y, z, w }; }
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static inline float2 make_float2(float x, float y) { return float2{ x, y }; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{ x, y }; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// ---- runtime API (memory is host memory, the stream is the calling thread) -------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "SIMT-EMU (CPU emulation of the sm_100a kernels, test infrastructure)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 1;      // one SM: a cooperative launch is one block (cooperative_groups.h)
    return cudaSuccess;
}
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc((void**)p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMallocHost(void** p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreate(cudaStream_t* s) { *s = (cudaStream_t)(uintptr_t)1; return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { return cudaStreamCreate(s); }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = malloc(sizeof(double)); return cudaSuccess; }
enum { cudaEventDisableTiming = 2 };
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }   // launches are synchronous here
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr);
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)((*(double*)b - *(double*)a) * 1e3); return cudaSuccess; }
template <class K> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* nb, K, int, size_t) { *nb = 1; return cudaSuccess; }

// ---- the SIMT machine -------------------------------------------------------------------------------------------------
namespace emu {
struct ThreadInfo { uint3 tIdx, bIdx; dim3 bDim, gDim; unsigned lane, warp; };
ThreadInfo& cur();
void launch(dim3 grid, dim3 block, const std::function<void()>& body);
// a warp collective: every lane of `mask` deposits `mine`; when all have arrived, out(lane, values, mask) is evaluated for every lane
typedef unsigned long long (*CollectiveFn)(unsigned lane, const unsigned long long* vals, const unsigned* aux, unsigned mask);
unsigned long long collective(unsigned mask, unsigned long long mine, unsigned aux, CollectiveFn out);
void syncthreads();
}  // namespace emu
#define threadIdx (emu::cur().tIdx)
#define blockIdx (emu::cur().bIdx)
#define blockDim (emu::cur().bDim)
#define gridDim (emu::cur().gDim)

// ---- bit casts, integer intrinsics -------------------------------------------------------------------------------------
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __clz(int v) { return __clz((unsigned)v); }
static inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
static inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
static inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
static inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
static inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
static inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
static inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
static inline unsigned max(unsigned a, int b) { return max(a, (unsigned)b); }
static inline unsigned max(int a, unsigned b) { return max((unsigned)a, b); }

// ---- float intrinsics (IEEE binary32; compile with -ffp-contract=off -frounding-math) ------------------------------------
static inline float __uint2float_rn(unsigned u) { return (float)u; }
static inline unsigned __float2uint_rz(float f) { if (!(f > 0.0f)) return 0u; if (f >= 4294967296.0f) return 0xFFFFFFFFu; return (unsigned)f; }
static inline int __float2int_rz(float f) { if (f != f) return 0; if (f >= 2147483648.0f) return 2147483647; if (f <= -2147483648.0f) return (int)0x80000000; return (int)f; }
#define RTB_EMU_DIRECTED(name, mode, expr) \
    static inline float name(float a, float b) { volatile float x = a, y = b; const int old = fegetround(); fesetround(mode); volatile float r = expr; fesetround(old); return r; }
RTB_EMU_DIRECTED(__fadd_ru, FE_UPWARD, x + y)
RTB_EMU_DIRECTED(__fadd_rd, FE_DOWNWARD, x + y)
RTB_EMU_DIRECTED(__fsub_ru, FE_UPWARD, x - y)
RTB_EMU_DIRECTED(__fsub_rd, FE_DOWNWARD, x - y)
RTB_EMU_DIRECTED(__fmul_ru, FE_UPWARD, x * y)
RTB_EMU_DIRECTED(__fmul_rd, FE_DOWNWARD, x * y)
RTB_EMU_DIRECTED(__fdiv_ru, FE_UPWARD, x / y)
RTB_EMU_DIRECTED(__fdiv_rd, FE_DOWNWARD, x / y)

// ---- memory ---------------------------------------------------------------------------------------------------------
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
template <class T> static inline void __stcg(T* p, T v) { *p = v; }
static inline void __threadfence() {}
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned* p, int v) { unsigned o = *p; *p = o + (unsigned)v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline unsigned atomicExch(unsigned* p, unsigned v) { unsigned o = *p; *p = v; return o; }
static inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { unsigned o = *p; if (o == cmp) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }

// ---- warp / block collectives ------------------------------------------------------------------------------------------
static inline void __syncthreads() { emu::syncthreads(); }
static inline unsigned __activemask() { return 1u << emu::cur().lane; }      // any subset containing the caller is a legal answer
namespace emu {
static inline unsigned long long c_ballot(unsigned, const unsigned long long* v, const unsigned*, unsigned mask) {
    unsigned r = 0; for (unsigned l = 0; l < 32; l++) if (((mask >> l) & 1u) && v[l]) r |= 1u << l; return r; }
static inline unsigned long long c_shfl(unsigned lane, const unsigned long long* v, const unsigned* aux, unsigned) { return v[aux[lane] & 31u]; }
static inline unsigned long long c_match(unsigned lane, const unsigned long long* v, const unsigned*, unsigned mask) {
    unsigned r = 0; for (unsigned l = 0; l < 32; l++) if (((mask >> l) & 1u) && v[l] == v[lane]) r |= 1u << l; return r; }
template <class T> static inline unsigned long long bits(T x) { unsigned long long u = 0; memcpy(&u, &x, sizeof(T)); return u; }
template <class T> static inline T unbits(unsigned long long u) { T x; memcpy(&x, &u, sizeof(T)); return x; }
}  // namespace emu
static inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)emu::collective(mask, pred ? 1ull : 0ull, 0, emu::c_ballot); }
static inline int __any_sync(unsigned mask, int pred) { return emu::collective(mask, pred ? 1ull : 0ull, 0, emu::c_ballot) != 0ull; }
static inline int __all_sync(unsigned mask, int pred) { return emu::collective(mask, pred ? 0ull : 1ull, 0, emu::c_ballot) == 0ull; }   // nobody voted "no"
static inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { (void)emu::collective(mask, 0ull, 0, emu::c_ballot); }
static inline unsigned __match_any_sync(unsigned mask, unsigned v) { return (unsigned)emu::collective(mask, v, 0, emu::c_match); }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) {       // every lane deposits (value, wanted source lane)
    return emu::unbits<T>(emu::collective(mask, emu::bits(v), (unsigned)src & 31u, emu::c_shfl));
}
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int laneMask) { return __shfl_sync(mask, v, (int)(emu::cur().lane ^ (unsigned)laneMask)); }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta) {
    const unsigned lane = emu::cur().lane;
    const T r = __shfl_sync(mask, v, (int)(lane >= delta ? lane - delta : lane));
    return r;
}
template <class T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta) {
    const unsigned lane = emu::cur().lane;
    return __shfl_sync(mask, v, (int)(lane + delta < 32 ? lane + delta : lane));
}
static inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { unsigned r = v; for (int o = 16; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(mask, r, o)); return r; }
static inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { unsigned r = v; for (int o = 16; o > 0; o >>= 1) r = max(r, __shfl_xor_sync(mask, r, o)); return r; }
static inline int __reduce_min_sync(unsigned mask, int v) { int r = v; for (int o = 16; o > 0; o >>= 1) r = min(r, __shfl_xor_sync(mask, r, o)); return r; }
static inline int __reduce_max_sync(unsigned mask, int v) { int r = v; for (int o = 16; o > 0; o >>= 1) r = max(r, __shfl_xor_sync(mask, r, o)); return r; }
