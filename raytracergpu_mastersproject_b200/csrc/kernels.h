// kernels.h -- internal launcher interface between capi.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace rtb {

struct TraceParams;

// bvh_build.cu
void launch_model_to_world(cudaStream_t st, const void* models, void* tris, uint32_t T, void* sphs, uint32_t S);
int launch_enclosing(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, uint32_t* red, void* enclosing,
                     int initInf, int smCount);
void launch_morton(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const void* enclosing, void* morton,
                   uint32_t* keys, uint32_t* vals);
void launch_morton_unpack(cudaStream_t st, const void* morton, uint32_t n, uint32_t T, uint32_t* keys, uint32_t* vals);
void launch_morton_repack(cudaStream_t st, const uint32_t* keys, const uint32_t* vals, uint32_t n, uint32_t T, void* morton);
void launch_hlbvh(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const uint32_t* codes,
                  uint32_t codeStrideWords, void* nodes, void* cinfo);
void launch_refit(cudaStream_t st, void* nodes, void* cinfo, uint32_t n, void* pairs, void* rootBox, float* etaNode);   // pairs != NULL: also emit traversal records; etaNode != NULL: also climb the hit-point slack
void launch_hlbvh_fused(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const uint32_t* codes, void* nodes,
                        void* cinfo, void* leafBox, void* ptris, void* psphs, void* sphMat, float* etaNode, const uint32_t* primBounds,
                        void* originRegion, const float* camPos, unsigned int* bigCount, uint32_t* bigList, uint32_t* smallBounds);
void launch_build_top_table(cudaStream_t st, const void* wide, uint32_t n, void* top, void* topGlobal, const unsigned int* flags);   // RTB_SMEM_TOP builds only
int launch_model_to_world_enclosing(cudaStream_t st, const void* models, void* tris, uint32_t T, void* sphs, uint32_t S, uint32_t* red,
                                    void* enclosing, int initInf);
void launch_pack_pairs(cudaStream_t st, const void* nodes, uint32_t n, void* pairs, void* rootBox);
void launch_pack_cnodes(cudaStream_t st, const void* nodes, uint32_t n, void* cnodes, void* leafBox);
void launch_pack_wide(cudaStream_t st, const void* nodes, uint32_t n, void* wide, const float* etaNode, unsigned int* flags);
// traversal_tree.cu: the hierarchy the order-free walk descends (built by rtb_build_bvh behind the reference's tree).  Returns #launches.
struct TraversalTreeBuffers {
    uint32_t *keys0, *vals0, *keys1, *vals1, *sortCounts;   // sort scratch (the reference build's, free again)
    uint2* child; uint32_t* parent; unsigned int* arrivals; // topology of the n2 - 1 internal nodes
    float4* box; float* eta;                                 // [n2-1][2] boxes, subtree slack
    unsigned int* flags;                                     // [0] cull allowed [1] start record [2] big count [3] n2
    const uint32_t* bigList; const uint32_t* primBounds; const uint32_t* smallBounds;
};
int launch_traversal_tree(cudaStream_t st, uint32_t N, const void* leafBox, const float* etaNode, const TraversalTreeBuffers& b, void* wide, cudaEvent_t etaRootReady);
int launch_eta(cudaStream_t st, const void* nodes, uint32_t n, const void* ptris, uint32_t T, const void* psphs, uint32_t S,
               void* rootBox /* [4]: box + origin region out */, const float* camPos, float* etaNode, uint32_t* parent, unsigned int* arrivals);
void launch_pack_prims(cudaStream_t st, const void* tris, uint32_t T, const void* sphs, uint32_t S, const void* mats, uint32_t M,
                       void* ptris, void* psphs, void* sphMat, void* pmats);

// radix_sort.cu
int launch_radix_sort(cudaStream_t st, uint32_t* keys0, uint32_t* vals0, uint32_t* keys1, uint32_t* vals1, uint32_t n, uint32_t* counts);
size_t radix_sort_counts_bytes(uint32_t n);

// trace.cu
void launch_trace(cudaStream_t st, TraceParams p, bool count, bool ext, bool linear, int smCount);
// second stream + fork / join events for the tail launch that runs concurrently with the main trace launch (aux == nullptr: same stream)
struct TailOverlap { cudaStream_t aux; cudaEvent_t fork, join; };
int launch_trace_wave(cudaStream_t st, TraceParams p, bool count, bool ext, bool cull, int nodesMode, int smCount, uint32_t samplesPerPass,
                      const TailOverlap& ov);   // trace_wave.cu
int launch_trace_stream(cudaStream_t st, TraceParams p, bool count, bool ext, int smCount, uint32_t samplesPerPass);   // trace_stream.cu
void launch_logistic(cudaStream_t st, void* points, uint32_t count, void* image, uint32_t W, uint32_t H, const float* pixelColor);
void launch_clear_image(cudaStream_t st, void* img, size_t pixels, int smCount);
void launch_resolve(cudaStream_t st, const void* img, size_t pixels, uint32_t rpp, void* out, int smCount);

}  // namespace rtb
