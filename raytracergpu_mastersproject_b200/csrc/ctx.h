// ctx.h -- the context object behind the C-ABI (include/rtb200.h) and the small helpers every translation unit that implements
// entry points shares (capi.cu: single-GPU path; comm.cu: the multi-GPU collectives).
#pragma once

#include <stdint.h>
#include <stdlib.h>

#include <string>

#include <cuda_runtime.h>

#include "../../include/rtb200.h"

namespace rtb {

std::string& last_error();          // thread-local, defined in capi.cu

inline int fail(const char* what, cudaError_t e = cudaSuccess) {
    std::string& g = last_error();
    g = what;
    if (e != cudaSuccess) { g += ": "; g += cudaGetErrorString(e); }
    return 1;
}
#define CK(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) return rtb::fail(#call, _e); } while (0)
#define REQUIRE(cond, msg) do { if (!(cond)) return rtb::fail(msg); } while (0)

struct Scratch {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace rtb

struct rtb_ctx {
    int device = 0;
    int smCount = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t auxStream = nullptr;             // the tail launch runs here, concurrently with the main trace launch
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaStream_t buildStream = nullptr;           // the reference tree's refit runs here, beside the first half of the traversal hierarchy's build
    cudaEvent_t evBuildFork = nullptr, evBuildJoin = nullptr;
    uint64_t launches = 0;
    char name[256] = { 0 };
    // build scratch (grow-only)
    rtb::Scratch sortKeys[2], sortVals[2], sortCounts, encRed, enclosing, cinfo, nodes;
    // the bound raytrace set: traversal records derived from the reference-layout arrays
    rtb::Scratch pairs, ptris, psphs, psphMat, pmats, rootBox, workCounter, errFlag, parkBuf;
    uint32_t parkEpoch = 0;           // last epoch handed to a launch (see TraceParams::parkEpoch)
    rtb::Scratch etaNode, etaParent, etaArrivals;   // per-node hit-point slack (launch_eta) and its scratch
    rtb::Scratch topTable, topGlobal; // A/B build RTB_SMEM_TOP
    rtb::Scratch walkFlag;            // device word: 1 = records grown by a finite slack, t-culling allowed (pack_wide_kernel)
    rtb::Scratch bigList, wideRef;    // traversal hierarchy (traversal_tree.cu): list of big leaves; records over the reference's tree for the reference-order walk
    rtb::Scratch ttChild, ttParent, ttArrivals, ttBox, ttEta;   // its topology, boxes and subtree slack
    bool wideRefReady = false, hoisted = false;
    rtb::Scratch cnodes, leafBox, wide;    // compressed 32-byte / wide 64-byte traversal records + exact leaf boxes
    rtb::Scratch activePix, activeXY, sampleBuf, primaryHits;     // wave kernel: active-pixel list and per-(sample, pixel) colour slots
    rtb::Scratch poolSlot, poolColor, poolAtt, poolOrg, poolDir, poolNrm, poolList, poolCnt;   // streaming kernel: path pool
    bool traced = false;              // a trace was submitted since the error flag was last read
    bool bound = false, boundNodes = false, cnodesReady = false, wideReady = false, leafBoxReady = false;
    // multi-GPU (comm.cu): NCCL communicator of this context's rank, scratch for the gathered bands
    void* comm = nullptr;             // ncclComm_t
    int commRank = 0, commSize = 1;
    rtb::Scratch gatherBuf, bandRgba8;
    const void* boundNodesPtr = nullptr;
    uint32_t bT = 0, bS = 0, bM = 0, bN = 0;
    // tuning knobs (none of them changes a result), read from the environment ONCE, at rtb_ctx_create
    struct Knobs {
        uint32_t tMin = 0;            // RTB_WAVE_TMIN: lanes needed to stay in the traverse phase (0 = kernel default)
        int sortedPush = -1;          // RTB_WAVE_SORTED_PUSH: -1 = by scene (sphere-majority scenes stack waiting entries farthest-first)
        uint32_t qGate = 4;           // RTB_WAVE_QGATE
        uint32_t mainCtas = 0;        // RTB_WAVE_MAIN_CTAS: resident CTAs per SM of the persistent trace launch (0 = all that fit, 7); fewer leave room for
                                      // the launches of another context's frame (frames in flight) to run beside it
        uint32_t tailThreads = 128;   // RTB_WAVE_TAIL_THREADS: CTA size of trace_tail_kernel (64 fits the slot RTB_WAVE_MAIN_CTAS=6 leaves)
        uint32_t wideMin = 512;       // RTB_WIDE_MIN: scenes of fewer primitives keep the exact child pairs in the reference's order (C1, 1.1 k primitives: 1.01 -> 0.96 ms over the hierarchy)
        uint32_t sMin = 1;            // RTB_WAVE_SMIN: lanes that must wait for the S phase before a warp enters it
        uint32_t coopMax = 8;         // RTB_WAVE_COOP: tail hand-over threshold (live lanes per warp)
        uint32_t coopTurns = 32;      // RTB_WAVE_COOP_TURNS: long-ray hand-over threshold (turns)
        uint32_t tailSpinUs = 20000;  // RTB_WAVE_TAIL_SPIN_US: polling bound of the experimental concurrent tail launch
        size_t sampleBufBytes = 4ull << 30;   // RTB_WAVE_SAMPLE_BUF_MB: per-(sample, pixel) slot budget
        size_t streamPool = 0;        // RTB_STREAM_POOL (A/B streaming kernel)
    } knobs;
};


namespace rtb {

inline int ensure(rtb_ctx* c, Scratch& s, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (s.cap >= bytes) return 0;
    if (s.p) { CK(cudaStreamSynchronize(c->stream)); CK(cudaFree(s.p)); s.p = nullptr; s.cap = 0; }
    const size_t want = bytes + bytes / 8;      // a little slack so small growth does not reallocate
    CK(cudaMalloc(&s.p, want));
    CK(cudaMemsetAsync(s.p, 0, want, c->stream));
    s.cap = want;
    return 0;
}
inline void release(Scratch& s) { if (s.p) cudaFree(s.p); s.p = nullptr; s.cap = 0; }

inline int check_launch(rtb_ctx* c, int n, const char* what) {
    c->launches += (uint64_t)n;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(what, e);
    return 0;
}

struct Activate {   // make the context's device current for the duration of a call
    int prev = -1;
    explicit Activate(const rtb_ctx* c) { cudaGetDevice(&prev); if (prev != c->device) cudaSetDevice(c->device); else prev = -1; }
    ~Activate() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace rtb
