// tests/emu/cooperative_groups.h -- the one cooperative-groups facility the kernels use (a grid-wide barrier), for the SIMT
// emulation (test infrastructure, see cuda_runtime.h).  The emulated device has ONE multiprocessor and reports one resident
// block per kernel, so a cooperative launch is a single block and grid.sync() is that block's barrier.
#pragma once
#include <cuda_runtime.h>
namespace cooperative_groups {
struct grid_group {
    void sync() const {
        if (gridDim.x * gridDim.y * gridDim.z != 1u) { fprintf(stderr, "emu: grid.sync() with more than one block\n"); abort(); }
        __syncthreads();
    }
};
static inline grid_group this_grid() { return grid_group(); }
}  // namespace cooperative_groups
