// engine_internal.hpp — macros shared by the engine's translation units.
#pragma once
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);      \
    } while (0)

// every kernel launch goes through LAUNCH: counts it and, when profiling, brackets it with CUDA events on the stream
#define LAUNCH(id, kernel, grid, block, smem, ...)                      \
    do {                                                                \
        prof_begin(id);                                                 \
        kernel<<<(grid), (block), (smem), stream>>>(__VA_ARGS__);       \
        prof_end();                                                     \
        ++launches;                                                     \
    } while (0)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }
