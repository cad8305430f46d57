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

// LAUNCH_PDL: the same, with programmatic stream serialisation allowed, for kernels whose first statement is pdl_enter()
// (mcl_device.cuh). ONLY for those: a kernel without it could run ahead of its predecessor's writes. Ordinary
// serialisation while profiling (the events sit between the launches), when MCL_PDL=0, and for the first kernel after an
// exchange (pdl_hold): a kernel that polls other shards must not have its successor's blocks parked on the SMs meanwhile,
// or shards that share one GPU (in-process tests) could keep each other's kernels from being scheduled.
#define LAUNCH_PDL(kid, kernel, grid, block, smem, ...)                                              \
    do {                                                                                            \
        prof_begin(kid);                                                                             \
        cudaLaunchConfig_t lc__ = {};                                                               \
        lc__.gridDim = dim3(grid); lc__.blockDim = dim3(block);                                     \
        lc__.dynamicSmemBytes = (smem); lc__.stream = stream;                                       \
        cudaLaunchAttribute la__[1];                                                                \
        la__[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                            \
        la__[0].val.programmaticStreamSerializationAllowed = (use_pdl && !profiling && !pdl_hold) ? 1 : 0; \
        pdl_hold = false;                                                                           \
        lc__.attrs = la__; lc__.numAttrs = 1;                                                       \
        cudaError_t le__ = cudaLaunchKernelEx(&lc__, kernel, __VA_ARGS__);                          \
        if (le__ != cudaSuccess) return cuda_fail(le__, "launch " #kernel);                         \
        prof_end();                                                                                 \
        ++launches;                                                                                 \
    } while (0)

static inline unsigned grid_for(int64_t n, int block) { return (unsigned)((n + block - 1) / block); }
