// Runtime glue: CUDA headers, launch macros, error handling.
// FLIP_CPU_EMU is defined only by tests/cpu_emu (dev tooling, see tests/cpu_emu/cuda_emu.h);
// the shipped library is always the nvcc/sm_100a build.
#pragma once
#ifdef FLIP_CPU_EMU
#include "cuda_emu.h"
#define FLIP_LAUNCH(kernel, grid, block, stream, ...) \
    emu::launch((unsigned)(grid), (unsigned)(block), false, [&]() { kernel(__VA_ARGS__); })
#define FLIP_LAUNCH_SYNC(kernel, grid, block, stream, ...) \
    emu::launch((unsigned)(grid), (unsigned)(block), true, [&]() { kernel(__VA_ARGS__); })
#define FLIP_LAUNCH_X(sync, kernel, grid, block, stream, ...) \
    emu::launch((unsigned)(grid), (unsigned)(block), (sync), [&]() { kernel(__VA_ARGS__); })
#else
#include <cuda_runtime.h>
// FLIP_LAUNCH: kernel without intra-block synchronisation; FLIP_LAUNCH_SYNC: kernel that uses
// __syncthreads / warp shuffles (the distinction only matters to the CPU emulator).
#define FLIP_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define FLIP_LAUNCH_SYNC(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// FLIP_LAUNCH_X: kernel whose only block-wide synchronisation is the multi-GPU hand-shake (xch.h); `sync` = handle is sharded
#define FLIP_LAUNCH_X(sync, kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

#define FLIP_HD __host__ __device__ __forceinline__
#define FLIP_D __device__ __forceinline__

struct FlipError : public std::runtime_error {
    explicit FlipError(const std::string &m) : std::runtime_error(m) {}
};

#define CUDA_CHECK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            throw FlipError(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                            __FILE__ + ":" + std::to_string(__LINE__));                        \
        }                                                                                      \
    } while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
