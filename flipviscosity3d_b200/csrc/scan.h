// Device-wide exclusive scan over ints (3 passes, tiles of 2048): cell binning and multigrid row lists.
#pragma once
#include "sim.h"

// ------------------------------------------------------------------------------------------
// exclusive scan over ints (3 passes, tiles of 2048)
// ------------------------------------------------------------------------------------------
#define SCAN_T 256
#define SCAN_E 8
#define SCAN_TILE (SCAN_T * SCAN_E)

static __global__ void __launch_bounds__(SCAN_T) k_scan_local(const int *__restrict__ in, int *__restrict__ out,
                                                       int *__restrict__ tile_sums, int n) {
    __shared__ int warp_sums[SCAN_T / 32];
    int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_E;
    int v[SCAN_E];
    int sum = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) {
        int id = base + e;
        v[e] = id < n ? in[id] : 0;
        sum += v[e];
    }
    // inclusive scan of per-thread sums across the block
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < SCAN_T / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    int excl = inc - sum + (wid > 0 ? warp_sums[wid - 1] : 0);
#pragma unroll
    for (int e = 0; e < SCAN_E; e++) {
        int id = base + e;
        if (id < n) out[id] = excl;
        excl += v[e];
    }
    if (threadIdx.x == SCAN_T - 1) tile_sums[blockIdx.x] = excl;
}

// single CTA: exclusive scan of the tile sums in place; total goes to *total_out
static __global__ void __launch_bounds__(1024) k_scan_tiles(int *__restrict__ tile_sums, int ntiles, int *__restrict__ total_out) {
    __shared__ int warp_sums[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < ntiles; base += 1024) {
        int id = base + threadIdx.x;
        int v = id < ntiles ? tile_sums[id] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sums[wid] = inc;
        __syncthreads();
        if (wid == 0) {
            int w = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        int carry = carry_s;
        int excl = carry + inc - v + (wid > 0 ? warp_sums[wid - 1] : 0);
        if (id < ntiles) tile_sums[id] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

static __global__ void k_scan_add(int *__restrict__ out, const int *__restrict__ tile_sums, int n) {
    int id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id < n) out[id] += tile_sums[id / SCAN_TILE];
}

// out[0..n) = exclusive scan of in[0..n); out[n] = total.  tmp holds >= n/2048+1 ints.
static void exclusive_scan(Sim &s, const int *in, int *out, int *tmp, int n) {
    int ntiles = cdiv(n, SCAN_TILE);
    FLIP_LAUNCH_SYNC(k_scan_local, ntiles, SCAN_T, s.stream, in, out, tmp, n);
    FLIP_LAUNCH_SYNC(k_scan_tiles, 1, 1024, s.stream, tmp, ntiles, out + n);
    FLIP_LAUNCH(k_scan_add, cdiv(n, 256), 256, s.stream, out, (const int *)tmp, n);
    s.kernel_launches += 3;
    KERNEL_CHECK();
}

