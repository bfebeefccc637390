// dev: micro-benchmark of grid-wide exchange variants for the persistent kernels (csrc/resident.h).
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o dev/grid_sync_bench dev/grid_sync_bench.cu
// Each variant runs ITERS exchanges of an all-reduce (sum of one double per CTA) inside one cooperative launch and
// checks the result; reported: microseconds per exchange.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
#define THREADS 512
#define MAXG 1024

__device__ __forceinline__ double cta_sum(double v, double *sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = sm[0];
    for (int w = 1; w < THREADS / 32; w++) r += sm[w];
    return r;
}
__device__ __forceinline__ unsigned long long ld_rlx(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_rlx(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acq(const unsigned *p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_rlx32(const unsigned *p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct Ctx { unsigned long long *slots; unsigned *ctr; double *part; double *out; int iters; };

// variant 0: atomic counter barrier with fences + partial array read-back (first resident.h version)
__global__ void __launch_bounds__(THREADS, 1) k_v0(Ctx c) {
    __shared__ double sm[THREADS / 32];
    unsigned gen = ((volatile unsigned *)c.ctr)[1];
    double acc = 0;
    for (int it = 0; it < c.iters; it++) {
        double v = cta_sum((double)(blockIdx.x + it), sm);
        if (threadIdx.x == 0) c.part[blockIdx.x] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&c.ctr[0], 1u) == gridDim.x - 1) { c.ctr[0] = 0; __threadfence(); atomicExch(&c.ctr[1], gen + 1); }
            else while (ld_acq(&c.ctr[1]) == gen) {}
            __threadfence();
        }
        gen++;
        __syncthreads();
        double a = 0;
        for (int q = threadIdx.x; q < gridDim.x; q += THREADS) a += __ldcg(c.part + q);
        acc += cta_sum(a, sm);
        // second barrier so that part[] can be rewritten (the real kernel has other phases in between)
        __syncthreads();
        if (threadIdx.x == 0) {
            if (atomicAdd(&c.ctr[0], 1u) == gridDim.x - 1) { c.ctr[0] = 0; __threadfence(); atomicExch(&c.ctr[1], gen + 1); }
            else while (ld_rlx32(&c.ctr[1]) == gen) {}
        }
        gen++;
        __syncthreads();
    }
    if (threadIdx.x == 0) c.out[blockIdx.x] = acc;
}

// variant 1: all-to-all, AoS slots of 4 words, one polling thread per slot (second resident.h version)
template <int WORDS, bool SOA, bool VIS>
__global__ void __launch_bounds__(THREADS, 1) k_v1(Ctx c) {
    __shared__ double sm[THREADS / 32];
    unsigned gen = 0;
    double acc = 0;
    for (int it = 0; it < c.iters; it++) {
        double v = cta_sum((double)(blockIdx.x + it), sm);
        ++gen;
        unsigned long long *buf = c.slots + (size_t)(gen & 1u) * MAXG * 4;
        if (threadIdx.x == 0) {
            if (VIS) __threadfence();
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            for (int k = 0; k < WORDS; k++) {
                unsigned long long w = (((k & 1) ? (u >> 32) : (u & 0xffffffffull)) << 32) | gen;
                st_rlx(SOA ? buf + (size_t)k * MAXG + blockIdx.x : buf + (size_t)blockIdx.x * 4 + k, w);
            }
        }
        __syncthreads();
        double a = 0;
        for (int q = threadIdx.x; q < gridDim.x; q += THREADS) {
            unsigned long long w[4];
            while (true) {
                bool ok = true;
                for (int k = 0; k < WORDS; k++) { w[k] = ld_rlx(SOA ? buf + (size_t)k * MAXG + q : buf + (size_t)q * 4 + k); ok = ok && (unsigned)w[k] == gen; }
                if (ok) break;
            }
            a += __longlong_as_double((long long)((w[0] >> 32) | ((w[1] >> 32) << 32)));
            if (VIS) __threadfence();
        }
        acc += cta_sum(a, sm);
    }
    if (threadIdx.x == 0) c.out[blockIdx.x] = acc;
}

// variant 2: tree - every CTA publishes {value, gen} (2 words); CTA 0's first warps gather and reduce in fixed order and
// publish the result; everybody polls the result slot with one thread
__global__ void __launch_bounds__(THREADS, 1) k_v2(Ctx c) {
    __shared__ double sm[THREADS / 32];
    __shared__ double res_s;
    unsigned gen = 0;
    double acc = 0;
    unsigned long long *res = c.slots + 2 * (size_t)MAXG * 4;   // result words
    for (int it = 0; it < c.iters; it++) {
        double v = cta_sum((double)(blockIdx.x + it), sm);
        ++gen;
        unsigned long long *buf = c.slots + (size_t)(gen & 1u) * MAXG * 4;
        if (threadIdx.x == 0) {
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            st_rlx(buf + blockIdx.x, ((u & 0xffffffffull) << 32) | gen);
            st_rlx(buf + MAXG + blockIdx.x, ((u >> 32) << 32) | gen);
        }
        if (blockIdx.x == 0) {
            double a = 0;
            for (int q = threadIdx.x; q < gridDim.x; q += THREADS) {
                unsigned long long w0, w1;
                do { w0 = ld_rlx(buf + q); w1 = ld_rlx(buf + MAXG + q); } while ((unsigned)w0 != gen || (unsigned)w1 != gen);
                a += __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
            }
            a = cta_sum(a, sm);
            if (threadIdx.x == 0) {
                unsigned long long u = (unsigned long long)__double_as_longlong(a);
                st_rlx(res + (gen & 1u) * 2, ((u & 0xffffffffull) << 32) | gen);
                st_rlx(res + (gen & 1u) * 2 + 1, ((u >> 32) << 32) | gen);
            }
        }
        if (threadIdx.x == 0) {
            unsigned long long w0, w1;
            do { w0 = ld_rlx(res + (gen & 1u) * 2); w1 = ld_rlx(res + (gen & 1u) * 2 + 1); } while ((unsigned)w0 != gen || (unsigned)w1 != gen);
            res_s = __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
        }
        __syncthreads();
        acc += res_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) c.out[blockIdx.x] = acc;
}

// variant 3: all-to-all SoA 2 words, polled by ONE warp-set with warp-shuffle reduction only (no block-wide reduce of
// the gathered values: warp 0 reduces and broadcasts through shared memory)
__global__ void __launch_bounds__(THREADS, 1) k_v3(Ctx c) {
    __shared__ double sm[THREADS / 32];
    __shared__ double res_s;
    unsigned gen = 0;
    double acc = 0;
    for (int it = 0; it < c.iters; it++) {
        double v = cta_sum((double)(blockIdx.x + it), sm);
        ++gen;
        unsigned long long *buf = c.slots + (size_t)(gen & 1u) * MAXG * 4;
        if (threadIdx.x == 0) {
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            st_rlx(buf + blockIdx.x, ((u & 0xffffffffull) << 32) | gen);
            st_rlx(buf + MAXG + blockIdx.x, ((u >> 32) << 32) | gen);
        }
        if (threadIdx.x < 32) {
            double a = 0;
            for (int q = threadIdx.x; q < gridDim.x; q += 32) {
                unsigned long long w0, w1;
                do { w0 = ld_rlx(buf + q); w1 = ld_rlx(buf + MAXG + q); } while ((unsigned)w0 != gen || (unsigned)w1 != gen);
                a += __longlong_as_double((long long)((w0 >> 32) | ((w1 >> 32) << 32)));
            }
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (threadIdx.x == 0) res_s = a;
        }
        __syncthreads();
        acc += res_s;
        __syncthreads();
    }
    if (threadIdx.x == 0) c.out[blockIdx.x] = acc;
}

// stale-read test: phase k: CTA b writes x[b] = k; exchange (v3-style + fences); then reads x[(b+1) % G] through
// (a) const __restrict__ pointer (may compile to LDG.CONSTANT / ld.global.nc), (b) plain pointer, (c) __ldcg
template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) k_stale(Ctx c, float *x, const float *__restrict__ xr, int *bad) {
    __shared__ double sm[THREADS / 32];
    unsigned gen = 0;
    int nbad = 0;
    for (int it = 1; it <= c.iters; it++) {
        for (int q = threadIdx.x; q < 64; q += THREADS) x[blockIdx.x * 64 + q] = (float)it;
        double v = cta_sum(1.0, sm);
        ++gen;
        unsigned long long *buf = c.slots + (size_t)(gen & 1u) * MAXG * 4;
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            st_rlx(buf + blockIdx.x, ((u & 0xffffffffull) << 32) | gen);
            st_rlx(buf + MAXG + blockIdx.x, ((u >> 32) << 32) | gen);
        }
        if (threadIdx.x < 32) {
            for (int q = threadIdx.x; q < gridDim.x; q += 32) {
                unsigned long long w0, w1;
                do { w0 = ld_rlx(buf + q); w1 = ld_rlx(buf + MAXG + q); } while ((unsigned)w0 != gen || (unsigned)w1 != gen);
            }
            __threadfence();
        }
        __syncthreads();
        const int nb = (blockIdx.x + 1) % gridDim.x;
        if (threadIdx.x < 64) {
            float got = MODE == 0 ? xr[nb * 64 + threadIdx.x] : (MODE == 1 ? x[nb * 64 + threadIdx.x] : __ldcg(x + nb * 64 + threadIdx.x));
            if (got != (float)it) nbad++;
        }
        __syncthreads();
    }
    if (nbad) atomicAdd(bad, nbad);
}

template <class K, class... A>
static float run(const char *name, K kern, int G, Ctx c, A... extra) {
    CK(cudaMemset(c.slots, 0, sizeof(unsigned long long) * (2 * MAXG * 4 + 8)));
    CK(cudaMemset(c.ctr, 0, 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    void *args[] = {(void *)&c, (void *)&extra...};
    CK(cudaLaunchCooperativeKernel((const void *)kern, dim3(G), dim3(THREADS), args, 0, 0));   // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaMemset(c.slots, 0, sizeof(unsigned long long) * (2 * MAXG * 4 + 8)));
    CK(cudaMemset(c.ctr, 0, 8));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchCooperativeKernel((const void *)kern, dim3(G), dim3(THREADS), args, 0, 0));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-58s %7.2f us per exchange\n", name, 1e3f * ms / c.iters);
    return ms;
}

int main() {
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int G = prop.multiProcessorCount;
    Ctx c;
    c.iters = 2000;
    CK(cudaMalloc(&c.slots, sizeof(unsigned long long) * (2 * MAXG * 4 + 8)));
    CK(cudaMalloc(&c.ctr, 8));
    CK(cudaMalloc(&c.part, sizeof(double) * MAXG));
    CK(cudaMalloc(&c.out, sizeof(double) * MAXG));
    double expect = 0;
    for (int it = 0; it < c.iters; it++) for (int b = 0; b < G; b++) expect += (double)THREADS * (b + it);
    auto check = [&](const char *n) {
        double h[MAXG];
        CK(cudaMemcpy(h, c.out, sizeof(double) * G, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int b = 0; b < G; b++) if (h[b] != expect) bad++;
        if (bad) printf("   %s: %d CTAs with a WRONG sum (%.17g vs %.17g)\n", n, bad, h[0], expect);
    };
    printf("%d SMs, %d threads per CTA, %d exchanges per launch\n", G, THREADS, c.iters);
    run("v0 atomic counter barrier x2 + fences + read-back", k_v0, G, c); check("v0");
    run("v1 all-to-all AoS 4 words, 1 poller per slot", k_v1<4, false, false>, G, c); check("v1");
    run("v1 all-to-all AoS 4 words, 1 poller per slot, VIS fences", k_v1<4, false, true>, G, c); check("v1vis");
    run("v1 all-to-all SoA 2 words, 1 poller per slot", k_v1<2, true, false>, G, c); check("v1soa");
    run("v1 all-to-all SoA 2 words, 1 poller per slot, VIS fences", k_v1<2, true, true>, G, c); check("v1soavis");
    run("v2 tree: CTA 0 gathers, everybody polls the result", k_v2, G, c); check("v2");
    run("v3 all-to-all SoA 2 words, one polling warp", k_v3, G, c); check("v3");
    float *x; int *bad;
    CK(cudaMalloc(&x, sizeof(float) * 64 * MAXG)); CK(cudaMalloc(&bad, 4));
    for (int mode = 0; mode < 3; mode++) {
        CK(cudaMemset(bad, 0, 4)); CK(cudaMemset(x, 0, sizeof(float) * 64 * MAXG));
        const char *names[3] = {"stale test: const __restrict__ loads", "stale test: plain loads", "stale test: __ldcg loads"};
        if (mode == 0) run(names[0], k_stale<0>, G, c, x, (const float *)x, bad);
        if (mode == 1) run(names[1], k_stale<1>, G, c, x, (const float *)x, bad);
        if (mode == 2) run(names[2], k_stale<2>, G, c, x, (const float *)x, bad);
        int h = 0; CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
        printf("   stale reads: %d\n", h);
    }
    return 0;
}
