// mk_microbench.cu — facts the persistent decode kernel is designed around (run on the B200, not part of the product):
//   1. latency of a grid barrier (one 64-bit counter, release/acquire) on an idle chip,
//   2. the same while every warp keeps a bulk-copy (TMA) ring streaming weights, for three arrive flavours,
//   3. round-trip latency of an L2-hit load (ld.global.cg) with and without the stream.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o mk_microbench tools/mk_microbench.cu && ./mk_microbench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include "../chatllm.cpp_b200/csrc/common.cuh"
using namespace b200;

__device__ __forceinline__ unsigned long long ld_acq(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned long long * p) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}
__device__ __forceinline__ void red_relaxed(unsigned long long * p) {
    asm volatile("red.relaxed.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

// mode 0: __threadfence + atomicAdd ; 1: red.release ; 2: red.relaxed (no ordering: latency floor)
template <int MODE>
__device__ __forceinline__ void barrier(unsigned long long * ctr, unsigned long long & target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        if (MODE == 0) { __threadfence(); atomicAdd(ctr, 1ull); }
        else if (MODE == 1) red_release(ctr);
        else red_relaxed(ctr);
        while (ld_acq(ctr) < target) { }
        if (MODE == 0) __threadfence();
    }
    __syncthreads();
}

struct P {
    unsigned long long * ctr;
    const uint8_t * W;
    size_t wbytes;
    float * probe;        // L2-resident words for the latency probe
    long long * out;      // [0] cycles in barriers, [1] #barriers, [2] cycles total, [3] probe cycles, [4] #probes
    int stream;           // keep the rings streaming
    int iters;            // ring items per warp between two barriers
    int nbar;
    int stage_bytes, stages;
};

template <int MODE>
__global__ void __launch_bounds__(256, 1) bench(const P p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem) + warp * p.stages;
    uint8_t * ring = smem + 1024 + (size_t) warp * p.stages * p.stage_bytes;
    if (lane == 0) { for (int s = 0; s < p.stages; ++s) mbar_init(&bars[s], 1); fence_mbar_init(); }
    __syncwarp();
    const uint64_t pol = make_evict_first_policy();
    const size_t gw = (size_t) blockIdx.x * 8 + warp, GW = (size_t) gridDim.x * 8;
    const size_t per = p.wbytes / GW / p.stage_bytes * p.stage_bytes;
    const uint8_t * base = p.W + gw * per;
    const size_t nitems = per / p.stage_bytes;
    size_t iss = 0, con = 0;
    auto issue = [&]() {
        if (lane == 0) { const int s = iss % p.stages; mbar_arrive_expect_tx(&bars[s], p.stage_bytes); bulk_g2s_hint(ring + (size_t) s * p.stage_bytes, base + (iss % nitems) * p.stage_bytes, p.stage_bytes, &bars[s], pol); }
        iss++;
    };
    if (p.stream) for (int i = 0; i < p.stages; ++i) issue();
    unsigned long long target = *p.ctr;  // all CTAs read the same start value before anyone arrives? (host zeroes it) 
    target = 0;
    long long tb = 0, tp = 0, np = 0;
    const long long t0 = clock64();
    float sink = 0.f;
    for (int b = 0; b < p.nbar; ++b) {
        if (p.stream) {
            for (int i = 0; i < p.iters; ++i) {
                const int s = con % p.stages;
                mbar_wait(&bars[s], (con / p.stages) & 1);
                sink += ring[(size_t) s * p.stage_bytes + lane * 16];
                __syncwarp();
                con++;
                issue();
            }
        }
        if (threadIdx.x == 32) {  // L2-hit probe: a dependent load chain of 4
            const long long a = clock64();
            int idx = (blockIdx.x * 64 + b) & 1023;
            for (int j = 0; j < 4; ++j) idx = ((int) __ldcg(p.probe + idx * 32)) & 1023;
            tp += clock64() - a; np += 4; sink += idx;
        }
        const long long a = clock64();
        barrier<MODE>(p.ctr, target);
        tb += clock64() - a;
    }
    const long long t1 = clock64();
    if (p.stream) for (size_t i = con; i < iss; ++i) mbar_wait(&bars[i % p.stages], (i / p.stages) & 1);  // drain before exit
    if (blockIdx.x == 0 && threadIdx.x == 0) { p.out[0] = tb; p.out[1] = p.nbar; p.out[2] = t1 - t0; }
    if (blockIdx.x == 0 && threadIdx.x == 32) { p.out[3] = tp; p.out[4] = np; }
    if (sink == 12345.678f) p.out[5] = 1;
}

template <int MODE>
static void run(const char * name, P p, int grid, size_t smem) {
    cudaMemset(p.ctr, 0, 8);
    cudaMemset(p.out, 0, 64);
    cudaFuncSetAttribute(bench<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    void * args[] = {&p};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((void *) bench<MODE>, dim3(grid), dim3(256), args, smem, 0);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    if (e != cudaSuccess || cudaGetLastError() != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(e)); return; }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[8]; cudaMemcpy(h, p.out, 64, cudaMemcpyDeviceToHost);
    const double ghz = h[2] / (ms * 1e6);  // SM cycles per ns over the kernel
    const double bytes = p.stream ? (double) p.nbar * p.iters * p.stage_bytes * grid * 8 : 0;
    printf("%-44s barrier %6.2f us  L2-hit load %5.0f ns  kernel %7.3f ms  stream %7.1f GB/s  (SM clock %.2f GHz)\n", name, h[0] / (double) h[1] / ghz / 1e3,
           h[4] ? h[3] / (double) h[4] / ghz : 0.0, ms, bytes / ms / 1e6, ghz);
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    P p{};
    p.wbytes = (size_t) 2 << 30;
    cudaMalloc((void **) &p.W, p.wbytes); cudaMemset((void *) p.W, 1, p.wbytes);
    cudaMalloc((void **) &p.ctr, 64); cudaMalloc((void **) &p.out, 64);
    cudaMalloc((void **) &p.probe, 1024 * 32 * 4);
    { float * h = (float *) malloc(1024 * 32 * 4); for (int i = 0; i < 1024; ++i) h[i * 32] = (float) ((i * 37 + 11) & 1023); cudaMemcpy(p.probe, h, 1024 * 32 * 4, cudaMemcpyHostToDevice); free(h); }
    p.nbar = 400; p.stage_bytes = 9216; p.stages = 2;
    const size_t smem = 1024 + (size_t) 8 * p.stages * p.stage_bytes;
    p.stream = 0; p.iters = 0;
    run<0>("idle   : threadfence + atomicAdd", p, sms, smem);
    run<1>("idle   : red.release", p, sms, smem);
    run<2>("idle   : red.relaxed (floor)", p, sms, smem);
    for (int iters : {1, 4, 16}) {
        p.stream = 1; p.iters = iters;
        char nm[96];
        snprintf(nm, sizeof nm, "stream : threadfence+atomicAdd, %2d items/bar", iters); run<0>(nm, p, sms, smem);
        snprintf(nm, sizeof nm, "stream : red.release,           %2d items/bar", iters); run<1>(nm, p, sms, smem);
        snprintf(nm, sizeof nm, "stream : red.relaxed,           %2d items/bar", iters); run<2>(nm, p, sms, smem);
    }
    p.stream = 1; p.iters = 64; p.nbar = 100;
    run<2>("stream : 64 items/bar (bandwidth reference)", p, sms, smem);
    return 0;
}
