// common.cuh — shared device helpers for the sm_100a kernels (PTX wrappers, block formats).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define B200_OK 0
#define B200_ERR_ARG -1
#define B200_ERR_UNSUPPORTED -2

#define B200_CUDA_CHECK(expr)                                   \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return (int) _e;                 \
    } while (0)

// ggml type ids (reference: ggml/include/ggml.h:389-405)
enum : int { B200_TYPE_F32 = 0, B200_TYPE_F16 = 1, B200_TYPE_Q4_0 = 2, B200_TYPE_Q8_0 = 8, B200_TYPE_Q4_K = 12 };

// block byte sizes (reference: ggml/src/ggml-common.h:170-176, :219-224, :288-306)
#define QK_K 256
#define Q4K_BLOCK_BYTES 144
#define Q4_0_BLOCK_BYTES 18
#define Q8_0_BLOCK_BYTES 34

namespace b200 {

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t * bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) { }
}

// ---- bulk async copy global -> shared (TMA engine, 1-D; SASS: UBLKCP) ------------------------------
// size and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void * smem_dst, const void * gmem_src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// same with an L2 evict-first policy: weights are streamed exactly once per token
__device__ __forceinline__ uint64_t make_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void * smem_dst, const void * gmem_src, uint32_t bytes, uint64_t * bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}

// bulk prefetch global -> L2 (TMA engine, fire and forget; size and address multiples of 16 bytes)
__device__ __forceinline__ void bulk_prefetch_l2(const void * gmem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
}

// ---- programmatic dependent launch -----------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// RULE (learned the hard way, r01): a pointer to data that a PREDECESSOR kernel writes must NOT be declared `const T * __restrict__`.
// nvcc turns such loads into invariant ld.global.nc and is then free to hoist them ABOVE the inline-asm griddepcontrol.wait
// (observed: the q loads of attn_scores_mma_kernel were scheduled before ACQBULK once the wait was no longer the first statement ->
// stale activations under real overlap, i.e. CUDA-graph replay / back-to-back stream launches, while eager launches looked fine).
// Under programmatic dependent launch the data is simply not read-only for the lifetime of the dependent grid.  Plain pointers make
// the "memory" clobber below a real barrier.  tools/sass_pdl_audit.py lists every global load that precedes the wait in the SASS.
// Host-side launch helper: every kernel of the decode step is launched with programmatic stream serialization so that
// its launch latency (and, for the GEMV, its weight prefetch) overlaps the predecessor's execution.  Contract for the
// kernels: call pdl_launch_dependents() at the top, and pdl_wait() before the first access to anything a predecessor
// may still be reading or writing.
#ifdef __CUDACC__
#include <cstdlib>
inline bool pdl_enabled() {
    static const bool on = !(getenv("B200_NO_PDL") && atoi(getenv("B200_NO_PDL")) != 0);
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- misc -------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// block-wide sum in DOUBLE (blockDim.x <= 1024; red: 32 doubles of shared memory).  The RMSNorm kernels accumulate the sum of squares
// exactly as the CPU does (float products added in double, ggml-cpu/ops.cpp:3736-3741): a fp32 tree reduction can differ in the last
// bit of `scale`, which moves EVERY normalised element by an ulp and flips int8 activation codes that sit on a rounding boundary.
__device__ __forceinline__ double block_sum_double(double v, double * red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nw; ++w) t += red[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ float half_bits_to_float(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short) h16)); }

}  // namespace b200
