// prefill_tc.cu — prompt-sized Q4_K matmul on the 5th-generation tensor cores (tcgen05.mma kind::i8, accumulators in TMEM), EXACT.
//
// Round 2: runs on the B200 and matches prefill.cu's mma.sync kernel (tests/test_gpu_blackwell.py::test_tcgen05_prefill_equals_mma_sync).
// The first version (no overlap at all) reached 146 TFLOP/s on the Qwen2.5-7B prompt matmuls against 159 for mma.sync
// (profiles/r02_config3_*); this one prefetches the next super-block's global data into registers behind the MMAs and the epilogue.
//
// Why this is exact.  Per 256-element super-block the reference computes (ggml-cpu/arch/x86/quants.c:1742-1916)
//        y += dx·d·Σ_j sc_j Σ_{k∈j} q_k x_k  −  dx·dmin·Σ_j m_j Σ_{k∈j} x_k          (sc_j, m_j: 6-bit, q: 4-bit, x: int8 codes)
// An int32 accumulator cannot run across sub-blocks with different sc_j — unless the scale is folded into the int8 operand:
//        sc_j = 8·hi_j + lo_j  (hi, lo < 8)   =>   hi_j·q ≤ 105 and lo_j·q ≤ 105 fit int8,
// so three int8 "planes" of the weight tile  A_hi = hi_j·q,  A_lo = lo_j·q,  A_m = m_j (broadcast over the sub-block)  give
//        Σ_j sc_j Σ q x = 8·(A_hi·X) + (A_lo·X),      Σ_j m_j bsum_j = (A_m·X)
// as three exact int32 GEMMs over K = 256, and ONE fp32 rescale per super-block — the same fp32 expression, in the same order over
// the super-blocks, as mmq_kernel.  TMEM is read back once per K = 256 (1/256 of the MMA work on the CUDA cores).
//
// Round-1 shape of the kernel (deliberately the simplest correct structure; the pipelined version is DESIGN.md §7.5):
//   CTA = 128 rows x 128 columns, 256 threads.  Per super-block: (1) every thread expands its row's nibbles into the three planes and
//   copies its share of the int8 activation tile, both straight into the canonical no-swizzle K-major core-matrix layout
//   (8 rows x 16 bytes per core matrix; cute/arch/mma_sm100_desc.hpp "INTERLEAVE": ((8,n),2):((1,SBO),LBO)); (2) fence.proxy.async +
//   barrier; (3) ONE thread issues 8 k-steps x 3 planes of tcgen05.mma (M=128, N=128, K=32) and commits to an mbarrier; (4) all
//   threads wait, tcgen05.ld their row's 64 columns of the three accumulators, rescale into 64 fp32 registers.  No overlap yet.
#include "common.cuh"
#include "kernels.h"
#include "prefill_tc_desc.h"

#include <cstdlib>

namespace b200 {

__host__ __device__ inline int64_t tc_al16(int64_t x) { return (x + 15) & ~(int64_t) 15; }

struct TcParams {
    const uint8_t * W;
    const uint8_t * pact;
    float * y;
    const float * bias;
    int64_t k, m, n, ldy;
    int nunits;
    size_t col_bytes;
};

__device__ __forceinline__ void tc_mma_i8_desc(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_c),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t accumulate) { tc_mma_i8_desc(tmem_c, da, db, TC_IDESC, accumulate); }
__device__ __forceinline__ void tc_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, int (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(256, 1) mmq_tc_q4k_kernel(const TcParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t * A_hi = sm, * A_lo = sm + TC_PLANE_BYTES, * A_m = sm + 2 * TC_PLANE_BYTES, * Bt = sm + 3 * TC_PLANE_BYTES;
    float * dxs = reinterpret_cast<float *>(sm + 4 * TC_PLANE_BYTES);  // [128] activation scale of this super-block per column
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wg = warp >> 2, q = warp & 3;  // warpgroup (K half in the expansion, column half in the epilogue), TMEM lane quarter
    const int row = q * 32 + lane;           // this thread's row of the tile = its TMEM lane
    const int64_t grow = min((int64_t) blockIdx.y * TC_M + row, p.m - 1);  // clamped: stores are masked
    const int64_t col_base = (int64_t) blockIdx.x * TC_N;
    const uint8_t * wrow = p.W + grow * (int64_t) p.nunits * 144;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t) TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    float yacc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) yacc[i] = 0.0f;
    const int64_t d_off = tc_al16(p.k);  // plain activation layout (prefill.cu): qs[k] | float d[k/256] | int bs[k/32]

    pdl_wait();  // the quantized activations come from the predecessor
    // software pipeline: the global loads of super-block u+1 (this thread's 80 weight bytes, 8 activation chunks, one scale) are issued
    // right after the MMAs of super-block u and land while the tensor core works and the epilogue runs — the first version exposed two
    // dependent global round trips per super-block (measured r02: 146 TFLOP/s, slower than the mma.sync kernel)
    const int coln = tid & 127, cb = (tid >> 7) * 8;
    const int64_t gc = min(col_base + coln, p.n - 1);
    const uint8_t * xcol = p.pact + (size_t) gc * p.col_bytes;
    uint4 hdr, wq[4], xq[8];
    float dx_n = 0.0f;
    auto prefetch = [&](int u) {
        const uint8_t * blk = wrow + (size_t) u * 144;
        hdr = *reinterpret_cast<const uint4 *>(blk);
#pragma unroll
        for (int h = 0; h < 4; ++h) wq[h] = *reinterpret_cast<const uint4 *>(blk + 16 + wg * 64 + h * 16);  // sub-blocks 4*wg .. 4*wg+3 share these 64 bytes
#pragma unroll
        for (int c = 0; c < 8; ++c) xq[c] = *reinterpret_cast<const uint4 *>(xcol + (size_t) u * 256 + (cb + c) * 16);
        if (tid < 128) dx_n = reinterpret_cast<const float *>(xcol + d_off)[u];
    };
    prefetch(0);
    for (int u = 0; u < p.nunits; ++u) {
        // ---- (1a) weights: this thread's row, sub-blocks 4*wg .. 4*wg+3 -> three int8 planes
        const float d_r = half_bits_to_float(hdr.x & 0xffffu), dmin_r = half_bits_to_float(hdr.x >> 16);
        // get_scale_min_k4 (ggml-quants.c:703-711), four at a time: sub-blocks 0..3 in *_a, 4..7 in *_b
        const uint32_t sc_a = hdr.y & 0x3f3f3f3fu, sc_b = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
        const uint32_t mn_a = hdr.z & 0x3f3f3f3fu, mn_b = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
        const uint32_t sc4 = wg ? sc_b : sc_a, mn4 = wg ? mn_b : mn_a;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * wg + jj;
            const uint32_t sc = (sc4 >> (8 * jj)) & 0xffu, mn = (mn4 >> (8 * jj)) & 0xffu;
            const uint32_t hi = sc >> 3, lo = sc & 7u, mrep = mn * 0x01010101u;
            // nibble order (ggml-quants.c:1352-1373): 64-element group j/2 = 32 bytes; low nibbles = sub-block 2*(j/2), high = the next one
            const int sh = (jj & 1) * 4;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const uint4 v = wq[(jj >> 1) * 2 + half];
                const uint32_t n0 = (v.x >> sh) & 0x0f0f0f0fu, n1 = (v.y >> sh) & 0x0f0f0f0fu, n2 = (v.z >> sh) & 0x0f0f0f0fu, n3 = (v.w >> sh) & 0x0f0f0f0fu;
                const uint32_t o = tc_off(row, j * 2 + half);
                // per-byte products <= 7 * 15 = 105: no carry between the four bytes of a word
                *reinterpret_cast<uint4 *>(A_hi + o) = make_uint4(n0 * hi, n1 * hi, n2 * hi, n3 * hi);
                *reinterpret_cast<uint4 *>(A_lo + o) = make_uint4(n0 * lo, n1 * lo, n2 * lo, n3 * lo);
                *reinterpret_cast<uint4 *>(A_m + o) = make_uint4(mrep, mrep, mrep, mrep);
            }
        }
        // ---- (1b) activations: column (tid & 127), K chunks 8*(tid >> 7) .. +7 of this super-block
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4 *>(Bt + tc_off(coln, cb + c)) = xq[c];
        if (tid < 128) dxs[coln] = dx_n;
        fence_proxy_async();  // generic-proxy stores above -> visible to the tensor core's (async proxy) reads
        __syncthreads();
        // ---- (2) one thread issues the 24 MMAs of this super-block and commits them to the mbarrier
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(A_hi), a1 = smem_u32(A_lo), a2 = smem_u32(A_m), b0 = smem_u32(Bt);
#pragma unroll
            for (int s = 0; s < 8; ++s) {  // k-step s = 16-byte chunks 2s, 2s+1
                const uint32_t koff = (uint32_t) s * 2u * TC_LBO;
                const uint64_t db = tc_desc(b0 + koff);
                tc_mma_i8(tmem + 0 * TC_N, tc_desc(a0 + koff), db, s > 0);
                tc_mma_i8(tmem + 1 * TC_N, tc_desc(a1 + koff), db, s > 0);
                tc_mma_i8(tmem + 2 * TC_N, tc_desc(a2 + koff), db, s > 0);
            }
            tc_commit(&mma_bar);  // implies tcgen05.fence::before_thread_sync
        }
        if (u + 1 < p.nunits) prefetch(u + 1);   // in flight during the MMAs and the epilogue below
        // ---- (3) everyone waits for the accumulators, then reads its row: columns 64*wg .. +63 of the three planes
        mbar_wait(&mma_bar, (uint32_t) u & 1u);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            const int c0 = wg * 64 + ch * 32;
            const uint32_t taddr = tmem + ((uint32_t) (q * 32) << 16) + (uint32_t) c0;
            int hi[32], lo[32], mm[32];
            tc_ld32(taddr + 0 * TC_N, hi);
            tc_ld32(taddr + 1 * TC_N, lo);
            tc_ld32(taddr + 2 * TC_N, mm);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float dx = dxs[c0 + i];
                // the same fp32 expression as mmq_kernel (prefill.cu): acc += dx*d*isum - dx*dmin*msum
                yacc[ch * 32 + i] += dx * d_r * (float) (8 * hi[i] + lo[i]) - dx * dmin_r * (float) mm[i];
            }
        }
        tc_fence_before();
        __syncthreads();  // TMEM accumulators and the smem tiles are free for the next super-block
    }
    // ---- store: y is column-major (ldy); this thread owns row `grow`, columns col_base + 64*wg .. +63
    const int64_t r = (int64_t) blockIdx.y * TC_M + row;
    if (r < p.m) {
        const float b = p.bias ? p.bias[r] : 0.0f;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
            const int64_t c = col_base + wg * 64 + i;
            if (c < p.n) p.y[c * p.ldy + r] = yacc[i] + b;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t) TC_TMEM_COLS) : "memory");
}


// ------------------------------------------------------------------------------------------------------------------
// Q4_0 / Q8_0 (one fp16 scale per 32-element block on BOTH operands): the rescale is inherently per K = 32, so the tile is
// 128 rows x 64 columns and every k-step gets ITS OWN TMEM accumulator (8 x 64 = 512 columns): 8 independent MMAs (K = 32,
// accumulate = false), one commit, one wait, then each thread applies  acc = fma(d_w[r][j] * d_x[c][j], (float) isum_j, acc)
// for j = 0..7 in order — the expression and order of mmq_kernel<1|2>.  FMT 1 = Q4_0, 2 = Q8_0 (device row layout: qs[nb][QB] | d[nb]).
// ------------------------------------------------------------------------------------------------------------------
template <int FMT>
__global__ void __launch_bounds__(256, 1) mmq_tc_blk32_kernel(const TcParams p) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t * At = sm, * Bt = sm + TC_PLANE_BYTES;                      // A: 128 x 256 B ; B: 64 x 256 B
    float * dxs = reinterpret_cast<float *>(sm + TC_PLANE_BYTES + TC_N64 * 256);  // [64][8] activation block scales of this unit
    __shared__ __align__(8) uint64_t mma_bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int QB = (FMT == 1) ? 16 : 32;
    pdl_launch_dependents();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wg = warp >> 2, q = warp & 3;
    const int row = q * 32 + lane;
    const int64_t grow = min((int64_t) blockIdx.y * TC_M + row, p.m - 1);
    const int64_t col_base = (int64_t) blockIdx.x * TC_N64;
    const int64_t row_bytes = (int64_t) p.nunits * (8 * QB + 16);
    const uint8_t * wrow = p.W + grow * row_bytes;
    const unsigned short * wd = reinterpret_cast<const unsigned short *>(wrow + (size_t) p.nunits * 8 * QB);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t) TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    float yacc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) yacc[i] = 0.0f;
    const int64_t d_off = tc_al16(p.k);  // plain activation layout: qs[k] | float d[k/32] | int bs[k/32]

    pdl_wait();
    // software pipeline (see mmq_tc_q4k_kernel): the next unit's weight blocks, block scales and activation chunks are prefetched into
    // registers while this unit's MMAs and epilogue run
    constexpr int WV = (FMT == 1) ? 4 : 8;           // uint4 of quants per thread per unit (4 blocks of 16 / 32 bytes)
    const int coln = tid & 63, cb = (tid >> 6) * 4;
    const int64_t gc = min(col_base + coln, p.n - 1);
    const uint8_t * xcol = p.pact + (size_t) gc * p.col_bytes;
    const int sc_col = tid >> 3, sc_j = tid & 7;     // block-scale word this thread stages ([64][8] floats = 512 words, 2 per thread)
    const int64_t gsc0 = min(col_base + sc_col, p.n - 1), gsc1 = min(col_base + sc_col + 32, p.n - 1);
    uint4 wq[WV], xq[4], wsc;
    float dxa = 0.0f, dxb = 0.0f;
    auto prefetch = [&](int u) {
#pragma unroll
        for (int h = 0; h < WV; ++h) wq[h] = *reinterpret_cast<const uint4 *>(wrow + ((size_t) (u * 8 + 4 * wg) * QB) + h * 16);
        wsc = *reinterpret_cast<const uint4 *>(wd + u * 8);   // 8 fp16 block scales of this row
#pragma unroll
        for (int c = 0; c < 4; ++c) xq[c] = *reinterpret_cast<const uint4 *>(xcol + (size_t) u * 256 + (cb + c) * 16);
        dxa = reinterpret_cast<const float *>(p.pact + (size_t) gsc0 * p.col_bytes + d_off)[u * 8 + sc_j];
        dxb = reinterpret_cast<const float *>(p.pact + (size_t) gsc1 * p.col_bytes + d_off)[u * 8 + sc_j];
    };
    prefetch(0);
    for (int u = 0; u < p.nunits; ++u) {
        float sw[8];
        {
            const uint32_t wsw[4] = {wsc.x, wsc.y, wsc.z, wsc.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) sw[j] = half_bits_to_float((wsw[j >> 1] >> (16 * (j & 1))) & 0xffffu);
        }
        // ---- weights: this thread's row, blocks 4*wg .. 4*wg+3 -> int8 (Q4_0: nibble - 8; low nibbles = elements 0..15, ggml-quants.c:307-325)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int j = 4 * wg + jj;
            if (FMT == 1) {
                const uint4 v = wq[jj];
                *reinterpret_cast<uint4 *>(At + tc_off(row, 2 * j)) =
                    make_uint4(__vsub4(v.x & 0x0f0f0f0fu, 0x08080808u), __vsub4(v.y & 0x0f0f0f0fu, 0x08080808u), __vsub4(v.z & 0x0f0f0f0fu, 0x08080808u),
                               __vsub4(v.w & 0x0f0f0f0fu, 0x08080808u));
                *reinterpret_cast<uint4 *>(At + tc_off(row, 2 * j + 1)) =
                    make_uint4(__vsub4((v.x >> 4) & 0x0f0f0f0fu, 0x08080808u), __vsub4((v.y >> 4) & 0x0f0f0f0fu, 0x08080808u),
                               __vsub4((v.z >> 4) & 0x0f0f0f0fu, 0x08080808u), __vsub4((v.w >> 4) & 0x0f0f0f0fu, 0x08080808u));
            } else {
                *reinterpret_cast<uint4 *>(At + tc_off(row, 2 * j)) = wq[(2 * jj) % WV];
                *reinterpret_cast<uint4 *>(At + tc_off(row, 2 * j + 1)) = wq[(2 * jj + 1) % WV];
            }
        }
        // ---- activations: column (tid & 63), K chunks 4*(tid >> 6) .. +3 ; block scales [64][8]
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4 *>(Bt + tc_off64(coln, cb + c)) = xq[c];
        dxs[sc_j * TC_N64 + sc_col] = dxa;          // [block][column]: the epilogue reads four columns per shared-memory load
        dxs[sc_j * TC_N64 + sc_col + 32] = dxb;
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(At), b0 = smem_u32(Bt);
#pragma unroll
            for (int s = 0; s < 8; ++s)   // block s = K chunks 2s, 2s+1 -> its own accumulator (columns 64 s .. 64 s + 63)
                tc_mma_i8_desc(tmem + (uint32_t) s * TC_N64, tc_desc_lbo(a0 + (uint32_t) s * 2u * TC_LBO, TC_LBO), tc_desc_lbo(b0 + (uint32_t) s * 2u * TC_LBO64, TC_LBO64),
                               TC_IDESC64, 0u);
            tc_commit(&mma_bar);
        }
        if (u + 1 < p.nunits) prefetch(u + 1);   // in flight during the MMAs and the epilogue below
        mbar_wait(&mma_bar, (uint32_t) u & 1u);
        tc_fence_after();
        const int c0 = wg * 32;  // this thread's 32 columns of the 64-column tile
#pragma unroll
        for (int s = 0; s < 8; ++s) {
            int isum[32];
            tc_ld32(tmem + ((uint32_t) (q * 32) << 16) + (uint32_t) (s * TC_N64 + c0), isum);
            tc_wait_ld();
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
                const float4 d4 = *reinterpret_cast<const float4 *>(dxs + s * TC_N64 + c0 + 4 * i4);
                yacc[4 * i4 + 0] = fmaf(sw[s] * d4.x, (float) isum[4 * i4 + 0], yacc[4 * i4 + 0]);   // mmq_kernel<1|2>'s expression
                yacc[4 * i4 + 1] = fmaf(sw[s] * d4.y, (float) isum[4 * i4 + 1], yacc[4 * i4 + 1]);
                yacc[4 * i4 + 2] = fmaf(sw[s] * d4.z, (float) isum[4 * i4 + 2], yacc[4 * i4 + 2]);
                yacc[4 * i4 + 3] = fmaf(sw[s] * d4.w, (float) isum[4 * i4 + 3], yacc[4 * i4 + 3]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    const int64_t r = (int64_t) blockIdx.y * TC_M + row;
    if (r < p.m) {
        const float b = p.bias ? p.bias[r] : 0.0f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int64_t c = col_base + wg * 32 + i;
            if (c < p.n) p.y[c * p.ldy + r] = yacc[i] + b;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t) TC_TMEM_COLS) : "memory");
}

bool mmq_tc_enabled() {
    // default ON since round 2 (measured on the B200, 2048-column prompt matmuls: Q4_K 378 vs 152 TFLOP/s, Q4_0 217 vs 159, Q8_0 157 vs 121
    // for the mma.sync kernel; profiles/r02_prefill_kernel_matrix.txt).  B200_MMQ_TCGEN05=0 selects prefill.cu's mma.sync kernel.
    static const bool on = !(getenv("B200_MMQ_TCGEN05") && atoi(getenv("B200_MMQ_TCGEN05")) == 0);
    return on;
}

// same contract as mul_mat_q_batched (prefill.cu)
template <typename Kern>
static int tc_launch(Kern kern, const TcParams & p, size_t smem, int ntile, bool * configured_dev, cudaStream_t st) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (!configured_dev[dev & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        configured_dev[dev & 15] = true;
    }
    dim3 grid((unsigned) ((p.n + ntile - 1) / ntile), (unsigned) ((p.m + TC_M - 1) / TC_M));
    launch_pdl(kern, grid, dim3(256), smem, st, p);
    return (int) cudaGetLastError();
}

int mul_mat_q_batched_tc(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                         size_t col_bytes, cudaStream_t st) {
    if (k <= 0 || m <= 0 || n <= 0) return B200_OK;
    if (k % 256) return B200_ERR_UNSUPPORTED;
    TcParams p;
    p.W = (const uint8_t *) W; p.pact = (const uint8_t *) pact; p.y = y; p.bias = bias;
    p.k = k; p.m = m; p.n = n; p.ldy = ldy; p.nunits = (int) (k / 256); p.col_bytes = col_bytes;
    static bool cfg_k[16] = {false}, cfg_40[16] = {false}, cfg_80[16] = {false};
    const size_t smem32 = (size_t) TC_PLANE_BYTES + TC_N64 * 256 + TC_N64 * 8 * sizeof(float);
    switch (wtype) {
        case B200_TYPE_Q4_K: return tc_launch(mmq_tc_q4k_kernel, p, 4 * (size_t) TC_PLANE_BYTES + TC_N * sizeof(float), TC_N, cfg_k, st);
        case B200_TYPE_Q4_0: return tc_launch(mmq_tc_blk32_kernel<1>, p, smem32, TC_N64, cfg_40, st);
        case B200_TYPE_Q8_0: return tc_launch(mmq_tc_blk32_kernel<2>, p, smem32, TC_N64, cfg_80, st);
        default: return B200_ERR_UNSUPPORTED;
    }
}

}  // namespace b200
