// prefill.cu — batched (prompt) quantized matmul with exact integer arithmetic on the tensor cores.
//
// Replaces, for n > 8 activation columns, the reference's prefill paths: CPU llamafile tinyBLAS / vec_dot loops
// (ggml/src/ggml-cpu/ggml-cpu.c:1229-1421, llamafile/sgemm.cpp:3676) and CUDA mul_mat_q (ggml/src/ggml-cuda/mmq.cuh:3464).
// Arithmetic is the reference's: activations are quantized to Q8_K / Q8_0 codes exactly like the CPU does, each
// 32-element block is an exact int8 x int8 -> int32 dot (mma.sync.m16n8k32.s8, SASS IMMA.16832), and the per-block
// scales are applied in fp32.  (A bf16 tcgen05 path would be faster but cannot reproduce the int8 activation
// quantization of the oracle)
//
// CTA tile 128 rows x BN columns; every warp owns 16 rows x BN columns.  A fragments (weights) are built straight from
// global memory — a 32-bit load of nibbles yields the a0 and a2 registers of two k-steps after masking, so weights need
// no shared-memory staging at all; B fragments (int8 activation codes, plain k-contiguous layout "pact") are staged per
// 256-element unit in shared memory with a 272-byte column stride (conflict-free 4-byte fragment loads).
#include "common.cuh"
#include "kernels.h"

namespace b200 {

__host__ __device__ inline int64_t pal16(int64_t x) { return (x + 15) & ~(int64_t) 15; }

// ---- plain activation layout for the batched path: per column  qs[k] | float d[k/G] | int32 bs[k/32]
size_t pact_col_bytes(int wtype, int64_t k) {
    const int64_t G = (wtype == B200_TYPE_Q4_K) ? 256 : 32;
    return (size_t) (pal16(k) + pal16(k / G * 4) + pal16(k / 32 * 4));
}

// one warp per 32-element block; Q8_K needs the block max over 256 -> two passes over 8 warps of a CTA
template <bool Q8K>
__global__ void __launch_bounds__(256) quantize_plain_kernel(const float * x, int64_t x_col_stride, int64_t k, uint8_t * pact,
                                                             size_t col_bytes) {
    __shared__ unsigned long long wkey[8];
    __shared__ float s_max;
    pdl_launch_dependents();
    pdl_wait();
    const int col = blockIdx.y, blk = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint8_t * base = pact + (size_t) col * col_bytes;
    int8_t * qs = (int8_t *) base;
    const float v = x[(int64_t) col * x_col_stride + (int64_t) blk * 256 + tid];
    int q;
    if (Q8K) {
        float * d = (float *) (base + pal16(k));
        int * bs = (int *) (base + pal16(k) + pal16(k / 256 * 4));
        unsigned long long key = ((unsigned long long) __float_as_uint(fabsf(v)) << 32) | (unsigned) (255 - tid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o); key = other > key ? other : key; }
        if (lane == 0) wkey[warp] = key;
        __syncthreads();
        if (tid < 32) {
            unsigned long long kk = tid < 8 ? wkey[tid] : 0ull;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) { const unsigned long long other = __shfl_xor_sync(0xffffffffu, kk, o); kk = other > kk ? other : kk; }
            if (tid == 0) s_max = x[(int64_t) col * x_col_stride + (int64_t) blk * 256 + (255 - (int) (kk & 0xffffffffu))];
        }
        __syncthreads();
        const float mx = s_max;
        q = 0;
        float dd = 0.0f;
        if (mx != 0.0f) {
            const float iscale = __fdiv_rn(-127.f, mx);
            q = min(127, __float2int_rn(__fmul_rn(iscale, v)));
            dd = __fdiv_rn(1.0f, iscale);
        }
        int s = q;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) bs[blk * 8 + warp] = s;
        if (tid == 0) d[blk] = dd;
    } else {
        float * d = (float *) (base + pal16(k));
        int * bs = (int *) (base + pal16(k) + pal16(k / 32 * 4));
        const float amax = warp_max(fabsf(v));
        const float dd = __fdiv_rn(amax, 127.f);
        const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
        q = __float2int_rn(__fmul_rn(v, id));
        int s = q;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { d[blk * 8 + warp] = __half2float(__float2half_rn(dd)); bs[blk * 8 + warp] = s; }
    }
    qs[(int64_t) blk * 256 + tid] = (int8_t) q;
}

int quantize_plain(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * pact, cudaStream_t st) {
    if (n <= 0 || k <= 0) return B200_OK;
    if (k % 256) return B200_ERR_UNSUPPORTED;
    const size_t cb = pact_col_bytes(wtype, k);
    dim3 grid((unsigned) (k / 256), (unsigned) n);
    if (wtype == B200_TYPE_Q4_K) launch_pdl(quantize_plain_kernel<true>, grid, dim3(256), 0, st, x, x_col_stride, k, (uint8_t *) pact, cb);
    else if (wtype == B200_TYPE_Q4_0 || wtype == B200_TYPE_Q8_0) launch_pdl(quantize_plain_kernel<false>, grid, dim3(256), 0, st, x, x_col_stride, k, (uint8_t *) pact, cb);
    else return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

__device__ __forceinline__ void imma16832(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct MmqParams {
    const uint8_t * W;
    const uint8_t * pact;
    float * y;
    const float * bias;
    int64_t k, m, n, ldy;
    int nunits;
    size_t col_bytes;
};

#define MMQ_BSTRIDE 272  // bytes per activation column per 256-element unit in smem (68 words: conflict-free)

// FMT: 0 = Q4_K (native blocks), 1 = Q4_0 (row SoA), 2 = Q8_0 (row SoA);  NT = n-tiles of 8 columns per CTA
template <int FMT, int NT>
__global__ void __launch_bounds__(256) mmq_kernel(const MmqParams p) {
    constexpr int BN = NT * 8;
    extern __shared__ __align__(16) uint8_t sm[];
    // per stage: B codes [BN][272] | dx [BN][8] float | bs [BN][8] int
    constexpr int STAGE = BN * MMQ_BSTRIDE + BN * 8 * 4 + BN * 8 * 4;
    pdl_launch_dependents();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int64_t row_base = (int64_t) blockIdx.y * 128 + warp * 16;
    const int64_t col_base = (int64_t) blockIdx.x * BN;
    const int64_t rA = min(row_base + g, p.m - 1), rB = min(row_base + g + 8, p.m - 1);  // clamped (stores are masked)
    const int64_t row_bytes = (FMT == 2) ? (int64_t) p.nunits * 272 : (int64_t) p.nunits * 144;
    const uint8_t * wA = p.W + rA * row_bytes;
    const uint8_t * wB = p.W + rB * row_bytes;
    const bool q8k = (FMT == 0);
    const int64_t d_off = pal16(p.k), bs_off = d_off + pal16(p.k / (q8k ? 256 : 32) * 4);

    float acc[NT][4];
#pragma unroll
    for (int i = 0; i < NT; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;

    pdl_wait();
    auto stage_load = [&](int u, int s) {
        uint8_t * st = sm + (size_t) s * STAGE;
        // codes: BN columns x 256 bytes, 16-byte chunks
        for (int c = threadIdx.x; c < BN * 16; c += 256) {
            const int col = c >> 4, ch = c & 15;
            const int64_t gc = min(col_base + col, p.n - 1);
            cp_async16(st + col * MMQ_BSTRIDE + ch * 16, p.pact + (size_t) gc * p.col_bytes + (size_t) u * 256 + ch * 16);
        }
        float * dxs = (float *) (st + BN * MMQ_BSTRIDE);
        int * bss = (int *) (st + BN * MMQ_BSTRIDE + BN * 8 * 4);
        for (int c = threadIdx.x; c < BN * 8; c += 256) {
            const int col = c >> 3, j = c & 7;
            const int64_t gc = min(col_base + col, p.n - 1);
            const uint8_t * cb = p.pact + (size_t) gc * p.col_bytes;
            dxs[c] = q8k ? ((const float *) (cb + d_off))[u] : ((const float *) (cb + d_off))[u * 8 + j];
            bss[c] = ((const int *) (cb + bs_off))[u * 8 + j];
        }
        cp_async_commit();
    };

    stage_load(0, 0);
    for (int u = 0; u < p.nunits; ++u) {
        const int s = u & 1;
        if (u + 1 < p.nunits) { stage_load(u + 1, s ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const uint8_t * st = sm + (size_t) s * STAGE;
        const float * dxs = (const float *) (st + BN * MMQ_BSTRIDE);
        const int * bss = (const int *) (st + BN * MMQ_BSTRIDE + BN * 8 * 4);

        // ---- per-row scales of this unit
        float swA[8], swB[8];            // Q4_0 / Q8_0: block scale per sub-block
        int scA[8], scB[8], mnA[8], mnB[8];  // Q4_K: 6-bit scale / min codes
        float dA = 0.f, dB = 0.f, dminA = 0.f, dminB = 0.f;
        if (FMT == 0) {
            auto decode = [&](const uint8_t * blk, int * sc, int * mn, float & d, float & dmin) {
                const uint4 hdr = *reinterpret_cast<const uint4 *>(blk);
                d = half_bits_to_float(hdr.x & 0xffffu); dmin = half_bits_to_float(hdr.x >> 16);
                // get_scale_min_k4 (ggml-quants.c:703-711)
                const uint32_t sc_a = hdr.y & 0x3f3f3f3fu, sc_b = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
                const uint32_t mn_a = hdr.z & 0x3f3f3f3fu, mn_b = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    sc[j] = (sc_a >> (8 * j)) & 0xff; sc[4 + j] = (sc_b >> (8 * j)) & 0xff;
                    mn[j] = (mn_a >> (8 * j)) & 0xff; mn[4 + j] = (mn_b >> (8 * j)) & 0xff;
                }
            };
            decode(wA + (size_t) u * 144, scA, mnA, dA, dminA);
            decode(wB + (size_t) u * 144, scB, mnB, dB, dminB);
        } else {
            const int QB = (FMT == 1) ? 16 : 32;
            const unsigned short * hA = reinterpret_cast<const unsigned short *>(wA + (size_t) p.nunits * 8 * QB) + u * 8;
            const unsigned short * hB = reinterpret_cast<const unsigned short *>(wB + (size_t) p.nunits * 8 * QB) + u * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) { swA[j] = half_bits_to_float(hA[j]); swB[j] = half_bits_to_float(hB[j]); }
        }
        int ai[NT][4], am[NT][4];  // Q4_K: integer sums over the 8 sub-blocks (exact), scaled once per super-block
        if (FMT == 0) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                ai[nt][0] = ai[nt][1] = ai[nt][2] = ai[nt][3] = 0;
                am[nt][0] = am[nt][1] = am[nt][2] = am[nt][3] = 0;
            }
        }

#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 32-element sub-block j of the unit = one k-step
            uint32_t a0, a1, a2, a3;
            if (FMT == 0) {
                // nibble order (ggml-quants.c:1352-1373): 64-element group j/2, low nibbles = sub-block 2*(j/2), high = +1
                const uint8_t * qa = wA + (size_t) u * 144 + 16 + (j >> 1) * 32, * qb = wB + (size_t) u * 144 + 16 + (j >> 1) * 32;
                const uint32_t wa0 = *reinterpret_cast<const uint32_t *>(qa + 4 * t), wa1 = *reinterpret_cast<const uint32_t *>(qa + 16 + 4 * t);
                const uint32_t wb0 = *reinterpret_cast<const uint32_t *>(qb + 4 * t), wb1 = *reinterpret_cast<const uint32_t *>(qb + 16 + 4 * t);
                const int sh = (j & 1) * 4;
                a0 = (wa0 >> sh) & 0x0f0f0f0fu; a2 = (wa1 >> sh) & 0x0f0f0f0fu; a1 = (wb0 >> sh) & 0x0f0f0f0fu; a3 = (wb1 >> sh) & 0x0f0f0f0fu;
            } else if (FMT == 1) {
                // block j: 16 bytes; low nibbles = elements 0..15, high = 16..31 (ggml-quants.c:307-325); value = nibble - 8
                const uint32_t wa = *reinterpret_cast<const uint32_t *>(wA + (size_t) (u * 8 + j) * 16 + 4 * t);
                const uint32_t wb = *reinterpret_cast<const uint32_t *>(wB + (size_t) (u * 8 + j) * 16 + 4 * t);
                a0 = __vsub4(wa & 0x0f0f0f0fu, 0x08080808u); a2 = __vsub4((wa >> 4) & 0x0f0f0f0fu, 0x08080808u);
                a1 = __vsub4(wb & 0x0f0f0f0fu, 0x08080808u); a3 = __vsub4((wb >> 4) & 0x0f0f0f0fu, 0x08080808u);
            } else {
                a0 = *reinterpret_cast<const uint32_t *>(wA + (size_t) (u * 8 + j) * 32 + 4 * t);
                a2 = *reinterpret_cast<const uint32_t *>(wA + (size_t) (u * 8 + j) * 32 + 16 + 4 * t);
                a1 = *reinterpret_cast<const uint32_t *>(wB + (size_t) (u * 8 + j) * 32 + 4 * t);
                a3 = *reinterpret_cast<const uint32_t *>(wB + (size_t) (u * 8 + j) * 32 + 16 + 4 * t);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const uint8_t * bc = st + (size_t) (nt * 8 + g) * MMQ_BSTRIDE + j * 32;
                const uint32_t b0 = *reinterpret_cast<const uint32_t *>(bc + 4 * t);
                const uint32_t b1 = *reinterpret_cast<const uint32_t *>(bc + 16 + 4 * t);
                int c[4] = {0, 0, 0, 0};
                imma16832(c, a0, a1, a2, a3, b0, b1);
                const int col0 = nt * 8 + 2 * t;
                if (FMT == 0) {
                    const int bs0 = bss[col0 * 8 + j], bs1 = bss[(col0 + 1) * 8 + j];
                    ai[nt][0] += scA[j] * c[0]; ai[nt][1] += scA[j] * c[1]; ai[nt][2] += scB[j] * c[2]; ai[nt][3] += scB[j] * c[3];
                    am[nt][0] += mnA[j] * bs0;  am[nt][1] += mnA[j] * bs1;  am[nt][2] += mnB[j] * bs0;  am[nt][3] += mnB[j] * bs1;
                } else {
                    const float dx0 = dxs[col0 * 8 + j], dx1 = dxs[(col0 + 1) * 8 + j];
                    acc[nt][0] = fmaf(swA[j] * dx0, (float) c[0], acc[nt][0]);
                    acc[nt][1] = fmaf(swA[j] * dx1, (float) c[1], acc[nt][1]);
                    acc[nt][2] = fmaf(swB[j] * dx0, (float) c[2], acc[nt][2]);
                    acc[nt][3] = fmaf(swB[j] * dx1, (float) c[3], acc[nt][3]);
                }
            }
        }
        if (FMT == 0) {
            // ggml_vec_dot_q4_K_q8_K (ggml-cpu/arch/x86/quants.c:1742-1916): sumf += d_x*d*isum - d_x*dmin*msum per super-block
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const int col0 = nt * 8 + 2 * t;
                const float dx0 = dxs[col0 * 8], dx1 = dxs[(col0 + 1) * 8];
                acc[nt][0] += dx0 * dA * (float) ai[nt][0] - dx0 * dminA * (float) am[nt][0];
                acc[nt][1] += dx1 * dA * (float) ai[nt][1] - dx1 * dminA * (float) am[nt][1];
                acc[nt][2] += dx0 * dB * (float) ai[nt][2] - dx0 * dminB * (float) am[nt][2];
                acc[nt][3] += dx1 * dB * (float) ai[nt][3] - dx1 * dminB * (float) am[nt][3];
            }
        }
        __syncthreads();
    }
    // ---- store: c0 = (row g, col 2t), c1 = (row g, col 2t+1), c2 = (row g+8, col 2t), c3 = (row g+8, col 2t+1); y is column-major (ldy)
    const int64_t r0 = row_base + g, r1 = row_base + g + 8;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
        const int64_t c0 = col_base + nt * 8 + 2 * t, c1 = c0 + 1;
        const float bA = (p.bias && r0 < p.m) ? p.bias[r0] : 0.0f, bB = (p.bias && r1 < p.m) ? p.bias[r1] : 0.0f;
        if (r0 < p.m && c0 < p.n) p.y[c0 * p.ldy + r0] = acc[nt][0] + bA;
        if (r0 < p.m && c1 < p.n) p.y[c1 * p.ldy + r0] = acc[nt][1] + bA;
        if (r1 < p.m && c0 < p.n) p.y[c0 * p.ldy + r1] = acc[nt][2] + bB;
        if (r1 < p.m && c1 < p.n) p.y[c1 * p.ldy + r1] = acc[nt][3] + bB;
    }
}

template <int FMT, int NT>
static int mmq_launch(const MmqParams & p, cudaStream_t st) {
    constexpr int BN = NT * 8;
    const size_t smem = 2 * (size_t) (BN * MMQ_BSTRIDE + BN * 8 * 4 + BN * 8 * 4);
    static bool configured_dev[16] = {false};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    bool & configured = configured_dev[dev & 15];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mmq_kernel<FMT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        if (e != cudaSuccess) return (int) e;
        configured = true;
    }
    dim3 grid((unsigned) ((p.n + BN - 1) / BN), (unsigned) ((p.m + 127) / 128));
    launch_pdl(mmq_kernel<FMT, NT>, grid, dim3(256), smem, st, p);
    return (int) cudaGetLastError();
}

// y[c*ldy + r] = sum_k W[r,k] * x_c[k] (+ bias[r]); pact = n columns quantized by quantize_plain
int mul_mat_q_batched(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                      cudaStream_t st) {
    if (k <= 0 || m <= 0 || n <= 0) return B200_OK;
    if (k % 256) return B200_ERR_UNSUPPORTED;
    MmqParams p;
    p.W = (const uint8_t *) W; p.pact = (const uint8_t *) pact; p.y = y; p.bias = bias;
    p.k = k; p.m = m; p.n = n; p.ldy = ldy; p.nunits = (int) (k / 256); p.col_bytes = pact_col_bytes(wtype, k);
    if (n >= 64 && mmq_tc_enabled()) return mul_mat_q_batched_tc(wtype, W, k, m, pact, n, y, ldy, bias, p.col_bytes, st);
    const bool wide = n > 32;
    switch (wtype) {
        case B200_TYPE_Q4_K: return wide ? mmq_launch<0, 8>(p, st) : mmq_launch<0, 4>(p, st);
        case B200_TYPE_Q4_0: return wide ? mmq_launch<1, 8>(p, st) : mmq_launch<1, 4>(p, st);
        case B200_TYPE_Q8_0: return wide ? mmq_launch<2, 8>(p, st) : mmq_launch<2, 4>(p, st);
        default: return B200_ERR_UNSUPPORTED;
    }
}

}  // namespace b200
