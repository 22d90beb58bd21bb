// gemv_fmt.cuh — per-format traits of the quantized GEMV (weight/activation fragment loads + exact integer dot) shared by gemv.cu
// (one launch per matmul) and decode_mk.cu (persistent whole-token kernel).  See gemv.cu for the design notes.
#pragma once
#include "actlayout.cuh"
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp4a_ss(uint32_t a_s8x4, uint32_t b_s8x4, int c) {
    int d;
    asm("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s8x4), "r"(b_s8x4), "r"(c));
    return d;
}
// a = 2 x s16, b = 4 x u8 (lo: bytes 0,1 ; hi: bytes 2,3)
__device__ __forceinline__ int dp2a_lo_su(uint32_t a_s16x2, uint32_t b_u8x4, int c) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_su(uint32_t a_s16x2, uint32_t b_u8x4, int c) {
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_s16x2), "r"(b_u8x4), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 lds128(const void * p) { return *reinterpret_cast<const uint4 *>(p); }
__device__ __forceinline__ int byte_of(uint32_t x, int i) { return (int) __byte_perm(x, 0, 0x4440 + i); }

// ======================================================================================================
// Format traits.  A "unit" is 256 consecutive k-elements for every format; LPU lanes cooperate on a unit.
// ======================================================================================================
struct FmtQ4K {
    static constexpr int A_UNIT = 144;  // one native block_q4_K per unit
    static constexpr int B_UNIT = 0;
    static constexpr bool Q8K = true;
    static constexpr int LPU = 2;  // lane h owns qs bytes [64h, 64h+64) = sub-blocks 4h .. 4h+3 (128 elements)
    struct Act { uint32_t a[32]; uint32_t bs01, bs23; float dx; };
    struct Wt { uint32_t w[16]; uint32_t sc4, mn4; float d, dmin; };

    __device__ static __forceinline__ void load_act(const uint8_t * col, const ActLayout & L, int64_t gu, int h, Act & A) {
        const uint8_t * p = col + (gu >> 2) * 1024 + (gu & 3) * 32 + h * 16;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 v = lds128(p + j * 128);
            A.a[4 * j + 0] = v.x; A.a[4 * j + 1] = v.y; A.a[4 * j + 2] = v.z; A.a[4 * j + 3] = v.w;
        }
        const uint2 b = *reinterpret_cast<const uint2 *>(col + L.bs_off + (gu * 8 + 4 * h) * 2);
        A.bs01 = b.x; A.bs23 = b.y;
        A.dx = *reinterpret_cast<const float *>(col + L.d_off + gu * 4);
    }
    __device__ static __forceinline__ void load_w(const uint8_t * a_row, const uint8_t *, int u, int h, Wt & W) {
        const uint8_t * blk = a_row + u * 144;
        const uint4 hdr = lds128(blk);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 v = lds128(blk + 16 + h * 64 + j * 16);
            W.w[4 * j + 0] = v.x; W.w[4 * j + 1] = v.y; W.w[4 * j + 2] = v.z; W.w[4 * j + 3] = v.w;
        }
        W.d = half_bits_to_float(hdr.x & 0xffffu);
        W.dmin = half_bits_to_float(hdr.x >> 16);
        // 6-bit scales / mins (reference get_scale_min_k4, ggml/src/ggml-quants.c:703-711), four at a time
        const uint32_t sc_a = hdr.y & 0x3f3f3f3fu;
        const uint32_t sc_b = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
        const uint32_t mn_a = hdr.z & 0x3f3f3f3fu;
        const uint32_t mn_b = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
        W.sc4 = h ? sc_b : sc_a;
        W.mn4 = h ? mn_b : mn_a;
    }
    __device__ static __forceinline__ float dot(const Wt & W, const Act & A, float acc) {
        int s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s0 = dp4a_us(W.w[i] & 0x0f0f0f0fu, A.a[i], s0);
            s1 = dp4a_us(W.w[i] & 0xf0f0f0f0u, A.a[8 + i], s1);       // 16 x
            s2 = dp4a_us(W.w[8 + i] & 0x0f0f0f0fu, A.a[16 + i], s2);
            s3 = dp4a_us(W.w[8 + i] & 0xf0f0f0f0u, A.a[24 + i], s3);  // 16 x
        }
        const int t = ((byte_of(W.sc4, 0) * s0 + byte_of(W.sc4, 2) * s2) << 4) + byte_of(W.sc4, 1) * s1 + byte_of(W.sc4, 3) * s3;
        const int ms = dp2a_hi_su(A.bs23, W.mn4, dp2a_lo_su(A.bs01, W.mn4, 0));
        // reference association (arch/x86/quants.c:1764-1815): d = y.d * fp16(x.d); dmin = y.d * fp16(x.dmin)
        acc = fmaf(A.dx * W.d * 0.0625f, (float) t, acc);
        acc = fmaf(-(A.dx * W.dmin), (float) ms, acc);
        return acc;
    }
};

// Q4_0 repacked rows: A = qs (16 B per 32-element block, 128 B per unit), B = fp16 d (2 B per block, 16 B per unit)
struct FmtQ40 {
    static constexpr int A_UNIT = 128;
    static constexpr int B_UNIT = 16;
    static constexpr bool Q8K = false;
    static constexpr int LPU = 8;  // one 32-element block per lane
    struct Act { uint32_t a[8]; int bs; float dx; };
    struct Wt { uint32_t w[4]; float d; };
    __device__ static __forceinline__ void load_act(const uint8_t * col, const ActLayout & L, int64_t gu, int g, Act & A) {
        const uint8_t * p = col + gu * 256 + g * 16;
        const uint4 v0 = lds128(p), v1 = lds128(p + 128);
        A.a[0] = v0.x; A.a[1] = v0.y; A.a[2] = v0.z; A.a[3] = v0.w;
        A.a[4] = v1.x; A.a[5] = v1.y; A.a[6] = v1.z; A.a[7] = v1.w;
        A.bs = *reinterpret_cast<const int *>(col + L.bs_off + (gu * 8 + g) * 4);
        A.dx = *reinterpret_cast<const float *>(col + L.d_off + (gu * 8 + g) * 4);
    }
    __device__ static __forceinline__ void load_w(const uint8_t * a_row, const uint8_t * b_row, int u, int g, Wt & W) {
        const uint4 q = lds128(a_row + u * 128 + g * 16);
        W.w[0] = q.x; W.w[1] = q.y; W.w[2] = q.z; W.w[3] = q.w;
        W.d = half_bits_to_float(*reinterpret_cast<const unsigned short *>(b_row + u * 16 + g * 2));
    }
    __device__ static __forceinline__ float dot(const Wt & W, const Act & A, float acc) {
        int slo = 0, shi = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            slo = dp4a_us(W.w[i] & 0x0f0f0f0fu, A.a[i], slo);      // elements 0..15 (low nibbles, ggml-quants.c:307-325)
            shi = dp4a_us(W.w[i] & 0xf0f0f0f0u, A.a[4 + i], shi);  // 16 x elements 16..31 (high nibbles)
        }
        const int t = (slo << 4) + shi - (A.bs << 7);  // 16 * sum (q-8)*a
        return fmaf(W.d * A.dx * 0.0625f, (float) t, acc);
    }
};

// Q8_0 repacked rows: A = qs (32 B per block, 256 B per unit), B = fp16 d (16 B per unit)
struct FmtQ80 {
    static constexpr int A_UNIT = 256;
    static constexpr int B_UNIT = 16;
    static constexpr bool Q8K = false;
    static constexpr int LPU = 8;
    struct Act { uint32_t a[8]; float dx; };
    struct Wt { uint32_t w[8]; float d; };
    __device__ static __forceinline__ void load_act(const uint8_t * col, const ActLayout & L, int64_t gu, int g, Act & A) {
        const uint8_t * p = col + gu * 256 + g * 16;
        const uint4 v0 = lds128(p), v1 = lds128(p + 128);
        A.a[0] = v0.x; A.a[1] = v0.y; A.a[2] = v0.z; A.a[3] = v0.w;
        A.a[4] = v1.x; A.a[5] = v1.y; A.a[6] = v1.z; A.a[7] = v1.w;
        A.dx = *reinterpret_cast<const float *>(col + L.d_off + (gu * 8 + g) * 4);
    }
    __device__ static __forceinline__ void load_w(const uint8_t * a_row, const uint8_t * b_row, int u, int g, Wt & W) {
        const uint4 q0 = lds128(a_row + u * 256 + g * 32);
        const uint4 q1 = lds128(a_row + u * 256 + g * 32 + 16);
        W.w[0] = q0.x; W.w[1] = q0.y; W.w[2] = q0.z; W.w[3] = q0.w;
        W.w[4] = q1.x; W.w[5] = q1.y; W.w[6] = q1.z; W.w[7] = q1.w;
        W.d = half_bits_to_float(*reinterpret_cast<const unsigned short *>(b_row + u * 16 + g * 2));
    }
    __device__ static __forceinline__ float dot(const Wt & W, const Act & A, float acc) {
        int s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s = dp4a_ss(W.w[i], A.a[i], s);
        return fmaf(W.d * A.dx, (float) s, acc);
    }
};

}  // namespace b200
