// fused.cu — fused decode-step kernels.  Same arithmetic as the unfused ops (ops.cu / quantize.cu), fewer launches:
//   add_rmsnorm_quant : [x += r] ; y = rms_norm(x) * w ; qact = Q(y)                (ADD + RMS_NORM + MUL + quantize)
//   rope_kv_store     : RoPE(q) in place ; RoPE(k) -> K cache row ; v -> V cache column  (2 x ROPE + SET_ROWS + CPY)
//   attn_scores2 / attn_softmax_pv : decode attention in two launches                    (7 graph nodes)
// Every kernel signals launch_dependents at its top (so the next GEMV can become resident and prefetch weights) and
// executes griddepcontrol.wait before touching data produced by its predecessor.
#include "actlayout.cuh"
#include "common.cuh"
#include "kernels.h"

#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

namespace b200 {

__device__ __forceinline__ float blk_sum(float v, float * red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) t = warp_sum(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    return red[0];
}
__device__ __forceinline__ float blk_max(float v, float * red) {
    v = warp_max(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    if (warp == 0) t = warp_max(t);
    if (threadIdx.x == 0) red[0] = t;
    __syncthreads();
    return red[0];
}

// ------------------------------------------------------------------------------------------------------------------
// add + rms_norm + mul + activation quantization.  One CTA per row; blockDim multiple of 64; thread t owns elements
// 4t + 4*blockDim*j (j < J).  Arithmetic: rms_norm as ops.cu (reference ops.cpp:3710-3758), quantizers as quantize.cu
// (reference ggml-quants.c:2555-2592 / arch/x86/quants.c:290-384) — bit-identical codes.
// ------------------------------------------------------------------------------------------------------------------
template <bool Q8K, int J>
__global__ void __launch_bounds__(1024) add_rmsnorm_quant_kernel(const float * x, const float * r, const float * w,
                                                                 float * x_out, float * y_out, uint8_t * qact,
                                                                 int64_t ne0, float eps, size_t col_bytes, int nsplit) {
    __shared__ double red[32];
    __shared__ unsigned long long keys[32];
    __shared__ float bmax[16];
    pdl_launch_dependents();
    pdl_wait();
    const int t = threadIdx.x, T = blockDim.x;
    const int64_t row = blockIdx.x;
    const float * xr = x + row * ne0;
    float v[J][4];
    double ss = 0.0;   // float products accumulated in double, like the CPU (common.cuh block_sum_double)
    if (nsplit > 0) {
        // reduction tail of the split decode attention (attn_pv_split_kernel): x = nsplit partial rows of ne0 floats, summed in
        // split order; no normalisation.  x_out receives the sum, qact its quantization (the o-projection's activations).
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (e < ne0) {
                // 8 partial rows in flight at a time (a plain loop would serialise one L2 round trip per split); summed in split order
                for (int c0 = 0; c0 < nsplit; c0 += 8) {
                    float4 b[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        b[i] = (c0 + i < nsplit) ? *reinterpret_cast<const float4 *>(x + (int64_t) (c0 + i) * ne0 + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { a.x += b[i].x; a.y += b[i].y; a.z += b[i].z; a.w += b[i].w; }
                }
                if (x_out) *reinterpret_cast<float4 *>(x_out + e) = a;
            }
            v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
        }
    } else {
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
        if (e < ne0) {
            float4 a = *reinterpret_cast<const float4 *>(xr + e);
            if (r) {
                const float4 b = *reinterpret_cast<const float4 *>(r + row * ne0 + e);
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (x_out) *reinterpret_cast<float4 *>(x_out + row * ne0 + e) = a;
            v[j][0] = a.x; v[j][1] = a.y; v[j][2] = a.z; v[j][3] = a.w;
            ss += (double) __fmul_rn(a.x, a.x); ss += (double) __fmul_rn(a.y, a.y); ss += (double) __fmul_rn(a.z, a.z); ss += (double) __fmul_rn(a.w, a.w);
        } else {
            v[j][0] = v[j][1] = v[j][2] = v[j][3] = 0.0f;
        }
    }
    // the norm weights do not depend on the reduction: fetch them before it (one L2 round trip off the serial chain)
    float4 wv[J];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
        wv[j] = (e < ne0) ? *reinterpret_cast<const float4 *>(w + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    ss = block_sum_double(ss, red);
    const float mean = (float) (ss / (double) ne0);
    const float scale = 1.0f / sqrtf(mean + eps);
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
        if (e < ne0) {
            const float4 ww = wv[j];
            v[j][0] = (v[j][0] * scale) * ww.x; v[j][1] = (v[j][1] * scale) * ww.y;
            v[j][2] = (v[j][2] * scale) * ww.z; v[j][3] = (v[j][3] * scale) * ww.w;
            if (y_out) *reinterpret_cast<float4 *>(y_out + row * ne0 + e) = make_float4(v[j][0], v[j][1], v[j][2], v[j][3]);
        }
    }
    }  // nsplit == 0
    if (!qact) return;
    const ActLayout L = act_layout(Q8K, ne0);
    uint8_t * base = qact + (size_t) row * col_bytes;
    const int lane = t & 31, warp = t >> 5;
    if (Q8K) {
        float * dd = (float *) (base + L.d_off);
        int16_t * bs = (int16_t *) (base + L.bs_off);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
            const bool on = e < ne0;
            // first-occurrence argmax |v| over the 256-element block (= 64 consecutive threads = 2 warps)
            unsigned long long key = 0ull;
            if (on) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned long long kk = ((unsigned long long) __float_as_uint(fabsf(v[j][i])) << 32) | (unsigned) (255 - (4 * (t & 63) + i));
                    key = kk > key ? kk : key;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            __syncthreads();
            if (lane == 0) keys[warp] = key;
            __syncthreads();
            const unsigned long long k2 = keys[warp ^ 1];
            key = k2 > key ? k2 : key;
            const int idx = 255 - (int) (key & 0xffffffffu);
            if (on && (idx >> 2) == (t & 63)) bmax[t >> 6] = v[j][idx & 3];
            __syncthreads();
            if (on) {
                const float mx = bmax[t >> 6];
                int q[4] = {0, 0, 0, 0};
                float dv = 0.0f;
                if (mx != 0.0f) {
                    const float iscale = __fdiv_rn(-127.f, mx);
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = min(127, __float2int_rn(__fmul_rn(iscale, v[j][i])));
                    dv = __fdiv_rn(1.0f, iscale);
                }
                const uint32_t packed = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
                *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e)) = packed;
                int s = q[0] + q[1] + q[2] + q[3];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                if ((t & 7) == 0) bs[e >> 5] = (int16_t) s;
                if ((t & 63) == 0) dd[e >> 8] = dv;
            }
        }
        // zero the padding of the last (partial) 1024-element group
        for (int64_t e = ne0 + 4 * (int64_t) t; e < L.qs_bytes; e += 4 * (int64_t) T) *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e)) = 0u;
    } else {
        float * dd = (float *) (base + L.d_off);
        int * bs = (int *) (base + L.bs_off);
#pragma unroll
        for (int j = 0; j < J; ++j) {
            const int64_t e = 4 * (int64_t) t + 4 * (int64_t) T * j;
            const bool on = e < ne0;
            float amax = on ? fmaxf(fmaxf(fabsf(v[j][0]), fabsf(v[j][1])), fmaxf(fabsf(v[j][2]), fabsf(v[j][3]))) : 0.0f;
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float dv = __fdiv_rn(amax, 127.f);
            const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = __float2int_rn(__fmul_rn(v[j][i], id));
            int s = q[0] + q[1] + q[2] + q[3];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (on) {
                const uint32_t packed = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
                *reinterpret_cast<uint32_t *>(base + act_qs_off_q80(e)) = packed;
                if ((t & 7) == 0) { dd[e >> 5] = __half2float(__float2half_rn(dv)); bs[e >> 5] = s; }
            }
        }
    }
}

int add_rmsnorm_quant(int wtype, const float * x, const float * r, const float * w, float * x_out, float * y_out, void * qact, int64_t ne0, int64_t nrows,
                      float eps, cudaStream_t st) {
    if (nrows <= 0) return B200_OK;
    if (ne0 % 256 || ne0 > 20480) return B200_ERR_UNSUPPORTED;
    const bool q8k = wtype == B200_TYPE_Q4_K;
    if (qact && !q8k && wtype != B200_TYPE_Q4_0 && wtype != B200_TYPE_Q8_0) return B200_ERR_UNSUPPORTED;
    int threads = (int) (ne0 / 4);
    if (threads > 1024) threads = 1024;
    const int J = (int) ((ne0 / 4 + threads - 1) / threads);
    const size_t cb = qact ? qact_col_bytes(wtype, ne0) : 0;
#define B200_ARQ(J_)                                                                                                                          \
    if (J == J_) {                                                                                                                            \
        if (q8k) launch_pdl(add_rmsnorm_quant_kernel<true, J_>, dim3((unsigned) nrows), dim3(threads), 0, st, x, r, w, x_out, y_out, (uint8_t *) qact, ne0, eps, cb, 0); \
        else launch_pdl(add_rmsnorm_quant_kernel<false, J_>, dim3((unsigned) nrows), dim3(threads), 0, st, x, r, w, x_out, y_out, (uint8_t *) qact, ne0, eps, cb, 0);     \
        return (int) cudaGetLastError();                                                                                                      \
    }
    B200_ARQ(1) B200_ARQ(2) B200_ARQ(3) B200_ARQ(4) B200_ARQ(5)
#undef B200_ARQ
    return B200_ERR_UNSUPPORTED;
}

// out[e] = sum_c partial[c][e] (c < nsplit, in order) ; qact = quantized out for weight type wtype (NULL: no quantization)
static int sum_partials_quant(int wtype, const float * partial, int nsplit, float * out, void * qact, int64_t ne0, cudaStream_t st) {
    if (ne0 % 256 || ne0 > 20480 || nsplit <= 0) return B200_ERR_UNSUPPORTED;
    const bool q8k = wtype == B200_TYPE_Q4_K;
    if (qact && !q8k && wtype != B200_TYPE_Q4_0 && wtype != B200_TYPE_Q8_0) return B200_ERR_UNSUPPORTED;
    int threads = (int) (ne0 / 4);
    if (threads > 1024) threads = 1024;
    const int J = (int) ((ne0 / 4 + threads - 1) / threads);
    const size_t cb = qact ? qact_col_bytes(wtype, ne0) : 0;
    const float * nul = nullptr;
    float * nulo = nullptr;
#define B200_SPQ(J_)                                                                                                                          \
    if (J == J_) {                                                                                                                            \
        if (q8k) launch_pdl(add_rmsnorm_quant_kernel<true, J_>, dim3(1), dim3(threads), 0, st, partial, nul, nul, out, nulo, (uint8_t *) qact, ne0, 0.0f, cb, nsplit); \
        else launch_pdl(add_rmsnorm_quant_kernel<false, J_>, dim3(1), dim3(threads), 0, st, partial, nul, nul, out, nulo, (uint8_t *) qact, ne0, 0.0f, cb, nsplit);     \
        return (int) cudaGetLastError();                                                                                                      \
    }
    B200_SPQ(1) B200_SPQ(2) B200_SPQ(3) B200_SPQ(4) B200_SPQ(5)
#undef B200_SPQ
    return B200_ERR_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------------------------
// RoPE(q), RoPE(k) + KV-cache append for ONE token.  grid = n_heads + kv_heads, block = head_dim/2 threads.
// angle recurrence exactly as rope_kernel (ops.cu).  mode 0 = adjacent pairs, 2 = NEOX.
// ------------------------------------------------------------------------------------------------------------------
__global__ void rope_kv_store_kernel(const float * q, float * q_out, const float * k, const float * v,
                                     const int32_t * pos, const float * ff, __half * kc, __half * vc,
                                     int n_heads, int kv_heads, int hd, int mode, float theta_scale, int64_t k_row_stride, int64_t v_row_stride, int v_col) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = threadIdx.x, half = hd / 2;
    const int p = pos[0];
    float theta = (float) p;
    for (int j = 0; j < i; ++j) theta *= theta_scale;
    const float th = theta / (ff ? ff[i] : 1.0f);
    const float c = cosf(th), s = sinf(th);
    const int i0 = (mode == 0) ? 2 * i : i, i1 = (mode == 0) ? 2 * i + 1 : i + half;
    const int b = blockIdx.x;
    if (b < n_heads) {
        const float * h = q + (int64_t) b * hd;
        float * o = q_out + (int64_t) b * hd;
        const float x0 = h[i0], x1 = h[i1];
        o[i0] = x0 * c - x1 * s;
        o[i1] = x0 * s + x1 * c;
    } else {
        const int g = b - n_heads;
        const float * h = k + (int64_t) g * hd;
        const float x0 = h[i0], x1 = h[i1];
        __half * krow = kc + (int64_t) p * k_row_stride + (int64_t) g * hd;
        krow[i0] = __float2half_rn(x0 * c - x1 * s);
        krow[i1] = __float2half_rn(x0 * s + x1 * c);
        const float * vv = v + (int64_t) g * hd;
        const int col = v_col >= 0 ? v_col : p;  // v_col >= 0: vc already points at this token's column
        vc[((int64_t) g * hd + i0) * v_row_stride + col] = __float2half_rn(vv[i0]);
        vc[((int64_t) g * hd + i1) * v_row_stride + col] = __float2half_rn(vv[i1]);
    }
}
int rope_kv_store2(const float * q, float * q_out, const float * k, const float * v, const int32_t * pos, const float * ff, void * kc, void * vc, int n_heads,
                   int kv_heads, int head_dim, int mode, float freq_base, int64_t k_row_stride, int64_t v_row_stride, int v_col, cudaStream_t st) {
    if (mode != 0 && mode != 2) return B200_ERR_UNSUPPORTED;
    if (head_dim % 2 || head_dim > 2048) return B200_ERR_ARG;
    const float theta_scale = powf(freq_base, -2.0f / head_dim);
    launch_pdl(rope_kv_store_kernel, dim3((unsigned) (n_heads + kv_heads)), dim3(head_dim / 2), 0, st, q, q_out, k, v, pos, ff, (__half *) kc, (__half *) vc, n_heads,
               kv_heads, head_dim, mode, theta_scale, k_row_stride, v_row_stride, v_col);
    return (int) cudaGetLastError();
}
int rope_kv_store(float * q, const float * k, const float * v, const int32_t * pos, const float * ff, void * kc, void * vc, int n_heads, int kv_heads,
                  int head_dim, int mode, float freq_base, int64_t k_row_stride, int64_t v_row_stride, cudaStream_t st) {
    return rope_kv_store2(q, q, k, v, pos, ff, kc, vc, n_heads, kv_heads, head_dim, mode, freq_base, k_row_stride, v_row_stride, -1, st);
}

// ------------------------------------------------------------------------------------------------------------------
// decode attention, 2 launches.  Layouts / semantics as ops.cu attn_* (reference src/layers.cpp:2541-2561).
// ------------------------------------------------------------------------------------------------------------------
// scores[h][t] = scale * sum_d K[t][g][d] * f16(q[h][d]);  a row (HD halves) is covered by HD/8 lanes with 16-byte loads.
// Each CTA covers CH = 128 positions of one KV group and also emits, per query head, the chunk's max and
// sum(exp(s - max)) so that the consumer can normalise without a pass over the whole row.
#define B200_ATTN_CH 128
template <int HD, int GQA>
__global__ void __launch_bounds__(256) attn_scores2_kernel(const float * q, const __half * kc, float * scores,
                                                           float2 * part, int n_kv, int64_t k_row_stride, float scale, int64_t s_stride,
                                                           int nchunks) {
    constexpr int LPR = HD / 8;               // lanes per row
    constexpr int RPW = 32 / LPR;             // rows per warp load
    constexpr int NLD = B200_ATTN_CH / 8 / RPW;  // loads per lane: each warp owns CH/8 = 16 consecutive positions
    __shared__ float sc[GQA][B200_ATTN_CH];
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int part_i = lane % LPR, rsel = lane / LPR;
    float qv[GQA][8];
#pragma unroll
    for (int h = 0; h < GQA; ++h)
#pragma unroll
        for (int e = 0; e < 8; ++e) qv[h][e] = __half2float(__float2half_rn(q[(int64_t) (g * GQA + h) * HD + part_i * 8 + e]));
    const int t_base = blockIdx.x * B200_ATTN_CH;
    const int t_end = min(n_kv, t_base + B200_ATTN_CH);
    // issue every load of this warp first (memory-level parallelism), then do the math
    uint4 raw[NLD];
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int t = t_base + warp * (B200_ATTN_CH / 8) + i * RPW + rsel;
        raw[i] = (t < t_end) ? *reinterpret_cast<const uint4 *>(kc + (int64_t) t * k_row_stride + (int64_t) g * HD + part_i * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
        const int tl = warp * (B200_ATTN_CH / 8) + i * RPW + rsel;
        const __half2 * hp = reinterpret_cast<const __half2 *>(&raw[i]);
        float kv[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { kv[2 * e] = __low2float(hp[e]); kv[2 * e + 1] = __high2float(hp[e]); }
#pragma unroll
        for (int h = 0; h < GQA; ++h) {
            float s = 0.0f;
#pragma unroll
            for (int e = 0; e < 8; ++e) s = fmaf(kv[e], qv[h][e], s);
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (part_i == 0) {
                const float v = (t_base + tl < t_end) ? s * scale : -INFINITY;
                sc[h][tl] = v;
                if (t_base + tl < t_end) scores[(int64_t) (g * GQA + h) * s_stride + t_base + tl] = v;
            }
        }
    }
    __syncthreads();
    // per-head chunk statistics: warp h handles head h (GQA <= 8 warps)
    if (warp < GQA) {
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < B200_ATTN_CH / 32; ++i) mx = fmaxf(mx, sc[warp][lane + 32 * i]);
        mx = warp_max(mx);
        float sum = 0.0f;
#pragma unroll
        for (int i = 0; i < B200_ATTN_CH / 32; ++i) sum += expf(sc[warp][lane + 32 * i] - mx);   // exp(-inf) = 0 for the padded tail
        sum = warp_sum(sum);
        if (lane == 0) part[(int64_t) (g * GQA + warp) * nchunks + blockIdx.x] = make_float2(mx, sum);
    }
}

// softmax normalisation (from the chunk statistics) fused with out = V . P
// grid (HD/8, kv_heads), block 256 = 8 warps, one V^T channel row per warp; P of the group's GQA heads lives in smem.
template <int GQA>
__global__ void __launch_bounds__(256) attn_softmax_pv_kernel(const float * scores, const float2 * part, const __half * vc,
                                                              float * out, int n_kv, int head_dim, int64_t v_row_stride, int64_t s_stride,
                                                              int nchunks) {
    extern __shared__ float P[];  // [GQA][s_stride]
    __shared__ float hmax[GQA], hinv[GQA];
    pdl_launch_dependents();
    const int g = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int d = blockIdx.x * 8 + warp;
    const int ch = g * head_dim + d;
    const __half * vr = vc + (int64_t) ch * v_row_stride;
    // V does not depend on the scores kernel (the KV cache was written two kernels earlier... but by THIS token's
    // rope_kv_store, which the scores kernel already waited for): safe to prefetch only after the dependency resolves
    pdl_wait();
    if (warp < GQA) {
        const float2 * pp = part + (int64_t) (g * GQA + warp) * nchunks;
        float mx = -INFINITY;
        for (int i = lane; i < nchunks; i += 32) mx = fmaxf(mx, pp[i].x);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int i = lane; i < nchunks; i += 32) { const float2 v = pp[i]; sum += v.y * expf(v.x - mx); }
        sum = warp_sum(sum);
        if (lane == 0) { hmax[warp] = mx; hinv[warp] = 1.0f / sum; }
    }
    // first V loads in flight while P is being built
    const int n8 = n_kv & ~7;
    uint4 pre[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int t = lane * 8 + u * 256;
        pre[u] = (t < n8) ? *reinterpret_cast<const uint4 *>(vr + t) : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    {
        // P[h][t] = f16(exp(s - max_h) / sum_h): the group's GQA score rows are contiguous -> one flat float4 stream,
        // loads batched 8 deep so the L2 round trips overlap (a one-load-at-a-time loop costs ~25 us here)
        const float4 * src = reinterpret_cast<const float4 *>(scores + (int64_t) g * GQA * s_stride);
        float4 * dst = reinterpret_cast<float4 *>(P);
        const int total4 = (int) (GQA * s_stride / 4);
        const int row4 = (int) (s_stride / 4);
        for (int j0 = threadIdx.x; j0 < total4; j0 += 8 * 256) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 256;
                v[u] = (j < total4) ? src[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int j = j0 + u * 256;
                if (j < total4) {
                    const int h = j / row4, t = (j - h * row4) * 4;
                    const float mx = hmax[h], inv = hinv[h];
                    float4 o;
                    o.x = (t + 0 < n_kv) ? __half2float(__float2half_rn(expf(v[u].x - mx) * inv)) : 0.0f;
                    o.y = (t + 1 < n_kv) ? __half2float(__float2half_rn(expf(v[u].y - mx) * inv)) : 0.0f;
                    o.z = (t + 2 < n_kv) ? __half2float(__float2half_rn(expf(v[u].z - mx) * inv)) : 0.0f;
                    o.w = (t + 3 < n_kv) ? __half2float(__float2half_rn(expf(v[u].w - mx) * inv)) : 0.0f;
                    dst[j] = o;
                }
            }
        }
    }
    __syncthreads();
    float acc[GQA];
#pragma unroll
    for (int h = 0; h < GQA; ++h) acc[h] = 0.0f;
    for (int t0 = 0; t0 < n8; t0 += 1024) {
        uint4 cur[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = pre[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // prefetch the next trip
            const int t = t0 + 1024 + lane * 8 + u * 256;
            pre[u] = (t < n8) ? *reinterpret_cast<const uint4 *>(vr + t) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = t0 + lane * 8 + u * 256;
            if (t < n8) {
                const __half2 * hp = reinterpret_cast<const __half2 *>(&cur[u]);
                float vv[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) { vv[2 * e] = __low2float(hp[e]); vv[2 * e + 1] = __high2float(hp[e]); }
#pragma unroll
                for (int h = 0; h < GQA; ++h) {
                    const float4 p0 = *reinterpret_cast<const float4 *>(P + (int64_t) h * s_stride + t);
                    const float4 p1 = *reinterpret_cast<const float4 *>(P + (int64_t) h * s_stride + t + 4);
                    acc[h] = fmaf(vv[0], p0.x, acc[h]); acc[h] = fmaf(vv[1], p0.y, acc[h]); acc[h] = fmaf(vv[2], p0.z, acc[h]); acc[h] = fmaf(vv[3], p0.w, acc[h]);
                    acc[h] = fmaf(vv[4], p1.x, acc[h]); acc[h] = fmaf(vv[5], p1.y, acc[h]); acc[h] = fmaf(vv[6], p1.z, acc[h]); acc[h] = fmaf(vv[7], p1.w, acc[h]);
                }
            }
        }
    }
    if (lane < n_kv - n8) {
        const int t = n8 + lane;
        const float vx = __half2float(vr[t]);
#pragma unroll
        for (int h = 0; h < GQA; ++h) acc[h] = fmaf(vx, P[(int64_t) h * s_stride + t], acc[h]);
    }
#pragma unroll
    for (int h = 0; h < GQA; ++h) {
        const float s = warp_sum(acc[h]);
        if (lane == 0) out[(int64_t) (g * GQA + h) * head_dim + d] = s;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// Tensor-core decode attention (mma.sync m16n8k16 f16 x f16 -> f32; HBM-bound op, the legacy warp MMA is plenty).
// Same arithmetic as the reference: f16 operands, fp32 accumulation, P rounded through f16 before V.P.
// The k index of each MMA is permuted (identically for A and B) so that every lane feeds its fragments from
// contiguous 16-byte global loads: lane t of a quad owns elements [8t + 32u, 8t + 32u + 8) of a 128-element k group.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    const __half2 h = __halves2half2(__float2half_rn(lo), __float2half_rn(hi));
    return *reinterpret_cast<const uint32_t *>(&h);
}

// scores[h][t] = scale * K[t][grp] . f16(q[h]) for the GQA heads of KV group `grp`; CTA = 128 positions, warp = 16.
// Also emits per (head, chunk) max and sum(exp(s - max)).
template <int HD, int GQA>
__global__ void __launch_bounds__(256) attn_scores_mma_kernel(const float * q, const __half * kc, float * scores,
                                                              float2 * part, int n_kv, int64_t k_row_stride, float scale, int64_t s_stride,
                                                              int nchunks, int preload) {
    constexpr int NU = HD / 32;  // 16-byte chunks per lane per row
    __shared__ float wmax[8][8], wsum[8][8];
    pdl_launch_dependents();
    // Only the newest position's K row (written by the kernel just before this one) depends on the predecessor: every older row was
    // written by an earlier decode step, hundreds of launches ago.  So all chunks but the last start their K loads BEFORE
    // griddepcontrol.wait and overlap the predecessor's tail; q (the predecessor's output) is read after the wait.
    const bool early = preload && (int) blockIdx.x + 1 < nchunks;
    if (!early) pdl_wait();
    const int grp = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // K fragments first (DRAM latency), then the q fragments (L2) — all loads of the thread in flight together
    const int t_base = blockIdx.x * 128 + warp * 16;
    const int rowA = t_base + g, rowB = rowA + 8;
    uint4 alo[NU], ahi[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        alo[u] = (rowA < n_kv) ? *reinterpret_cast<const uint4 *>(kc + (int64_t) rowA * k_row_stride + (int64_t) grp * HD + 8 * t + 32 * u) : make_uint4(0, 0, 0, 0);
        ahi[u] = (rowB < n_kv) ? *reinterpret_cast<const uint4 *>(kc + (int64_t) rowB * k_row_stride + (int64_t) grp * HD + 8 * t + 32 * u) : make_uint4(0, 0, 0, 0);
    }
    if (early) pdl_wait();
    // B fragments: q of head n = g (zero for the padding heads)
    uint32_t bq[NU][4];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        if (g < GQA) {
            const float * qp = q + (int64_t) (grp * GQA + g) * HD + 8 * t + 32 * u;
            const float4 x0 = *reinterpret_cast<const float4 *>(qp), x1 = *reinterpret_cast<const float4 *>(qp + 4);
            bq[u][0] = pack_h2(x0.x, x0.y); bq[u][1] = pack_h2(x0.z, x0.w); bq[u][2] = pack_h2(x1.x, x1.y); bq[u][3] = pack_h2(x1.z, x1.w);
        } else {
            bq[u][0] = bq[u][1] = bq[u][2] = bq[u][3] = 0u;
        }
    }
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < NU; ++u) {
        mma16816(c, alo[u].x, ahi[u].x, alo[u].y, ahi[u].y, bq[u][0], bq[u][1]);
        mma16816(c, alo[u].z, ahi[u].z, alo[u].w, ahi[u].w, bq[u][2], bq[u][3]);
    }
    // c0 = (rowA, head 2t), c1 = (rowA, head 2t+1), c2 = (rowB, head 2t), c3 = (rowB, head 2t+1)
    const int h0 = 2 * t, h1 = 2 * t + 1;
    float v[4];
    v[0] = (rowA < n_kv) ? c[0] * scale : -INFINITY; v[1] = (rowA < n_kv) ? c[1] * scale : -INFINITY;
    v[2] = (rowB < n_kv) ? c[2] * scale : -INFINITY; v[3] = (rowB < n_kv) ? c[3] * scale : -INFINITY;
    if (h0 < GQA) {
        if (rowA < n_kv) scores[(int64_t) (grp * GQA + h0) * s_stride + rowA] = v[0];
        if (rowB < n_kv) scores[(int64_t) (grp * GQA + h0) * s_stride + rowB] = v[2];
    }
    if (h1 < GQA) {
        if (rowA < n_kv) scores[(int64_t) (grp * GQA + h1) * s_stride + rowA] = v[1];
        if (rowB < n_kv) scores[(int64_t) (grp * GQA + h1) * s_stride + rowB] = v[3];
    }
    // chunk statistics per head: max over the CTA's 128 positions, then sum exp(s - max)
    float m0 = fmaxf(v[0], v[2]), m1 = fmaxf(v[1], v[3]);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o)); }
    if (g == 0) { wmax[warp][h0] = m0; wmax[warp][h1] = m1; }
    __syncthreads();
    float M0 = -INFINITY, M1 = -INFINITY;
#pragma unroll
    for (int w = 0; w < 8; ++w) { M0 = fmaxf(M0, wmax[w][h0]); M1 = fmaxf(M1, wmax[w][h1]); }
    float s0 = (M0 == -INFINITY) ? 0.0f : expf(v[0] - M0) + expf(v[2] - M0);
    float s1 = (M1 == -INFINITY) ? 0.0f : expf(v[1] - M1) + expf(v[3] - M1);
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if (g == 0) { wsum[warp][h0] = s0; wsum[warp][h1] = s1; }
    __syncthreads();
    if (threadIdx.x < GQA) {
        const int h = threadIdx.x;
        float M = -INFINITY, S = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) M = fmaxf(M, wmax[w][h]);
#pragma unroll
        for (int w = 0; w < 8; ++w) S += wsum[w][h];
        part[(int64_t) (grp * GQA + h) * nchunks + blockIdx.x] = make_float2(M, S);
    }
}

// out[h][d] = sum_t Vt[grp*HD + d][t] * P[h][t];  CTA = 16 channels of one KV group, 16 warps split the positions.
// The kernel is latency-bound (few CTAs, short dependent chains), so every phase issues all of its loads up front.
#define B200_PV_WARPS 16
template <int GQA>
__global__ void __launch_bounds__(B200_PV_WARPS * 32) attn_pv_mma_kernel(const float * scores, const float2 * part,
                                                                         const __half * vc, float * out, int n_kv, int head_dim,
                                                                         int64_t v_row_stride, int64_t s_stride, int nchunks, int sp) {
    constexpr int NT = B200_PV_WARPS * 32;
    extern __shared__ __align__(16) unsigned char smraw[];
    __half * Ph = reinterpret_cast<__half *>(smraw);                       // [8][sp]  (rows >= GQA are zero)
    float * red = reinterpret_cast<float *>(smraw + (size_t) 8 * sp * 2);  // [warps][128]
    __shared__ float hmax[8], hinv[8];
    pdl_launch_dependents();
    pdl_wait();
    const int grp = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int npos = ((n_kv + 127) / 128) * 128;
    const int per_row4 = npos / 4;
    // ---- raw scores of the group's heads: all loads of this thread in flight at once (<= 3 float4 per head at 4K context)
    constexpr int MAXV = 3;
    float4 sv[GQA][MAXV];
#pragma unroll
    for (int h = 0; h < GQA; ++h)
#pragma unroll
        for (int u = 0; u < MAXV; ++u) {
            const int i = threadIdx.x + u * NT;
            sv[h][u] = (i < per_row4 && 4 * i < n_kv) ? *reinterpret_cast<const float4 *>(scores + (int64_t) (grp * GQA + h) * s_stride + 4 * i)
                                                      : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }
    if (warp < GQA) {
        const float2 * pp = part + (int64_t) (grp * GQA + warp) * nchunks;
        float mx = -INFINITY;
        for (int i = lane; i < nchunks; i += 32) mx = fmaxf(mx, pp[i].x);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int i = lane; i < nchunks; i += 32) { const float2 pv = pp[i]; sum += pv.y * expf(pv.x - mx); }
        sum = warp_sum(sum);
        if (lane == 0) { hmax[warp] = mx; hinv[warp] = 1.0f / sum; }
    }
    // ---- first V fragments (two position groups per warp) in flight while P is built
    const int c0 = blockIdx.x * 16;
    const __half * rowA = vc + (int64_t) (grp * head_dim + c0 + g) * v_row_stride;
    const __half * rowB = rowA + 8 * v_row_stride;
    const int ngroups = npos / 128;
    auto ldv = [&](const __half * row, int p0) -> uint4 {
        if (p0 >= n_kv) return make_uint4(0, 0, 0, 0);
        uint4 v = *reinterpret_cast<const uint4 *>(row + p0);
        const int r = n_kv - p0;  // valid halves in this chunk
        if (r < 8) {              // tail: zero what lies beyond n_kv (P is zero there, but the cache may hold anything)
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (2 * j + 1 < r) ? w[j] : ((2 * j < r) ? (w[j] & 0xffffu) : 0u);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    };
    uint4 alo[2][4], ahi[2][4];
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
        const int gi = warp + s2 * B200_PV_WARPS;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            alo[s2][u] = (gi < ngroups) ? ldv(rowA, gi * 128 + 8 * t + 32 * u) : make_uint4(0, 0, 0, 0);
            ahi[s2][u] = (gi < ngroups) ? ldv(rowB, gi * 128 + 8 * t + 32 * u) : make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    // ---- P build: Ph[h][t] = f16(exp(s - max_h) * inv_h); zero beyond n_kv and for the padding heads
#pragma unroll
    for (int h = 0; h < GQA; ++h) {
        const float mx = hmax[h], inv = hinv[h];
#pragma unroll
        for (int u = 0; u < MAXV; ++u) {
            const int i = threadIdx.x + u * NT;
            if (i < per_row4) {
                const int p4 = 4 * i;
                // __expf (ex2.approx, ~2 ulp): the value is rounded to f16 (11 bits) right away
                const float e0 = (p4 + 0 < n_kv) ? __expf(sv[h][u].x - mx) * inv : 0.0f;
                const float e1 = (p4 + 1 < n_kv) ? __expf(sv[h][u].y - mx) * inv : 0.0f;
                const float e2 = (p4 + 2 < n_kv) ? __expf(sv[h][u].z - mx) * inv : 0.0f;
                const float e3 = (p4 + 3 < n_kv) ? __expf(sv[h][u].w - mx) * inv : 0.0f;
                uint2 pk;
                pk.x = pack_h2(e0, e1); pk.y = pack_h2(e2, e3);
                *reinterpret_cast<uint2 *>(Ph + (size_t) h * sp + p4) = pk;
            }
        }
    }
    for (int h = GQA; h < 8; ++h)
        for (int i = threadIdx.x; i < per_row4; i += NT) *reinterpret_cast<uint2 *>(Ph + (size_t) h * sp + 4 * i) = make_uint2(0u, 0u);
    // contexts longer than MAXV * NT * 4 positions: remaining score columns (rare; keeps the kernel general)
    for (int h = 0; h < GQA; ++h)
        for (int i = threadIdx.x + MAXV * NT; i < per_row4; i += NT) {
            const int p4 = 4 * i;
            const float mx = hmax[h], inv = hinv[h];
            float e[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) e[j] = (p4 + j < n_kv) ? __expf(scores[(int64_t) (grp * GQA + h) * s_stride + p4 + j] - mx) * inv : 0.0f;
            uint2 pk;
            pk.x = pack_h2(e[0], e[1]); pk.y = pack_h2(e[2], e[3]);
            *reinterpret_cast<uint2 *>(Ph + (size_t) h * sp + p4) = pk;
        }
    __syncthreads();
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
        const int gi = warp + s2 * B200_PV_WARPS;
        if (gi < ngroups) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 b = *reinterpret_cast<const uint4 *>(Ph + (size_t) g * sp + gi * 128 + 8 * t + 32 * u);
                mma16816(c, alo[s2][u].x, ahi[s2][u].x, alo[s2][u].y, ahi[s2][u].y, b.x, b.y);
                mma16816(c, alo[s2][u].z, ahi[s2][u].z, alo[s2][u].w, ahi[s2][u].w, b.z, b.w);
            }
        }
    }
    for (int gi = warp + 2 * B200_PV_WARPS; gi < ngroups; gi += B200_PV_WARPS) {   // contexts beyond 4096 positions
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint4 lo = ldv(rowA, gi * 128 + 8 * t + 32 * u), hi = ldv(rowB, gi * 128 + 8 * t + 32 * u);
            const uint4 b = *reinterpret_cast<const uint4 *>(Ph + (size_t) g * sp + gi * 128 + 8 * t + 32 * u);
            mma16816(c, lo.x, hi.x, lo.y, hi.y, b.x, b.y);
            mma16816(c, lo.z, hi.z, lo.w, hi.w, b.z, b.w);
        }
    }
    // c0 = (ch g, head 2t), c1 = (ch g, head 2t+1), c2 = (ch g+8, head 2t), c3 = (ch g+8, head 2t+1)
    float * rw = red + warp * 128;
    rw[g * 8 + 2 * t] = c[0]; rw[g * 8 + 2 * t + 1] = c[1]; rw[(g + 8) * 8 + 2 * t] = c[2]; rw[(g + 8) * 8 + 2 * t + 1] = c[3];
    __syncthreads();
    if (threadIdx.x < 128) {
        const int m = threadIdx.x >> 3, n = threadIdx.x & 7;
        if (n < GQA) {
            float s = 0.0f;
#pragma unroll
            for (int w = 0; w < B200_PV_WARPS; ++w) s += red[w * 128 + threadIdx.x];
            out[(int64_t) (grp * GQA + n) * head_dim + c0 + m] = s;
        }
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Position-split V.P (third design; replaces attn_pv_mma_kernel on the default path).  The channel-split kernel above uses
// only HD/16 x kv_heads = 64 CTAs and every one of them rebuilds P for ALL positions of its group from the raw scores.
// Here a CTA owns `span` consecutive positions of one KV group: it builds P only for those (nothing is recomputed anywhere),
// multiplies the [HD x span] slab of the transposed V cache with it on the tensor cores (warp w = channels 16w .. 16w+15) and
// writes an fp32 partial [GQA heads][HD].  nsplit x kv_heads CTAs (136 for Llama-3-8B at 4K context: one wave of the 148
// SMs); the partials are summed in split order by the tail kernel (sum_partials_quant), which also emits the quantized
// activations of the o-projection, so the launch count of the layer does not change.
// Arithmetic is unchanged: P = f16(exp(s - max) / sum) with the GLOBAL max / sum (from the scores kernel's chunk statistics).
// ------------------------------------------------------------------------------------------------------------------
#define B200_PVS_ROUND 256  // positions per inner round (two 128-position k groups held in registers)
template <int HD, int GQA>
__global__ void __launch_bounds__(HD * 2) attn_pv_split_kernel(const float * scores, const float2 * part,
                                                               const __half * vc, float * partial, int n_kv,
                                                               int64_t v_row_stride, int64_t s_stride, int nchunks, int span, int n_heads, int preload) {
    constexpr int NT = HD * 2;
    constexpr int SP = B200_PVS_ROUND + 32;  // +64 bytes: the 8 head rows land in different bank groups (conflict-free 16-byte B loads)
    __shared__ __align__(16) __half Ph[8 * SP];
    __shared__ float hmax[8], hinv[8];
    pdl_launch_dependents();
    const int grp = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int p_begin = blockIdx.x * span;
    const int p_end = min(n_kv, p_begin + span);
    // padding head rows of P stay zero for the whole kernel
    for (int i = threadIdx.x; i < (8 - GQA) * SP / 2; i += NT) reinterpret_cast<uint32_t *>(Ph + GQA * SP)[i] = 0u;

    const int c0 = warp * 16;
    const __half * rowA = vc + (int64_t) (grp * HD + c0 + g) * v_row_stride;
    const __half * rowB = rowA + 8 * v_row_stride;
    auto ldv = [&](const __half * row, int p0) -> uint4 {
        if (p0 >= p_end) return make_uint4(0, 0, 0, 0);
        uint4 v = *reinterpret_cast<const uint4 *>(row + p0);
        const int r = p_end - p0;  // valid halves in this 16-byte chunk
        if (r < 8) {               // tail: zero what lies beyond n_kv (P is zero there, but the cache may hold anything, NaN included)
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (2 * j + 1 < r) ? w[j] : ((2 * j < r) ? (w[j] & 0xffffu) : 0u);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    };
    uint4 alo[2][4], ahi[2][4];
    auto load_round = [&](int p0) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                alo[s2][u] = ldv(rowA, p0 + s2 * 128 + 8 * t + 32 * u);
                ahi[s2][u] = ldv(rowB, p0 + s2 * 128 + 8 * t + 32 * u);
            }
    };

    // the scores / statistics come from the predecessor and the newest V column from the kernel before it; every older column was
    // written by an earlier decode step, so all splits but the last stream their V slab while the scores kernel is still running
    const bool early = preload && p_end < n_kv;
    if (early) load_round(p_begin);
    pdl_wait();
    if (!early) load_round(p_begin);  // DRAM loads first, the (L2-resident) scores behind them
    for (int h = warp; h < GQA; h += NT / 32) {  // HD = 64 has 4 warps for up to 8 heads
        const float2 * pp = part + (int64_t) (grp * GQA + h) * nchunks;
        float mx = -INFINITY;
        for (int i = lane; i < nchunks; i += 32) mx = fmaxf(mx, pp[i].x);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int i = lane; i < nchunks; i += 32) { const float2 pv = pp[i]; sum += pv.y * expf(pv.x - mx); }
        sum = warp_sum(sum);
        if (lane == 0) { hmax[h] = mx; hinv[h] = 1.0f / sum; }
    }
    __syncthreads();

    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = p_begin; p0 < p_end; p0 += B200_PVS_ROUND) {
        if (p0 != p_begin) { __syncthreads(); load_round(p0); }  // every warp is done with the previous round's P
        // ---- P of this round: Ph[h][i] = f16(exp(s - max_h) * inv_h), zero beyond n_kv
        for (int i = threadIdx.x; i < B200_PVS_ROUND; i += NT) {
            const int pp0 = p0 + i;
#pragma unroll
            for (int h = 0; h < GQA; ++h) {
                // __expf (ex2.approx, ~2 ulp): the value is rounded to f16 (11 bits) right away
                const float e = (pp0 < p_end) ? __expf(scores[(int64_t) (grp * GQA + h) * s_stride + pp0] - hmax[h]) * hinv[h] : 0.0f;
                Ph[h * SP + i] = __float2half_rn(e);
            }
        }
        __syncthreads();
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
            if (p0 + s2 * 128 < p_end) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint4 b = *reinterpret_cast<const uint4 *>(Ph + g * SP + s2 * 128 + 8 * t + 32 * u);
                    mma16816(c, alo[s2][u].x, ahi[s2][u].x, alo[s2][u].y, ahi[s2][u].y, b.x, b.y);
                    mma16816(c, alo[s2][u].z, ahi[s2][u].z, alo[s2][u].w, ahi[s2][u].w, b.z, b.w);
                }
            }
        }
    }
    // c0 = (ch g, head 2t), c1 = (ch g, head 2t+1), c2 = (ch g+8, head 2t), c3 = (ch g+8, head 2t+1)
    float * po = partial + ((int64_t) blockIdx.x * n_heads + (int64_t) grp * GQA) * HD + c0;
    if (2 * t < GQA) { po[(int64_t) (2 * t) * HD + g] = c[0]; po[(int64_t) (2 * t) * HD + g + 8] = c[2]; }
    if (2 * t + 1 < GQA) { po[(int64_t) (2 * t + 1) * HD + g] = c[1]; po[(int64_t) (2 * t + 1) * HD + g + 8] = c[3]; }
}


// ------------------------------------------------------------------------------------------------------------------
// Cluster V.P (OPT-IN, B200_ATTN_CLUSTER=1; written at the end of round 1, not yet run on a GPU).  The position splits of one KV group
// form a THREAD-BLOCK CLUSTER (up to 16 CTAs, non-portable size allowed on sm_100): each CTA multiplies its 512-position slab of V with
// its slice of P exactly like attn_pv_split_kernel (16 warps: two 256-position rounds side by side, every V load in flight at once),
// leaves its fp32 partial [GQA][HD] in its own shared memory, and after a cluster barrier rank 0 sums the partials of all ranks IN RANK
// ORDER through distributed shared memory, writes the group's slice of the attention output and — because a KV group owns whole
// 256-element (Q8_K) / 32-element (Q8_0) quantization blocks of that vector (GQA*HD % 256 == 0) — also its slice of the o-projection's
// quantized activations.  The single-CTA summing/quantizing tail launch (3.4 us of the 55 us layer) disappears: attention = 2 launches.
// Requires GQA * HD % 256 == 0 and n_kv <= 16 * 512 (longer contexts keep the scratch-partial path).
// ------------------------------------------------------------------------------------------------------------------
#define B200_PVC_SPAN 512
template <int HD, int GQA, bool Q8K>
__global__ void __launch_bounds__(HD * 4) attn_pv_cluster_kernel(const float * scores, const float2 * part, const __half * vc, float * out, uint8_t * qact,
                                                                 int n_kv, int64_t v_row_stride, int64_t s_stride, int nchunks, int n_heads, int preload) {
    namespace cg = cooperative_groups;
    constexpr int NT = HD * 4;        // 2 rounds x HD/16 channel warps x 32
    constexpr int NWC = HD / 16;      // channel warps per round
    constexpr int SP = B200_PVC_SPAN + 32;  // row stride in halves: 64 bytes past a multiple of 128 -> conflict-free 16-byte B loads
    constexpr int NE = GQA * HD;      // outputs of this KV group
    static_assert(NE % 256 == 0, "a KV group must own whole quantization blocks");
    __shared__ __align__(16) __half Ph[8 * SP];
    __shared__ __align__(16) float part_s[NE];   // this CTA's partial [head][channel]; read by rank 0 through DSMEM
    __shared__ float hmax[8], hinv[8];
    __shared__ unsigned long long wkey[NT / 32];
    __shared__ float bmaxs[NE / 256];
    cg::cluster_group cluster = cg::this_cluster();
    pdl_launch_dependents();
    const int grp = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int rnd = warp / NWC, cw = warp % NWC;  // which 256-position round of the slab, which 16 channels
    const int p_begin = blockIdx.x * B200_PVC_SPAN;
    const int p_end = min(n_kv, p_begin + B200_PVC_SPAN);
    for (int i = threadIdx.x; i < (8 - GQA) * SP / 2; i += NT) reinterpret_cast<uint32_t *>(Ph + GQA * SP)[i] = 0u;

    const int c0 = cw * 16;
    const __half * rowA = vc + (int64_t) (grp * HD + c0 + g) * v_row_stride;
    const __half * rowB = rowA + 8 * v_row_stride;
    auto ldv = [&](const __half * row, int p0) -> uint4 {
        if (p0 >= p_end) return make_uint4(0, 0, 0, 0);
        uint4 v = *reinterpret_cast<const uint4 *>(row + p0);
        const int r = p_end - p0;
        if (r < 8) {
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (2 * j + 1 < r) ? w[j] : ((2 * j < r) ? (w[j] & 0xffffu) : 0u);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    };
    uint4 alo[2][4], ahi[2][4];
    const int pr = p_begin + rnd * 256;  // this warp's 256 positions
    auto load_round = [&]() {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                alo[s2][u] = ldv(rowA, pr + s2 * 128 + 8 * t + 32 * u);
                ahi[s2][u] = ldv(rowB, pr + s2 * 128 + 8 * t + 32 * u);
            }
    };
    const bool early = preload && p_end < n_kv;  // only the last slab can hold the newest position (see attn_pv_split_kernel)
    if (early) load_round();
    pdl_wait();
    if (!early) load_round();
    for (int h = warp; h < GQA; h += NT / 32) {
        const float2 * pp = part + (int64_t) (grp * GQA + h) * nchunks;
        float mx = -INFINITY;
        for (int i = lane; i < nchunks; i += 32) mx = fmaxf(mx, pp[i].x);
        mx = warp_max(mx);
        float sum = 0.0f;
        for (int i = lane; i < nchunks; i += 32) { const float2 pv = pp[i]; sum += pv.y * expf(pv.x - mx); }
        sum = warp_sum(sum);
        if (lane == 0) { hmax[h] = mx; hinv[h] = 1.0f / sum; }
    }
    __syncthreads();
    // ---- P of the slab: Ph[h][i] = f16(exp(s - max_h) * inv_h), zero beyond n_kv  (same expression as attn_pv_split_kernel)
    for (int i = threadIdx.x; i < B200_PVC_SPAN; i += NT) {
        const int pp0 = p_begin + i;
#pragma unroll
        for (int h = 0; h < GQA; ++h) {
            const float e = (pp0 < p_end) ? __expf(scores[(int64_t) (grp * GQA + h) * s_stride + pp0] - hmax[h]) * hinv[h] : 0.0f;
            Ph[h * SP + i] = __float2half_rn(e);
        }
    }
    __syncthreads();
    float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
        if (pr + s2 * 128 < p_end) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint4 b = *reinterpret_cast<const uint4 *>(Ph + g * SP + rnd * 256 + s2 * 128 + 8 * t + 32 * u);
                mma16816(c, alo[s2][u].x, ahi[s2][u].x, alo[s2][u].y, ahi[s2][u].y, b.x, b.y);
                mma16816(c, alo[s2][u].z, ahi[s2][u].z, alo[s2][u].w, ahi[s2][u].w, b.z, b.w);
            }
        }
    }
    // ---- CTA partial: round 1 deposits, round 0 adds.  c0 = (ch g, head 2t), c1 = (ch g, head 2t+1), c2 = (ch g+8, head 2t), c3 = (ch g+8, head 2t+1)
    float * pw = part_s + c0;
    if (rnd == 1) {
        if (2 * t < GQA) { pw[(2 * t) * HD + g] = c[0]; pw[(2 * t) * HD + g + 8] = c[2]; }
        if (2 * t + 1 < GQA) { pw[(2 * t + 1) * HD + g] = c[1]; pw[(2 * t + 1) * HD + g + 8] = c[3]; }
    }
    __syncthreads();
    if (rnd == 0) {
        if (2 * t < GQA) { pw[(2 * t) * HD + g] += c[0]; pw[(2 * t) * HD + g + 8] += c[2]; }
        if (2 * t + 1 < GQA) { pw[(2 * t + 1) * HD + g] += c[1]; pw[(2 * t + 1) * HD + g + 8] += c[3]; }
    }
    cluster.sync();   // every rank's partial is complete and visible cluster-wide
    if (cluster.block_rank() == 0) {
        const unsigned nr = cluster.num_blocks();
        const ActLayout L = act_layout(Q8K, (int64_t) n_heads * HD);
        const int64_t ebase = (int64_t) grp * NE;   // first element of this group in the [n_heads * HD] output vector
        for (int e0 = 0; e0 < NE; e0 += NT) {        // NT == NE for GQA = 4 (one pass); the barriers below are uniform
            const int e = e0 + threadIdx.x;
            const bool on = e < NE;
            float v = 0.0f;
            if (on) {
                for (unsigned r = 0; r < nr; ++r) v += cluster.map_shared_rank(part_s, r)[e];   // rank order: deterministic
                out[ebase + e] = v;
            }
            if (qact == nullptr) continue;
            if (Q8K) {
                // quantize_row_q8_K_ref (ggml-quants.c:2555-2592) as quantize_q8_K_kernel: 256 consecutive threads = one block
                unsigned long long key = on ? (((unsigned long long) __float_as_uint(fabsf(v)) << 32) | (unsigned) (255 - (e & 255))) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                    key = other > key ? other : key;
                }
                __syncthreads();
                if (lane == 0) wkey[warp] = key;
                __syncthreads();
                const int b8 = (threadIdx.x >> 8) * 8;  // first warp of this thread's 256-element block
                unsigned long long kk = wkey[b8];
#pragma unroll
                for (int w = 1; w < 8; ++w) kk = wkey[b8 + w] > kk ? wkey[b8 + w] : kk;
                const int idx = 255 - (int) (kk & 0xffffffffu);
                if (on && (e & 255) == idx) bmaxs[e >> 8] = v;
                __syncthreads();
                if (on) {
                    const float mx = bmaxs[e >> 8];
                    int qv = 0;
                    float dd = 0.0f;
                    if (mx != 0.0f) {
                        const float iscale = __fdiv_rn(-127.f, mx);
                        qv = min(127, __float2int_rn(__fmul_rn(iscale, v)));
                        dd = __fdiv_rn(1.0f, iscale);
                    }
                    const int64_t ge = ebase + e;
                    reinterpret_cast<int8_t *>(qact)[act_qs_off_q8k(ge)] = (int8_t) qv;
                    int sq = qv;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    if (lane == 0) reinterpret_cast<int16_t *>(qact + L.bs_off)[ge >> 5] = (int16_t) sq;
                    if ((e & 255) == 0) reinterpret_cast<float *>(qact + L.d_off)[ge >> 8] = dd;
                }
            } else {
                // x86 quantize_row_q8_0 (arch/x86/quants.c:290-384) as add_rmsnorm_quant_kernel: one warp = one 32-element block
                float amax = on ? fabsf(v) : 0.0f;
                amax = warp_max(amax);
                const float dv = __fdiv_rn(amax, 127.f);
                const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
                const int qv = __float2int_rn(__fmul_rn(v, id));
                int sq = qv;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (on) {
                    const int64_t ge = ebase + e;
                    reinterpret_cast<int8_t *>(qact)[act_qs_off_q80(ge)] = (int8_t) qv;
                    if (lane == 0) {
                        reinterpret_cast<float *>(qact + L.d_off)[ge >> 5] = __half2float(__float2half_rn(dv));
                        reinterpret_cast<int *>(qact + L.bs_off)[ge >> 5] = sq;
                    }
                }
            }
        }
        // the last group also zeroes the padding of the code area (k not a multiple of 1024) so the GEMV can read it blindly
        if (qact != nullptr && Q8K && grp == (int) gridDim.y - 1) {
            const int64_t k = (int64_t) n_heads * HD;
            for (int64_t e = k + threadIdx.x; e < L.qs_bytes; e += NT) reinterpret_cast<int8_t *>(qact)[act_qs_off_q8k(e)] = 0;
        }
    }
    cluster.sync();   // no rank may exit while rank 0 is still reading its shared memory
}

// positions per CTA of the split V.P kernel: 256 up to 8K context (<= 32 splits), then grown in steps of 256
static int pv_span(int n_kv) {
    int span = B200_PVS_ROUND;
    while ((n_kv + span - 1) / span > 32) span += B200_PVS_ROUND;
    return span;
}

template <int HD, int GQA>
static int attn_decode_mma_t(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_kv, int kv_heads, int64_t k_row_stride,
                             int64_t v_row_stride, float scale, int wtype, void * qact, int preload_arg, int cluster_arg, cudaStream_t st) {
    const int64_t s_stride = (n_kv + 7) & ~7;
    const int nchunks = (n_kv + 127) / 128;
    float2 * part = reinterpret_cast<float2 *>(scratch + (int64_t) kv_heads * GQA * s_stride);
    // B200_ATTN_PRELOAD=0: every CTA waits for its predecessor before its first load (A/B aid).  Contract of the default: cache positions
    // below n_kv - 1 were written by earlier decode steps, not by the launches immediately preceding this call (include/chatllm_b200.h).
    static const int preload_env = getenv("B200_ATTN_PRELOAD") ? atoi(getenv("B200_ATTN_PRELOAD")) : 1;
    const int preload = preload_arg >= 0 ? (preload_arg && preload_env) : preload_env;  // the caller knows when older cache rows were just rewritten
    launch_pdl(attn_scores_mma_kernel<HD, GQA>, dim3((unsigned) nchunks, (unsigned) kv_heads), dim3(256), 0, st, q, (const __half *) kc, scratch, part, n_kv,
               k_row_stride, scale, s_stride, nchunks, preload);
    static const int old_pv = getenv("B200_ATTN_OLD_PV") ? atoi(getenv("B200_ATTN_OLD_PV")) : 0;  // A/B aid: the channel-split kernel
    const int n_heads = kv_heads * GQA;
    if constexpr ((GQA * HD) % 256 == 0) {
        static const int use_cluster = getenv("B200_ATTN_CLUSTER") ? atoi(getenv("B200_ATTN_CLUSTER")) : 1;  // measured r02: 9.45 vs 12.1 us per layer at 4K context (B200_ATTN_CLUSTER=0: split V.P + tail launch)
        const int nsplit_c = (n_kv + B200_PVC_SPAN - 1) / B200_PVC_SPAN;
        const bool q8k = wtype == B200_TYPE_Q4_K;
        if ((cluster_arg >= 0 ? cluster_arg : use_cluster) && nsplit_c <= 16 && (!qact || q8k || wtype == B200_TYPE_Q4_0 || wtype == B200_TYPE_Q8_0)) {
            auto kern = q8k ? attn_pv_cluster_kernel<HD, GQA, true> : attn_pv_cluster_kernel<HD, GQA, false>;
            static bool nonportable_dev[16] = {false};
            int dev = 0;
            cudaGetDevice(&dev);
            if (nsplit_c > 8 && !nonportable_dev[dev & 15]) {  // clusters of 9..16 CTAs need the opt-in attribute (both quantizer variants)
                cudaError_t e = cudaFuncSetAttribute(attn_pv_cluster_kernel<HD, GQA, true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_pv_cluster_kernel<HD, GQA, false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                if (e != cudaSuccess) return (int) e;
                nonportable_dev[dev & 15] = true;
            }
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned) nsplit_c, (unsigned) kv_heads);
            cfg.blockDim = dim3(HD * 4);
            cfg.dynamicSmemBytes = 0;
            cfg.stream = st;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = (unsigned) nsplit_c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = pdl_enabled() ? 2 : 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const float *) scratch, (const float2 *) part, (const __half *) vc, out, (uint8_t *) qact, n_kv, v_row_stride,
                                               s_stride, nchunks, n_heads, preload);
            return (int) e;
        }
    }
    if (!old_pv && ((int64_t) n_heads * HD) % 256 == 0 && (int64_t) n_heads * HD <= 20480) {
        const int span = pv_span(n_kv);
        const int nsplit = (n_kv + span - 1) / span;
        // partial sums live behind the scores and the chunk statistics (attn_decode2_scratch_bytes)
        float * partial = scratch + (((int64_t) n_heads * s_stride + 2 * (int64_t) n_heads * nchunks + 15) & ~(int64_t) 15);
        launch_pdl(attn_pv_split_kernel<HD, GQA>, dim3((unsigned) nsplit, (unsigned) kv_heads), dim3(HD * 2), 0, st, (const float *) scratch, (const float2 *) part,
                   (const __half *) vc, partial, n_kv, v_row_stride, s_stride, nchunks, span, n_heads, preload);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int) e;
        return sum_partials_quant(wtype, partial, nsplit, out, qact, (int64_t) n_heads * HD, st);
    }
    const int npos = nchunks * 128;
    const int sp = npos + 32;  // +64 bytes: the 8 head rows land in different bank groups (conflict-free 16-byte B loads)
    const size_t smem = (size_t) 8 * sp * 2 + B200_PV_WARPS * 128 * 4;
    static size_t configured_dev[16] = {0};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t & configured = configured_dev[dev & 15];
    if (smem > configured) {
        const size_t want = smem > 160 * 1024 ? smem : 160 * 1024;
        cudaError_t e = cudaFuncSetAttribute(attn_pv_mma_kernel<GQA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) want);
        if (e != cudaSuccess) return (int) e;
        configured = want;
    }
    if (smem > 227 * 1024) return B200_ERR_UNSUPPORTED;
    launch_pdl(attn_pv_mma_kernel<GQA>, dim3((unsigned) (HD / 16), (unsigned) kv_heads), dim3(B200_PV_WARPS * 32), smem, st, (const float *) scratch, (const float2 *) part,
               (const __half *) vc, out, n_kv, HD, v_row_stride, s_stride, nchunks, sp);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int) e;
    if (qact) return quantize_act(wtype, out, (int64_t) n_heads * HD, (int64_t) n_heads * HD, 1, qact, st);
    return B200_OK;
}

template <int HD, int GQA>
static int attn_decode2_t(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_kv, int kv_heads, int64_t k_row_stride,
                          int64_t v_row_stride, float scale, cudaStream_t st) {
    const int64_t s_stride = (n_kv + 7) & ~7;
    const int nchunks = (n_kv + B200_ATTN_CH - 1) / B200_ATTN_CH;
    float2 * part = reinterpret_cast<float2 *>(scratch + (int64_t) kv_heads * GQA * s_stride);
    dim3 g1((unsigned) nchunks, (unsigned) kv_heads);
    launch_pdl(attn_scores2_kernel<HD, GQA>, g1, dim3(256), 0, st, q, (const __half *) kc, scratch, part, n_kv, k_row_stride, scale, s_stride, nchunks);
    const size_t smem = (size_t) GQA * s_stride * 4;
    static size_t configured_dev[16] = {0};  // function attributes are per device
    int dev = 0;
    cudaGetDevice(&dev);
    size_t & configured = configured_dev[dev & 15];
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_softmax_pv_kernel<GQA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (smem > 200 * 1024 ? smem : 200 * 1024));
        if (e != cudaSuccess) return (int) e;
        configured = smem > 200 * 1024 ? smem : 200 * 1024;
    }
    if (smem > 227 * 1024) return B200_ERR_UNSUPPORTED;
    dim3 g2((unsigned) (HD / 8), (unsigned) kv_heads);
    launch_pdl(attn_softmax_pv_kernel<GQA>, g2, dim3(256), smem, st, scratch, (const float2 *) part, (const __half *) vc, out, n_kv, HD, v_row_stride, s_stride, nchunks);
    return (int) cudaGetLastError();
}

size_t attn_decode2_scratch_bytes(int n_heads, int n_kv) {
    // raw scores [n_heads][n_kv] + chunk statistics + (split V.P) up to 32 partial outputs of at most 20480 floats
    return (size_t) n_heads * (size_t) ((n_kv + 7) & ~7) * 4 + (size_t) n_heads * (size_t) ((n_kv + B200_ATTN_CH - 1) / B200_ATTN_CH) * 8 + 128 +
           (size_t) 32 * 20480 * 4;
}

int attn_decode2(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                 int64_t k_row_stride, int64_t v_row_stride, float scale, cudaStream_t st) {
    return attn_decode3(q, kc, vc, out, scratch, n_heads, kv_heads, head_dim, n_kv, k_row_stride, v_row_stride, scale, 0, nullptr, st);
}

// as attn_decode2, and additionally qact = the output quantized as the activations of a following matmul with weight type wtype
// the (head_dim, heads per KV head) pairs attn_decode3 is instantiated for
bool attn_decode3_supported(int n_heads, int kv_heads, int head_dim, int64_t k_row_stride, int64_t v_row_stride) {
    if (kv_heads <= 0 || n_heads % kv_heads || (k_row_stride % 8) || (v_row_stride % 8)) return false;
    const int gqa = n_heads / kv_heads;
    if (head_dim == 128) return gqa == 4 || gqa == 7 || gqa == 1 || gqa == 8 || gqa == 2;
    if (head_dim == 64) return gqa == 8 || gqa == 4 || gqa == 2 || gqa == 1;
    return false;
}

int attn_decode3(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                 int64_t k_row_stride, int64_t v_row_stride, float scale, int wtype, void * qact, cudaStream_t st, int preload, int cluster) {
    if (n_kv <= 0) return B200_OK;
    if (n_heads % kv_heads) return B200_ERR_ARG;
    if ((k_row_stride % 8) || (v_row_stride % 8)) return B200_ERR_UNSUPPORTED;  // 16-byte row loads
    const int gqa = n_heads / kv_heads;
    static const int no_mma = getenv("B200_ATTN_NO_MMA") ? atoi(getenv("B200_ATTN_NO_MMA")) : 0;
#define B200_ATTN2(HD_, G_)                                                                                                                       \
    if (head_dim == HD_ && gqa == G_) {                                                                                                           \
        if (!no_mma) return attn_decode_mma_t<HD_, G_>(q, kc, vc, out, scratch, n_kv, kv_heads, k_row_stride, v_row_stride, scale, wtype, qact, preload, cluster, st); \
        const int rc_ = attn_decode2_t<HD_, G_>(q, kc, vc, out, scratch, n_kv, kv_heads, k_row_stride, v_row_stride, scale, st);                  \
        if (rc_ || !qact) return rc_;                                                                                                             \
        return quantize_act(wtype, out, (int64_t) n_heads * head_dim, (int64_t) n_heads * head_dim, 1, qact, st);                                 \
    }
    B200_ATTN2(128, 4) B200_ATTN2(128, 7) B200_ATTN2(128, 1) B200_ATTN2(128, 8) B200_ATTN2(128, 2)
    B200_ATTN2(64, 8) B200_ATTN2(64, 4) B200_ATTN2(64, 2) B200_ATTN2(64, 1)
#undef B200_ATTN2
    return B200_ERR_UNSUPPORTED;
}

}  // namespace b200
