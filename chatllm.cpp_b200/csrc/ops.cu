// ops.cu — the fp32 glue ops around the quantized matmuls (RMSNorm, RoPE, softmax, SwiGLU, residual add,
// embedding gather, F16 attention matmuls, KV-cache writes).  Each mirrors the arithmetic of the reference CPU
// op (file:line cited per kernel) closely enough that decode logits agree to ~1e-5 relative.
#include "common.cuh"
#include "dequant.cuh"
#include "kernels.h"

#include <math.h>

namespace b200 {

__device__ __forceinline__ float block_sum(float v, float * red /*[32]*/) {
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
__device__ __forceinline__ float block_max(float v, float * red) {
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

// ---- RMSNorm (+ weight multiply) ------------------------------------------------------------------------
// reference: ggml_compute_forward_rms_norm_f32 ggml/src/ggml-cpu/ops.cpp:3710-3758 followed by ggml_mul
// (RMSNorm::forward src/layers.cpp:2216-2225):  y = (x * (1/sqrt(mean(x^2)+eps))) * w
__global__ void __launch_bounds__(1024) rms_norm_mul_kernel(const float * x, const float * w, float * y,
                                                            int64_t ne0, float eps) {
    __shared__ double red[32];
    pdl_launch_dependents();
    pdl_wait();
    const float * xr = x + (int64_t) blockIdx.x * ne0;
    float * yr = y + (int64_t) blockIdx.x * ne0;
    double s = 0.0;   // sum += (ggml_float)(x*x): float product, double accumulation (ops.cpp:3736-3741)
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) { const float v = xr[i]; s += (double) __fmul_rn(v, v); }
    s = block_sum_double(s, red);
    const float mean = (float) (s / (double) ne0);
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) {
        const float v = xr[i] * scale;
        yr[i] = w ? v * w[i] : v;
    }
}
int rms_norm_mul(const float * x, const float * w, float * y, int64_t ne0, int64_t nrows, float eps, cudaStream_t st) {
    if (nrows <= 0) return B200_OK;
    const int threads = ne0 >= 4096 ? 1024 : (ne0 >= 1024 ? 512 : 256);
    launch_pdl(rms_norm_mul_kernel, dim3((unsigned) nrows), dim3(threads), 0, st, x, w, y, ne0, eps);
    return (int) cudaGetLastError();
}

// ---- elementwise ----------------------------------------------------------------------------------------
__global__ void add_kernel(const float * a, const float * b, float * y, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a[i] + b[i];
}
int add_f32(const float * a, const float * b, float * y, int64_t n, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(add_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, a, b, y, n);
    return (int) cudaGetLastError();
}
// SwiGLU of BaseMLP::forward (src/layers.cpp:2475-2483): silu(gate) * up, silu = x/(1+exp(-x)) (vec.h:1061)
__global__ void silu_mul_kernel(const float * g, const float * u, float * y, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = g[i]; y[i] = (v / (1.0f + expf(-v))) * u[i]; }
}
int silu_mul(const float * gate, const float * up, float * y, int64_t n, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(silu_mul_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, gate, up, y, n);
    return (int) cudaGetLastError();
}

// ---- RoPE -------------------------------------------------------------------------------------------------
// reference: ggml_compute_forward_rope_flt ggml/src/ggml-cpu/ops.cpp:5720-5865; cache init :5613-5628 builds the
// angle by an fp32 recurrence seeded with the position (theta = pos; use; theta *= theta_scale) — replayed here per
// (token, pair) so cos/sin see the same fp32 angle as the CPU (SURVEY.md §8a parity note); YaRN :5587-5611.
struct RopeParams {
    int64_t ne0, n_heads, n_tokens, xs_h, xs_t, ys_h, ys_t;
    int n_dims, mode;
    float theta_scale, freq_scale, ext_factor, attn_factor, corr0, corr1;
};
__global__ void rope_kernel(const float * x, float * y, const int32_t * pos,
                            const float * ff, const RopeParams p) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t t = blockIdx.x;
    const int half = p.n_dims / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        float theta = (float) pos[t];
        for (int j = 0; j < i; ++j) theta *= p.theta_scale;
        const float f = ff ? ff[i] : 1.0f;
        const float theta_extrap = theta / f;
        const float theta_interp = p.freq_scale * theta_extrap;
        float th = theta_interp, mscale = p.attn_factor;
        if (p.ext_factor != 0.0f) {
            const float yv = ((float) i - p.corr0) / fmaxf(0.001f, p.corr1 - p.corr0);
            const float ramp = (1.0f - fminf(1.0f, fmaxf(0.0f, yv))) * p.ext_factor;
            th = theta_interp * (1.0f - ramp) + theta_extrap * ramp;
            mscale *= 1.0f + 0.1f * logf(1.0f / p.freq_scale);
        }
        const float c = cosf(th) * mscale, s = sinf(th) * mscale;
        const int64_t i0 = (p.mode == 0) ? 2 * (int64_t) i : i;
        const int64_t i1 = (p.mode == 0) ? i0 + 1 : i + half;
        {
            const int64_t h = blockIdx.y;
            const float * src = x + t * p.xs_t + h * p.xs_h;
            float * dst = y + t * p.ys_t + h * p.ys_h;
            const float x0 = src[i0], x1 = src[i1];
            dst[i0] = x0 * c - x1 * s;
            dst[i1] = x0 * s + x1 * c;
        }
    }
    // pass-through of the un-rotated tail (n_dims < ne0), ops.cpp:5849-5858
    if (p.n_dims < p.ne0 && x != y) {
        const int64_t h = blockIdx.y;
        for (int64_t i = p.n_dims + threadIdx.x; i < p.ne0; i += blockDim.x)
            y[t * p.ys_t + h * p.ys_h + i] = x[t * p.xs_t + h * p.xs_h + i];
    }
}
static float yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float) M_PI)) / (2 * logf(base));
}
int rope_f32(const float * x, float * y, const int32_t * pos, const float * freq_factors, int64_t ne0, int64_t n_heads, int64_t n_tokens,
             int64_t xs_h, int64_t xs_t, int64_t ys_h, int64_t ys_t, int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale,
             float ext_factor, float attn_factor, float beta_fast, float beta_slow, cudaStream_t st) {
    if (n_tokens <= 0) return B200_OK;
    if (mode != 0 && mode != 2) return B200_ERR_UNSUPPORTED;
    RopeParams p;
    p.ne0 = ne0; p.n_heads = n_heads; p.n_tokens = n_tokens; p.xs_h = xs_h; p.xs_t = xs_t; p.ys_h = ys_h; p.ys_t = ys_t;
    p.n_dims = n_dims; p.mode = mode;
    p.theta_scale = powf(freq_base, -2.0f / n_dims);  // host libm, as the reference (ops.cpp:5783)
    p.freq_scale = freq_scale; p.ext_factor = ext_factor; p.attn_factor = attn_factor;
    // ggml_rope_yarn_corr_dims (ggml/src/ggml.c)
    const float start = floorf(yarn_corr_dim(n_dims, n_ctx_orig, beta_fast, freq_base));
    const float end = ceilf(yarn_corr_dim(n_dims, n_ctx_orig, beta_slow, freq_base));
    p.corr0 = fmaxf(0.0f, start);
    p.corr1 = fminf((float) (n_dims - 1), end);
    const int threads = n_dims / 2 >= 64 ? 64 : 32;
    launch_pdl(rope_kernel, dim3((unsigned) n_tokens, (unsigned) n_heads), dim3(threads), 0, st, x, y, pos, freq_factors, p);
    return (int) cudaGetLastError();
}

// ---- softmax ------------------------------------------------------------------------------------------------
// reference: ggml_compute_forward_soft_max_f32 ggml/src/ggml-cpu/ops.cpp:5225-5335 (scale, optional mask; no ALiBi):
// p = exp(x*scale + mask - max) / sum
__global__ void __launch_bounds__(1024) soft_max_kernel(const float * x, const float * mask, float * y,
                                                        int64_t ne0, float scale) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const float * xr = x + (int64_t) blockIdx.x * ne0;
    const float * mr = mask ? mask + (int64_t) blockIdx.x * ne0 : nullptr;
    float * yr = y + (int64_t) blockIdx.x * ne0;
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) {
        float v = xr[i] * scale;
        if (mr) v += mr[i];
        mx = fmaxf(mx, v);
    }
    mx = block_max(mx, red);
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) {
        float v = xr[i] * scale;
        if (mr) v += mr[i];
        const float e = expf(v - mx);
        yr[i] = e;
        s += e;
    }
    s = block_sum(s, red);
    const float inv = 1.0f / s;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) yr[i] *= inv;
}
int soft_max_f32(const float * x, const float * mask, float * y, int64_t ne0, int64_t nrows, float scale, cudaStream_t st) {
    if (nrows <= 0) return B200_OK;
    const int threads = ne0 >= 2048 ? 1024 : (ne0 >= 512 ? 256 : 128);
    launch_pdl(soft_max_kernel, dim3((unsigned) nrows), dim3(threads), 0, st, x, mask, y, ne0, scale);
    return (int) cudaGetLastError();
}

// ---- embedding gather from a quantized table ------------------------------------------------------------------
// reference: ggml_compute_forward_get_rows ggml/src/ggml-cpu/ops.cpp:4820 + dequantize_row_* ggml-quants.c:307/401/1352.
// Q4_0 / Q8_0 tables are in the repacked SoA row layout (quantize.cu).
__global__ void get_rows_kernel(int type, const uint8_t * table, int64_t k, const int32_t * ids, int64_t n_rows,
                                float * y) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t r = blockIdx.x;
    const int64_t row = ids[r];
    float * out = y + r * k;
    if (row < 0 || (n_rows > 0 && row >= n_rows)) {  // the reference asserts 0 <= i01 < ne01 (ops.cpp get_rows); never read out of bounds
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) out[e] = 0.0f;
        return;
    }
    const uint8_t * base = table + row * dequant_row_bytes(type, k);
    for (int64_t e = threadIdx.x; e < k; e += blockDim.x) out[e] = dequant_row_elem(type, base, k, e);
}
int get_rows_q(int type, const void * table, int64_t k, const int32_t * ids, int64_t n, float * y, cudaStream_t st, int64_t n_rows) {
    if (n <= 0) return B200_OK;
    if (type != B200_TYPE_Q4_K && type != B200_TYPE_Q4_0 && type != B200_TYPE_Q8_0 && type != B200_TYPE_F32 && type != B200_TYPE_F16)
        return B200_ERR_UNSUPPORTED;
    launch_pdl(get_rows_kernel, dim3((unsigned) n), dim3(256), 0, st, type, (const uint8_t *) table, k, ids, n_rows, y);
    return (int) cudaGetLastError();
}

}  // namespace b200

// ======================================================================================================
// Decode attention over the reference's F16 KV-cache layouts (one new token, GQA).
//   K cache: [max_len][kv_heads*head_dim] f16, one row per position          (src/layers.cpp:2933)
//   V cache: [kv_heads*head_dim][max_len] f16, TRANSPOSED: one row per channel (src/layers.cpp:2937)
// Semantics = the reference's unfused graph (src/layers.cpp:2541-2561): scores = K.Q with both operands in f16 and
// fp32 accumulation (ggml-cpu.c:213-219, vec.cpp:264), * scale, causal mask (all n_kv positions visible for the last
// token), softmax (ops.cpp:5225-5335), then V.P with P rounded to f16.  Three kernels instead of the graph's seven,
// each K / V element is read from HBM once for all query heads of its KV group.
// ======================================================================================================
namespace b200 {

// scores[h][t] = scale * sum_d K[t][g][d] * f16(q[h][d])      grid (ceil(n_kv/TPB), kv_heads), block 256 (8 warps)
template <int HD, int GQA>
__global__ void __launch_bounds__(256) attn_scores_kernel(const float * q, const __half * kc, float * scores,
                                                          int n_kv, int kv_heads, int64_t k_row_stride, float scale, int64_t s_stride) {
    __shared__ float qs[GQA][HD];
    pdl_launch_dependents();
    pdl_wait();
    const int g = blockIdx.y;
    for (int i = threadIdx.x; i < GQA * HD; i += blockDim.x) {
        const int h = i / HD, d = i % HD;
        qs[h][d] = __half2float(__float2half_rn(q[(int64_t) (g * GQA + h) * HD + d]));
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int EPL = HD / 32;  // elements per lane (HD = 64 -> 2, 128 -> 4)
    for (int t = blockIdx.x * 64 + warp; t < min(n_kv, (int) (blockIdx.x + 1) * 64); t += 8) {
        const __half * kr = kc + (int64_t) t * k_row_stride + (int64_t) g * HD + lane * EPL;
        float kv[EPL];
        if (EPL == 4) {
            const uint2 raw = *reinterpret_cast<const uint2 *>(kr);
            const __half2 a = *reinterpret_cast<const __half2 *>(&raw.x), b = *reinterpret_cast<const __half2 *>(&raw.y);
            kv[0] = __low2float(a); kv[1] = __high2float(a); kv[2] = __low2float(b); kv[3] = __high2float(b);
        } else {
#pragma unroll
            for (int e = 0; e < EPL; ++e) kv[e] = __half2float(kr[e]);
        }
#pragma unroll
        for (int h = 0; h < GQA; ++h) {
            float s = 0.0f;
#pragma unroll
            for (int e = 0; e < EPL; ++e) s = fmaf(kv[e], qs[h][lane * EPL + e], s);
            s = warp_sum(s);
            if (lane == 0) scores[(int64_t) (g * GQA + h) * s_stride + t] = s * scale;
        }
    }
}

// in-place softmax over scores[h][0..n_kv), result rounded through f16 (it is the f16 operand of V.P)
__global__ void __launch_bounds__(1024) attn_softmax_kernel(float * scores, int n_kv, int64_t s_stride) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    float * s = scores + (int64_t) blockIdx.x * s_stride;
    float mx = -INFINITY;
    for (int i = threadIdx.x; i < n_kv; i += blockDim.x) mx = fmaxf(mx, s[i]);
    mx = block_max(mx, red);
    float sum = 0.0f;
    for (int i = threadIdx.x; i < n_kv; i += blockDim.x) { const float e = expf(s[i] - mx); s[i] = e; sum += e; }
    sum = block_sum(sum, red);
    const float inv = 1.0f / sum;
    for (int i = threadIdx.x; i < n_kv; i += blockDim.x) s[i] = __half2float(__float2half_rn(s[i] * inv));
}

// out[h][d] = sum_t Vt[g*HD+d][t] * P[h][t]          grid (HD*kv_heads/8), block 256: one warp per channel row
template <int GQA>
__global__ void __launch_bounds__(256) attn_pv_kernel(const float * P, const __half * vc, float * out, int n_kv,
                                                      int head_dim, int64_t v_row_stride, int64_t s_stride) {
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * 8 + (threadIdx.x >> 5);  // channel = g*HD + d
    const int g = ch / head_dim, d = ch % head_dim;
    const __half * vr = vc + (int64_t) ch * v_row_stride;
    float acc[GQA];
#pragma unroll
    for (int h = 0; h < GQA; ++h) acc[h] = 0.0f;
    for (int t = lane * 2; t < n_kv; t += 64) {
        float v0, v1 = 0.0f;
        if (t + 1 < n_kv && ((v_row_stride & 1) == 0)) {
            const __half2 hv = *reinterpret_cast<const __half2 *>(vr + t);
            v0 = __low2float(hv); v1 = __high2float(hv);
        } else {
            v0 = __half2float(vr[t]);
            if (t + 1 < n_kv) v1 = __half2float(vr[t + 1]);
        }
#pragma unroll
        for (int h = 0; h < GQA; ++h) {
            const float * p = P + (int64_t) (g * GQA + h) * s_stride + t;
            acc[h] = fmaf(v0, p[0], acc[h]);
            if (t + 1 < n_kv) acc[h] = fmaf(v1, p[1], acc[h]);
        }
    }
#pragma unroll
    for (int h = 0; h < GQA; ++h) {
        const float s = warp_sum(acc[h]);
        if (lane == 0) out[(int64_t) (g * GQA + h) * head_dim + d] = s;
    }
}

template <int HD, int GQA>
static int attn_decode_t(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_kv, int kv_heads, int64_t k_row_stride,
                         int64_t v_row_stride, float scale, cudaStream_t st) {
    const int64_t s_stride = (n_kv + 3) & ~3;
    dim3 g1((unsigned) ((n_kv + 63) / 64), (unsigned) kv_heads);
    launch_pdl(attn_scores_kernel<HD, GQA>, dim3(g1), dim3(256), 0, st, q, (const __half *) kc, scratch, n_kv, kv_heads, k_row_stride, scale, s_stride);
    launch_pdl(attn_softmax_kernel, dim3((unsigned) (kv_heads * GQA)), dim3(n_kv >= 2048 ? 1024 : 256), 0, st, scratch, n_kv, s_stride);
    launch_pdl(attn_pv_kernel<GQA>, dim3((unsigned) (kv_heads * HD / 8)), dim3(256), 0, st, scratch, (const __half *) vc, out, n_kv, HD, v_row_stride, s_stride);
    return (int) cudaGetLastError();
}

size_t attn_decode_scratch_bytes(int n_heads, int n_kv) { return (size_t) n_heads * (size_t) ((n_kv + 3) & ~3) * 4; }

int attn_decode(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                int64_t k_row_stride, int64_t v_row_stride, float scale, cudaStream_t st) {
    if (n_kv <= 0) return B200_OK;
    const int gqa = n_heads / kv_heads;
    if (n_heads % kv_heads) return B200_ERR_ARG;
#define B200_ATTN_CASE(HD_, G_) \
    if (head_dim == HD_ && gqa == G_) return attn_decode_t<HD_, G_>(q, kc, vc, out, scratch, n_kv, kv_heads, k_row_stride, v_row_stride, scale, st);
    B200_ATTN_CASE(128, 4) B200_ATTN_CASE(128, 7) B200_ATTN_CASE(128, 1) B200_ATTN_CASE(128, 8) B200_ATTN_CASE(128, 2)
    B200_ATTN_CASE(64, 8) B200_ATTN_CASE(64, 4) B200_ATTN_CASE(64, 2) B200_ATTN_CASE(64, 1)
#undef B200_ATTN_CASE
    return B200_ERR_UNSUPPORTED;
}

// write one token's K (row) and V (column of the transposed cache) as f16 — KVCacheAttention::save_to_cache
__global__ void kv_store_kernel(const float * k, const float * v, __half * kc, __half * vc, int kv_hidden,
                                int64_t k_row_stride, int64_t v_row_stride, int pos) {
    pdl_launch_dependents();
    pdl_wait();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= kv_hidden) return;
    kc[(int64_t) pos * k_row_stride + i] = __float2half_rn(k[i]);
    vc[(int64_t) i * v_row_stride + pos] = __float2half_rn(v[i]);
}
int kv_store(const float * k, const float * v, void * kc, void * vc, int kv_hidden, int64_t k_row_stride, int64_t v_row_stride, int pos, cudaStream_t st) {
    launch_pdl(kv_store_kernel, dim3((unsigned) ((kv_hidden + 255) / 256)), dim3(256), 0, st, k, v, (__half *) kc, (__half *) vc, kv_hidden, k_row_stride, v_row_stride, pos);
    return (int) cudaGetLastError();
}


// ======================================================================================================
// Layer-sharded multi-GPU hand-off without the host (the reference copies the hidden row with cpy_tensor_async + events,
// ggml/src/ggml-cuda/ggml-cuda.cu:2806-2866, and synchronizes the scheduler around it): two one-CTA kernels that sit at the two ends of a
// rank's CUDA graph.  peer_wait spins on a flag in LOCAL memory that the previous rank raises through NVLink; peer_send stores the
// hidden row (or, from the last rank, the next token) straight into the next rank's CUDA-IPC mapped mailbox and raises its flag.  Flags
// are 64-bit launch counters (never reset): launch t of a rank waits for flag >= t - 1 + offset.
// ======================================================================================================
__global__ void peer_wait_kernel(const unsigned long long * flag, const unsigned long long * seq, int offset, unsigned long long * status) {
    pdl_launch_dependents();
    pdl_wait();
    if (threadIdx.x == 0) {
        const unsigned long long want = *seq + (unsigned long long) offset;
        unsigned long long v;
        unsigned long long t0;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if ((++spins & 4095u) == 0) {
                unsigned long long t1;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
                if (t1 - t0 > 20000000000ull) { if (status) *status = 1ull; break; }   // a lost peer must not hang the GPU
            }
        } while (v < want);
    }
}
__global__ void peer_send_kernel(const float * x, float * peer_x, int64_t n, const int32_t * tok, int32_t * peer_tok, unsigned long long * peer_flag,
                                 unsigned long long * seq, int32_t * pos) {
    pdl_launch_dependents();
    pdl_wait();
    if (peer_x) for (int64_t e = 4 * (int64_t) threadIdx.x; e < n; e += 4 * (int64_t) blockDim.x) *reinterpret_cast<float4 *>(peer_x + e) = *reinterpret_cast<const float4 *>(x + e);
    if (peer_tok && threadIdx.x == 0) peer_tok[0] = tok[0];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long s = *seq + 1ull;
        if (peer_flag) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag), "l"(s) : "memory");
        *seq = s;
        if (pos) pos[0] += 1;
    }
}
int peer_wait(const void * flag, const void * seq, int offset, void * status, cudaStream_t st) {
    launch_pdl(peer_wait_kernel, dim3(1), dim3(32), 0, st, (const unsigned long long *) flag, (const unsigned long long *) seq, offset, (unsigned long long *) status);
    return (int) cudaGetLastError();
}
int peer_send(const float * x, float * peer_x, int64_t n, const int32_t * tok, int32_t * peer_tok, void * peer_flag, void * seq, int32_t * pos, cudaStream_t st) {
    if (n % 4) return B200_ERR_UNSUPPORTED;
    launch_pdl(peer_send_kernel, dim3(1), dim3(256), 0, st, x, peer_x, n, tok, peer_tok, (unsigned long long *) peer_flag, (unsigned long long *) seq, pos);
    return (int) cudaGetLastError();
}

// out[0] = index of the FIRST maximum of x[0..n)  (greedy sampling on the device: src/models.cpp:1026-1031 reads 0.5 MB of logits back
// and scans them on the host every token).  64 CTAs fold their slices into one 64-bit key with atomicMax; the last CTA to finish decodes it
// and re-arms the scratch (per-device module globals; one stream at a time).
__device__ unsigned long long g_argmax_key = 0ull;
__device__ unsigned int g_argmax_done = 0u;
__global__ void __launch_bounds__(256) argmax_kernel(const float * x, int64_t n, int32_t * out) {
    __shared__ unsigned long long red[8];
    pdl_launch_dependents();
    pdl_wait();
    unsigned long long best = 0ull;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x) {
        const uint32_t b = __float_as_uint(x[i]);
        const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // order-preserving float -> uint
        const unsigned long long key = ((unsigned long long) ord << 32) | (unsigned) (0xffffffffu - (unsigned) i);   // ties: the smaller index wins
        best = key > best ? key : best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long v = __shfl_xor_sync(0xffffffffu, best, o); best = v > best ? v : best; }
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int) (blockDim.x >> 5); ++w) best = red[w] > best ? red[w] : best;
        atomicMax(&g_argmax_key, best);
        __threadfence();
        if (atomicAdd(&g_argmax_done, 1u) == gridDim.x - 1) {   // last CTA: every key has been folded in
            __threadfence();
            const unsigned long long v = atomicExch(&g_argmax_key, 0ull);
            out[0] = (int32_t) (0xffffffffu - (unsigned) (v & 0xffffffffu));
            g_argmax_done = 0u;
        }
    }
}
int argmax_f32(const float * x, int64_t n, int32_t * out, cudaStream_t st) {
    if (n <= 0 || n > 0x7fffffff) return B200_ERR_ARG;
    const int grid = (int) (n >= 65536 ? 64 : (n >= 4096 ? 8 : 1));
    launch_pdl(argmax_kernel, dim3(grid), dim3(256), 0, st, x, n, out);
    return (int) cudaGetLastError();
}

}  // namespace b200
