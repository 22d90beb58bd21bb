// ops.cu — the fp32 glue ops around the quantized matmuls (RMSNorm, RoPE, softmax, SwiGLU, residual add,
// embedding gather, F16 attention matmuls, KV-cache writes).  Each mirrors the arithmetic of the reference CPU
// op (file:line cited per kernel) closely enough that decode logits agree to ~1e-5 relative.
#include "common.cuh"
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
__global__ void __launch_bounds__(1024) rms_norm_mul_kernel(const float * __restrict__ x, const float * __restrict__ w, float * __restrict__ y,
                                                            int64_t ne0, float eps) {
    __shared__ float red[32];
    pdl_wait();
    const float * xr = x + (int64_t) blockIdx.x * ne0;
    float * yr = y + (int64_t) blockIdx.x * ne0;
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) { const float v = xr[i]; s = fmaf(v, v, s); }
    s = block_sum(s, red);
    const float mean = s / (float) ne0;
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = threadIdx.x; i < ne0; i += blockDim.x) {
        const float v = xr[i] * scale;
        yr[i] = w ? v * w[i] : v;
    }
}
int rms_norm_mul(const float * x, const float * w, float * y, int64_t ne0, int64_t nrows, float eps, cudaStream_t st) {
    if (nrows <= 0) return B200_OK;
    const int threads = ne0 >= 4096 ? 1024 : (ne0 >= 1024 ? 512 : 256);
    rms_norm_mul_kernel<<<(unsigned) nrows, threads, 0, st>>>(x, w, y, ne0, eps);
    return (int) cudaGetLastError();
}

// ---- elementwise ----------------------------------------------------------------------------------------
__global__ void add_kernel(const float * __restrict__ a, const float * __restrict__ b, float * __restrict__ y, int64_t n) {
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = a[i] + b[i];
}
int add_f32(const float * a, const float * b, float * y, int64_t n, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    add_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(a, b, y, n);
    return (int) cudaGetLastError();
}
// SwiGLU of BaseMLP::forward (src/layers.cpp:2475-2483): silu(gate) * up, silu = x/(1+exp(-x)) (vec.h:1061)
__global__ void silu_mul_kernel(const float * __restrict__ g, const float * __restrict__ u, float * __restrict__ y, int64_t n) {
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = g[i]; y[i] = (v / (1.0f + expf(-v))) * u[i]; }
}
int silu_mul(const float * gate, const float * up, float * y, int64_t n, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    silu_mul_kernel<<<(unsigned) ((n + 255) / 256), 256, 0, st>>>(gate, up, y, n);
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
__global__ void rope_kernel(const float * __restrict__ x, float * __restrict__ y, const int32_t * __restrict__ pos,
                            const float * __restrict__ ff, const RopeParams p) {
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
        for (int64_t h = 0; h < p.n_heads; ++h) {
            const float * src = x + t * p.xs_t + h * p.xs_h;
            float * dst = y + t * p.ys_t + h * p.ys_h;
            const float x0 = src[i0], x1 = src[i1];
            dst[i0] = x0 * c - x1 * s;
            dst[i1] = x0 * s + x1 * c;
        }
    }
    // pass-through of the un-rotated tail (n_dims < ne0), ops.cpp:5849-5858
    if (p.n_dims < p.ne0 && x != y) {
        for (int64_t h = 0; h < p.n_heads; ++h)
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
    rope_kernel<<<(unsigned) n_tokens, threads, 0, st>>>(x, y, pos, freq_factors, p);
    return (int) cudaGetLastError();
}

// ---- softmax ------------------------------------------------------------------------------------------------
// reference: ggml_compute_forward_soft_max_f32 ggml/src/ggml-cpu/ops.cpp:5225-5335 (scale, optional mask; no ALiBi):
// p = exp(x*scale + mask - max) / sum
__global__ void __launch_bounds__(1024) soft_max_kernel(const float * __restrict__ x, const float * __restrict__ mask, float * __restrict__ y,
                                                        int64_t ne0, float scale) {
    __shared__ float red[32];
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
    soft_max_kernel<<<(unsigned) nrows, threads, 0, st>>>(x, mask, y, ne0, scale);
    return (int) cudaGetLastError();
}

// ---- embedding gather from a quantized table ------------------------------------------------------------------
// reference: ggml_compute_forward_get_rows ggml/src/ggml-cpu/ops.cpp:4820 + dequantize_row_* ggml-quants.c:307/401/1352.
// Q4_0 / Q8_0 tables are in the repacked SoA row layout (quantize.cu).
__global__ void get_rows_kernel(int type, const uint8_t * __restrict__ table, int64_t k, const int32_t * __restrict__ ids,
                                float * __restrict__ y) {
    pdl_wait();
    const int64_t r = blockIdx.x;
    const int64_t row = ids[r];
    float * out = y + r * k;
    if (type == B200_TYPE_Q4_K) {
        const uint8_t * base = table + row * (k / 256) * 144;
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) {
            const int64_t b = e >> 8;
            const int w = (int) (e & 255);
            const uint8_t * blk = base + b * 144;
            const float d = half_bits_to_float(blk[0] | (blk[1] << 8)), dmin = half_bits_to_float(blk[2] | (blk[3] << 8));
            const int j = w >> 5;  // sub-block
            const uint8_t * q = blk + 4;
            int sc, mn;
            if (j < 4) { sc = q[j] & 63; mn = q[j + 4] & 63; }
            else { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); mn = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
            const uint8_t byte = blk[16 + (j >> 1) * 32 + (w & 31)];
            const int qv = (j & 1) ? (byte >> 4) : (byte & 0xF);
            out[e] = (d * sc) * qv - (dmin * mn);
        }
    } else if (type == B200_TYPE_Q4_0) {
        const int64_t nb = k / 32;
        const uint8_t * base = table + row * nb * 18;
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) {
            const int64_t b = e >> 5;
            const int w = (int) (e & 31);
            const uint8_t * dp = base + nb * 16 + b * 2;
            const float d = half_bits_to_float(dp[0] | (dp[1] << 8));
            const uint8_t byte = base[b * 16 + (w & 15)];
            const int qv = (w < 16 ? (byte & 0xF) : (byte >> 4)) - 8;
            out[e] = qv * d;
        }
    } else if (type == B200_TYPE_Q8_0) {
        const int64_t nb = k / 32;
        const uint8_t * base = table + row * nb * 34;
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) {
            const int64_t b = e >> 5;
            const uint8_t * dp = base + nb * 32 + b * 2;
            const float d = half_bits_to_float(dp[0] | (dp[1] << 8));
            out[e] = (float) ((const int8_t *) base)[e] * d;
        }
    } else if (type == B200_TYPE_F32) {
        const float * src = (const float *) table + row * k;
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) out[e] = src[e];
    } else if (type == B200_TYPE_F16) {
        const __half * src = (const __half *) table + row * k;
        for (int64_t e = threadIdx.x; e < k; e += blockDim.x) out[e] = __half2float(src[e]);
    }
}
int get_rows_q(int type, const void * table, int64_t k, const int32_t * ids, int64_t n, float * y, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    if (type != B200_TYPE_Q4_K && type != B200_TYPE_Q4_0 && type != B200_TYPE_Q8_0 && type != B200_TYPE_F32 && type != B200_TYPE_F16)
        return B200_ERR_UNSUPPORTED;
    get_rows_kernel<<<(unsigned) n, 256, 0, st>>>(type, (const uint8_t *) table, k, ids, y);
    return (int) cudaGetLastError();
}

}  // namespace b200
