// ops_generic.cu — strided / broadcasting versions of the glue ops, as the ggml graph presents them to the backend
// (views, permutes, in-place results).  Shapes follow ggml's convention: ne[0] is the fastest dimension, nb[] are
// byte strides.  Each kernel cites the reference CPU routine whose arithmetic it mirrors.
#include "common.cuh"
#include "kernels.h"

#include <cstdint>
#include <cstdlib>

namespace b200 {

struct TVd {  // device-side copy of a TV
    char * data;
    int64_t ne0, ne1, ne2, ne3;
    int64_t nb0, nb1, nb2, nb3;
};
static inline TVd dv(const TV & t) {
    TVd d;
    d.data = (char *) t.data;
    d.ne0 = t.ne[0]; d.ne1 = t.ne[1]; d.ne2 = t.ne[2]; d.ne3 = t.ne[3];
    d.nb0 = t.nb[0]; d.nb1 = t.nb[1]; d.nb2 = t.nb[2]; d.nb3 = t.nb[3];
    return d;
}
static inline int64_t nelem(const TV & t) { return t.ne[0] * t.ne[1] * t.ne[2] * t.ne[3]; }
static inline int64_t nrows(const TV & t) { return t.ne[1] * t.ne[2] * t.ne[3]; }

__device__ __forceinline__ void unravel(int64_t i, const TVd & t, int64_t & i0, int64_t & i1, int64_t & i2, int64_t & i3) {
    i0 = i % t.ne0; i /= t.ne0;
    i1 = i % t.ne1; i /= t.ne1;
    i2 = i % t.ne2;
    i3 = i / t.ne2;
}
__device__ __forceinline__ char * at(const TVd & t, int64_t i0, int64_t i1, int64_t i2, int64_t i3) {
    return t.data + i0 * t.nb0 + i1 * t.nb1 + i2 * t.nb2 + i3 * t.nb3;
}

template <typename T> __device__ __forceinline__ float ldf(const char * p);
template <> __device__ __forceinline__ float ldf<float>(const char * p) { return *(const float *) p; }
template <> __device__ __forceinline__ float ldf<__half>(const char * p) { return __half2float(*(const __half *) p); }
template <typename T> __device__ __forceinline__ void stf(char * p, float v);
template <> __device__ __forceinline__ void stf<float>(char * p, float v) { *(float *) p = v; }
template <> __device__ __forceinline__ void stf<__half>(char * p, float v) { *(__half *) p = __float2half_rn(v); }  // GGML_FP32_TO_FP16 (RNE)

// ---- ADD / MUL / DIV with ggml broadcasting of src1 (binary-ops.cpp apply_binary_op) --------------------------
template <int OP>
__global__ void bin_bcast_kernel(const TVd a, const TVd b, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t i0, i1, i2, i3;
    unravel(i, d, i0, i1, i2, i3);
    const float x = *(const float *) at(a, i0, i1, i2, i3);
    const float y = *(const float *) at(b, i0 % b.ne0, i1 % b.ne1, i2 % b.ne2, i3 % b.ne3);
    float r;
    if (OP == 0) r = x + y;
    else if (OP == 1) r = x * y;
    else r = x / y;
    *(float *) at(d, i0, i1, i2, i3) = r;
}
int op_bin(int op, const TV & a, const TV & b, const TV & d, cudaStream_t st) {
    const int64_t n = nelem(d);
    if (n <= 0) return B200_OK;
    const unsigned grid = (unsigned) ((n + 255) / 256);
    switch (op) {
        case 0: launch_pdl(bin_bcast_kernel<0>, dim3(grid), dim3(256), 0, st, dv(a), dv(b), dv(d), n); break;
        case 1: launch_pdl(bin_bcast_kernel<1>, dim3(grid), dim3(256), 0, st, dv(a), dv(b), dv(d), n); break;
        case 2: launch_pdl(bin_bcast_kernel<2>, dim3(grid), dim3(256), 0, st, dv(a), dv(b), dv(d), n); break;
        default: return B200_ERR_ARG;
    }
    return (int) cudaGetLastError();
}

// ---- CPY / DUP / CONT: same number of elements, arbitrary shapes & strides, optional F32<->F16 conversion
//      (ggml_compute_forward_dup, ops.cpp:4637: elements are matched by logical (row-major-in-ne) index)
template <typename TS, typename TD>
__global__ void cpy_kernel(const TVd s, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t a0, a1, a2, a3, b0, b1, b2, b3;
    unravel(i, s, a0, a1, a2, a3);
    unravel(i, d, b0, b1, b2, b3);
    stf<TD>(at(d, b0, b1, b2, b3), ldf<TS>(at(s, a0, a1, a2, a3)));
}
__global__ void cpy_i32_kernel(const TVd s, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t a0, a1, a2, a3, b0, b1, b2, b3;
    unravel(i, s, a0, a1, a2, a3);
    unravel(i, d, b0, b1, b2, b3);
    *(int32_t *) at(d, b0, b1, b2, b3) = *(const int32_t *) at(s, a0, a1, a2, a3);
}
int op_cpy(const TV & s, const TV & d, cudaStream_t st) {
    const int64_t n = nelem(s);
    if (n != nelem(d)) return B200_ERR_ARG;
    if (n <= 0) return B200_OK;
    const unsigned grid = (unsigned) ((n + 255) / 256);
    if (s.type == B200_TYPE_F32 && d.type == B200_TYPE_F32) launch_pdl(cpy_kernel<float, float>, dim3(grid), dim3(256), 0, st, dv(s), dv(d), n);
    else if (s.type == B200_TYPE_F32 && d.type == B200_TYPE_F16) launch_pdl(cpy_kernel<float, __half>, dim3(grid), dim3(256), 0, st, dv(s), dv(d), n);
    else if (s.type == B200_TYPE_F16 && d.type == B200_TYPE_F16) launch_pdl(cpy_kernel<__half, __half>, dim3(grid), dim3(256), 0, st, dv(s), dv(d), n);
    else if (s.type == B200_TYPE_F16 && d.type == B200_TYPE_F32) launch_pdl(cpy_kernel<__half, float>, dim3(grid), dim3(256), 0, st, dv(s), dv(d), n);
    else if (s.type == 26 && d.type == 26) launch_pdl(cpy_i32_kernel, dim3(grid), dim3(256), 0, st, dv(s), dv(d), n);
    else return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

// ---- SET_ROWS: dst[:, ids[i1, i2 % ne11, i3 % ne12], i2, i3] = convert(src[:, i1, i2, i3])  (ops.cpp:4942)
template <typename TD, typename TI>
__global__ void set_rows_kernel(const TVd s, const TVd ids, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t i0, i1, i2, i3;
    unravel(i, s, i0, i1, i2, i3);
    const int64_t r = (int64_t) * (const TI *) at(ids, i1, i2 % ids.ne1, i3 % ids.ne2, 0);
    stf<TD>(at(d, i0, r, i2, i3), *(const float *) at(s, i0, i1, i2, i3));
}
int op_set_rows(const TV & s, const TV & ids, const TV & d, cudaStream_t st) {
    const int64_t n = nelem(s);
    if (n <= 0) return B200_OK;
    const unsigned grid = (unsigned) ((n + 255) / 256);
    const bool i64 = ids.type == 27;
    if (d.type == B200_TYPE_F16) {
        if (i64) launch_pdl(set_rows_kernel<__half, int64_t>, dim3(grid), dim3(256), 0, st, dv(s), dv(ids), dv(d), n);
        else launch_pdl(set_rows_kernel<__half, int32_t>, dim3(grid), dim3(256), 0, st, dv(s), dv(ids), dv(d), n);
    } else if (d.type == B200_TYPE_F32) {
        if (i64) launch_pdl(set_rows_kernel<float, int64_t>, dim3(grid), dim3(256), 0, st, dv(s), dv(ids), dv(d), n);
        else launch_pdl(set_rows_kernel<float, int32_t>, dim3(grid), dim3(256), 0, st, dv(s), dv(ids), dv(d), n);
    } else return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

// ---- SCALE (y = x*s + b, ops.cpp:4374), CLAMP, SILU, DIAG_MASK_INF on contiguous tensors -----------------------
__global__ void scale_kernel(const float * x, float * y, int64_t n, float s, float b) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = (b == 0.0f) ? x[i] * s : fmaf(x[i], s, b);  // ggml_vec_scale_f32 / ggml_vec_mad1_f32
}
int op_scale(const float * x, float * y, int64_t n, float s, float b, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(scale_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, x, y, n, s, b);
    return (int) cudaGetLastError();
}
__global__ void clamp_kernel(const float * x, float * y, int64_t n, float lo, float hi) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = fminf(fmaxf(x[i], lo), hi);
}
int op_clamp(const float * x, float * y, int64_t n, float lo, float hi, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(clamp_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, x, y, n, lo, hi);
    return (int) cudaGetLastError();
}
__global__ void silu_kernel(const float * x, float * y, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const float v = x[i]; y[i] = v / (1.0f + expf(-v)); }
}
int op_silu(const float * x, float * y, int64_t n, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(silu_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, x, y, n);
    return (int) cudaGetLastError();
}
// ggml_compute_forward_diag_mask_f32: for k in z, j in rows, i >= n_past: if (i > n_past + j) -inf
__global__ void diag_mask_inf_kernel(const float * x, float * y, int64_t ne0, int64_t ne1, int64_t n, int n_past) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t c = i % ne0, j = (i / ne0) % ne1;
    y[i] = (c > n_past + j) ? -INFINITY : x[i];
}
int op_diag_mask_inf(const float * x, float * y, int64_t ne0, int64_t ne1, int64_t n, int n_past, cudaStream_t st) {
    if (n <= 0) return B200_OK;
    launch_pdl(diag_mask_inf_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, x, y, ne0, ne1, n, n_past);
    return (int) cudaGetLastError();
}

// ---- SOFT_MAX with optional F32/F16 mask broadcast over dims 2,3 (ops.cpp:5225-5335; max_bias == 0 only) -------
template <typename TM>
__global__ void __launch_bounds__(1024) soft_max_ext_kernel(const TVd x, const TVd m, const TVd y, float scale, bool has_mask) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % x.ne1, i2 = (row / x.ne1) % x.ne2, i3 = row / (x.ne1 * x.ne2);
    const float * xr = (const float *) at(x, 0, i1, i2, i3);
    float * yr = (float *) at(y, 0, i1, i2, i3);
    const char * mr = has_mask ? at(m, 0, i1, i2 % m.ne2, i3 % m.ne3) : nullptr;
    float mx = -INFINITY;
    for (int64_t i = threadIdx.x; i < x.ne0; i += blockDim.x) {
        float v = xr[i] * scale;
        if (mr) v += ldf<TM>(mr + i * sizeof(TM));
        mx = fmaxf(mx, v);
    }
    // block max
    mx = warp_max(mx);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
    if (warp == 0) { t = warp_max(t); if (lane == 0) red[0] = t; }
    __syncthreads();
    mx = red[0];
    __syncthreads();
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < x.ne0; i += blockDim.x) {
        float v = xr[i] * scale;
        if (mr) v += ldf<TM>(mr + i * sizeof(TM));
        const float e = expf(v - mx);
        yr[i] = e;
        s += e;
    }
    s = warp_sum(s);
    if (lane == 0) red[warp] = s;
    __syncthreads();
    t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) { t = warp_sum(t); if (lane == 0) red[0] = t; }
    __syncthreads();
    const float inv = 1.0f / red[0];
    for (int64_t i = threadIdx.x; i < x.ne0; i += blockDim.x) yr[i] *= inv;
}
int op_soft_max(const TV & x, const TV * mask, const TV & y, float scale, cudaStream_t st) {
    const int64_t rows = nrows(x);
    if (rows <= 0) return B200_OK;
    const int threads = x.ne[0] >= 2048 ? 1024 : (x.ne[0] >= 512 ? 256 : 128);
    TVd m = mask ? dv(*mask) : dv(x);
    if (mask && mask->type == B200_TYPE_F16) launch_pdl(soft_max_ext_kernel<__half>, dim3((unsigned) rows), dim3(threads), 0, st, dv(x), m, dv(y), scale, true);
    else launch_pdl(soft_max_ext_kernel<float>, dim3((unsigned) rows), dim3(threads), 0, st, dv(x), m, dv(y), scale, mask != nullptr);
    return (int) cudaGetLastError();
}

// ---- RMS_NORM on strided rows (ops.cpp:3710-3758) ----------------------------------------------------------------
__global__ void __launch_bounds__(1024) rms_norm_strided_kernel(const TVd x, const TVd y, float eps) {
    __shared__ float red[32];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % x.ne1, i2 = (row / x.ne1) % x.ne2, i3 = row / (x.ne1 * x.ne2);
    const float * xr = (const float *) at(x, 0, i1, i2, i3);
    float * yr = (float *) at(y, 0, i1, i2, i3);
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < x.ne0; i += blockDim.x) { const float v = xr[i]; s = fmaf(v, v, s); }
    s = warp_sum(s);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = s;
    __syncthreads();
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
    if (warp == 0) { t = warp_sum(t); if (lane == 0) red[0] = t; }
    __syncthreads();
    const float mean = red[0] / (float) x.ne0;
    const float scale = 1.0f / sqrtf(mean + eps);
    for (int64_t i = threadIdx.x; i < x.ne0; i += blockDim.x) yr[i] = xr[i] * scale;
}
int op_rms_norm(const TV & x, const TV & y, float eps, cudaStream_t st) {
    const int64_t rows = nrows(x);
    if (rows <= 0) return B200_OK;
    const int threads = x.ne[0] >= 4096 ? 1024 : (x.ne[0] >= 1024 ? 512 : 256);
    launch_pdl(rms_norm_strided_kernel, dim3((unsigned) rows), dim3(threads), 0, st, dv(x), dv(y), eps);
    return (int) cudaGetLastError();
}

// ---- MUL_MAT with a float src0 (F16 / F32, arbitrary strides except nb00 == element size) ---------------------------
// dst[i0, i1, i2, i3] = sum_k src0[k, i0, i2/r2, i3/r3] * src1[k, i1, i2, i3]; when src0 is F16 the reference first
// rounds src1 to F16 (ggml-cpu.c:213-219, :1291-1326) and accumulates in fp32 (vec.cpp:264).  One warp per output.
template <typename T0>
__global__ void mul_mat_f_kernel(const TVd a, const TVd b, const TVd d, int64_t nout) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t o = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (o >= nout) return;
    int64_t i0, i1, i2, i3;
    unravel(o, d, i0, i1, i2, i3);
    const int64_t r2 = b.ne2 / a.ne2, r3 = b.ne3 / a.ne3;
    const char * ar = at(a, 0, i0, i2 / r2, i3 / r3);
    const char * br = at(b, 0, i1, i2, i3);
    float s = 0.0f;
    for (int64_t k = lane; k < a.ne0; k += 32) {
        float bv = *(const float *) (br + k * b.nb0);
        if (sizeof(T0) == 2) bv = __half2float(__float2half_rn(bv));
        s = fmaf(ldf<T0>(ar + k * sizeof(T0)), bv, s);
    }
    s = warp_sum(s);
    if (lane == 0) *(float *) at(d, i0, i1, i2, i3) = s;
}

// ---- the same MUL_MAT (F16 src0) for prompt-sized src1 (>= 16 columns): tiled tensor-core GEMM --------------------------------
// The warp-per-output kernel above re-reads a whole src0 row for every output element; for the prefill attention matmuls
// (scores = K^T.Q with [n_kv x qlen] outputs per head, ctx = V.P) that is tens of GB of L2 traffic per layer.  Here a CTA owns a
// 64 x 64 output tile of one (i2, i3) slice: both operands are staged through shared memory as f16 (src1 is rounded to f16 exactly
// as the reference does before its f16 dot, ggml-cpu.c:213-219, :1291-1326) and multiplied with mma.sync m16n8k16 f16 x f16 -> f32
// (f16 products are exact in fp32; only the fp32 summation order differs from vec.cpp:264).  4 warps, each a 32 x 32 sub-tile.
// Arbitrary strides / ragged edges: 16-byte vector loads where a chunk is aligned and in bounds, guarded scalar loads elsewhere.
#define MMF_TM 64
#define MMF_TN 64
#define MMF_TK 32
#define MMF_LD 40  // halves per smem row: 80-byte stride -> the 8 rows of a fragment load hit 8 distinct 4-bank groups
__global__ void __launch_bounds__(128) mul_mat_f16_mma_kernel(const TVd a, const TVd b, const TVd d) {
    __shared__ __align__(16) __half As[MMF_TM * MMF_LD];
    __shared__ __align__(16) __half Bs[MMF_TN * MMF_LD];
    pdl_launch_dependents();
    pdl_wait();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm = (warp & 1) * 32, wn = (warp >> 1) * 32;
    const int64_t m0 = (int64_t) blockIdx.x * MMF_TM, n0 = (int64_t) blockIdx.y * MMF_TN;
    const int64_t i2 = blockIdx.z % d.ne2, i3 = blockIdx.z / d.ne2;
    const int64_t r2 = b.ne2 / a.ne2, r3 = b.ne3 / a.ne3;
    const char * abase = a.data + (i2 / r2) * a.nb2 + (i3 / r3) * a.nb3;
    const char * bbase = b.data + i2 * b.nb2 + i3 * b.nb3;
    const int64_t K = a.ne0, M = a.ne1, N = b.ne1;
    const bool a_vec = (a.nb1 % 16 == 0) && (((uintptr_t) abase) % 16 == 0);
    const bool b_vec = (b.nb1 % 16 == 0) && (((uintptr_t) bbase) % 16 == 0);

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = acc[i][j][2] = acc[i][j][3] = 0.0f;

    for (int64_t k0 = 0; k0 < K; k0 += MMF_TK) {
        // ---- A tile: 64 rows x 32 halves = 256 chunks of 8 halves
        for (int c = threadIdx.x; c < MMF_TM * 4; c += 128) {
            const int r = c >> 2, ch = c & 3;
            const int64_t row = m0 + r, kk = k0 + ch * 8;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (row < M) {
                const char * src = abase + row * a.nb1 + kk * 2;
                if (a_vec && kk + 8 <= K) {
                    v = *reinterpret_cast<const uint4 *>(src);
                } else {
                    __align__(16) __half h[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[e] = (kk + e < K) ? *reinterpret_cast<const __half *>(src + 2 * e) : __float2half_rn(0.0f);
                    v = *reinterpret_cast<const uint4 *>(h);
                }
            }
            *reinterpret_cast<uint4 *>(As + r * MMF_LD + ch * 8) = v;
        }
        // ---- B tile: 64 columns x 32 floats -> f16 (RNE, as GGML_FP32_TO_FP16)
        for (int c = threadIdx.x; c < MMF_TN * 8; c += 128) {
            const int r = c >> 3, ch = c & 7;
            const int64_t col = n0 + r, kk = k0 + ch * 4;
            float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col < N) {
                const char * src = bbase + col * b.nb1 + kk * 4;
                if (b_vec && kk + 4 <= K) {
                    f = *reinterpret_cast<const float4 *>(src);
                } else {
                    if (kk + 0 < K) f.x = *reinterpret_cast<const float *>(src + 0);
                    if (kk + 1 < K) f.y = *reinterpret_cast<const float *>(src + 4);
                    if (kk + 2 < K) f.z = *reinterpret_cast<const float *>(src + 8);
                    if (kk + 3 < K) f.w = *reinterpret_cast<const float *>(src + 12);
                }
            }
            const __half2 lo = __halves2half2(__float2half_rn(f.x), __float2half_rn(f.y)), hi = __halves2half2(__float2half_rn(f.z), __float2half_rn(f.w));
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t *>(&lo); pk.y = *reinterpret_cast<const uint32_t *>(&hi);
            *reinterpret_cast<uint2 *>(Bs + r * MMF_LD + ch * 4) = pk;
        }
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < MMF_TK; ks += 16) {
            uint32_t af[2][4], bf[4][2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const __half * ap = As + (wm + i * 16 + g) * MMF_LD + ks + 2 * t;
                af[i][0] = *reinterpret_cast<const uint32_t *>(ap);
                af[i][1] = *reinterpret_cast<const uint32_t *>(ap + 8 * MMF_LD);
                af[i][2] = *reinterpret_cast<const uint32_t *>(ap + 8);
                af[i][3] = *reinterpret_cast<const uint32_t *>(ap + 8 * MMF_LD + 8);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const __half * bp = Bs + (wn + j * 8 + g) * MMF_LD + ks + 2 * t;
                bf[j][0] = *reinterpret_cast<const uint32_t *>(bp);
                bf[j][1] = *reinterpret_cast<const uint32_t *>(bp + 8);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                                 : "+f"(acc[i][j][0]), "+f"(acc[i][j][1]), "+f"(acc[i][j][2]), "+f"(acc[i][j][3])
                                 : "r"(af[i][0]), "r"(af[i][1]), "r"(af[i][2]), "r"(af[i][3]), "r"(bf[j][0]), "r"(bf[j][1]));
        }
        __syncthreads();
    }
    // c0 = (row g, col 2t), c1 = (row g, col 2t+1), c2 = (row g+8, col 2t), c3 = (row g+8, col 2t+1); row -> i0, col -> i1
    char * dbase = d.data + i2 * d.nb2 + i3 * d.nb3;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t rA = m0 + wm + i * 16 + g, rB = rA + 8, cA = n0 + wn + j * 8 + 2 * t, cB = cA + 1;
            if (rA < M && cA < N) *reinterpret_cast<float *>(dbase + rA * d.nb0 + cA * d.nb1) = acc[i][j][0];
            if (rA < M && cB < N) *reinterpret_cast<float *>(dbase + rA * d.nb0 + cB * d.nb1) = acc[i][j][1];
            if (rB < M && cA < N) *reinterpret_cast<float *>(dbase + rB * d.nb0 + cA * d.nb1) = acc[i][j][2];
            if (rB < M && cB < N) *reinterpret_cast<float *>(dbase + rB * d.nb0 + cB * d.nb1) = acc[i][j][3];
        }
}

int op_mul_mat_f(const TV & a, const TV & b, const TV & d, cudaStream_t st) {
    const int64_t nout = nelem(d);
    if (nout <= 0) return B200_OK;
    static const bool no_mma = getenv("B200_MMF_NO_MMA") != nullptr;  // bisect aid
    if (!no_mma && a.type == B200_TYPE_F16 && b.ne[1] >= 16 && a.nb[0] == 2 && b.nb[0] == 4 && d.ne[2] * d.ne[3] <= 65535 && a.ne[0] > 0) {
        const dim3 grid((unsigned) ((a.ne[1] + MMF_TM - 1) / MMF_TM), (unsigned) ((b.ne[1] + MMF_TN - 1) / MMF_TN), (unsigned) (d.ne[2] * d.ne[3]));
        if (grid.y <= 65535) {
            launch_pdl(mul_mat_f16_mma_kernel, grid, dim3(128), 0, st, dv(a), dv(b), dv(d));
            return (int) cudaGetLastError();
        }
    }
    const unsigned grid = (unsigned) ((nout * 32 + 255) / 256);
    if (a.type == B200_TYPE_F16) launch_pdl(mul_mat_f_kernel<__half>, dim3(grid), dim3(256), 0, st, dv(a), dv(b), dv(d), nout);
    else if (a.type == B200_TYPE_F32) launch_pdl(mul_mat_f_kernel<float>, dim3(grid), dim3(256), 0, st, dv(a), dv(b), dv(d), nout);
    else return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

// ---- SUM_ROWS, REPEAT -----------------------------------------------------------------------------------------------
__global__ void sum_rows_kernel(const TVd x, const TVd y) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % x.ne1, i2 = (row / x.ne1) % x.ne2, i3 = row / (x.ne1 * x.ne2);
    const float * xr = (const float *) at(x, 0, i1, i2, i3);
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < x.ne0; i += 32) s += xr[i];
    s = warp_sum(s);
    if (threadIdx.x == 0) *(float *) at(y, 0, i1, i2, i3) = s;
}
int op_sum_rows(const TV & x, const TV & y, cudaStream_t st) {
    const int64_t rows = nrows(x);
    if (rows <= 0) return B200_OK;
    launch_pdl(sum_rows_kernel, dim3((unsigned) rows), dim3(32), 0, st, dv(x), dv(y));
    return (int) cudaGetLastError();
}
__global__ void repeat_kernel(const TVd s, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t i0, i1, i2, i3;
    unravel(i, d, i0, i1, i2, i3);
    *(float *) at(d, i0, i1, i2, i3) = *(const float *) at(s, i0 % s.ne0, i1 % s.ne1, i2 % s.ne2, i3 % s.ne3);
}
int op_repeat(const TV & s, const TV & d, cudaStream_t st) {
    const int64_t n = nelem(d);
    if (n <= 0) return B200_OK;
    launch_pdl(repeat_kernel, dim3((unsigned) ((n + 255) / 256)), dim3(256), 0, st, dv(s), dv(d), n);
    return (int) cudaGetLastError();
}

// ---- TOP_K / ARGSORT on short rows (router logits: ne0 = number of experts <= 1024) ---------------------------------
// one block per row; rank of element i = #{j : x[j] > x[i] or (x[j] == x[i] and j < i)} (stable, descending)
__global__ void argsort_desc_kernel(const TVd x, const TVd y, int k_out, bool ascending, bool swap01) {
    extern __shared__ float sx[];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t row = blockIdx.x;
    const int64_t i1 = row % x.ne1, i2 = (row / x.ne1) % x.ne2, i3 = row / (x.ne1 * x.ne2);
    const float * xr = (const float *) at(x, 0, i1, i2, i3);
    int32_t * yr = (int32_t *) at(y, 0, i1, i2, i3);
    const int n = (int) x.ne0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) sx[i] = xr[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float v = sx[i];
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const float u = sx[j];
            const bool before = ascending ? (u < v || (u == v && j < i)) : (u > v || (u == v && j < i));
            rank += before ? 1 : 0;
        }
        // ggml_compute_forward_top_k_f32 swaps the first two results ("order is not important", ops.cpp:8089-8092)
        if (swap01 && rank < 2) rank ^= 1;
        if (rank < k_out) yr[rank] = i;
    }
}
int op_argsort(const TV & x, const TV & y, int k_out, bool ascending, bool swap01, cudaStream_t st) {
    const int64_t rows = nrows(x);
    if (rows <= 0) return B200_OK;
    if (x.ne[0] > 8192) return B200_ERR_UNSUPPORTED;
    const int threads = x.ne[0] >= 256 ? 256 : 32;
    launch_pdl(argsort_desc_kernel, dim3((unsigned) rows), dim3(threads), (size_t) x.ne[0] * 4, st, dv(x), dv(y), k_out, ascending, swap01 && k_out > 1);
    return (int) cudaGetLastError();
}

// ---- GET_ROWS for float tables with strides (quantized tables: get_rows_q in ops.cu) ---------------------------------
template <typename T0>
__global__ void get_rows_f_kernel(const TVd a, const TVd ids, const TVd d, int64_t n) {
    pdl_launch_dependents();
    pdl_wait();
    const int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t i0, i1, i2, i3;
    unravel(i, d, i0, i1, i2, i3);
    const int64_t r = *(const int32_t *) at(ids, i1, i2, i3, 0);
    *(float *) at(d, i0, i1, i2, i3) = ldf<T0>(at(a, i0, r, i2 % a.ne2, i3 % a.ne3));
}
int op_get_rows_f(const TV & a, const TV & ids, const TV & d, cudaStream_t st) {
    const int64_t n = nelem(d);
    if (n <= 0) return B200_OK;
    const unsigned grid = (unsigned) ((n + 255) / 256);
    if (a.type == B200_TYPE_F32) launch_pdl(get_rows_f_kernel<float>, dim3(grid), dim3(256), 0, st, dv(a), dv(ids), dv(d), n);
    else if (a.type == B200_TYPE_F16) launch_pdl(get_rows_f_kernel<__half>, dim3(grid), dim3(256), 0, st, dv(a), dv(ids), dv(d), n);
    else return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

}  // namespace b200
