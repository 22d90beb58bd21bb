// quantize.cu — on-device activation quantization that reproduces the reference CPU backend bit-for-bit.
//
// The reference computes every quantized matmul as  dequant(W) . dequant(Q(x))  where x is first converted
// to the weight type's vec_dot_type (ggml/src/ggml-cpu/ggml-cpu.c:1291-1326, type table :207-308):
//   Q4_K weights           -> Q8_K activations: quantize_row_q8_K_ref (ggml/src/ggml-quants.c:2555-2592)
//   Q4_0 / Q8_0 weights    -> Q8_0 activations: x86 quantize_row_q8_0 (ggml/src/ggml-cpu/arch/x86/quants.c:290-384)
// Matching the integer codes exactly makes the integer dot products exact; only fp32 summation order differs.
//
// Device layout of one quantized activation column ("qact", ours — not the reference's AoS blocks): actlayout.cuh.
// Columns are laid out back to back with stride qact_col_bytes().
#include "actlayout.cuh"
#include "common.cuh"
#include "kernels.h"

namespace b200 {

size_t qact_col_bytes(int wtype, int64_t k) { return (size_t) act_layout(wtype == B200_TYPE_Q4_K, k).col_bytes; }

// ---- Q8_K: one warp-group of 256 threads per 256-element block ----------------------------------------
// max = the element with the largest |x| (FIRST occurrence on ties, ggml-quants.c:2563-2567), iscale = -127/max,
// q = nearest_int(iscale * x) clamped to 127 (RNE, :444-449), d = 1/iscale.
__global__ void __launch_bounds__(256) quantize_q8_K_kernel(const float * x, int64_t x_col_stride, int64_t k,
                                                            uint8_t * qact, size_t col_bytes) {
    pdl_launch_dependents();
    pdl_wait();
    const int blk = blockIdx.x;
    const int col = blockIdx.y;
    const int tid = threadIdx.x;
    const float v = x[(int64_t) col * x_col_stride + (int64_t) blk * 256 + tid];

    // argmax |v| with smallest index on ties: pack (|v| bits, 255 - idx) into a 64-bit key and take the max.
    // |v| as uint32 bits is monotone for non-negative floats (NaN not expected).
    unsigned long long key = ((unsigned long long) __float_as_uint(fabsf(v)) << 32) | (unsigned) (255 - tid);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    __shared__ unsigned long long wkey[8];
    __shared__ float s_max;
    if ((tid & 31) == 0) wkey[tid >> 5] = key;
    __syncthreads();
    if (tid < 32) {
        unsigned long long kk = tid < 8 ? wkey[tid] : 0ull;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, kk, o);
            kk = other > kk ? other : kk;
        }
        if (tid == 0) {
            const int idx = 255 - (int) (kk & 0xffffffffu);
            s_max = x[(int64_t) col * x_col_stride + (int64_t) blk * 256 + idx];
        }
    }
    __syncthreads();
    const float mx = s_max;

    const ActLayout L = act_layout(true, k);
    uint8_t * base = qact + (size_t) col * col_bytes;
    int8_t * qs = (int8_t *) base;
    float * d = (float *) (base + L.d_off);
    int16_t * bs = (int16_t *) (base + L.bs_off);

    int q = 0;
    float dd = 0.0f;
    if (mx != 0.0f) {  // amax != 0  (|mx| == amax)
        const float iscale = __fdiv_rn(-127.f, mx);
        q = min(127, __float2int_rn(__fmul_rn(iscale, v)));
        dd = __fdiv_rn(1.0f, iscale);
    }
    qs[act_qs_off_q8k((int64_t) blk * 256 + tid)] = (int8_t) q;
    int s = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((tid & 31) == 0) bs[blk * 8 + (tid >> 5)] = (int16_t) s;
    if (tid == 0) d[blk] = dd;
    // zero the padding chunks of the last (partial) 1024-element group so the GEMV can read them blindly
    if (blk == (int) (k / 256) - 1) {
        for (int64_t e = k + tid; e < L.qs_bytes; e += 256) qs[act_qs_off_q8k(e)] = 0;
    }
}

// ---- Q8_0 (x86 variant): one warp per 32-element block --------------------------------------------------
// amax = max|x|, d = amax/127 stored as fp16, id = 127/amax (0 if amax == 0), q = RNE(x*id).
__global__ void __launch_bounds__(256) quantize_q8_0_kernel(const float * x, int64_t x_col_stride, int64_t k,
                                                            uint8_t * qact, size_t col_bytes) {
    pdl_launch_dependents();
    pdl_wait();
    const int col = blockIdx.y;
    const int64_t blk = (int64_t) blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (blk * 32 >= k) return;
    const float v = x[(int64_t) col * x_col_stride + blk * 32 + lane];
    const float amax = warp_max(fabsf(v));
    const float dd = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q = __float2int_rn(__fmul_rn(v, id));

    const ActLayout L = act_layout(false, k);
    uint8_t * base = qact + (size_t) col * col_bytes;
    int8_t * qs = (int8_t *) base;
    float * d = (float *) (base + L.d_off);
    int * bs = (int *) (base + L.bs_off);
    qs[act_qs_off_q80(blk * 32 + lane)] = (int8_t) q;
    int s = q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        d[blk] = __half2float(__float2half_rn(dd));
        bs[blk] = s;
    }
}

int quantize_act(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * qact, cudaStream_t st) {
    if (n <= 0 || k <= 0) return B200_OK;
    const size_t cb = qact_col_bytes(wtype, k);
    if (wtype == B200_TYPE_Q4_K) {
        if (k % 256) return B200_ERR_ARG;
        dim3 grid((unsigned) (k / 256), (unsigned) n);
        launch_pdl(quantize_q8_K_kernel, dim3(grid), dim3(256), 0, st, x, x_col_stride, k, (uint8_t *) qact, cb);
    } else if (wtype == B200_TYPE_Q4_0 || wtype == B200_TYPE_Q8_0) {
        if (k % 32) return B200_ERR_ARG;
        dim3 grid((unsigned) ((k / 32 + 7) / 8), (unsigned) n);
        launch_pdl(quantize_q8_0_kernel, dim3(grid), dim3(256), 0, st, x, x_col_stride, k, (uint8_t *) qact, cb);
    } else {
        return B200_ERR_UNSUPPORTED;
    }
    return (int) cudaGetLastError();
}

// ---- weight repack (Q4_0 / Q8_0): AoS blocks -> per-row SoA so that every quant word is 16-byte aligned ----
// native row:  nb x { fp16 d ; qs[QB] }                (18 / 34 bytes per block: only 2-byte aligned)
// device row:  qs[nb][QB]  followed by  d[nb] (fp16)    (same byte count; row stride unchanged)
// The mapping is per byte, so arbitrary (offset, size) windows written by ggml_backend_tensor_set
// (1 MiB chunks, src/chat.cpp:1322-1338) can be converted independently.
template <int QB>  // quant bytes per block: 16 (Q4_0) or 32 (Q8_0)
__global__ void repack_bytes_kernel(const uint8_t * src, uint8_t * dst_tensor, int64_t tensor_off,
                                    int64_t nbytes, int64_t nb_row, bool inverse) {
    // launched through launch_pdl like every kernel of the library: a kernel of a PDL chain that never waits could finish while its
    // predecessor is still running and break the transitive ordering every later kernel relies on (common.cuh)
    pdl_launch_dependents();
    pdl_wait();
    const int BB = QB + 2;
    const int64_t row_bytes = nb_row * BB;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < nbytes; i += (int64_t) gridDim.x * blockDim.x) {
        const int64_t o = tensor_off + i;  // byte offset in the NATIVE layout
        const int64_t row = o / row_bytes;
        const int64_t r = o - row * row_bytes;
        const int64_t b = r / BB;
        const int w = (int) (r - b * BB);
        const int64_t p = row * row_bytes + (w < 2 ? nb_row * QB + b * 2 + w : b * QB + (w - 2));  // offset in device layout
        if (!inverse) dst_tensor[p] = src[i];
        else ((uint8_t *) src)[i] = dst_tensor[p];  // inverse: `src` is the native-layout output window
    }
}

int repack_window(int wtype, const void * host_layout_window, void * dev_tensor, int64_t tensor_off, int64_t nbytes, int64_t k,
                  bool inverse, cudaStream_t st) {
    if (nbytes <= 0) return B200_OK;
    const int64_t nb_row = k / 32;
    const int threads = 256;
    const int blocks = (int) ((nbytes + threads - 1) / threads > 4096 ? 4096 : (nbytes + threads - 1) / threads);
    if (wtype == B200_TYPE_Q4_0)
        launch_pdl(repack_bytes_kernel<16>, dim3(blocks), dim3(threads), 0, st, (const uint8_t *) host_layout_window, (uint8_t *) dev_tensor, tensor_off, nbytes, nb_row, inverse);
    else if (wtype == B200_TYPE_Q8_0)
        launch_pdl(repack_bytes_kernel<32>, dim3(blocks), dim3(threads), 0, st, (const uint8_t *) host_layout_window, (uint8_t *) dev_tensor, tensor_off, nbytes, nb_row, inverse);
    else
        return B200_ERR_UNSUPPORTED;
    return (int) cudaGetLastError();
}

}  // namespace b200
