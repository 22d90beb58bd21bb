// dequant.cuh — element-wise dequantization of one row of a quantized table (embedding gather), shared by ops.cu (get_rows_kernel)
// and decode_mk.cu.  Reference: dequantize_row_q4_K / q4_0 / q8_0 (ggml/src/ggml-quants.c:1352-1373, :307-325, :401-414) behind
// ggml_compute_forward_get_rows (ggml/src/ggml-cpu/ops.cpp:4820).  Q4_0 / Q8_0 rows are in the repacked SoA layout (quantize.cu).
#pragma once
#include "common.cuh"

namespace b200 {

__host__ __device__ inline int64_t dequant_row_bytes(int type, int64_t k) {
    switch (type) {
        case B200_TYPE_Q4_K: return k / 256 * 144;
        case B200_TYPE_Q4_0: return k / 32 * 18;
        case B200_TYPE_Q8_0: return k / 32 * 34;
        case B200_TYPE_F16: return k * 2;
        default: return k * 4;
    }
}

__device__ __forceinline__ float dequant_row_elem(int type, const uint8_t * base, int64_t k, int64_t e) {
    if (type == B200_TYPE_Q4_K) {
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
        return (d * sc) * qv - (dmin * mn);
    }
    if (type == B200_TYPE_Q4_0) {
        const int64_t nb = k / 32;
        const int64_t b = e >> 5;
        const int w = (int) (e & 31);
        const uint8_t * dp = base + nb * 16 + b * 2;
        const float d = half_bits_to_float(dp[0] | (dp[1] << 8));
        const uint8_t byte = base[b * 16 + (w & 15)];
        const int qv = (w < 16 ? (byte & 0xF) : (byte >> 4)) - 8;
        return qv * d;
    }
    if (type == B200_TYPE_Q8_0) {
        const int64_t nb = k / 32;
        const int64_t b = e >> 5;
        const uint8_t * dp = base + nb * 32 + b * 2;
        const float d = half_bits_to_float(dp[0] | (dp[1] << 8));
        return (float) ((const int8_t *) base)[e] * d;
    }
    if (type == B200_TYPE_F16) return __half2float(((const __half *) base)[e]);
    return ((const float *) base)[e];
}

}  // namespace b200
