// actlayout.cuh — device layout of a quantized activation column ("qact"), shared by quantize.cu (writer) and
// gemv.cu (reader).  The int8 codes are stored in 16-byte chunks permuted so that the GEMV kernel's per-lane
// 16-byte shared-memory loads are bank-conflict free (each quarter-warp reads 128 contiguous bytes):
//
//  Q8_K codes (for Q4_K weights; 2 lanes per 256-element unit, lane h owns elements [128h, 128h+128) = chunks j=0..7)
//      byte offset(e) = (u>>2)*1024 + j*128 + (u&3)*32 + h*16 + (e&15),   u = e>>8, h = (e>>7)&1, j = (e>>4)&7
//      followed by  float d[k/256]  and  int16 bs[k/32]  (sum of codes per 32-element sub-block = reference bsums pairs)
//  Q8_0 codes (for Q4_0 / Q8_0 weights; one lane per 32-element block, chunks j=0..1)
//      byte offset(e) = (b>>3)*256 + j*128 + (b&7)*16 + (e&15),           b = e>>5, j = (e>>4)&1
//      followed by  float d[k/32] (rounded through fp16 like block_q8_0.d)  and  int32 bs[k/32]
#pragma once
#include <stdint.h>

namespace b200 {

__host__ __device__ inline int64_t al16(int64_t x) { return (x + 15) & ~(int64_t) 15; }

struct ActLayout {
    int64_t qs_bytes, d_off, bs_off, col_bytes;
};

__host__ __device__ inline ActLayout act_layout(bool q8k, int64_t k) {
    ActLayout L;
    if (q8k) {
        L.qs_bytes = (k + 1023) / 1024 * 1024;
        L.d_off = L.qs_bytes;
        L.bs_off = L.d_off + al16(k / 256 * 4);
        L.col_bytes = L.bs_off + al16(k / 32 * 2);
    } else {
        L.qs_bytes = al16(k);
        L.d_off = L.qs_bytes;
        L.bs_off = L.d_off + al16(k / 32 * 4);
        L.col_bytes = L.bs_off + al16(k / 32 * 4);
    }
    return L;
}

__host__ __device__ inline int64_t act_qs_off_q8k(int64_t e) {
    const int64_t u = e >> 8;
    const int h = (int) (e >> 7) & 1, j = (int) (e >> 4) & 7, b = (int) e & 15;
    return (u >> 2) * 1024 + j * 128 + (u & 3) * 32 + h * 16 + b;
}
__host__ __device__ inline int64_t act_qs_off_q80(int64_t e) {
    const int64_t blk = e >> 5;
    const int j = (int) (e >> 4) & 1, b = (int) e & 15;
    return (blk >> 3) * 256 + j * 128 + (blk & 7) * 16 + b;
}

}  // namespace b200
