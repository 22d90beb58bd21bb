// prefill_tc_desc.h — constants and descriptor encodings of csrc/prefill_tc.cu, shared with the host-side CuTe check
// (tools/cute_layout_check.cu) so that what is verified there is exactly what the kernel uses.
#pragma once
#include <stdint.h>

#ifndef TC_HD
#ifdef __CUDACC__
#define TC_HD __host__ __device__ __forceinline__
#else
#define TC_HD inline
#endif
#endif

#define TC_M 128
#define TC_N 128
#define TC_PLANE_BYTES (TC_M * 256)  // one 128 x 256-byte int8 tile
#define TC_LBO 2048u                 // bytes between the two 16-byte K chunks of a core-matrix column: (TC_M / 8) core matrices of 128 bytes
#define TC_SBO 128u                  // bytes between consecutive 8-row groups
#define TC_TMEM_COLS 512             // 3 accumulators x 128 columns -> next power of two

// shared-memory matrix descriptor, no swizzle, K-major (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version = 1 [46,48), base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0) [61,64)
TC_HD uint64_t tc_desc_lbo(uint32_t saddr, uint32_t lbo_bytes) {
    return (uint64_t) ((saddr >> 4) & 0x3fffu) | ((uint64_t) ((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t) ((TC_SBO >> 4) & 0x3fffu) << 32) | ((uint64_t) 1 << 46);
}
TC_HD uint64_t tc_desc(uint32_t saddr) { return tc_desc_lbo(saddr, TC_LBO); }
// instruction descriptor (mma_sm100_desc.hpp InstrDescriptor): c_format S32 (2) [4,6), a_format / b_format signed 8 bit (1) [7,10) / [10,13),
// K-major A and B, n_dim = N >> 3 [17,23), m_dim = M >> 4 [24,29)
#define TC_IDESC ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (TC_N >> 3) << 17) | ((uint32_t) (TC_M >> 4) << 24))

// the same for an N = 64 tile (block-scaled formats: one accumulator per 32-element block, 8 x 64 = 512 TMEM columns)
#define TC_N64 64
#define TC_LBO64 1024u  // (64 / 8) core matrices of 128 bytes between the two K chunks
#define TC_IDESC64 ((2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (TC_N64 >> 3) << 17) | ((uint32_t) (TC_M >> 4) << 24))
TC_HD uint32_t tc_off64(int i, int c) { return (uint32_t) ((c * (TC_N64 / 8) + (i >> 3)) * 128 + (i & 7) * 16); }

// byte offset of (row-or-column i, 16-byte K chunk c) inside a 128 x 256-byte tile in the canonical no-swizzle K-major layout
TC_HD uint32_t tc_off(int i, int c) { return (uint32_t) ((c * (TC_M / 8) + (i >> 3)) * 128 + (i & 7) * 16); }
