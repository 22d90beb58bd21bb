// Host-side check of csrc/prefill_tc.cu's shared-memory layout against CuTe's canonical UMMA K-major no-swizzle (INTERLEAVE) layout
// and of the LBO / SBO values make_umma_desc would derive for it.
#include <cstdio>
#include <cute/tensor.hpp>
#include <cute/atom/mma_traits_sm90_gmma.hpp>
#include <cute/arch/mma_sm100_desc.hpp>
#include "../chatllm.cpp_b200/csrc/prefill_tc_desc.h"   // tc_off, tc_desc, TC_IDESC: exactly what the kernel uses
using namespace cute;

int main() {
    // canonical atom for int8, K-major, no swizzle: 8 rows x 16 bytes
    auto atom = GMMA::Layout_K_INTER_Atom<int8_t>{};
    print("atom  : "); print(atom); print("\n");
    // our tile: 128 rows x 256 K-bytes, stored as [K chunk of 16 B][row group of 8][8 rows][16 B]
    // build it explicitly as a CuTe layout: (row, k) -> byte offset
    auto mine = make_layout(make_shape(make_shape(Int<8>{}, Int<16>{}), make_shape(Int<16>{}, Int<16>{})),
                            make_stride(make_stride(Int<16>{}, Int<128>{}), make_stride(Int<1>{}, Int<2048>{})));
    print("mine  : "); print(mine); print("\n");
    int bad = 0;
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 256; ++k) {
            unsigned want = tc_off(r, k / 16) + (k % 16);
            unsigned got = (unsigned) mine(r, k);
            if (want != got) { if (bad < 5) printf("mismatch r=%d k=%d want=%u got=%u\n", r, k, want, got); ++bad; }
        }
    printf("tc_off vs explicit layout mismatches: %d\n", bad);
    // what make_umma_desc<Major::K> reads off such a tensor (mma_traits_sm100.hpp): in uint128 units
    auto u128 = recast_layout<int8_t, uint128_t>(mine);   // (128, 16) in 16-byte units
    print("u128  : "); print(u128); print("\n");
    auto canonical = logical_divide(u128, Tile<Layout<_8, _1>, Layout<_2, _1>>{});
    print("canon : "); print(canonical); print("\n");
    printf("stride<0,0> (rows inside a core matrix, must be 1 for SWIZZLE_NONE... SwizzleAtomMNSize=1) = %d\n", (int) stride<0, 0>(canonical));
    printf("stride<0,1> -> SBO (uint128 units) = %d  => %d bytes\n", (int) stride<0, 1>(canonical), 16 * (int) stride<0, 1>(canonical));
    printf("stride<1,0> -> LBO (uint128 units) = %d  => %d bytes\n", (int) stride<1, 0>(canonical), 16 * (int) stride<1, 0>(canonical));
    // the k-step advance used by the kernel: descriptor start moves by 2 K-chunks = 2 * 2048 bytes per K = 32
    printf("offset of (row 0, k 32) = %d bytes (kernel: s * 2 * TC_LBO with TC_LBO = 2048)\n", (int) mine(0, 32));
    // compare with tiling the canonical atom to the same shape
    auto tiled = tile_to_shape(atom, make_shape(Int<128>{}, Int<256>{}));
    print("tiled : "); print(tiled); print("\n");
    // descriptor encodings against CuTe's own bit-field unions
    UMMA::SmemDescriptor sd;
    sd.version_ = 1; sd.lbo_mode_ = 0; sd.layout_type_ = uint8_t(UMMA::LayoutType::SWIZZLE_NONE); sd.base_offset_ = 0;
    const uint32_t saddr = 0x12340u;   // some 16-byte aligned shared-memory address
    sd.start_address_ = uint16_t(saddr >> 4); sd.stride_byte_offset_ = 8; sd.leading_byte_offset_ = 128;
    printf("smem descriptor  cute=%016llx  kernel=%016llx  %s\n", (unsigned long long) uint64_t(sd), (unsigned long long) tc_desc(saddr),
           uint64_t(sd) == tc_desc(saddr) ? "EQUAL" : "DIFFERENT");
    auto id = UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 128, UMMA::Major::K, UMMA::Major::K>();
    printf("instr descriptor cute=%08x  kernel=%08x  %s\n", uint32_t(id), (uint32_t) TC_IDESC, uint32_t(id) == (uint32_t) TC_IDESC ? "EQUAL" : "DIFFERENT");
    // ---- the N = 64 activation tile of the block-scaled kernels (mmq_tc_blk32_kernel)
    auto tiled64 = tile_to_shape(atom, make_shape(Int<64>{}, Int<256>{}));
    int bad64 = 0;
    for (int r = 0; r < 64; ++r)
        for (int k = 0; k < 256; ++k) bad64 += (tc_off64(r, k / 16) + (k % 16)) != (unsigned) tiled64.layout_b()(r, k);
    printf("tc_off64 vs tiled canonical layout mismatches: %d\n", bad64);
    auto canon64 = logical_divide(recast_layout<int8_t, uint128_t>(tiled64.layout_b()), Tile<Layout<_8, _1>, Layout<_2, _1>>{});
    printf("N=64 tile: SBO (uint128 units) = %d , LBO (uint128 units) = %d , kernel TC_LBO64 = %u bytes\n", (int) stride<0, 1>(canon64), (int) stride<1, 0>(canon64), TC_LBO64);
    auto id64 = UMMA::make_instr_desc<int8_t, int8_t, int32_t, 128, 64, UMMA::Major::K, UMMA::Major::K>();
    printf("instr descriptor N=64 cute=%08x  kernel=%08x  %s\n", uint32_t(id64), (uint32_t) TC_IDESC64, uint32_t(id64) == (uint32_t) TC_IDESC64 ? "EQUAL" : "DIFFERENT");
    sd.leading_byte_offset_ = uint16_t(stride<1, 0>(canon64));
    printf("smem descriptor N=64 cute=%016llx  kernel=%016llx  %s\n", (unsigned long long) uint64_t(sd), (unsigned long long) tc_desc_lbo(saddr, TC_LBO64),
           uint64_t(sd) == tc_desc_lbo(saddr, TC_LBO64) ? "EQUAL" : "DIFFERENT");
    return 0;
}
