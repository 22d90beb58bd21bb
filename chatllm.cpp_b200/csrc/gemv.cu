// gemv.cu — quantized GEMV / skinny GEMM for single-token decode on sm_100a.
//
// Replaces, for src0 in {Q4_K, Q4_0, Q8_0} and 1..8 activation columns, the reference's
//   CPU:  ggml_compute_forward_mul_mat      ggml/src/ggml-cpu/ggml-cpu.c:1229-1421 (+ vec_dot kernels quants.c:115/305/550)
//   CUDA: mul_mat_vec_q                     ggml/src/ggml-cuda/mmvq.cu:140-356
// with a design built around the B200 memory system instead of the reference's one-row-per-block dp4a kernel:
//
//   * HBM-bound (Q4_K: 0.5625 B/weight, ~3.6 FLOP/B).  The weight matrix is a contiguous byte stream, so every
//     warp owns a contiguous range of rows and streams it through its own ring of shared-memory stages with 1-D
//     bulk async copies (cp.async.bulk -> TMA engine, SASS UBLKCP) completing on mbarriers.  No warp ever waits on
//     a global load; bytes in flight per SM = warps x (stages-1) x stage bytes (~100 KB), independent of occupancy.
//   * warps are fully independent pipelines (the same warp produces and consumes its stages): no block-wide
//     barriers in the main loop, no producer/consumer hand-off, perfectly even byte split over all SMs
//     (rows_per_warp differs by at most one row).
//   * activations are pre-quantized exactly like the reference (quantize.cu) and live in shared memory for the whole
//     kernel; integer dot products use dp4a (u8 x s8); RG rows share each activation fetch.
//   * weights are read from HBM exactly once with an L2 evict-first policy (they are not re-used within a token).
//
// Weight layouts in HBM: Q4_K native ggml blocks (144 B, 16-byte aligned).  Q4_0 / Q8_0 per-row SoA
// ("qs[nb][QB] then d[nb]", see quantize.cu repack_window) because 18 / 34-byte AoS blocks cannot be read with
// aligned 16-byte shared-memory loads.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace b200 {

__host__ __device__ inline int64_t a16(int64_t x) { return (x + 15) & ~(int64_t) 15; }

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
__device__ __forceinline__ uint4 lds128(const void * p) { return *reinterpret_cast<const uint4 *>(p); }

// view of one quantized activation column in shared memory (layout: quantize.cu)
struct ActView {
    const int8_t * qs;
    const float * d;
    const int * bs;
};

// ======================================================================================================
// Format traits.  A "unit" is 256 consecutive k-elements for every format.
//   stage layout in smem: region A = RG rows x (KS units x A_UNIT bytes), region B = RG rows x (KS units x B_UNIT bytes)
// ======================================================================================================
struct FmtQ4K {
    static constexpr int A_UNIT = 144;  // one native block_q4_K per unit
    static constexpr int B_UNIT = 0;
    static constexpr int ACT_G = 256;  // activation scale granularity (Q8_K)
    static constexpr int LANES_PER_UNIT = 4;  // each lane handles 64 elements (one 32-byte qs group = sub-blocks 2g, 2g+1)
    __device__ static __forceinline__ const uint8_t * rowA(const uint8_t * W, int64_t row, int64_t nunits) { return W + row * nunits * 144; }
    __device__ static __forceinline__ const uint8_t * rowB(const uint8_t * W, int64_t row, int64_t nunits) { return nullptr; }

    // one row, one unit, one lane-group g: returns this lane's contribution
    template <int NC>
    __device__ static __forceinline__ void dot(const uint8_t * a_row, const uint8_t * /*b_row*/, int u, int g, int64_t gu, const ActView * act,
                                               float * acc /*[NC]*/) {
        const uint8_t * blk = a_row + u * 144;
        const uint4 hdr = lds128(blk);
        const uint4 q0 = lds128(blk + 16 + g * 32);
        const uint4 q1 = lds128(blk + 32 + g * 32);
        const float d = half_bits_to_float(hdr.x & 0xffffu);
        const float dmin = half_bits_to_float(hdr.x >> 16);
        // 6-bit scales / mins (reference get_scale_min_k4, ggml/src/ggml-quants.c:703-711), 4 at a time
        const uint32_t sc_a = hdr.y & 0x3f3f3f3fu;
        const uint32_t sc_b = (hdr.w & 0x0f0f0f0fu) | (((hdr.y >> 6) & 0x03030303u) << 4);
        const uint32_t mn_a = hdr.z & 0x3f3f3f3fu;
        const uint32_t mn_b = ((hdr.w >> 4) & 0x0f0f0f0fu) | (((hdr.z >> 6) & 0x03030303u) << 4);
        const uint32_t scw = (g < 2) ? sc_a : sc_b;
        const uint32_t mnw = (g < 2) ? mn_a : mn_b;
        const int sh = (g & 1) * 16;
        const int sc0 = (scw >> sh) & 0xff, sc1 = (scw >> (sh + 8)) & 0xff;
        const int mn0 = (mnw >> sh) & 0xff, mn1 = (mnw >> (sh + 8)) & 0xff;
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int8_t * aq = act[c].qs + gu * 256 + g * 64;
            const uint4 a0 = lds128(aq), a1 = lds128(aq + 16), a2 = lds128(aq + 32), a3 = lds128(aq + 48);
            const uint32_t al[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const uint32_t ah[8] = {a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
            int slo = 0, shi = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                slo = dp4a_us(w[i] & 0x0f0f0f0fu, al[i], slo);
                shi = dp4a_us(w[i] & 0xf0f0f0f0u, ah[i], shi);  // = 16 * sum(hi nibble * a)
            }
            const int2 bs = *reinterpret_cast<const int2 *>(act[c].bs + gu * 8 + g * 2);
            const int t = ((sc0 * slo) << 4) + sc1 * shi;  // 16 * sum_j sc_j * isum_j
            const int ms = mn0 * bs.x + mn1 * bs.y;
            const float dx = act[c].d[gu];
            // reference association (arch/x86/quants.c:1764-1815): d = y.d * fp16(x.d); dmin = y.d * fp16(x.dmin)
            acc[c] = fmaf(dx * d * 0.0625f, (float) t, acc[c]);
            acc[c] = fmaf(-(dx * dmin), (float) ms, acc[c]);
        }
    }
};

// Q4_0 repacked rows: A = qs (16 B per 32-element block, 128 B per unit), B = fp16 d (2 B per block, 16 B per unit)
struct FmtQ40 {
    static constexpr int A_UNIT = 128;
    static constexpr int B_UNIT = 16;
    static constexpr int ACT_G = 32;
    static constexpr int LANES_PER_UNIT = 8;  // one 32-element block per lane
    __device__ static __forceinline__ const uint8_t * rowA(const uint8_t * W, int64_t row, int64_t nunits) { return W + row * nunits * 144; }
    __device__ static __forceinline__ const uint8_t * rowB(const uint8_t * W, int64_t row, int64_t nunits) {
        return W + row * nunits * 144 + nunits * 128;
    }
    template <int NC>
    __device__ static __forceinline__ void dot(const uint8_t * a_row, const uint8_t * b_row, int u, int g, int64_t gu, const ActView * act,
                                               float * acc) {
        const uint4 q = lds128(a_row + u * 128 + g * 16);
        const float d = half_bits_to_float(*reinterpret_cast<const unsigned short *>(b_row + u * 16 + g * 2));
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int64_t gb = gu * 8 + g;  // global 32-block index
            const int8_t * aq = act[c].qs + gb * 32;
            const uint4 a0 = lds128(aq), a1 = lds128(aq + 16);
            const uint32_t al[4] = {a0.x, a0.y, a0.z, a0.w};
            const uint32_t ah[4] = {a1.x, a1.y, a1.z, a1.w};
            int slo = 0, shi = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                slo = dp4a_us(w[i] & 0x0f0f0f0fu, al[i], slo);   // elements 0..15  (low nibbles, ggml-quants.c:307-325)
                shi = dp4a_us(w[i] & 0xf0f0f0f0u, ah[i], shi);   // 16 * elements 16..31 (high nibbles)
            }
            // sum (q-8)*a = sum q*a - 8*sum a
            const int t = (slo << 4) + shi - (act[c].bs[gb] << 7);
            acc[c] = fmaf(d * act[c].d[gb] * 0.0625f, (float) t, acc[c]);
        }
    }
};

// Q8_0 repacked rows: A = qs (32 B per block, 256 B per unit), B = fp16 d (16 B per unit)
struct FmtQ80 {
    static constexpr int A_UNIT = 256;
    static constexpr int B_UNIT = 16;
    static constexpr int ACT_G = 32;
    static constexpr int LANES_PER_UNIT = 8;
    __device__ static __forceinline__ const uint8_t * rowA(const uint8_t * W, int64_t row, int64_t nunits) { return W + row * nunits * 272; }
    __device__ static __forceinline__ const uint8_t * rowB(const uint8_t * W, int64_t row, int64_t nunits) {
        return W + row * nunits * 272 + nunits * 256;
    }
    template <int NC>
    __device__ static __forceinline__ void dot(const uint8_t * a_row, const uint8_t * b_row, int u, int g, int64_t gu, const ActView * act,
                                               float * acc) {
        const uint4 q0 = lds128(a_row + u * 256 + g * 32);
        const uint4 q1 = lds128(a_row + u * 256 + g * 32 + 16);
        const float d = half_bits_to_float(*reinterpret_cast<const unsigned short *>(b_row + u * 16 + g * 2));
        const uint32_t w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int64_t gb = gu * 8 + g;
            const int8_t * aq = act[c].qs + gb * 32;
            const uint4 a0 = lds128(aq), a1 = lds128(aq + 16);
            const uint32_t a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            int s = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = dp4a_ss(w[i], a[i], s);
            acc[c] = fmaf(d * act[c].d[gb], (float) s, acc[c]);
        }
    }
};

// ======================================================================================================
// The kernel: every warp is an independent bulk-copy pipeline over its own contiguous row range.
// ======================================================================================================
struct GemvParams {
    const uint8_t * W;
    const uint8_t * qact;  // global, n columns
    float * y;
    const float * bias;
    int64_t k, m, ldy;
    int n;
    int nunits;     // k / 256
    int ks;         // units per stage
    int stages;
    uint32_t act_col_bytes;
};

template <class F, int RG, int NC>
__global__ void __launch_bounds__(512) gemv_q_kernel(const GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    // ---- shared memory carve-up: [activations][per-warp: barriers | stages]
    const uint32_t act_bytes = (uint32_t) a16((int64_t) p.act_col_bytes * NC);
    const uint32_t stageA = (uint32_t) RG * p.ks * F::A_UNIT;
    const uint32_t stageB = (uint32_t) RG * p.ks * F::B_UNIT;
    const uint32_t stage_bytes = stageA + stageB;
    uint8_t * act_s = smem;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem + act_bytes) + (size_t) warp * p.stages;
    uint8_t * ring = smem + act_bytes + a16((int64_t) nwarps * p.stages * 8) + (size_t) warp * p.stages * stage_bytes;

    // ---- this warp's rows and work items (row-group x k-segment)
    const int64_t gw = (int64_t) blockIdx.x * nwarps + warp;
    const int64_t GW = (int64_t) gridDim.x * nwarps;
    const int64_t r0 = p.m * gw / GW, r1 = p.m * (gw + 1) / GW;
    const int nseg = (p.nunits + p.ks - 1) / p.ks;
    const int ngroups = (int) ((r1 - r0 + RG - 1) / RG);
    const int nitems = ngroups * nseg;

    if (lane == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    __syncwarp();

    uint64_t pol = 0;
    if (lane == 0) pol = make_evict_first_policy();

    auto issue = [&](int it) {  // lane 0 only
        const int s = it % p.stages;
        const int grp = it / nseg, seg = it - grp * nseg;
        const int64_t row0 = r0 + (int64_t) grp * RG;
        const int nr = (int) min((int64_t) RG, r1 - row0);
        const int u0 = seg * p.ks;
        const int nu = min(p.ks, p.nunits - u0);
        uint8_t * st = ring + (size_t) s * stage_bytes;
        mbar_arrive_expect_tx(&bars[s], (uint32_t) nr * nu * (F::A_UNIT + F::B_UNIT));
        for (int r = 0; r < nr; ++r) {
            bulk_g2s_hint(st + (size_t) r * p.ks * F::A_UNIT, F::rowA(p.W, row0 + r, p.nunits) + (size_t) u0 * F::A_UNIT, (uint32_t) nu * F::A_UNIT,
                          &bars[s], pol);
            if (F::B_UNIT)
                bulk_g2s_hint(st + stageA + (size_t) r * p.ks * F::B_UNIT, F::rowB(p.W, row0 + r, p.nunits) + (size_t) u0 * F::B_UNIT,
                              (uint32_t) nu * F::B_UNIT, &bars[s], pol);
        }
    };

    // weights do not depend on the producer kernel: start streaming them before the PDL dependency resolves
    if (lane == 0) {
        const int pre = min(p.stages, nitems);
        for (int it = 0; it < pre; ++it) issue(it);
    }

    // ---- activations: wait for the producer kernel (PDL), then stage the quantized columns in smem
    pdl_wait();
    {
        const uint4 * src = reinterpret_cast<const uint4 *>(p.qact);
        uint4 * dst = reinterpret_cast<uint4 *>(act_s);
        const int n16 = (int) (((size_t) p.act_col_bytes * NC) >> 4);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    pdl_launch_dependents();

    ActView act[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const uint8_t * b = act_s + (size_t) c * p.act_col_bytes;
        act[c].qs = reinterpret_cast<const int8_t *>(b);
        act[c].d = reinterpret_cast<const float *>(b + a16(p.k));
        act[c].bs = reinterpret_cast<const int *>(b + a16(p.k) + a16(p.k / F::ACT_G * 4));
    }

    constexpr int LPU = F::LANES_PER_UNIT;
    constexpr int UPS = 32 / LPU;  // units per warp step
    const int g = lane % LPU;
    const int ul = lane / LPU;

    float acc[RG][NC];
    for (int it = 0; it < nitems; ++it) {
        const int s = it % p.stages;
        const uint32_t parity = (uint32_t) (it / p.stages) & 1u;
        const int grp = it / nseg, seg = it - grp * nseg;
        const int64_t row0 = r0 + (int64_t) grp * RG;
        const int nr = (int) min((int64_t) RG, r1 - row0);
        const int u0 = seg * p.ks;
        const int nu = min(p.ks, p.nunits - u0);
        if (seg == 0) {
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[r][c] = 0.0f;
        }
        mbar_wait(&bars[s], parity);
        const uint8_t * st = ring + (size_t) s * stage_bytes;
        for (int u = ul; u < nu; u += UPS) {
#pragma unroll
            for (int r = 0; r < RG; ++r) {
                if (r < nr) F::template dot<NC>(st + (size_t) r * p.ks * F::A_UNIT, st + stageA + (size_t) r * p.ks * F::B_UNIT, u, g, u0 + u, act, acc[r]);
            }
        }
        __syncwarp();
        if (lane == 0 && it + p.stages < nitems) issue(it + p.stages);

        if (seg == nseg - 1) {
#pragma unroll
            for (int r = 0; r < RG; ++r) {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const float v = warp_sum(acc[r][c]);
                    if (lane == 0 && r < nr && c < p.n) {
                        float o = v;
                        if (p.bias) o += p.bias[row0 + r];
                        p.y[(int64_t) c * p.ldy + row0 + r] = o;
                    }
                }
            }
        }
    }
}

// ======================================================================================================
// host side
// ======================================================================================================
static int g_sms = 0;
int sm_count() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_sms <= 0) g_sms = 148;
    }
    return g_sms;
}

static int env_int(const char * name, int dflt) {
    const char * v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <class F, int RG, int NC>
static int launch(const GemvParams & p, int warps, int grid, size_t smem_bytes, cudaStream_t st) {
    auto kern = gemv_q_kernel<F, RG, NC>;
    static size_t configured = 0;  // per template instantiation
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
        if (e != cudaSuccess) return (int) e;
        configured = smem_bytes;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) grid);
    cfg.blockDim = dim3((unsigned) warps * 32);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return (int) cudaLaunchKernelEx(&cfg, kern, p);
}

template <class F, int RG>
static int launch_nc(const GemvParams & p, int nc, int warps, int grid, size_t smem, cudaStream_t st) {
    switch (nc) {
        case 1: return launch<F, RG, 1>(p, warps, grid, smem, st);
        case 2: return launch<F, RG, 2>(p, warps, grid, smem, st);
        case 4: return launch<F, RG, 4>(p, warps, grid, smem, st);
        default: return B200_ERR_ARG;
    }
}

template <class F>
static int launch_rg(const GemvParams & p, int rg, int nc, int warps, int grid, size_t smem, cudaStream_t st) {
    switch (rg) {
        case 1: return launch_nc<F, 1>(p, nc, warps, grid, smem, st);
        case 2: return launch_nc<F, 2>(p, nc, warps, grid, smem, st);
        case 4: return launch_nc<F, 4>(p, nc, warps, grid, smem, st);
        default: return B200_ERR_ARG;
    }
}

// y[c*ldy + i] = sum_k W[i,k] * x_c[k]   for c < n (n <= 8 handled in column groups of <= 4)
int mul_mat_q(int wtype, const void * W, int64_t k, int64_t m, const void * qact, int64_t n, float * y, int64_t ldy, const float * bias,
              const GemvTuning * tune, cudaStream_t st) {
    if (k <= 0 || m <= 0 || n <= 0) return B200_OK;
    if (k % 256) return B200_ERR_UNSUPPORTED;
    int a_unit, b_unit;
    switch (wtype) {
        case B200_TYPE_Q4_K: a_unit = FmtQ4K::A_UNIT; b_unit = FmtQ4K::B_UNIT; break;
        case B200_TYPE_Q4_0: a_unit = FmtQ40::A_UNIT; b_unit = FmtQ40::B_UNIT; break;
        case B200_TYPE_Q8_0: a_unit = FmtQ80::A_UNIT; b_unit = FmtQ80::B_UNIT; break;
        default: return B200_ERR_UNSUPPORTED;
    }
    const size_t acb = qact_col_bytes(wtype, k);
    const int nunits = (int) (k / 256);
    const int sms = sm_count();

    for (int64_t c0 = 0; c0 < n; c0 += 4) {
        const int ncols = (int) ((n - c0) < 4 ? (n - c0) : 4);
        const int nc = ncols == 3 ? 4 : ncols;  // template width (1, 2, 4); the 4th column of a 3-wide group is masked

        GemvTuning t;
        t.rg = tune && tune->rg ? tune->rg : env_int("B200_GEMV_RG", 4);
        t.warps = tune && tune->warps ? tune->warps : env_int("B200_GEMV_WARPS", 8);
        t.stages = tune && tune->stages ? tune->stages : env_int("B200_GEMV_STAGES", 4);
        t.ks = tune && tune->ks ? tune->ks : env_int("B200_GEMV_KS", 8);
        t.grid = tune && tune->grid ? tune->grid : env_int("B200_GEMV_GRID", 0);
        if (t.ks > nunits) t.ks = nunits;
        // few rows: keep at least one full row-group per warp
        while (t.rg > 1 && m < (int64_t) sms * t.warps * t.rg / 2) t.rg >>= 1;
        int grid = t.grid > 0 ? t.grid : sms;
        {
            const int64_t max_warps = (m + t.rg - 1) / t.rg;
            const int64_t max_grid = (max_warps + t.warps - 1) / t.warps;
            if (grid > max_grid) grid = (int) max_grid;
        }
        const size_t act_bytes = (size_t) a16((int64_t) acb * nc);
        auto smem_for = [&](const GemvTuning & q) {
            return act_bytes + (size_t) a16((int64_t) q.warps * q.stages * 8) + (size_t) q.warps * q.stages * q.rg * q.ks * (a_unit + b_unit);
        };
        const size_t limit = 227 * 1024;
        while (smem_for(t) > limit && t.stages > 2) t.stages--;
        while (smem_for(t) > limit && t.ks > 1) t.ks = (t.ks + 1) / 2;
        while (smem_for(t) > limit && t.warps > 1) t.warps >>= 1;
        if (smem_for(t) > limit) return B200_ERR_UNSUPPORTED;

        GemvParams p;
        p.W = (const uint8_t *) W;
        // a 3-wide group reads one column past the end of qact: the caller's buffer always holds >= n columns; the
        // extra column (if any) is garbage-in and masked at the store (c < p.n).  For the very last group we must not
        // read out of bounds, so narrow to what exists.
        p.qact = (const uint8_t *) qact + (size_t) c0 * acb;
        p.y = y + c0 * ldy;
        p.bias = bias;
        p.k = k; p.m = m; p.ldy = ldy;
        p.n = ncols;
        p.nunits = nunits;
        p.ks = t.ks;
        p.stages = t.stages;
        p.act_col_bytes = (uint32_t) acb;
        int rc;
        if (ncols == 3) {
            // run as 2 + 1 to stay inside the qact buffer
            GemvParams p2 = p; p2.n = 2;
            const size_t s2 = (size_t) a16((int64_t) acb * 2) + (smem_for(t) - act_bytes);
            switch (wtype) {
                case B200_TYPE_Q4_K: rc = launch_rg<FmtQ4K>(p2, t.rg, 2, t.warps, grid, s2, st); break;
                case B200_TYPE_Q4_0: rc = launch_rg<FmtQ40>(p2, t.rg, 2, t.warps, grid, s2, st); break;
                default: rc = launch_rg<FmtQ80>(p2, t.rg, 2, t.warps, grid, s2, st); break;
            }
            if (rc) return rc;
            GemvParams p1 = p; p1.n = 1; p1.qact += 2 * acb; p1.y += 2 * ldy;
            const size_t s1 = (size_t) a16((int64_t) acb) + (smem_for(t) - act_bytes);
            switch (wtype) {
                case B200_TYPE_Q4_K: rc = launch_rg<FmtQ4K>(p1, t.rg, 1, t.warps, grid, s1, st); break;
                case B200_TYPE_Q4_0: rc = launch_rg<FmtQ40>(p1, t.rg, 1, t.warps, grid, s1, st); break;
                default: rc = launch_rg<FmtQ80>(p1, t.rg, 1, t.warps, grid, s1, st); break;
            }
            if (rc) return rc;
            continue;
        }
        switch (wtype) {
            case B200_TYPE_Q4_K: rc = launch_rg<FmtQ4K>(p, t.rg, nc, t.warps, grid, smem_for(t), st); break;
            case B200_TYPE_Q4_0: rc = launch_rg<FmtQ40>(p, t.rg, nc, t.warps, grid, smem_for(t), st); break;
            default: rc = launch_rg<FmtQ80>(p, t.rg, nc, t.warps, grid, smem_for(t), st); break;
        }
        if (rc) return rc;
    }
    return B200_OK;
}

}  // namespace b200
