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
//     kernel in a chunk-permuted layout (actlayout.cuh) so that every 16-byte lane load is bank-conflict free;
//     integer dot products use dp4a (u8 x s8) / dp2a; RG rows share each activation fetch.
//   * weights are read from HBM exactly once with an L2 evict-first policy (they are not re-used within a token).
//   * launched with programmatic dependent launch: the weight prefetch of kernel N+1 overlaps the tail of kernel N;
//     only the activation read waits (griddepcontrol.wait).
//
// Weight layouts in HBM: Q4_K native ggml blocks (144 B, 16-byte aligned).  Q4_0 / Q8_0 per-row SoA
// ("qs[nb][QB] then d[nb]", see quantize.cu repack_window) because 18 / 34-byte AoS blocks cannot be read with
// aligned 16-byte shared-memory loads.
#include "actlayout.cuh"
#include "common.cuh"
#include "gemv_fmt.cuh"
#include "kernels.h"

#include <cstdlib>

namespace b200 {


// ======================================================================================================
// The kernel: every warp is an independent bulk-copy pipeline over its own contiguous row range.
//   MODE 0 (concat): up to 3 matrices that share the activation vector are treated as one tall matrix (QKV in one launch);
//                    every m_i is a multiple of RG so a row group never straddles two matrices.
//   MODE 1 (paired): two matrices of equal shape (gate, up); a row group holds RG/2 gate rows and the SAME RG/2 up rows,
//                    and the epilogue writes silu(gate.x) * (up.x)  (SwiGLU of BaseMLP::forward, src/layers.cpp:2475-2483).
//   MODE 2 / 3 (expert-indexed concat / paired): ggml_mul_mat_id for ONE token (ggml/src/ggml.c:3225-3240, CPU
//                    ggml-cpu.c:1503-1700; caller MultiLinear::forward src/layers.cpp:2145-2151).  mat[0] (and mat[1] in the
//                    paired mode) is a stack of n_expert matrices of m rows; logical row = slot * m + row, slot < n_ids, and slot
//                    s streams the rows of expert ids[s] (read from DEVICE memory: the router's top_k runs on the GPU just before).
//                    Every slot either shares activation column 0 (gate/up: src1 is broadcast over the slots) or owns column s
//                    (down: src1 has one column per slot).  y[slot * ldy + row].
// ======================================================================================================
struct GemvMat {
    const uint8_t * W;
    float * y;
    const float * bias;
    int64_t m;    // rows
    int64_t ldy;  // column stride of y
};
struct GemvParams {
    GemvMat mat[3];
    int nmat;
    const uint8_t * qact;  // global, NC columns back to back
    int64_t k, m_total;    // MODE 0: sum of m_i ; MODE 1: m of one matrix (number of pairs)
    int n;                 // valid columns (<= NC)
    int nunits;            // k / 256
    int ks;                // units per stage (== nunits -> whole rows per stage)
    int stages;
    uint32_t act_col_bytes;
    // expert-indexed modes only
    const int32_t * ids;    // device, n_ids expert indices of this token
    int n_ids, n_expert;
    int act_cols;           // 1 = every slot reads activation column 0, n_ids = slot s reads column s
    int64_t expert_bytes;   // bytes between consecutive experts of a stack
};

template <class F, int RG, int NC, int MODE>
__global__ void __launch_bounds__(256) gemv_q_kernel(const GemvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    pdl_launch_dependents();  // let the next kernel in the stream become resident and prefetch its own weights
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;

    constexpr int UNIT = F::A_UNIT + F::B_UNIT;
    constexpr int HALF = RG / 2;
    constexpr bool PAIRED = (MODE & 1) != 0;
    constexpr bool IDX = MODE >= 2;
    constexpr int GROUP = PAIRED ? HALF : RG;  // logical rows advanced per row group
    static_assert(!PAIRED || RG >= 2, "paired mode needs RG >= 2");
    static_assert(!IDX || NC == 1, "expert-indexed modes process one token");
    const bool whole = (p.ks == p.nunits);
    const int stage_cols = IDX ? p.act_cols : NC;  // activation columns kept in shared memory
    const uint32_t act_bytes = (uint32_t) al16((int64_t) p.act_col_bytes * stage_cols);
    const uint32_t stage_bytes = (uint32_t) RG * p.ks * UNIT;
    // row r of a stage: A part at st + r*rsA, B part at st + offB + r*rsB
    const uint32_t rsA = whole ? p.ks * UNIT : p.ks * F::A_UNIT;
    const uint32_t rsB = whole ? p.ks * UNIT : p.ks * F::B_UNIT;
    const uint32_t offB = whole ? p.ks * F::A_UNIT : RG * p.ks * F::A_UNIT;

    uint8_t * act_s = smem;
    uint64_t * bars = reinterpret_cast<uint64_t *>(smem + act_bytes) + (size_t) warp * p.stages;
    uint64_t * act_bar = reinterpret_cast<uint64_t *>(smem + act_bytes) + (size_t) nwarps * p.stages;  // completion of the activation copy
    uint8_t * ring = smem + act_bytes + al16((int64_t) (nwarps * p.stages + 1) * 8) + (size_t) warp * p.stages * stage_bytes;

    // ---- this warp's logical rows (aligned to GROUP so groups never straddle a matrix) and its work items
    const int64_t gw = (int64_t) blockIdx.x * nwarps + warp;
    const int64_t GW = (int64_t) gridDim.x * nwarps;
    const int64_t ngrp_total = (p.m_total + GROUP - 1) / GROUP;
    const int64_t r0 = (ngrp_total * gw / GW) * GROUP, r1 = min(p.m_total, (ngrp_total * (gw + 1) / GW) * GROUP);
    const int nseg = (p.nunits + p.ks - 1) / p.ks;
    const int ngroups = (int) ((r1 - r0 + GROUP - 1) / GROUP);
    const int nitems = ngroups * nseg;
    const int64_t row_bytes = (int64_t) p.nunits * UNIT;

    // stage row i of the group starting at logical row `row0` -> global pointer of that weight row
    auto row_src = [&](int64_t row0, int i) -> const uint8_t * {
        if constexpr (IDX) {
            // a group never straddles two slots (m % GROUP == 0, checked on the host)
            const int64_t slot = row0 / p.mat[0].m;
            const int64_t lr = row0 - slot * p.mat[0].m;
            const int e = min(max(p.ids[slot], 0), p.n_expert - 1);  // plain load: produced by the predecessor (see common.cuh, PDL rule)
            const int64_t eoff = (int64_t) e * p.expert_bytes;
            if (PAIRED) return (i < HALF) ? p.mat[0].W + eoff + (lr + i) * row_bytes : p.mat[1].W + eoff + (lr + i - HALF) * row_bytes;
            return p.mat[0].W + eoff + (lr + i) * row_bytes;
        } else if (PAIRED) {
            return (i < HALF) ? p.mat[0].W + (row0 + i) * row_bytes : p.mat[1].W + (row0 + i - HALF) * row_bytes;
        } else {
            int64_t r = row0 + i;
            if (p.nmat > 1 && r >= p.mat[0].m) {
                r -= p.mat[0].m;
                if (p.nmat > 2 && r >= p.mat[1].m) return p.mat[2].W + (r - p.mat[1].m) * row_bytes;
                return p.mat[1].W + r * row_bytes;
            }
            return p.mat[0].W + r * row_bytes;
        }
    };

    if (lane == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&bars[s], 1);
        if (warp == 0) mbar_init(act_bar, 1);
        fence_mbar_init();
    }
    __syncwarp();

    uint64_t pol = 0;
    if (lane == 0) pol = make_evict_first_policy();

    auto issue = [&](int it) {  // lane 0 only
        const int s = it % p.stages;
        const int grp = it / nseg, seg = it - grp * nseg;
        const int64_t row0 = r0 + (int64_t) grp * GROUP;
        const int nlog = (int) min((int64_t) GROUP, r1 - row0);  // valid logical rows (pairs in MODE 1)
        uint8_t * st = ring + (size_t) s * stage_bytes;
        if (whole) {
            if (PAIRED) {
                const uint32_t bytes = (uint32_t) (nlog * row_bytes);
                mbar_arrive_expect_tx(&bars[s], 2 * bytes);
                bulk_g2s_hint(st, row_src(row0, 0), bytes, &bars[s], pol);
                bulk_g2s_hint(st + (size_t) HALF * rsA, row_src(row0, HALF), bytes, &bars[s], pol);
            } else {
                const uint32_t bytes = (uint32_t) (nlog * row_bytes);
                mbar_arrive_expect_tx(&bars[s], bytes);
                bulk_g2s_hint(st, row_src(row0, 0), bytes, &bars[s], pol);
            }
        } else {
            const int u0 = seg * p.ks;
            const int nu = min(p.ks, p.nunits - u0);
            const int nrows = PAIRED ? 2 * nlog : nlog;
            mbar_arrive_expect_tx(&bars[s], (uint32_t) nrows * nu * UNIT);
            for (int r = 0; r < RG; ++r) {
                const bool valid = PAIRED ? ((r < HALF ? r : r - HALF) < nlog) : (r < nlog);
                if (!valid) continue;
                const uint8_t * grow = row_src(row0, r);
                bulk_g2s_hint(st + (size_t) r * rsA, grow + (size_t) u0 * F::A_UNIT, (uint32_t) nu * F::A_UNIT, &bars[s], pol);
                if (F::B_UNIT)
                    bulk_g2s_hint(st + offB + (size_t) r * rsB, grow + (size_t) p.nunits * F::A_UNIT + (size_t) u0 * F::B_UNIT,
                                  (uint32_t) nu * F::B_UNIT, &bars[s], pol);
            }
        }
    };

    // weights do not depend on the producer kernel: start streaming them before the PDL dependency resolves
    // (expert-indexed modes: WHICH weights are streamed is the producer's output, so they wait first)
    if constexpr (IDX) pdl_wait();
    if (lane == 0) {
        const int pre = min(p.stages, nitems);
        for (int it = 0; it < pre; ++it) issue(it);
    }

    // ---- activations: wait for the producer kernel (PDL), then stage the quantized columns in smem with ONE bulk copy (a per-thread
    // copy loop costs one L2 round trip per 4 KB: 4 serial trips for k = 14336)
    if constexpr (!IDX) pdl_wait();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t) (((size_t) p.act_col_bytes * stage_cols) & ~(size_t) 15);
        asm volatile("fence.proxy.async.global;" ::: "memory");
        mbar_arrive_expect_tx(act_bar, bytes);
        bulk_g2s(act_s, p.qact, bytes, act_bar);
    }
    __syncthreads();  // act_bar's initialisation is visible to every warp
    mbar_wait(act_bar, 0);

    const ActLayout L = act_layout(F::Q8K, p.k);
    constexpr int LPU = F::LPU;
    constexpr int UPS = 32 / LPU;  // units per warp step
    const int g = lane % LPU;
    const int ul = lane / LPU;

    float acc[RG][NC];
    for (int it = 0; it < nitems; ++it) {
        const int s = it % p.stages;
        const uint32_t parity = (uint32_t) (it / p.stages) & 1u;
        const int grp = it / nseg, seg = it - grp * nseg;
        const int64_t row0 = r0 + (int64_t) grp * GROUP;
        const int nlog = (int) min((int64_t) GROUP, r1 - row0);
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
        const uint8_t * acol = act_s;
        if constexpr (IDX) { if (p.act_cols > 1) acol += (size_t) (row0 / p.mat[0].m) * p.act_col_bytes; }
        if (nlog == GROUP) {
            // full group (all but the last group of a warp's range): no per-row branches, so the RG rows' shared-memory loads and
            // dot-product chains interleave — the per-stage latency of a warp, not the issue rate, is what short GEMVs wait for
            for (int u = ul; u < nu; u += UPS) {
                typename F::Wt Wr[RG];
#pragma unroll
                for (int r = 0; r < RG; ++r) F::load_w(st + (size_t) r * rsA, st + offB + (size_t) r * rsB, u, g, Wr[r]);
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    typename F::Act A;
                    F::load_act(acol + (size_t) c * p.act_col_bytes, L, u0 + u, g, A);
#pragma unroll
                    for (int r = 0; r < RG; ++r) acc[r][c] = F::dot(Wr[r], A, acc[r][c]);
                }
            }
        } else {
            for (int u = ul; u < nu; u += UPS) {
#pragma unroll
                for (int r = 0; r < RG; ++r) {
                    const bool valid = PAIRED ? ((r < HALF ? r : r - HALF) < nlog) : (r < nlog);
                    if (valid) {
                        typename F::Wt Wr;
                        F::load_w(st + (size_t) r * rsA, st + offB + (size_t) r * rsB, u, g, Wr);
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            typename F::Act A;
                            F::load_act(acol + (size_t) c * p.act_col_bytes, L, u0 + u, g, A);
                            acc[r][c] = F::dot(Wr, A, acc[r][c]);
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0 && it + p.stages < nitems) issue(it + p.stages);

        if (seg == nseg - 1) {
            // full butterfly: every lane ends up with every sum
#pragma unroll
            for (int r = 0; r < RG; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c) acc[r][c] = warp_sum(acc[r][c]);
            if (lane == 0) {
                if constexpr (IDX) {
                    const int64_t slot = row0 / p.mat[0].m;
                    float * yo = p.mat[0].y + slot * p.mat[0].ldy + (row0 - slot * p.mat[0].m);
                    if (PAIRED) {
#pragma unroll
                        for (int q = 0; q < HALF; ++q) {
                            if (q < nlog) {
                                const float gv = acc[q][0], uv = acc[HALF + q][0];
                                yo[q] = (gv / (1.0f + expf(-gv))) * uv;
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < RG; ++r) if (r < nlog) yo[r] = acc[r][0];
                    }
                } else if (PAIRED) {
#pragma unroll
                    for (int q = 0; q < HALF; ++q) {
                        if (q < nlog) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                if (c < p.n) {
                                    const float gv = acc[q][c], uv = acc[HALF + q][c];
                                    p.mat[0].y[(int64_t) c * p.mat[0].ldy + row0 + q] = (gv / (1.0f + expf(-gv))) * uv;
                                }
                            }
                        }
                    }
                } else {
                    // the whole group lies in one matrix
                    int mi = 0;
                    int64_t lrow = row0;
                    if (p.nmat > 1 && lrow >= p.mat[0].m) { lrow -= p.mat[0].m; mi = 1; if (p.nmat > 2 && lrow >= p.mat[1].m) { lrow -= p.mat[1].m; mi = 2; } }
                    const GemvMat & M = p.mat[mi];
#pragma unroll
                    for (int r = 0; r < RG; ++r) {
                        if (r < nlog) {
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                if (c < p.n) {
                                    float o = acc[r][c];
                                    if (M.bias) o += M.bias[lrow + r];
                                    M.y[(int64_t) c * M.ldy + lrow + r] = o;
                                }
                            }
                        }
                    }
                }
            }
        }
    }

}

// ======================================================================================================
// host side
// ======================================================================================================
int sm_count() {  // of the CURRENT device (cached per device ordinal)
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int & n = sms[dev & 63];
    if (!n) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static int env_int(const char * name, int dflt) {
    const char * v = getenv(name);
    return v ? atoi(v) : dflt;
}
// pipeline shape defaults: the B200_GEMV_* environment is read ONCE (the plugin passes tune == NULL ~225 times per decoded token)
static const GemvTuning & env_tuning() {
    static const GemvTuning t = {env_int("B200_GEMV_KS", 16), env_int("B200_GEMV_STAGES", 2), env_int("B200_GEMV_WARPS", 8), env_int("B200_GEMV_RG", 4),
                                 env_int("B200_GEMV_GRID", 0)};
    return t;
}

template <class F, int RG, int NC, int MODE>
static int launch(const GemvParams & p, int warps, int grid, size_t smem_bytes, cudaStream_t st) {
    auto kern = gemv_q_kernel<F, RG, NC, MODE>;
    static size_t configured[64] = {0};  // per template instantiation AND per device (function attributes are per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return B200_ERR_UNSUPPORTED;
    if (smem_bytes > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_bytes);
        if (e != cudaSuccess) return (int) e;
        configured[dev] = smem_bytes;
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
    static const int no_pdl = env_int("B200_NO_PDL", 0);
    cfg.numAttrs = no_pdl ? 0 : 1;
    return (int) cudaLaunchKernelEx(&cfg, kern, p);
}

template <class F, int MODE>
static int launch_rg_nc(const GemvParams & p, int rg, int nc, int warps, int grid, size_t smem, cudaStream_t st) {
#define B200_CASE(RG_, NC_) if (rg == RG_ && nc == NC_) return launch<F, RG_, NC_, MODE>(p, warps, grid, smem, st);
    if constexpr (MODE >= 2) {  // expert-indexed: one token
        if constexpr (MODE == 2) { B200_CASE(1, 1) }
        B200_CASE(2, 1) B200_CASE(4, 1)
    } else {
        if constexpr (MODE == 0) { B200_CASE(1, 1) B200_CASE(1, 2) B200_CASE(1, 4) }
        B200_CASE(2, 1) B200_CASE(2, 2) B200_CASE(2, 4) B200_CASE(4, 1) B200_CASE(4, 2) B200_CASE(4, 4)
    }
#undef B200_CASE
    return B200_ERR_ARG;
}

template <class F>
static int launch_mode(int mode, const GemvParams & p, int rg, int nc, int warps, int grid, size_t smem, cudaStream_t st) {
    switch (mode) {
        case 0: return launch_rg_nc<F, 0>(p, rg, nc, warps, grid, smem, st);
        case 1: return launch_rg_nc<F, 1>(p, rg, nc, warps, grid, smem, st);
        case 2: return launch_rg_nc<F, 2>(p, rg, nc, warps, grid, smem, st);
        case 3: return launch_rg_nc<F, 3>(p, rg, nc, warps, grid, smem, st);
        default: return B200_ERR_ARG;
    }
}
static int launch_fmt(int wtype, int mode, const GemvParams & p, int rg, int nc, int warps, int grid, size_t smem, cudaStream_t st) {
    switch (wtype) {
        case B200_TYPE_Q4_K: return launch_mode<FmtQ4K>(mode, p, rg, nc, warps, grid, smem, st);
        case B200_TYPE_Q4_0: return launch_mode<FmtQ40>(mode, p, rg, nc, warps, grid, smem, st);
        case B200_TYPE_Q8_0: return launch_mode<FmtQ80>(mode, p, rg, nc, warps, grid, smem, st);
        default: return B200_ERR_UNSUPPORTED;
    }
}

static int unit_bytes(int wtype) {
    switch (wtype) {
        case B200_TYPE_Q4_K: return FmtQ4K::A_UNIT + FmtQ4K::B_UNIT;
        case B200_TYPE_Q4_0: return FmtQ40::A_UNIT + FmtQ40::B_UNIT;
        case B200_TYPE_Q8_0: return FmtQ80::A_UNIT + FmtQ80::B_UNIT;
        default: return 0;
    }
}

// mode 0: y_i[c*ldy_i + r] = sum_k W_i[r,k] x_c[k] (+ bias_i[r]) for up to 3 matrices sharing x  (QKV)
// mode 1: y_0[c*ldy_0 + r] = silu(W_0[r,:].x_c) * (W_1[r,:].x_c)                                   (gate/up)
int mul_mat_q_multi(int wtype, int mode, int nmat, const void * const * W, const int64_t * m, float * const * y, const int64_t * ldy,
                    const float * const * bias, int64_t k, const void * qact, int64_t n, const GemvTuning * tune, cudaStream_t st) {
    if (k <= 0 || n <= 0 || nmat <= 0) return B200_OK;
    int unit;
    switch (wtype) {
        case B200_TYPE_Q4_K: unit = FmtQ4K::A_UNIT + FmtQ4K::B_UNIT; break;
        case B200_TYPE_Q4_0: unit = FmtQ40::A_UNIT + FmtQ40::B_UNIT; break;
        case B200_TYPE_Q8_0: unit = FmtQ80::A_UNIT + FmtQ80::B_UNIT; break;
        default: return B200_ERR_UNSUPPORTED;
    }
    const size_t acb = qact_col_bytes(wtype, k);
    const int nunits = (int) (k / 256);
    const int sms = sm_count();
    int64_t m_total = 0, m_phys = 0;
    for (int i = 0; i < nmat; ++i) m_phys += m[i];
    m_total = mode == 1 ? m[0] : m_phys;
    if (m_total <= 0) return B200_OK;

    int64_t c0 = 0;
    while (c0 < n) {
        const int64_t left = n - c0;
        const int nc = left >= 4 ? 4 : (left >= 2 ? 2 : 1);

        GemvTuning t;
        t.rg = tune && tune->rg ? tune->rg : env_tuning().rg;
        t.warps = tune && tune->warps ? tune->warps : env_tuning().warps;
    if (t.warps > 8) t.warps = 8;
        if (t.warps > 8) t.warps = 8;  // __launch_bounds__(256): the full-group path keeps RG rows of weights in registers
        t.stages = tune && tune->stages ? tune->stages : env_tuning().stages;
        t.ks = tune && tune->ks ? tune->ks : env_tuning().ks;
        t.grid = tune && tune->grid ? tune->grid : env_tuning().grid;
        if (t.ks > nunits) t.ks = nunits;
        // few rows: keep at least one row-group per warp
        while (t.rg > 1 && m_phys < (int64_t) sms * t.warps * t.rg / 2) t.rg >>= 1;
        if (mode == 1 && t.rg < 2) t.rg = 2;
        if (mode == 0) {  // groups must not straddle matrices
            for (int i = 0; i + 1 < nmat; ++i) while (t.rg > 1 && m[i] % t.rg) t.rg >>= 1;
        }
        if (nc > 1 && t.rg > 2) t.rg = 2;  // register budget of the multi-column variants
        const int group = mode == 1 ? t.rg / 2 : t.rg;
        int grid = t.grid > 0 ? t.grid : sms;
        {
            const int64_t max_warps = (m_total + group - 1) / group;
            const int64_t max_grid = (max_warps + t.warps - 1) / t.warps;
            if (grid > max_grid) grid = (int) max_grid;
        }
        const size_t act_bytes = (size_t) al16((int64_t) acb * nc);
        auto smem_for = [&](const GemvTuning & q) {
            return act_bytes + (size_t) al16((int64_t) (q.warps * q.stages + 1) * 8) + (size_t) q.warps * q.stages * q.rg * q.ks * unit;
        };
        const size_t limit = 227 * 1024;
        while (smem_for(t) > limit && t.stages > 3) t.stages--;
        while (smem_for(t) > limit && t.ks > 4) t.ks = (t.ks + 1) / 2;
        while (smem_for(t) > limit && t.stages > 2) t.stages--;
        while (smem_for(t) > limit && t.warps > 1) t.warps >>= 1;
        while (smem_for(t) > limit && t.ks > 1) t.ks = (t.ks + 1) / 2;
        if (smem_for(t) > limit) return B200_ERR_UNSUPPORTED;

        GemvParams p;
        p.nmat = nmat;
        for (int i = 0; i < 3; ++i) {
            if (i < nmat) {
                p.mat[i].W = (const uint8_t *) W[i];
                p.mat[i].y = y[i] ? y[i] + c0 * ldy[i] : nullptr;
                p.mat[i].bias = bias ? bias[i] : nullptr;
                p.mat[i].m = m[i];
                p.mat[i].ldy = ldy[i];
            } else {
                p.mat[i] = GemvMat{nullptr, nullptr, nullptr, 0, 0};
            }
        }
        p.qact = (const uint8_t *) qact + (size_t) c0 * acb;
        p.k = k; p.m_total = m_total;
        p.n = nc;
        p.nunits = nunits;
        p.ks = t.ks;
        p.stages = t.stages;
        p.act_col_bytes = (uint32_t) acb;
        p.ids = nullptr; p.n_ids = 0; p.n_expert = 0; p.act_cols = 0; p.expert_bytes = 0;
        const int rc = launch_fmt(wtype, mode, p, t.rg, nc, t.warps, grid, smem_for(t), st);
        if (rc) return rc;
        c0 += nc;
    }
    return B200_OK;
}

// ggml_mul_mat_id for one token (see MODE 2 / 3 above):  paired = 0:  y[s*ldy + r] = W0[ids[s]][r,:] . x_{c(s)}
//                                                         paired = 1:  y[s*ldy + r] = silu(W0[ids[s]][r,:] . x) * (W1[ids[s]][r,:] . x)
// c(s) = s if act_cols == n_ids, 0 if act_cols == 1.  W0 / W1: stacks of n_expert matrices [m, k] in the device layout.
int mul_mat_q_id(int wtype, int paired, const void * W0, const void * W1, int64_t k, int64_t m, int n_expert, const int32_t * ids, int n_ids,
                 const void * qact, int act_cols, float * y, int64_t ldy, const GemvTuning * tune, cudaStream_t st) {
    if (k <= 0 || m <= 0 || n_ids <= 0) return B200_OK;
    const int unit = unit_bytes(wtype);
    if (!unit || k % 256) return B200_ERR_UNSUPPORTED;
    if (!ids || n_expert <= 0 || (act_cols != 1 && act_cols != n_ids) || (paired && !W1)) return B200_ERR_ARG;
    const size_t acb = qact_col_bytes(wtype, k);
    const int nunits = (int) (k / 256);
    const int sms = sm_count();
    const int64_t m_total = (int64_t) n_ids * m, m_phys = m_total * (paired ? 2 : 1);

    GemvTuning t;
    t.rg = tune && tune->rg ? tune->rg : env_tuning().rg;
    t.warps = tune && tune->warps ? tune->warps : env_tuning().warps;
    t.stages = tune && tune->stages ? tune->stages : env_tuning().stages;
    t.ks = tune && tune->ks ? tune->ks : env_tuning().ks;
    t.grid = tune && tune->grid ? tune->grid : env_tuning().grid;
    if (t.ks > nunits) t.ks = nunits;
    while (t.rg > 1 && m_phys < (int64_t) sms * t.warps * t.rg / 2) t.rg >>= 1;
    if (paired && t.rg < 2) t.rg = 2;
    // a row group must not straddle two slots
    while (t.rg > (paired ? 2 : 1) && m % (paired ? t.rg / 2 : t.rg)) t.rg >>= 1;
    const int group = paired ? t.rg / 2 : t.rg;
    if (m % group) return B200_ERR_UNSUPPORTED;
    int grid = t.grid > 0 ? t.grid : sms;
    {
        const int64_t max_warps = (m_total + group - 1) / group;
        const int64_t max_grid = (max_warps + t.warps - 1) / t.warps;
        if (grid > max_grid) grid = (int) max_grid;
    }
    const size_t act_bytes = (size_t) al16((int64_t) acb * act_cols);
    auto smem_for = [&](const GemvTuning & q) {
        return act_bytes + (size_t) al16((int64_t) (q.warps * q.stages + 1) * 8) + (size_t) q.warps * q.stages * q.rg * q.ks * unit;
    };
    const size_t limit = 227 * 1024;
    while (smem_for(t) > limit && t.stages > 3) t.stages--;
    while (smem_for(t) > limit && t.ks > 4) t.ks = (t.ks + 1) / 2;
    while (smem_for(t) > limit && t.stages > 2) t.stages--;
    while (smem_for(t) > limit && t.warps > 1) t.warps >>= 1;
    while (smem_for(t) > limit && t.ks > 1) t.ks = (t.ks + 1) / 2;
    if (smem_for(t) > limit) return B200_ERR_UNSUPPORTED;

    GemvParams p;
    p.nmat = paired ? 2 : 1;
    p.mat[0] = GemvMat{(const uint8_t *) W0, y, nullptr, m, ldy};
    p.mat[1] = paired ? GemvMat{(const uint8_t *) W1, nullptr, nullptr, m, ldy} : GemvMat{nullptr, nullptr, nullptr, 0, 0};
    p.mat[2] = GemvMat{nullptr, nullptr, nullptr, 0, 0};
    p.qact = (const uint8_t *) qact;
    p.k = k; p.m_total = m_total;
    p.n = 1;
    p.nunits = nunits;
    p.ks = t.ks;
    p.stages = t.stages;
    p.act_col_bytes = (uint32_t) acb;
    p.ids = ids; p.n_ids = n_ids; p.n_expert = n_expert; p.act_cols = act_cols;
    p.expert_bytes = m * (int64_t) nunits * unit;
    return launch_fmt(wtype, paired ? 3 : 2, p, t.rg, 1, t.warps, grid, smem_for(t), st);
}

int mul_mat_q(int wtype, const void * W, int64_t k, int64_t m, const void * qact, int64_t n, float * y, int64_t ldy, const float * bias,
              const GemvTuning * tune, cudaStream_t st) {
    const void * Ws[1] = {W};
    float * ys[1] = {y};
    const float * bs[1] = {bias};
    return mul_mat_q_multi(wtype, 0, 1, Ws, &m, ys, &ldy, bs, k, qact, n, tune, st);
}

}  // namespace b200
