// decode_mk.cu — the whole single-token decode step of a dense Llama-family model as ONE persistent kernel on sm_100a.
//
// Why (round-1 profile, profiles/r01f_*): the decode step was 355 launches; every quantized GEMV of a layer is a 1.4 - 10 us stream that
// pays ~2 us of launch ramp / pipeline fill, so the in-sequence GEMVs ran at 0.48 of the HBM copy peak while the same kernel reaches
// 0.95 - 0.98 on the one long matrix (lm_head).  A kernel boundary drains every SM's shared-memory ring; nothing in the dependency chain
// of a token needs that.  The WEIGHTS of GEMV i+1 do not depend on the activations of GEMV i.
//
// Design: one CTA per SM (cooperative launch, 8 warps), alive for the whole token.
//   * Every warp is a private bulk-copy (TMA, UBLKCP + mbarrier) pipeline over ITS rows of EVERY quantized matmul of the token, in
//     execution order: q/k/v (one concatenated logical matrix), o, gate+up (paired rows, SwiGLU epilogue), down, ..., lm_head.  The
//     issue cursor runs ahead of the consume cursor across phase boundaries, so while the grid is waiting at a barrier, quantizing
//     activations or doing attention, the rings keep ~20 MB of weight requests in flight and HBM never idles.
//   * Steps are separated by a grid barrier (one 64-bit counter in L2, release/acquire; monotonic across launches so no reset is needed).
//   * RMSNorm + activation quantization are recomputed by every CTA in the prologue of the consuming GEMV straight into shared memory
//     (16 - 57 KB from L2); residual adds and SwiGLU live in GEMV epilogues.
//   * Decode attention runs inside as three steps (tensor-core scores per 16-position tile with the RoPE of q / the new K row folded in,
//     position-split V.P with the new V column folded in, partial-sum + quantization of the o-projection's activations), with the
//     reference's arithmetic: f16 operands, fp32 accumulation, GLOBAL-max softmax, P rounded through f16 before V.P
//     (src/layers.cpp:2541-2561, ggml-cpu.c:213-219).
// Same integer dot products and fp32 expressions as gemv.cu / fused.cu (format traits shared through gemv_fmt.cuh), so the per-op
// parity properties carry over; replaces for one token what HeterogeneousModel::forward (src/models.cpp:1399-1424) ->
// LMBlock1Forward::forward (src/layers.cpp:2719-2761) -> LMFinalSteps::forward (src/models.cpp:1736-1785) build as a ~1000-node graph.
#include "actlayout.cuh"
#include "common.cuh"
#include "dequant.cuh"
#include "gemv_fmt.cuh"
#include "kernels.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace b200 {

constexpr int MK_WARPS = 8;
constexpr int MK_THREADS = MK_WARPS * 32;
constexpr int MK_RG = 4;      // rows per row group (2 gate + 2 up rows in the paired phases)
constexpr int MK_PV_ROUND = 256;
constexpr int MK_PV_SP = MK_PV_ROUND + 32;

enum : int { MK_PRO_COPY = 0, MK_PRO_NORM = 1, MK_PRO_QUANT = 2 };
enum : int { MK_EPI_STORE = 0, MK_EPI_RESID = 1, MK_EPI_SWIGLU = 2, MK_EPI_STORE_ARGMAX = 3 };
// step kinds inside a layer (7 per layer): 0 qkv, 1 scores, 2 pv, 3 tail, 4 o, 5 gate/up, 6 down
enum : int { MK_FLAG_ADVANCE = 1, MK_FLAG_ADVANCE_POS = 2 };

struct MKPhase {  // one quantized-matmul phase; table built on the host, read-only on the device
    const uint8_t * W[3];
    float * y[3];
    const float * bias[3];
    int m[3];
    int nmat, paired, k, m_total;  // m_total: logical rows (pairs in a paired phase)
    int pro, epi;
    const float * pro_x;
    const float * pro_w;
    const uint8_t * pro_q;
};

struct MKLayerKV {
    __half * kc;
    __half * vc;
};

struct MKParams {
    const MKPhase * phases;
    const MKLayerKV * kv;
    int n_phases, n_layers;
    int hidden, heads, kv_heads, head_dim, ffn, vocab;
    int rope_mode;
    float theta_scale, eps, attn_scale;
    const float * rope_ff;
    int64_t k_row_stride, v_row_stride;
    const int32_t * tok;
    const int32_t * pos;
    int n_kv_arg;   // < 0: pos[0] + 1
    int v_col_arg;  // < 0: pos[0]
    const uint8_t * embed;  // NULL: x holds the incoming hidden state
    int embed_type;
    float * x;
    int32_t * next_tok;
    int has_head, flags;
    // layer-sharded multi-GPU hand-off over NVLink peer memory (all NULL on one GPU)
    const unsigned long long * wait_flag;  // LOCAL flag a peer raises when this launch's input (hidden row / token) has arrived
    int wait_offset;                       // proceed when *wait_flag >= launches completed so far + wait_offset
    float * send_x;                        // PEER: the next shard's residual-stream buffer
    unsigned long long * send_flag;        // PEER: its wait_flag
    int32_t * send_tok;                    // PEER (last shard): the first shard's token buffer
    // workspace
    unsigned long long * sync;  // [0] barrier counter, [1] base of the next launch, [2] argmax key, [3] status
    float * q, * kbuf, * vbuf, * scores, * partial, * gate;
    float2 * part;
    uint8_t * att_q;
    int64_t s_stride;
    int max_tiles;
    int step_begin, step_end;
    long long * times;  // profiling aid (B200_MK_TIMES=1): SM clock of CTA 0 at kernel start and after every step's barrier
    int ks, stages;
    uint32_t act_bytes, stage_bytes;
    int kv_prefetch;     // pull the layer's K / V rows into L2 during its q/k/v projection
    int l2_ahead;        // items per warp pulled into L2 ahead of the shared-memory ring (0 = off)
    uint32_t tab_bytes;  // > 0: both tables live at the start of dynamic shared memory
    int merged_tail;     // the last V.P CTA of a KV group sums + quantizes the group's slice (no separate tail step)
};

// the phase table / KV-cache table as the device code sees them: copied into shared memory at kernel start when they fit (a first-touch
// miss on a table entry costs a DRAM round trip behind ~20 MB of queued weight requests — it used to stall every warp at every phase change)
// block-reduction scratch (double-precision sum of squares of the RMSNorm prologue)
struct NormQuantSmem {
    float red[32];
};

struct MKTabs {
    const MKPhase * ph;
    const MKLayerKV * kv;
};

// ---- loads of data produced by OTHER CTAs of this launch: L2 only (never a stale L1 line) ----------------------------------------
__device__ __forceinline__ float ldcg_f(const float * p) { return __ldcg(p); }
__device__ __forceinline__ float4 ldcg_f4(const float * p) { return __ldcg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ uint4 ldcg_u4(const void * p) { return __ldcg(reinterpret_cast<const uint4 *>(p)); }

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long * p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Grid barrier: every CTA of the (co-resident, cooperative) grid arrives on one monotonic counter.  A watchdog turns a lost arrival into
// status != 0 instead of a hung GPU; once status is set every later barrier falls through.
__device__ __forceinline__ void grid_barrier(const MKParams & p, unsigned long long & target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        // release-arrive: orders every write this CTA made before the bar.sync above (cumulativity) ahead of the increment; measured on the
        // B200 (tools/mk_microbench.cu) 0.4 us cheaper than __threadfence() + atomicAdd
        asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p.sync) : "memory");
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_u64(p.sync) < target) {
            if ((++spins & 1023u) == 0) {
                if (__ldcg(p.sync + 3) != 0ull) break;
                if (globaltimer_ns() - t0 > 2000000000ull) { atomicExch(p.sync + 3, 1ull); break; }
            }
        }
    }
    __syncthreads();  // data of other CTAs is read with ld.global.cg (L2) after this point: no stale L1 line can be hit
}

// block sum in double (blockDim.x == 256); red: 8 doubles of shared memory
__device__ __forceinline__ double block_sum_f64(double v, double * red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < MK_WARPS; ++w) t += red[w];
    __syncthreads();
    return t;
}

// ---- activation quantization into a column (shared or global memory), one WARP per 256-element block ----------------------------------
// Same expressions as quantize.cu / fused.cu add_rmsnorm_quant_kernel (reference quantize_row_q8_K_ref ggml-quants.c:2555-2592, x86 quantize_row_q8_0
// arch/x86/quants.c:290-384): bit-identical codes.  Lane l of the warp owns elements [4l, 4l+4) and [128+4l, 128+4l+4) of its block, so
// the block maximum / the 32-element sums are warp shuffles: NO block-wide barrier (the round-1 prologue's 3 barriers per 1024 elements
// made it cost ~6 us per GEMV, gpurun_out/r02_bringup).
template <bool Q8K>
__device__ __forceinline__ void quantize_block_warp(uint8_t * base, const ActLayout & L, int64_t blk, const float4 & a, const float4 & b) {
    const int lane = threadIdx.x & 31;
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    const int64_t e0 = blk * 256 + 4 * lane, e1 = e0 + 128;
    if (Q8K) {
        // first-occurrence argmax |v| over the block: key = (|v| bits, 255 - index)
        unsigned long long key = 0ull;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = (i < 4) ? 4 * lane + i : 128 + 4 * lane + (i - 4);
            const unsigned long long kk = ((unsigned long long) __float_as_uint(fabsf(v[i])) << 32) | (unsigned) (255 - idx);
            key = kk > key ? kk : key;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        const int idx = 255 - (int) (key & 0xffffffffu);
        const int sel = ((idx >> 7) << 2) | (idx & 3);  // which of this lane's 8 values, if this lane owns idx
        float cand = v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) cand = (sel == i) ? v[i] : cand;
        const float mx = __shfl_sync(0xffffffffu, cand, (idx & 127) >> 2);
        int q[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float dv = 0.0f;
        if (mx != 0.0f) {
            const float iscale = __fdiv_rn(-127.f, mx);
#pragma unroll
            for (int i = 0; i < 8; ++i) q[i] = min(127, __float2int_rn(__fmul_rn(iscale, v[i])));
            dv = __fdiv_rn(1.0f, iscale);
        }
        *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e0)) = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
        *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e1)) = (uint32_t) (q[4] & 0xff) | ((uint32_t) (q[5] & 0xff) << 8) | ((uint32_t) (q[6] & 0xff) << 16) | ((uint32_t) (q[7] & 0xff) << 24);
        int s0 = q[0] + q[1] + q[2] + q[3], s1 = q[4] + q[5] + q[6] + q[7];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        int16_t * bs = (int16_t *) (base + L.bs_off);
        if ((lane & 7) == 0) { bs[e0 >> 5] = (int16_t) s0; bs[e1 >> 5] = (int16_t) s1; }
        if (lane == 0) ((float *) (base + L.d_off))[blk] = dv;
    } else {
        float * dd = (float *) (base + L.d_off);
        int * bs = (int *) (base + L.bs_off);
#pragma unroll
        for (int h = 0; h < 2; ++h) {  // two independent 32-element blocks per 8-lane group
            const float * w = v + 4 * h;
            const int64_t e = h ? e1 : e0;
            float amax = fmaxf(fmaxf(fabsf(w[0]), fabsf(w[1])), fmaxf(fabsf(w[2]), fabsf(w[3])));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float dv = __fdiv_rn(amax, 127.f);
            const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = __float2int_rn(__fmul_rn(w[i], id));
            int sq = q[0] + q[1] + q[2] + q[3];
            sq += __shfl_xor_sync(0xffffffffu, sq, 1);
            sq += __shfl_xor_sync(0xffffffffu, sq, 2);
            sq += __shfl_xor_sync(0xffffffffu, sq, 4);
            *reinterpret_cast<uint32_t *>(base + act_qs_off_q80(e)) = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
            if ((lane & 7) == 0) { dd[e >> 5] = __half2float(__float2half_rn(dv)); bs[e >> 5] = sq; }
        }
    }
}

// blocks [blk_begin, blk_end) of a k-element column spread over the CTA's warps; val(e) returns 4 consecutive elements.  NP blocks per warp
// are loaded before any of them is quantized (one L2 round trip per NP blocks).
template <bool Q8K, int NP, class ValFn>
__device__ __forceinline__ void quantize_blocks(uint8_t * base, int64_t k, int blk_begin, int blk_end, ValFn val) {
    const ActLayout L = act_layout(Q8K, k);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b0 = blk_begin + warp; b0 < blk_end; b0 += MK_WARPS * NP) {
        float4 pa[NP], pb[NP];
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            const int blk = b0 + MK_WARPS * r;
            if (blk < blk_end) { pa[r] = val((int64_t) blk * 256 + 4 * lane); pb[r] = val((int64_t) blk * 256 + 128 + 4 * lane); }
        }
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            const int blk = b0 + MK_WARPS * r;
            if (blk < blk_end) quantize_block_warp<Q8K>(base, L, blk, pa[r], pb[r]);
        }
    }
}

// ---- the per-warp weight pipeline -----------------------------------------------------------------------------------------------
struct MKCursor {
    int ph;         // current phase (== ph_end: exhausted)
    int grp, seg;   // next item: row group of this warp, k-segment of the group (advanced incrementally: no divisions on the issue path)
    int ngrp;
    int r0, r1;     // this warp's logical rows
    int nseg, nunits, ks;
    int nmat, paired;
    int m0, m1;
    const uint8_t * W0;
    const uint8_t * W1;
    const uint8_t * W2;
};

__device__ __forceinline__ void cursor_seek(MKCursor & c, const MKParams & p, const MKTabs & T, int ph, int ph_end, unsigned gw, unsigned GW) {
    for (; ph < ph_end; ++ph) {
        const MKPhase * P = T.ph + ph;
        const unsigned m_total = (unsigned) P->m_total;   // m_total * GW < 2^32 (checked on the host): 32-bit divisions only
        const int paired = P->paired;
        const int group = paired ? MK_RG / 2 : MK_RG;
        const int r0 = (int) (m_total * gw / GW), r1 = (int) (m_total * (gw + 1) / GW);
        if (r1 <= r0) continue;
        c.nunits = P->k / 256;
        c.ks = min(p.ks, c.nunits);
        c.nseg = (c.nunits + c.ks - 1) / c.ks;
        c.r0 = r0; c.r1 = r1;
        c.ngrp = (r1 - r0 + group - 1) / group;
        c.nmat = P->nmat; c.paired = paired;
        c.m0 = P->m[0]; c.m1 = P->m[1];
        c.W0 = P->W[0]; c.W1 = P->W[1]; c.W2 = P->W[2];
        c.grp = 0; c.seg = 0;
        break;
    }
    c.ph = ph;
}
// next item of a REGISTER cursor
__device__ __forceinline__ void cursor_next(MKCursor & c, const MKParams & p, const MKTabs & T, int ph_end, unsigned gw, unsigned GW) {
    if (++c.seg == c.nseg) { c.seg = 0; ++c.grp; }
    if (c.grp >= c.ngrp) cursor_seek(c, p, T, c.ph + 1, ph_end, gw, GW);
}
// next item of a SHARED-MEMORY cursor (one per warp): every lane computes the same new state from what it read, lane 0 stores it
__device__ __forceinline__ void cursor_next_shared(MKCursor & c, const MKParams & p, const MKTabs & T, int ph_end, unsigned gw, unsigned GW) {
    MKCursor t = c;
    __syncwarp();
    cursor_next(t, p, T, ph_end, gw, GW);
    if ((threadIdx.x & 31) == 0) c = t;
    __syncwarp();
}

// global pointer of stage row i (0 .. RG-1) of the group whose first logical row is row0
template <int UNIT>
__device__ __forceinline__ const uint8_t * cursor_row_src(const MKCursor & c, int row0, int i) {
    const int64_t row_bytes = (int64_t) c.nunits * UNIT;
    if (c.paired) return (i < MK_RG / 2) ? c.W0 + (int64_t) (row0 + i) * row_bytes : c.W1 + (int64_t) (row0 + i - MK_RG / 2) * row_bytes;
    int r = row0 + i;
    if (c.nmat > 1 && r >= c.m0) {
        r -= c.m0;
        if (c.nmat > 2 && r >= c.m1) return c.W2 + (int64_t) (r - c.m1) * row_bytes;
        return c.W1 + (int64_t) r * row_bytes;
    }
    return c.W0 + (int64_t) r * row_bytes;
}

// !PREFETCH (lane 0 only): one bulk copy into the ring, completing on `bar`.  PREFETCH (ALL lanes): the same bytes pulled into L2 only, one
// 128-byte line per lane and instruction (LSU prefetches: a bulk prefetch would queue in the SM's in-order TMA unit ahead of the demand copies)
template <bool PREFETCH>
__device__ __forceinline__ void mk_copy(void * smem_dst, const uint8_t * gmem_src, uint32_t bytes, uint64_t * bar, uint64_t pol) {
    if (PREFETCH) {
        for (uint32_t off = (threadIdx.x & 31) * 128u; off < bytes; off += 32u * 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(gmem_src + off));
    } else {
        bulk_g2s_hint(smem_dst, gmem_src, bytes, bar, pol);
    }
}

// the bulk copies of the cursor's current item into stage `st` (completing on `bar`) — or, PREFETCH, the same byte ranges into L2
template <class F, bool PREFETCH = false>
__device__ __forceinline__ void issue_item(const MKCursor & c, uint8_t * st, uint64_t * bar, uint64_t pol) {
    constexpr int UNIT = F::A_UNIT + F::B_UNIT;
    constexpr int HALF = MK_RG / 2;
    const int group = c.paired ? HALF : MK_RG;
    const int row0 = c.r0 + c.grp * group;
    const int nlog = min(group, c.r1 - row0);
    const int64_t row_bytes = (int64_t) c.nunits * UNIT;
    if (c.nseg == 1) {  // whole rows: row r of the stage at st + r * row_bytes (A part then B part, as in HBM)
        if (c.paired) {
            const uint32_t bytes = (uint32_t) (nlog * row_bytes);
            if (!PREFETCH) mbar_arrive_expect_tx(bar, 2 * bytes);
            mk_copy<PREFETCH>(st, cursor_row_src<UNIT>(c, row0, 0), bytes, bar, pol);
            mk_copy<PREFETCH>(st + (size_t) HALF * row_bytes, cursor_row_src<UNIT>(c, row0, HALF), bytes, bar, pol);
        } else {
            const uint8_t * first = cursor_row_src<UNIT>(c, row0, 0);
            const uint8_t * last = (c.nmat > 1) ? cursor_row_src<UNIT>(c, row0, nlog - 1) : first + (int64_t) (nlog - 1) * row_bytes;
            if (last == first + (int64_t) (nlog - 1) * row_bytes) {  // the group lies in one matrix: one copy
                const uint32_t bytes = (uint32_t) (nlog * row_bytes);
                if (!PREFETCH) mbar_arrive_expect_tx(bar, bytes);
                mk_copy<PREFETCH>(st, first, bytes, bar, pol);
            } else {
                if (!PREFETCH) mbar_arrive_expect_tx(bar, (uint32_t) (nlog * row_bytes));
                for (int r = 0; r < nlog; ++r) mk_copy<PREFETCH>(st + (size_t) r * row_bytes, cursor_row_src<UNIT>(c, row0, r), (uint32_t) row_bytes, bar, pol);
            }
        }
    } else {  // k-segment: A parts [RG][ks * A_UNIT], then B parts [RG][ks * B_UNIT]
        const int u0 = c.seg * c.ks;
        const int nu = min(c.ks, c.nunits - u0);
        const int nrows = c.paired ? 2 * nlog : nlog;
        const uint32_t rsA = c.ks * F::A_UNIT, rsB = c.ks * F::B_UNIT, offB = MK_RG * c.ks * F::A_UNIT;
        if (!PREFETCH) mbar_arrive_expect_tx(bar, (uint32_t) nrows * nu * UNIT);
        for (int r = 0; r < MK_RG; ++r) {
            const bool valid = c.paired ? ((r < HALF ? r : r - HALF) < nlog) : (r < nlog);
            if (!valid) continue;
            const uint8_t * grow = cursor_row_src<UNIT>(c, row0, r);
            mk_copy<PREFETCH>(st + (size_t) r * rsA, grow + (size_t) u0 * F::A_UNIT, (uint32_t) nu * F::A_UNIT, bar, pol);
            if (F::B_UNIT)
                mk_copy<PREFETCH>(st + offB + (size_t) r * rsB, grow + (size_t) c.nunits * F::A_UNIT + (size_t) u0 * F::B_UNIT, (uint32_t) nu * F::B_UNIT, bar, pol);
        }
    }
}

struct MKRing {
    uint8_t * ring;   // this warp's stages
    uint64_t * bars;  // this warp's full barriers
    int iss_stage;            // stage the next issued item goes to
    int con_stage;            // stage of the next consumed item ...
    uint32_t con_parity;      // ... and the mbarrier phase parity it completes
    uint64_t pol;
};
__device__ __forceinline__ void ring_issue_advance(MKRing & rg, int stages) { if (++rg.iss_stage == stages) rg.iss_stage = 0; }
__device__ __forceinline__ void ring_consume_advance(MKRing & rg, int stages) { if (++rg.con_stage == stages) { rg.con_stage = 0; rg.con_parity ^= 1u; } }

// ---- one GEMV step: stage the activations (prologue), then consume this warp's items of phase `ph` ---------------------------------
template <class F>
__device__ __forceinline__ void gemv_step(const MKParams & p, const MKTabs & T, int ph, int ph_end, MKCursor & con, MKCursor & iss, MKCursor & pre, MKRing & rg, uint8_t * act_s,
                                          NormQuantSmem & nq, unsigned gw, unsigned GW, bool argmax = false) {
    constexpr int UNIT = F::A_UNIT + F::B_UNIT;
    constexpr int HALF = MK_RG / 2;
    constexpr int LPU = F::LPU;
    constexpr int UPS = 32 / LPU;
    const MKPhase * P = T.ph + ph;
    const int lane = threadIdx.x & 31;
    const int k = P->k;
    // pull the norm weights of the next two phases into L2 now (16 KB each, first touched by all 148 CTAs at once otherwise)
    if (threadIdx.x == 0) {
        for (int nx = ph + 1; nx <= ph + 2 && nx < p.n_phases; ++nx) {
            const MKPhase * N = T.ph + nx;
            if (N->pro == MK_PRO_NORM && (int) blockIdx.x < N->k / 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(N->pro_w + (size_t) blockIdx.x * 32));
        }
    }
    const int pro = P->pro;
    const ActLayout L = act_layout(F::Q8K, k);

    // ---- prologue: the quantized activation column of this phase in shared memory
    const long long t_pro0 = p.times ? clock64() : 0;
    if (pro == MK_PRO_COPY) {
        const uint8_t * src = P->pro_q;
        const int n16 = (int) (L.col_bytes >> 4);
        for (int i = threadIdx.x; i < n16; i += MK_THREADS) reinterpret_cast<uint4 *>(act_s)[i] = ldcg_u4(src + (size_t) i * 16);
    } else {
        const float * x = P->pro_x;
        const int nblk = k / 256;
        if (pro == MK_PRO_NORM) {
            // y = rms_norm(x) * w.  Sum of squares exactly as the CPU does it (ops.cpp:3736-3741): float products accumulated in DOUBLE,
            // mean rounded to float — a fp32 tree reduction can differ in the last bit of `scale`, which flips activation codes downstream.
            const float * w = P->pro_w;
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            if (nblk <= 3 * MK_WARPS) {  // k <= 6144: every element of the row lives in registers, one L2 round trip for x and w together
                float4 xa[3], xb[3], wa[3], wb[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int blk = warp + MK_WARPS * r;
                    const int64_t e = (int64_t) blk * 256 + 4 * lane;
                    const bool on = blk < nblk;
                    xa[r] = on ? ldcg_f4(x + e) : make_float4(0.f, 0.f, 0.f, 0.f);
                    xb[r] = on ? ldcg_f4(x + e + 128) : make_float4(0.f, 0.f, 0.f, 0.f);
                    wa[r] = on ? __ldg(reinterpret_cast<const float4 *>(w + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    wb[r] = on ? __ldg(reinterpret_cast<const float4 *>(w + e + 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                double ss = 0.0;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    ss += (double) __fmul_rn(xa[r].x, xa[r].x); ss += (double) __fmul_rn(xa[r].y, xa[r].y); ss += (double) __fmul_rn(xa[r].z, xa[r].z); ss += (double) __fmul_rn(xa[r].w, xa[r].w);
                    ss += (double) __fmul_rn(xb[r].x, xb[r].x); ss += (double) __fmul_rn(xb[r].y, xb[r].y); ss += (double) __fmul_rn(xb[r].z, xb[r].z); ss += (double) __fmul_rn(xb[r].w, xb[r].w);
                }
                ss = block_sum_f64(ss, reinterpret_cast<double *>(nq.red));
                const float mean = (float) (ss / (double) k);
                const float scale = 1.0f / sqrtf(mean + p.eps);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int blk = warp + MK_WARPS * r;
                    if (blk < nblk) {
                        const float4 ya = make_float4((xa[r].x * scale) * wa[r].x, (xa[r].y * scale) * wa[r].y, (xa[r].z * scale) * wa[r].z, (xa[r].w * scale) * wa[r].w);
                        const float4 yb = make_float4((xb[r].x * scale) * wb[r].x, (xb[r].y * scale) * wb[r].y, (xb[r].z * scale) * wb[r].z, (xb[r].w * scale) * wb[r].w);
                        quantize_block_warp<F::Q8K>(act_s, L, blk, ya, yb);
                    }
                }
            } else {
                double ss = 0.0;
                for (int64_t e = 4 * (int64_t) threadIdx.x; e < k; e += 4 * MK_THREADS) {
                    const float4 a = ldcg_f4(x + e);
                    ss += (double) __fmul_rn(a.x, a.x); ss += (double) __fmul_rn(a.y, a.y); ss += (double) __fmul_rn(a.z, a.z); ss += (double) __fmul_rn(a.w, a.w);
                }
                ss = block_sum_f64(ss, reinterpret_cast<double *>(nq.red));
                const float mean = (float) (ss / (double) k);
                const float scale = 1.0f / sqrtf(mean + p.eps);
                quantize_blocks<F::Q8K, 2>(act_s, k, 0, nblk, [&](int64_t e) {
                    const float4 a = ldcg_f4(x + e);
                    const float4 ww = __ldg(reinterpret_cast<const float4 *>(w + e));
                    return make_float4((a.x * scale) * ww.x, (a.y * scale) * ww.y, (a.z * scale) * ww.z, (a.w * scale) * ww.w);
                });
            }
        } else {
            quantize_blocks<F::Q8K, 4>(act_s, k, 0, nblk, [&](int64_t e) { return ldcg_f4(x + e); });
        }
        if (F::Q8K) {  // zero the padding of the last (partial) 1024-element group so the main loop can read it blindly
            for (int64_t e = k + 4 * (int64_t) threadIdx.x; e < L.qs_bytes; e += 4 * MK_THREADS) *reinterpret_cast<uint32_t *>(act_s + act_qs_off_q8k(e)) = 0u;
        }
    }
    __syncthreads();
    const long long t_pro = p.times ? clock64() - t_pro0 : 0;
    if (con.ph != ph) return;  // this warp owns no row of the phase

    const int epi = argmax ? (int) MK_EPI_STORE_ARGMAX : P->epi;
    const int g = lane % LPU, ul = lane / LPU;
    const bool PAIRED = con.paired != 0;
    const int group = PAIRED ? HALF : MK_RG;
    const bool whole = con.nseg == 1;
    const uint32_t rsA = whole ? con.nunits * UNIT : con.ks * F::A_UNIT;
    const uint32_t rsB = whole ? con.nunits * UNIT : con.ks * F::B_UNIT;
    const uint32_t offB = whole ? con.nunits * F::A_UNIT : MK_RG * con.ks * F::A_UNIT;
    float best = -INFINITY;
    int best_i = 0;
    const bool prof = p.times != nullptr && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (threadIdx.x >> 5) == 3;
    long long t_wait = 0, t_comp = 0, t_iss = 0, t_mark = prof ? clock64() : 0;
    int n_it = 0;

    float acc[MK_RG];
    while (con.ph == ph) {
        const int seg = con.seg;
        const int row0 = con.r0 + con.grp * group;
        const int nlog = min(group, con.r1 - row0);
        const int u0 = seg * con.ks;
        const int nu = min(con.ks, con.nunits - u0);
        if (seg == 0) {
#pragma unroll
            for (int r = 0; r < MK_RG; ++r) acc[r] = 0.0f;
        }
        mbar_wait(&rg.bars[rg.con_stage], rg.con_parity);
        if (prof) { const long long t = clock64(); t_wait += t - t_mark; t_mark = t; ++n_it; }
        const uint8_t * st = rg.ring + (size_t) rg.con_stage * p.stage_bytes;
        if (nlog == group) {
            // full group (the common case): straight-line code, the four rows' loads and dot products interleave freely
            for (int u = ul; u < nu; u += UPS) {
                typename F::Act A;
                F::load_act(act_s, L, u0 + u, g, A);
#pragma unroll
                for (int r = 0; r < MK_RG; r += 2) {  // two rows in flight: their shared-memory loads overlap the other's dp4a chains
                    typename F::Wt Wa, Wb;
                    F::load_w(st + (size_t) r * rsA, st + offB + (size_t) r * rsB, u, g, Wa);
                    F::load_w(st + (size_t) (r + 1) * rsA, st + offB + (size_t) (r + 1) * rsB, u, g, Wb);
                    acc[r] = F::dot(Wa, A, acc[r]);
                    acc[r + 1] = F::dot(Wb, A, acc[r + 1]);
                }
            }
        } else {
            for (int u = ul; u < nu; u += UPS) {
                typename F::Act A;
                F::load_act(act_s, L, u0 + u, g, A);
#pragma unroll
                for (int r = 0; r < MK_RG; ++r) {
                    const bool valid = PAIRED ? ((r < HALF ? r : r - HALF) < nlog) : (r < nlog);
                    if (valid) {
                        typename F::Wt Wr;
                        F::load_w(st + (size_t) r * rsA, st + offB + (size_t) r * rsB, u, g, Wr);
                        acc[r] = F::dot(Wr, A, acc[r]);
                    }
                }
            }
        }
        __syncwarp();
        if (prof) { const long long t = clock64(); t_comp += t - t_mark; t_mark = t; }
        ring_consume_advance(rg, p.stages);
        // the stage is free again: refill it with this warp's next item (possibly of a later phase)
        if (iss.ph < ph_end) {
            if (lane == 0) issue_item<F>(iss, rg.ring + (size_t) rg.iss_stage * p.stage_bytes, &rg.bars[rg.iss_stage], rg.pol);
            ring_issue_advance(rg, p.stages);
            cursor_next_shared(iss, p, T, ph_end, gw, GW);
        }
        // ... and keep HBM busy further ahead than shared memory can hold: the item l2_ahead positions later goes to L2 now, so the stream
        // does not stop while this warp sits at a grid barrier or quantizes activations (the ring alone covers ~3 us)
        if (p.l2_ahead > 0 && pre.ph < ph_end) {
            issue_item<F, true>(pre, nullptr, nullptr, 0);
            cursor_next_shared(pre, p, T, ph_end, gw, GW);
        }
        if (prof) { const long long t = clock64(); t_iss += t - t_mark; t_mark = t; }

        if (seg == con.nseg - 1) {
#pragma unroll
            for (int r = 0; r < MK_RG; ++r) acc[r] = warp_sum(acc[r]);
            if (lane == 0) {
                if (PAIRED) {
                    float * y = P->y[0];
#pragma unroll
                    for (int q = 0; q < HALF; ++q) {
                        if (q < nlog) {
                            const float gv = acc[q], uv = acc[HALF + q];
                            y[row0 + q] = (gv / (1.0f + expf(-gv))) * uv;
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < MK_RG; ++r) {
                        if (r < nlog) {
                            int mi = 0, lrow = row0 + r;
                            if (con.nmat > 1 && lrow >= con.m0) { lrow -= con.m0; mi = 1; if (con.nmat > 2 && lrow >= con.m1) { lrow -= con.m1; mi = 2; } }
                            float * y = P->y[mi];
                            const float * bias = P->bias[mi];
                            float o = acc[r];
                            if (bias) o += __ldg(bias + lrow);
                            if (epi == MK_EPI_RESID) o += ldcg_f(y + lrow);
                            y[lrow] = o;
                            if (epi == MK_EPI_STORE_ARGMAX && o > best) { best = o; best_i = lrow; }  // rows ascend: the first maximum wins
                        }
                    }
                }
            }
        }
        cursor_next(con, p, T, ph_end, gw, GW);
    }
    if (epi == MK_EPI_STORE_ARGMAX && lane == 0 && best > -INFINITY) {
        // order-preserving key: larger value first, then the SMALLER index (the host's argmax keeps the first maximum)
        const uint32_t b = __float_as_uint(best);
        const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        atomicMax(p.sync + 2, ((unsigned long long) ord << 32) | (unsigned) (0xffffffffu - (unsigned) best_i));
    }
    if (prof && lane == 0) {
        long long * a = p.times + 1 + 7 * p.n_layers + 3 + ((blockIdx.x ? 5 : 0) + (ph == 4 * p.n_layers ? 4 : (ph & 3))) * 8;
        a[0] += t_wait; a[1] += t_comp; a[2] += t_iss; a[3] += n_it; a[4] += 1; a[5] += t_pro;
    }
}

// ---- attention -------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mk_mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// scores[h][t] = scale * K[t][grp] . f16(rope(q[h]))  per (16-position tile, KV group) unit, one unit per warp at a time; also the per
// (head, tile) max and sum(exp(s - max)).  The warp that owns the tile of the NEW position first ropes k and appends the K-cache row.
template <int HD>
__device__ __forceinline__ void scores_step(const MKParams & p, const MKTabs & T, int layer, int n_kv, int kpos, uint8_t * sm) {
    constexpr int NU = HD / 32;
    constexpr int HP = HD / 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int gqa = p.heads / p.kv_heads;
    __half * qh = reinterpret_cast<__half *>(sm);                       // [heads][HD] roped q as f16
    float2 * cs = reinterpret_cast<float2 *>(sm + (size_t) p.heads * HD * 2);  // [HD/2] cos, sin of this position
    // RoPE angles: the CPU's fp32 recurrence seeded with the position (ggml-cpu ops.cpp:5613-5628), as rope_kv_store_kernel (fused.cu)
    if (threadIdx.x < HP) {
        const int i = threadIdx.x;
        float theta = (float) kpos;
        for (int j = 0; j < i; ++j) theta *= p.theta_scale;
        const float th = theta / (p.rope_ff ? __ldg(p.rope_ff + i) : 1.0f);
        cs[i] = make_float2(cosf(th), sinf(th));
    }
    __syncthreads();
    {
        // RoPE(q) -> f16 in shared memory; every load of the thread is issued before the first one is used (one L2 round trip, not eight)
        constexpr int NPT = 8;   // pairs per thread issued together: covers hidden <= 4096 (the loop below takes the rest)
        float x0[NPT], x1[NPT];
        const int total = p.heads * HP;
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const int idx = threadIdx.x + j * MK_THREADS;
            if (idx < total) {
                const int h = idx / HP, i = idx - h * HP;
                const int i0 = (p.rope_mode == 0) ? 2 * i : i, i1 = (p.rope_mode == 0) ? 2 * i + 1 : i + HP;
                x0[j] = ldcg_f(p.q + (int64_t) h * HD + i0); x1[j] = ldcg_f(p.q + (int64_t) h * HD + i1);
            }
        }
#pragma unroll
        for (int j = 0; j < NPT; ++j) {
            const int idx = threadIdx.x + j * MK_THREADS;
            if (idx < total) {
                const int h = idx / HP, i = idx - h * HP;
                const int i0 = (p.rope_mode == 0) ? 2 * i : i, i1 = (p.rope_mode == 0) ? 2 * i + 1 : i + HP;
                const float c = cs[i].x, sn = cs[i].y;
                qh[h * HD + i0] = __float2half_rn(x0[j] * c - x1[j] * sn);
                qh[h * HD + i1] = __float2half_rn(x0[j] * sn + x1[j] * c);
            }
        }
        for (int idx = threadIdx.x + NPT * MK_THREADS; idx < total; idx += MK_THREADS) {  // hidden > 4096
            const int h = idx / HP, i = idx - h * HP;
            const int i0 = (p.rope_mode == 0) ? 2 * i : i, i1 = (p.rope_mode == 0) ? 2 * i + 1 : i + HP;
            const float a = ldcg_f(p.q + (int64_t) h * HD + i0), b = ldcg_f(p.q + (int64_t) h * HD + i1);
            qh[h * HD + i0] = __float2half_rn(a * cs[i].x - b * cs[i].y);
            qh[h * HD + i1] = __float2half_rn(a * cs[i].y + b * cs[i].x);
        }
    }
    __syncthreads();

    __half * kc = T.kv[layer].kc;
    const int ntiles = (n_kv + 15) / 16;
    const int total = ntiles * p.kv_heads;
    const int ustride = gridDim.x * MK_WARPS;
    uint4 alo[2][NU], ahi[2][NU];
    auto load_unit = [&](int u, uint4 (&lo)[NU], uint4 (&hi)[NU]) {
        const int grp = u % p.kv_heads, tile = u / p.kv_heads;
        if (tile == (kpos >> 4)) {
            // append the K row of the new position: RoPE(k) -> f16 (KVCacheAttention::save_to_cache, src/layers.cpp:3044-3123)
            for (int i = lane; i < HP; i += 32) {
                const int i0 = (p.rope_mode == 0) ? 2 * i : i, i1 = (p.rope_mode == 0) ? 2 * i + 1 : i + HP;
                const float x0 = ldcg_f(p.kbuf + (int64_t) grp * HD + i0), x1 = ldcg_f(p.kbuf + (int64_t) grp * HD + i1);
                const float c = cs[i].x, s = cs[i].y;
                __half * krow = kc + (int64_t) kpos * p.k_row_stride + (int64_t) grp * HD;
                krow[i0] = __float2half_rn(x0 * c - x1 * s);
                krow[i1] = __float2half_rn(x0 * s + x1 * c);
            }
            __threadfence();  // the row is read back through L2 by the other lanes of this warp
            __syncwarp();
        }
        const int rowA = tile * 16 + g, rowB = rowA + 8;
#pragma unroll
        for (int q = 0; q < NU; ++q) {
            lo[q] = (rowA < n_kv) ? ldcg_u4(kc + (int64_t) rowA * p.k_row_stride + (int64_t) grp * HD + 8 * t + 32 * q) : make_uint4(0, 0, 0, 0);
            hi[q] = (rowB < n_kv) ? ldcg_u4(kc + (int64_t) rowB * p.k_row_stride + (int64_t) grp * HD + 8 * t + 32 * q) : make_uint4(0, 0, 0, 0);
        }
    };
    auto compute_unit = [&](int u, const uint4 (&lo)[NU], const uint4 (&hi)[NU]) {
        const int grp = u % p.kv_heads, tile = u / p.kv_heads;
        const int rowA = tile * 16 + g, rowB = rowA + 8;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int q = 0; q < NU; ++q) {
            uint4 b = make_uint4(0, 0, 0, 0);
            if (g < gqa) b = *reinterpret_cast<const uint4 *>(qh + (size_t) (grp * gqa + g) * HD + 8 * t + 32 * q);
            mk_mma16816(c, lo[q].x, hi[q].x, lo[q].y, hi[q].y, b.x, b.y);
            mk_mma16816(c, lo[q].z, hi[q].z, lo[q].w, hi[q].w, b.z, b.w);
        }
        // c0 = (rowA, head 2t), c1 = (rowA, head 2t+1), c2 = (rowB, head 2t), c3 = (rowB, head 2t+1)
        const int h0 = 2 * t, h1 = 2 * t + 1;
        float v[4];
        v[0] = (rowA < n_kv) ? c[0] * p.attn_scale : -INFINITY; v[1] = (rowA < n_kv) ? c[1] * p.attn_scale : -INFINITY;
        v[2] = (rowB < n_kv) ? c[2] * p.attn_scale : -INFINITY; v[3] = (rowB < n_kv) ? c[3] * p.attn_scale : -INFINITY;
        if (h0 < gqa) {
            if (rowA < n_kv) p.scores[(int64_t) (grp * gqa + h0) * p.s_stride + rowA] = v[0];
            if (rowB < n_kv) p.scores[(int64_t) (grp * gqa + h0) * p.s_stride + rowB] = v[2];
        }
        if (h1 < gqa) {
            if (rowA < n_kv) p.scores[(int64_t) (grp * gqa + h1) * p.s_stride + rowA] = v[1];
            if (rowB < n_kv) p.scores[(int64_t) (grp * gqa + h1) * p.s_stride + rowB] = v[3];
        }
        float m0 = fmaxf(v[0], v[2]), m1 = fmaxf(v[1], v[3]);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o)); }
        float s0 = (m0 == -INFINITY) ? 0.0f : expf(v[0] - m0) + expf(v[2] - m0);
        float s1 = (m1 == -INFINITY) ? 0.0f : expf(v[1] - m1) + expf(v[3] - m1);
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
        if (g == 0) {
            if (h0 < gqa) p.part[(int64_t) (grp * gqa + h0) * p.max_tiles + tile] = make_float2(m0, s0);
            if (h1 < gqa) p.part[(int64_t) (grp * gqa + h1) * p.max_tiles + tile] = make_float2(m1, s1);
        }
    };
    // units of this warp: blockIdx.x + gridDim.x * (warp + 8 j) — consecutive CTAs share a tile (2 KB contiguous K rows), SMs stay balanced
    int u = blockIdx.x + gridDim.x * warp;
    if (u < total) load_unit(u, alo[0], ahi[0]);
    while (u < total) {
        const int u1 = u + ustride;
        if (u1 < total) load_unit(u1, alo[1], ahi[1]);
        compute_unit(u, alo[0], ahi[0]);
        if (u1 >= total) break;
        const int u2 = u1 + ustride;
        if (u2 < total) load_unit(u2, alo[0], ahi[0]);
        compute_unit(u1, alo[1], ahi[1]);
        u = u2;
    }
}

// positions per CTA of the V.P step: the KV groups share the grid, spans are multiples of 16
__device__ __forceinline__ int mk_pv_span(int n_kv, int per_grp) {
    int span = (n_kv + per_grp - 1) / per_grp;
    span = (span + 15) & ~15;
    return span < 16 ? 16 : span;
}

// att[e] = sum_split partial[split][e] (in split order) for the 256-element blocks [blk_begin, blk_end), quantized as the o-projection's
// activations into global memory (one warp per block)
template <bool Q8K>
__device__ __forceinline__ void sum_partials_quantize(const MKParams & p, int nsplit, int blk_begin, int blk_end) {
    const int64_t ne = (int64_t) p.heads * p.head_dim;
    quantize_blocks<Q8K, 1>(p.att_q, ne, blk_begin, blk_end, [&](int64_t e) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c0 = 0; c0 < nsplit; c0 += 8) {  // 8 partial rows in flight at a time; summed in split order
            float4 b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = (c0 + i < nsplit) ? ldcg_f4(p.partial + (int64_t) (c0 + i) * ne + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { a.x += b[i].x; a.y += b[i].y; a.z += b[i].z; a.w += b[i].w; }
        }
        return a;
    });
}
// separate tail step (KV groups that do not own whole 256-element blocks, e.g. GQA 7 x 128): CTA b takes 8 blocks
template <bool Q8K>
__device__ __forceinline__ void tail_step(const MKParams & p, int n_kv) {
    const int nblk = p.heads * p.head_dim / 256;
    const int b0 = (int) blockIdx.x * MK_WARPS;
    if (b0 >= nblk) return;
    const int per_grp = max(1, (int) gridDim.x / p.kv_heads);
    const int span = mk_pv_span(n_kv, per_grp);
    sum_partials_quantize<Q8K>(p, (n_kv + span - 1) / span, b0, min(nblk, b0 + MK_WARPS));
}

// partial[split][h][d] = sum over the split's positions of Vt[grp*HD + d][t] * P[h][t],  P = f16(exp(s - max_h) / sum_h) with the GLOBAL
// max / sum of the row (from the per-tile statistics).  The CTA whose span holds the new position first appends the V-cache column.
template <int HD, bool Q8K>
__device__ __forceinline__ void pv_step(const MKParams & p, const MKTabs & T, int layer, int n_kv, int kpos, int v_col, uint8_t * sm) {
    constexpr int MW = HD / 16;  // warps that own 16 channels each
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int gqa = p.heads / p.kv_heads;
    const int per_grp = max(1, (int) gridDim.x / p.kv_heads);
    const int span = mk_pv_span(n_kv, per_grp);
    const int nsplit = (n_kv + span - 1) / span;
    const int grp = blockIdx.x % p.kv_heads, split = blockIdx.x / p.kv_heads;
    if (split >= nsplit || (int) blockIdx.x >= per_grp * p.kv_heads) return;
    __half * Ph = reinterpret_cast<__half *>(sm);  // [8][SP]
    float * hmax = reinterpret_cast<float *>(sm + 8 * MK_PV_SP * 2);
    float * hinv = hmax + 8;
    const int ntiles = (n_kv + 15) / 16;
    const int p_begin = split * span, p_end = min(n_kv, p_begin + span);
    for (int i = threadIdx.x; i < (8 - gqa) * MK_PV_SP / 2; i += MK_THREADS) reinterpret_cast<uint32_t *>(Ph + gqa * MK_PV_SP)[i] = 0u;

    __half * vc = T.kv[layer].vc;
    const int c0 = warp * 16;
    const __half * rowA = vc + (int64_t) (grp * HD + c0 + g) * p.v_row_stride;
    const __half * rowB = rowA + 8 * p.v_row_stride;
    if (warp < MW && v_col >= p_begin && v_col < p_end) {
        // append the V column of the new token (transposed cache: one f16 per channel row, src/layers.cpp:3095-3110)
        if (lane < 16) vc[(int64_t) (grp * HD + c0 + lane) * p.v_row_stride + v_col] = __float2half_rn(ldcg_f(p.vbuf + (int64_t) grp * HD + c0 + lane));
        __threadfence();
        __syncwarp();
    }
    auto ldv = [&](const __half * row, int p0) -> uint4 {
        if (p0 >= p_end) return make_uint4(0, 0, 0, 0);
        uint4 v = ldcg_u4(row + p0);
        const int r = p_end - p0;  // valid halves in this 16-byte chunk
        if (r < 8) {               // tail: zero what lies beyond n_kv (P is zero there, but the cache may hold anything, NaN included)
            uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (2 * j + 1 < r) ? w[j] : ((2 * j < r) ? (w[j] & 0xffffu) : 0u);
            v = make_uint4(w[0], w[1], w[2], w[3]);
        }
        return v;
    };
    uint4 alo[2][4], ahi[2][4];
    auto load_round = [&](int p0) {
        if (warp < MW) {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    alo[s2][u] = ldv(rowA, p0 + s2 * 128 + 8 * t + 32 * u);
                    ahi[s2][u] = ldv(rowB, p0 + s2 * 128 + 8 * t + 32 * u);
                }
        }
    };
    load_round(p_begin);  // DRAM loads first, the (L2-resident) statistics and scores behind them
    for (int h = warp; h < gqa; h += MK_WARPS) {
        // global max / sum of the row from the per-tile statistics: each lane folds its tiles (8 loads in flight at a time) into a running
        // (max, sum), then the warp combines the 32 pairs
        const float2 * pp = p.part + (int64_t) (grp * gqa + h) * p.max_tiles;
        float m_l = -INFINITY, s_l = 0.0f;
        for (int i0 = lane; i0 < ntiles; i0 += 32 * 8) {
            float2 st[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) st[j] = (i0 + 32 * j < ntiles) ? __ldcg(pp + i0 + 32 * j) : make_float2(-INFINITY, 0.0f);
            float mc = m_l;
#pragma unroll
            for (int j = 0; j < 8; ++j) mc = fmaxf(mc, st[j].x);
            if (mc > -INFINITY) {
                float sc = (m_l > -INFINITY) ? s_l * expf(m_l - mc) : 0.0f;
#pragma unroll
                for (int j = 0; j < 8; ++j) sc += (st[j].x > -INFINITY) ? st[j].y * expf(st[j].x - mc) : 0.0f;
                m_l = mc; s_l = sc;
            }
        }
        const float mx = warp_max(m_l);
        const float sum = warp_sum((m_l > -INFINITY) ? s_l * expf(m_l - mx) : 0.0f);
        if (lane == 0) { hmax[h] = mx; hinv[h] = 1.0f / sum; }
    }
    __syncthreads();

    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int p0 = p_begin; p0 < p_end; p0 += MK_PV_ROUND) {
        if (p0 != p_begin) { __syncthreads(); load_round(p0); }  // every warp is done with the previous round's P
        for (int i = threadIdx.x; i < MK_PV_ROUND; i += MK_THREADS) {
            const int pp0 = p0 + i;
            for (int h = 0; h < gqa; ++h) {
                // __expf (ex2.approx, ~2 ulp): the value is rounded to f16 (11 bits) right away
                const float e = (pp0 < p_end) ? __expf(ldcg_f(p.scores + (int64_t) (grp * gqa + h) * p.s_stride + pp0) - hmax[h]) * hinv[h] : 0.0f;
                Ph[h * MK_PV_SP + i] = __float2half_rn(e);
            }
        }
        __syncthreads();
        if (warp < MW) {
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
                if (p0 + s2 * 128 < p_end) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const uint4 b = *reinterpret_cast<const uint4 *>(Ph + g * MK_PV_SP + s2 * 128 + 8 * t + 32 * u);
                        mk_mma16816(c, alo[s2][u].x, ahi[s2][u].x, alo[s2][u].y, ahi[s2][u].y, b.x, b.y);
                        mk_mma16816(c, alo[s2][u].z, ahi[s2][u].z, alo[s2][u].w, ahi[s2][u].w, b.z, b.w);
                    }
                }
            }
        }
    }
    if (warp < MW) {
        // c0 = (ch g, head 2t), c1 = (ch g, head 2t+1), c2 = (ch g+8, head 2t), c3 = (ch g+8, head 2t+1)
        float * po = p.partial + ((int64_t) split * p.heads + (int64_t) grp * gqa) * HD + c0;
        if (2 * t < gqa) { po[(int64_t) (2 * t) * HD + g] = c[0]; po[(int64_t) (2 * t) * HD + g + 8] = c[2]; }
        if (2 * t + 1 < gqa) { po[(int64_t) (2 * t + 1) * HD + g] = c[1]; po[(int64_t) (2 * t + 1) * HD + g + 8] = c[3]; }
    }
    if (p.merged_tail) {
        // the LAST split of this KV group to finish sums the group's partials (in split order: deterministic whoever does it) and emits the
        // quantized activations of the o-projection for the group's slice — the separate tail step and its grid barrier disappear
        int * flag = reinterpret_cast<int *>(hinv + 8);
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned long long old = atomicAdd(p.sync + 8 + grp, 1ull);
            const int last = old == (unsigned long long) (nsplit - 1);
            if (last) { p.sync[8 + grp] = 0ull; __threadfence(); }
            flag[0] = last;
        }
        __syncthreads();
        if (flag[0]) {
            const int bpg = gqa * HD / 256;
            sum_partials_quantize<Q8K>(p, nsplit, grp * bpg, (grp + 1) * bpg);
        }
    }
}

// pull this layer's K rows and V channel rows [0, n_kv) into L2 while the q/k/v projection streams its weights: the attention steps are
// latency chains and would otherwise start with a DRAM round trip behind ~20 MB of queued weight requests
__device__ __forceinline__ void prefetch_kv_l2(const MKParams & p, const MKTabs & T, int layer, int n_kv) {
    const int kv_hidden = p.kv_heads * p.head_dim;
    const uint8_t * kc = reinterpret_cast<const uint8_t *>(T.kv[layer].kc);
    const uint8_t * vc = reinterpret_cast<const uint8_t *>(T.kv[layer].vc);
    const int64_t gtid = (int64_t) blockIdx.x * MK_THREADS + threadIdx.x, gthreads = (int64_t) gridDim.x * MK_THREADS;
    // one 128-byte line per thread and instruction (LSU prefetches; 16.8 MB at 4K context = 3-4 per thread)
    const int64_t klines_row = ((int64_t) kv_hidden * 2 + 127) / 128;
    for (int64_t i = gtid; i < (int64_t) n_kv * klines_row; i += gthreads) {
        const int64_t r = i / klines_row, l = i - r * klines_row;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(kc + r * p.k_row_stride * 2 + l * 128));
    }
    const int64_t vlines_row = ((int64_t) n_kv * 2 + 127) / 128;
    for (int64_t i = gtid; i < (int64_t) kv_hidden * vlines_row; i += gthreads) {
        const int64_t r = i / vlines_row, l = i - r * vlines_row;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vc + r * p.v_row_stride * 2 + l * 128));
    }
}

// ---- the kernel -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int mk_phases_before(int step) {  // quantized-matmul phases among the steps [0, step)
    if (step <= 1) return 0;
    const int tt = step - 1, l = tt / 7, r = tt % 7;
    return 4 * l + (r > 0) + (r > 4) + (r > 5);
}

template <class F, int HD>
__global__ void __launch_bounds__(MK_THREADS, 1) decode_mk_kernel(const MKParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ NormQuantSmem nq;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // dynamic shared memory: [phase table | KV table] [activation column / attention scratch] [mbarriers] [8 warps x stages x stage]
    MKTabs T{p.phases, p.kv};
    if (p.tab_bytes) {
        const size_t nph = (size_t) p.n_phases * sizeof(MKPhase), nkv = (size_t) p.n_layers * sizeof(MKLayerKV);
        unsigned long long * dst = reinterpret_cast<unsigned long long *>(smem);
        const unsigned long long * s0 = reinterpret_cast<const unsigned long long *>(p.phases), * s1 = reinterpret_cast<const unsigned long long *>(p.kv);
        for (size_t i = threadIdx.x; i < nph / 8; i += MK_THREADS) dst[i] = s0[i];
        for (size_t i = threadIdx.x; i < nkv / 8; i += MK_THREADS) dst[nph / 8 + i] = s1[i];
        T.ph = reinterpret_cast<const MKPhase *>(smem);
        T.kv = reinterpret_cast<const MKLayerKV *>(smem + nph);
        __syncthreads();
    }
    uint8_t * act_s = smem + p.tab_bytes;
    MKRing rg;
    rg.bars = reinterpret_cast<uint64_t *>(act_s + p.act_bytes) + (size_t) warp * p.stages;
    rg.ring = act_s + p.act_bytes + al16((int64_t) MK_WARPS * p.stages * 8) + (size_t) warp * p.stages * p.stage_bytes;
    rg.iss_stage = rg.con_stage = 0;
    rg.con_parity = 0;
    rg.pol = 0;
    if (lane == 0) {
        for (int s = 0; s < p.stages; ++s) mbar_init(&rg.bars[s], 1);
        fence_mbar_init();
        rg.pol = make_evict_first_policy();
    }
    __syncwarp();

    const unsigned gw = blockIdx.x * MK_WARPS + warp, GW = gridDim.x * MK_WARPS;
    const int ph_begin = mk_phases_before(p.step_begin), ph_end = min(p.n_phases, mk_phases_before(p.step_end));
    // the consume cursor lives in registers (hot loop); the issue / L2-prefetch cursors are only touched around a bulk copy and live in shared
    // memory (every lane reads the same words; lane 0 writes), which keeps the kernel under the 255-register ceiling
    __shared__ MKCursor s_iss[MK_WARPS], s_pre[MK_WARPS];
    MKCursor con;
    MKCursor & iss = s_iss[warp];
    MKCursor & pre = s_pre[warp];
    cursor_seek(con, p, T, ph_begin, ph_end, gw, GW);
    if (lane == 0) iss = con;
    __syncwarp();
    for (int i = 0; i < p.stages && iss.ph < ph_end; ++i) {
        if (lane == 0) issue_item<F>(iss, rg.ring + (size_t) rg.iss_stage * p.stage_bytes, &rg.bars[rg.iss_stage], rg.pol);
        ring_issue_advance(rg, p.stages);
        cursor_next_shared(iss, p, T, ph_end, gw, GW);
    }
    if (lane == 0) pre = iss;  // the L2 prefetch cursor runs l2_ahead items in front of the ring's issue cursor
    __syncwarp();
    for (int i = 0; i < p.l2_ahead && pre.ph < ph_end; ++i) {
        issue_item<F, true>(pre, nullptr, nullptr, 0);
        cursor_next_shared(pre, p, T, ph_end, gw, GW);
    }

    // layer-sharded multi-GPU: the previous shard (or, for the first shard, the last one with the next token) raises our flag through
    // NVLink when this launch's input is in place.  The weight rings above are already streaming while we wait.
    const unsigned long long seq = p.sync[4];  // launches of this plan completed so far
    if (p.wait_flag) {
        if (threadIdx.x == 0) {
            const unsigned long long want = seq + (unsigned long long) p.wait_offset;
            const unsigned long long t0 = globaltimer_ns();
            unsigned spins = 0;
            while (ld_acquire_sys_u64(p.wait_flag) < want) {
                if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 20000000000ull) { atomicExch(p.sync + 3, 2ull); break; }
            }
        }
        __syncthreads();
    }
    const int kpos = __ldcg(p.pos);
    const int n_kv = p.n_kv_arg >= 0 ? p.n_kv_arg : kpos + 1;
    const int v_col = p.v_col_arg >= 0 ? p.v_col_arg : kpos;
    unsigned long long bar_target = p.sync[1];  // value of the counter when this launch started (written by the previous launch's last CTA 0)
    const int step_head = 1 + 7 * p.n_layers;

    if (p.times && blockIdx.x == 0 && threadIdx.x == 0) p.times[0] = clock64();
    for (int s = p.step_begin; s < p.step_end; ++s) {
        // which quantized-matmul phase (if any) this step is: one call site for the GEMV keeps the kernel small
        int gph = -1, layer = 0, kind = -1;
        if (s == step_head) { if (p.has_head) gph = 4 * p.n_layers; }
        else if (s > 0 && s < step_head) {
            layer = (s - 1) / 7; kind = (s - 1) % 7;
            if (kind == 0) gph = 4 * layer; else if (kind >= 4) gph = 4 * layer + kind - 3;
        }
        if (gph >= 0) {
            if (kind == 0 && p.kv_prefetch) prefetch_kv_l2(p, T, layer, n_kv);
            gemv_step<F>(p, T, gph, ph_end, con, iss, pre, rg, act_s, nq, gw, GW, s == step_head && p.next_tok != nullptr);
        } else if (s == 0) {
            if (p.embed) {  // Embedding::forward (src/layers.cpp:2038-2067): x = dequant(table[tok])
                const int64_t row = __ldcg(p.tok);  // (a peer may have just written it)
                const uint8_t * base = p.embed + row * dequant_row_bytes(p.embed_type, p.hidden);
                for (int64_t e = (int64_t) blockIdx.x * MK_THREADS + threadIdx.x; e < p.hidden; e += (int64_t) gridDim.x * MK_THREADS)
                    p.x[e] = dequant_row_elem(p.embed_type, base, p.hidden, e);
            }
        } else if (s == step_head + 1) {
            int idx = 0;
            if (p.has_head && p.next_tok && blockIdx.x == 0 && threadIdx.x == 0) {
                const unsigned long long key = __ldcg(p.sync + 2);
                idx = (int) (0xffffffffu - (unsigned) (key & 0xffffffffu));
                p.next_tok[0] = idx;
                p.sync[2] = 0ull;
                if (p.flags & MK_FLAG_ADVANCE) {
                    const_cast<int32_t *>(p.tok)[0] = idx;
                    const_cast<int32_t *>(p.pos)[0] = kpos + 1;
                }
            }
            // ---- hand-off to the next shard: the hidden row (or, from the last shard, the next token) is stored straight into the peer's
            // memory over NVLink, then its flag is raised with a system-scope release (the grid barrier before this step made every
            // CTA's writes to x visible to this CTA)
            if (blockIdx.x == 0 && (p.send_x || p.send_tok)) {
                if (p.send_x) {
                    for (int e = 4 * threadIdx.x; e < p.hidden; e += 4 * MK_THREADS) *reinterpret_cast<float4 *>(p.send_x + e) = ldcg_f4(p.x + e);
                }
                if (p.send_tok && threadIdx.x == 0) p.send_tok[0] = idx;
                __threadfence_system();
                __syncthreads();
                if (threadIdx.x == 0 && p.send_flag) st_release_sys_u64(p.send_flag, seq + 1);
            }
            if (blockIdx.x == 0 && threadIdx.x == 0 && (p.flags & MK_FLAG_ADVANCE_POS) && !(p.flags & MK_FLAG_ADVANCE)) const_cast<int32_t *>(p.pos)[0] = kpos + 1;
        } else if (kind == 1) {
            scores_step<HD>(p, T, layer, n_kv, kpos, act_s);
        } else if (kind == 2) {
            pv_step<HD, F::Q8K>(p, T, layer, n_kv, kpos, v_col, act_s);
        } else if (kind == 3) {
            if (p.merged_tail) continue;  // done by the last V.P CTA of every KV group: no step, no barrier
            tail_step<F::Q8K>(p, n_kv);
        }
        if (s + 1 < p.step_end) grid_barrier(p, bar_target);
        if (p.times && blockIdx.x == 0 && threadIdx.x == 0) p.times[1 + s - p.step_begin] = clock64();
    }
    // the counter value the next launch starts from (every CTA read sync[1] before its first barrier, which CTA 0 has passed)
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.step_end - p.step_begin > 1) {
        p.sync[1] = bar_target;
        if (p.step_begin == 0 && p.step_end == step_head + 2) p.sync[4] = seq + 1;  // a whole token completed
    }
}

// ======================================================================================================
// host side: the plan (phase table + workspace) and the launch
// ======================================================================================================
struct DecodePlan {
    int device = 0;
    MKParams prm{};
    MKPhase * d_phases = nullptr;
    MKLayerKV * d_kv = nullptr;
    void * ws = nullptr;
    size_t ws_bytes = 0;
    int max_ctx = 0;
    int grid = 0;
    size_t smem = 0;
    int wtype = 0;
    std::vector<MKPhase> h_phases;
    std::vector<MKLayerKV> h_kv;
    bool coop = true;
    bool dirty = true;            // host tables differ from the device copies
    float * up_x = nullptr;       // residual stream / logits the device tables currently point at
    float * up_logits = nullptr;
};

static int mk_env_int(const char * name, int dflt) {
    const char * v = getenv(name);
    return v ? atoi(v) : dflt;
}

template <class F, int HD>
static cudaError_t mk_launch_t(const DecodePlan & pl, const MKParams & prm, cudaStream_t st) {
    auto kern = decode_mk_kernel<F, HD>;
    static size_t configured[16] = {0};  // per template instantiation and per device (function attributes are per device)
    if (pl.smem > configured[pl.device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pl.smem);
        if (e != cudaSuccess) return e;
        configured[pl.device & 15] = pl.smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) pl.grid);
    cfg.blockDim = dim3(MK_THREADS);
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barrier cannot deadlock
    attr[0].val.cooperative = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pl.coop ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, prm);
}

static cudaError_t mk_launch(const DecodePlan & pl, const MKParams & prm, cudaStream_t st) {
    const bool hd128 = prm.head_dim == 128;
    switch (pl.wtype) {
        case B200_TYPE_Q4_K: return hd128 ? mk_launch_t<FmtQ4K, 128>(pl, prm, st) : mk_launch_t<FmtQ4K, 64>(pl, prm, st);
        case B200_TYPE_Q4_0: return hd128 ? mk_launch_t<FmtQ40, 128>(pl, prm, st) : mk_launch_t<FmtQ40, 64>(pl, prm, st);
        default: return hd128 ? mk_launch_t<FmtQ80, 128>(pl, prm, st) : mk_launch_t<FmtQ80, 64>(pl, prm, st);
    }
}

static int mk_unit_bytes(int wtype) {
    switch (wtype) {
        case B200_TYPE_Q4_K: return FmtQ4K::A_UNIT + FmtQ4K::B_UNIT;
        case B200_TYPE_Q4_0: return FmtQ40::A_UNIT + FmtQ40::B_UNIT;
        case B200_TYPE_Q8_0: return FmtQ80::A_UNIT + FmtQ80::B_UNIT;
        default: return 0;
    }
}

void decode_plan_destroy(void * h) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl) return;
    cudaSetDevice(pl->device);
    if (pl->d_phases) cudaFree(pl->d_phases);
    if (pl->d_kv) cudaFree(pl->d_kv);
    if (pl->ws) cudaFree(pl->ws);
    delete pl;
}

// Build the phase table + workspace for a model description (device pointers inside; `layers` itself is a HOST array).
void * decode_plan_create(const DecodeModel & m, int max_ctx, int * err) {
    auto fail = [&](int e) -> void * { if (err) *err = e; return nullptr; };
    const int unit = mk_unit_bytes(m.wtype);
    if (!unit) return fail(B200_ERR_UNSUPPORTED);
    if (m.n_layers < 0 || m.hidden <= 0 || m.hidden % 256 || m.ffn % 256 || (m.head_dim != 64 && m.head_dim != 128) || m.kv_heads <= 0 || m.heads % m.kv_heads ||
        m.heads / m.kv_heads > 8 || m.heads * m.head_dim != m.hidden || m.hidden > 20480 || m.ffn > 20480 || (m.rope_mode != 0 && m.rope_mode != 2) ||
        (m.k_row_stride % 8) || (m.v_row_stride % 8) || max_ctx <= 0)
        return fail(B200_ERR_UNSUPPORTED);
    if (m.n_layers > 0 && !m.layers) return fail(B200_ERR_ARG);
    if (m.lm_head && (m.vocab <= 0 || !m.final_norm)) return fail(B200_ERR_ARG);
    if ((int64_t) std::max(std::max(m.vocab, m.ffn), m.hidden + 2 * m.kv_heads * m.head_dim) * 148 * MK_WARPS >= ((int64_t) 1 << 32)) return fail(B200_ERR_UNSUPPORTED);
    DecodePlan * pl = new DecodePlan;
    cudaGetDevice(&pl->device);
    pl->wtype = m.wtype;
    pl->max_ctx = max_ctx;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
    pl->grid = mk_env_int("B200_MK_GRID", sms > 0 ? sms : 148);
    pl->coop = mk_env_int("B200_MK_COOP", 1) != 0;
    MKParams & P = pl->prm;
    P.n_layers = m.n_layers; P.hidden = m.hidden; P.heads = m.heads; P.kv_heads = m.kv_heads; P.head_dim = m.head_dim; P.ffn = m.ffn; P.vocab = m.vocab;
    P.rope_mode = m.rope_mode; P.theta_scale = powf(m.rope_theta, -2.0f / m.head_dim); P.eps = m.eps; P.attn_scale = m.attn_scale; P.rope_ff = m.rope_freq_factors;
    P.k_row_stride = m.k_row_stride; P.v_row_stride = m.v_row_stride;
    P.embed = (const uint8_t *) m.embed; P.embed_type = m.embed_type ? m.embed_type : m.wtype;
    P.has_head = m.lm_head ? 1 : 0;

    // ---- workspace
    const int kvh = m.kv_heads, kv_hidden = kvh * m.head_dim;
    const int64_t s_stride = ((int64_t) max_ctx + 7) & ~(int64_t) 7;
    const int max_tiles = (max_ctx + 15) / 16;
    const int per_grp = std::max(1, pl->grid / kvh);
    const size_t acb_h = qact_col_bytes(m.wtype, m.hidden);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t) 255; return o; };
    const size_t o_sync = take(1024), o_q = take((size_t) m.hidden * 4), o_k = take((size_t) kv_hidden * 4), o_v = take((size_t) kv_hidden * 4),
                 o_sc = take((size_t) m.heads * s_stride * 4), o_part = take((size_t) m.heads * max_tiles * 8),
                 o_partial = take((size_t) per_grp * m.hidden * 4), o_attq = take(acb_h), o_gate = take((size_t) m.ffn * 4), o_x = take((size_t) m.hidden * 4),
                 o_times = take((size_t) (1 + 7 * m.n_layers + 3 + 80) * 8);
    pl->ws_bytes = off;
    if (cudaMalloc(&pl->ws, pl->ws_bytes) != cudaSuccess) { cudaGetLastError(); delete pl; return fail((int) cudaErrorMemoryAllocation); }
    cudaMemset(pl->ws, 0, pl->ws_bytes);  // barrier counter / argmax key / the padding of att_q start at zero
    cudaDeviceSynchronize();               // (the first launch may be on a non-blocking stream)
    uint8_t * w = (uint8_t *) pl->ws;
    P.sync = (unsigned long long *) (w + o_sync);
    P.q = (float *) (w + o_q); P.kbuf = (float *) (w + o_k); P.vbuf = (float *) (w + o_v); P.scores = (float *) (w + o_sc);
    P.part = (float2 *) (w + o_part); P.partial = (float *) (w + o_partial); P.att_q = w + o_attq; P.gate = (float *) (w + o_gate);
    P.x = (float *) (w + o_x);  // default residual stream (the caller may pass its own)
    P.s_stride = s_stride; P.max_tiles = max_tiles;
    P.times = mk_env_int("B200_MK_TIMES", 0) ? (long long *) (w + o_times) : nullptr;

    // ---- phase table (pro_x / the residual y / the logits pointer are filled in by plan_upload)
    auto mat = [](MKPhase & ph, int i, const void * W, float * y, const float * bias, int mm) { ph.W[i] = (const uint8_t *) W; ph.y[i] = y; ph.bias[i] = bias; ph.m[i] = mm; };
    for (int l = 0; l < m.n_layers; ++l) {
        const DecodeLayer & L = m.layers[l];
        if (!L.wq || !L.wk || !L.wv || !L.wo || !L.wgate || !L.wup || !L.wdown || !L.attn_norm || !L.ffn_norm) { decode_plan_destroy(pl); return fail(B200_ERR_ARG); }
        MKPhase a{};  // q/k/v = W . Q(rms_norm(x) * attn_norm) (+ bias)
        mat(a, 0, L.wq, P.q, L.bq, m.hidden); mat(a, 1, L.wk, P.kbuf, L.bk, kv_hidden); mat(a, 2, L.wv, P.vbuf, L.bv, kv_hidden);
        a.nmat = 3; a.paired = 0; a.k = m.hidden; a.m_total = m.hidden + 2 * kv_hidden; a.pro = MK_PRO_NORM; a.epi = MK_EPI_STORE; a.pro_w = L.attn_norm;
        MKPhase b{};  // x += W_o . att
        mat(b, 0, L.wo, nullptr, nullptr, m.hidden);
        b.nmat = 1; b.k = m.hidden; b.m_total = m.hidden; b.pro = MK_PRO_COPY; b.epi = MK_EPI_RESID; b.pro_q = P.att_q;
        MKPhase c{};  // gate = silu(W_g . n) * (W_u . n),  n = Q(rms_norm(x) * ffn_norm)
        mat(c, 0, L.wgate, P.gate, nullptr, m.ffn); mat(c, 1, L.wup, nullptr, nullptr, m.ffn);
        c.nmat = 2; c.paired = 1; c.k = m.hidden; c.m_total = m.ffn; c.pro = MK_PRO_NORM; c.epi = MK_EPI_SWIGLU; c.pro_w = L.ffn_norm;
        MKPhase d{};  // x += W_down . Q(gate)
        mat(d, 0, L.wdown, nullptr, nullptr, m.hidden);
        d.nmat = 1; d.k = m.ffn; d.m_total = m.hidden; d.pro = MK_PRO_QUANT; d.epi = MK_EPI_RESID; d.pro_x = P.gate;
        pl->h_phases.push_back(a); pl->h_phases.push_back(b); pl->h_phases.push_back(c); pl->h_phases.push_back(d);
        pl->h_kv.push_back(MKLayerKV{(__half *) L.k_cache, (__half *) L.v_cache});
    }
    if (m.lm_head) {
        MKPhase h{};
        mat(h, 0, m.lm_head, nullptr, nullptr, m.vocab);
        h.nmat = 1; h.k = m.hidden; h.m_total = m.vocab; h.pro = MK_PRO_NORM; h.epi = MK_EPI_STORE; h.pro_w = m.final_norm;
        pl->h_phases.push_back(h);
    }
    P.n_phases = (int) pl->h_phases.size();
    if (cudaMalloc((void **) &pl->d_phases, sizeof(MKPhase) * std::max<size_t>(1, pl->h_phases.size())) != cudaSuccess ||
        cudaMalloc((void **) &pl->d_kv, sizeof(MKLayerKV) * std::max<size_t>(1, pl->h_kv.size())) != cudaSuccess) {
        cudaGetLastError(); decode_plan_destroy(pl); return fail((int) cudaErrorMemoryAllocation);
    }
    P.phases = pl->d_phases; P.kv = pl->d_kv;

    // ---- shared-memory layout: [activation column / attention scratch][mbarriers][8 warps x stages x stage]
    P.l2_ahead = mk_env_int("B200_MK_L2AHEAD", 0);
    P.kv_prefetch = mk_env_int("B200_MK_KVPREFETCH", 1);
    P.ks = mk_env_int("B200_MK_KS", 16);
    P.stages = mk_env_int("B200_MK_STAGES", 2);
    if (P.ks < 1) P.ks = 1;
    if (P.stages < 1) P.stages = 1;
    const size_t act_need = std::max(qact_col_bytes(m.wtype, m.hidden), qact_col_bytes(m.wtype, m.ffn));
    const size_t attn_need = std::max((size_t) m.heads * m.head_dim * 2 + (size_t) m.head_dim * 4, (size_t) 8 * MK_PV_SP * 2 + 128);
    P.merged_tail = ((m.heads / m.kv_heads) * m.head_dim) % 256 == 0 && mk_env_int("B200_MK_MERGED_TAIL", 1) ? 1 : 0;
    if (m.kv_heads > 100) P.merged_tail = 0;
    {   // both tables in shared memory when they leave room for a 2-stage ring
        const size_t tb = ((size_t) P.n_phases * sizeof(MKPhase) + (size_t) m.n_layers * sizeof(MKLayerKV) + 127) & ~(size_t) 127;
        P.tab_bytes = (tb <= 48 * 1024 && mk_env_int("B200_MK_TAB_SMEM", 1)) ? (uint32_t) tb : 0;
    }
    P.act_bytes = (uint32_t) ((std::max(act_need, attn_need) + 127) & ~(size_t) 127);
    auto smem_for = [&]() {
        P.stage_bytes = (uint32_t) (MK_RG * P.ks * unit);
        return (size_t) P.tab_bytes + (size_t) P.act_bytes + (size_t) al16((int64_t) MK_WARPS * P.stages * 8) + (size_t) MK_WARPS * P.stages * P.stage_bytes;
    };
    const size_t limit = 224 * 1024;  // + ~1.5 KB static
    while (smem_for() > limit && P.stages > 2) P.stages--;
    while (smem_for() > limit && P.ks > 1) P.ks = (P.ks + 1) / 2;
    if (smem_for() > limit) { decode_plan_destroy(pl); return fail(B200_ERR_UNSUPPORTED); }
    pl->smem = smem_for();
    if (err) *err = 0;
    return pl;
}

// the KV cache of a layer moved (plugin: the host application owns the cache tensors)
int decode_plan_set_kv(void * h, int layer, void * k_cache, void * v_cache) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl || layer < 0 || layer >= (int) pl->h_kv.size()) return B200_ERR_ARG;
    if (pl->h_kv[layer].kc != (__half *) k_cache || pl->h_kv[layer].vc != (__half *) v_cache) {
        pl->h_kv[layer] = MKLayerKV{(__half *) k_cache, (__half *) v_cache};
        pl->dirty = true;
    }
    return B200_OK;
}

// status word of the last launches: 0 = fine, 1 = a grid barrier timed out (the plan is then reset)
int decode_plan_status(void * h, cudaStream_t st) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl) return B200_ERR_ARG;
    unsigned long long s[4] = {0, 0, 0, 0};
    if (cudaMemcpyAsync(s, pl->prm.sync, sizeof(s), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return (int) cudaGetLastError();
    if (s[3] != 0) { cudaMemsetAsync(pl->prm.sync, 0, 1024, st); cudaStreamSynchronize(st); }
    return (int) s[3];
}

int decode_step(void * h, const DecodeIO & io, cudaStream_t st) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl) return B200_ERR_ARG;
    MKParams prm = pl->prm;
    float * x = io.x ? io.x : pl->prm.x;
    if (prm.has_head && !io.logits) return B200_ERR_ARG;
    if (!io.pos) return B200_ERR_ARG;
    if (io.n_kv > pl->max_ctx) return B200_ERR_ARG;
    if (pl->dirty || x != pl->up_x || io.logits != pl->up_logits) {
        // point the table at this call's residual stream / logits buffer, then (re)upload both tables
        for (size_t i = 0; i < pl->h_phases.size(); ++i) {
            MKPhase & ph = pl->h_phases[i];
            const bool head = prm.has_head && i + 1 == pl->h_phases.size();
            const int kind = head ? 4 : (int) (i % 4);
            if (kind == 0 || kind == 2 || kind == 4) ph.pro_x = x;
            if (kind == 1 || kind == 3) ph.y[0] = x;
            if (kind == 4) ph.y[0] = io.logits;
        }
        // in-stream copies (ordered behind a launch that may still be reading the tables); the source is pageable memory, so the runtime
        // has taken its copy by the time the call returns and the host vectors may be edited again
        if (!pl->h_phases.empty()) B200_CUDA_CHECK(cudaMemcpyAsync(pl->d_phases, pl->h_phases.data(), sizeof(MKPhase) * pl->h_phases.size(), cudaMemcpyHostToDevice, st));
        if (!pl->h_kv.empty()) B200_CUDA_CHECK(cudaMemcpyAsync(pl->d_kv, pl->h_kv.data(), sizeof(MKLayerKV) * pl->h_kv.size(), cudaMemcpyHostToDevice, st));
        pl->dirty = false; pl->up_x = x; pl->up_logits = io.logits;
    }
    prm.x = x;
    if (!io.tok) prm.embed = nullptr;  // the caller provides the hidden state in x (a later layer shard)
    prm.tok = io.tok; prm.pos = io.pos; prm.n_kv_arg = io.n_kv; prm.v_col_arg = io.v_col;
    prm.next_tok = io.next_tok; prm.flags = io.flags;
    prm.wait_flag = (const unsigned long long *) io.wait_flag; prm.wait_offset = io.wait_offset;
    prm.send_x = io.send_x; prm.send_flag = (unsigned long long *) io.send_flag; prm.send_tok = io.send_tok;
    const int last = 1 + 7 * prm.n_layers + 2;
    prm.step_begin = io.step_begin > 0 ? io.step_begin : 0;
    prm.step_end = (io.step_end > 0 && io.step_end < last) ? io.step_end : last;
    if (prm.step_begin >= prm.step_end) return B200_OK;
    return (int) mk_launch(*pl, prm, st);
}

// profiling aid: SM-clock stamps of the last whole-token launch (needs B200_MK_TIMES=1 at plan creation); returns the number of stamps
int decode_plan_times(void * h, long long * out, int cap, cudaStream_t st) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl || !pl->prm.times) return 0;
    const int n = 1 + 7 * pl->prm.n_layers + 3 + 80;
    const int c = n < cap ? n : cap;
    if (cudaMemcpyAsync(out, pl->prm.times, (size_t) c * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) return 0;
    return c;
}

int decode_plan_info(void * h, int * grid, int * smem_bytes, int * n_steps, int * stages, int * ks) {
    DecodePlan * pl = (DecodePlan *) h;
    if (!pl) return B200_ERR_ARG;
    if (grid) *grid = pl->grid;
    if (smem_bytes) *smem_bytes = (int) pl->smem;
    if (n_steps) *n_steps = 1 + 7 * pl->prm.n_layers + 2;
    if (stages) *stages = pl->prm.stages;
    if (ks) *ks = pl->prm.ks;
    return B200_OK;
}

}  // namespace b200
