// normquant.cuh — RMSNorm + activation quantization INTO SHARED MEMORY, as the prologue of a consumer GEMV (gemv.cu, FX bit 0).
//
// Round-1 profile (profiles/r01f_*): the single-CTA add_rmsnorm_quant_kernel is 2 x 3.45 us of a 55 us layer and sits on the serial
// chain between two GEMVs.  With this prologue every GEMV CTA recomputes  y = rms_norm(x) * w  and its Q8_K / Q8_0 quantization
// straight into its activation column (k <= 20480 floats: 16-80 KB from L2, 16-80 elements per thread) while its weight stream is
// already in flight, and the residual add moves into the producer GEMV's epilogue (FX bit 1) — two launches per layer disappear.
// Arithmetic is the stand-alone kernel's (fused.cu add_rmsnorm_quant_kernel; reference ops.cpp:3710-3758, ggml-quants.c:2555-2592,
// arch/x86/quants.c:290-384): same fp32 expressions, bit-identical codes for identical y; only the block-reduction order of the
// sum of squares differs (256 instead of 1024 threads).
// OPT-IN in round 1 (B200 time ran out before it could be measured): DecodeSession(fused=2), b200_gemv_fused().
#pragma once
#include "actlayout.cuh"
#include "common.cuh"

namespace b200 {

struct NormQuantSmem {
    float red[32];
    unsigned long long keys[16];
    float bmax[8];
};

__device__ __forceinline__ float nq_block_sum(float v, float * red) {
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

// x, w: k floats in global memory (x is written by the predecessor kernel: plain loads, after griddepcontrol.wait).
// base: the CTA's shared-memory activation column (layout actlayout.cuh).  blockDim.x must be a multiple of 64 and <= 512.
template <bool Q8K>
__device__ __forceinline__ void norm_quant_to_smem(const float * x, const float * w, float eps, int64_t k, uint8_t * base, NormQuantSmem & sm) {
    const int t = threadIdx.x, T = blockDim.x;
    const int64_t step = 4 * (int64_t) T;
    float ss = 0.0f;
    for (int64_t e = 4 * (int64_t) t; e < k; e += step) {
        const float4 a = *reinterpret_cast<const float4 *>(x + e);
        ss = fmaf(a.x, a.x, ss); ss = fmaf(a.y, a.y, ss); ss = fmaf(a.z, a.z, ss); ss = fmaf(a.w, a.w, ss);
    }
    ss = nq_block_sum(ss, sm.red);
    const float mean = ss / (float) k;
    const float scale = 1.0f / sqrtf(mean + eps);
    const ActLayout L = act_layout(Q8K, k);
    const int lane = t & 31, warp = t >> 5;
    const int64_t k_pad = (k + step - 1) / step * step;  // every thread runs the same number of rounds (block barriers inside)
    for (int64_t e = 4 * (int64_t) t; e < k_pad; e += step) {
        const bool on = e < k;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (on) {
            const float4 a = *reinterpret_cast<const float4 *>(x + e);
            const float4 ww = *reinterpret_cast<const float4 *>(w + e);
            v[0] = (a.x * scale) * ww.x; v[1] = (a.y * scale) * ww.y; v[2] = (a.z * scale) * ww.z; v[3] = (a.w * scale) * ww.w;
        }
        if (Q8K) {
            float * dd = (float *) (base + L.d_off);
            int16_t * bs = (int16_t *) (base + L.bs_off);
            // first-occurrence argmax |v| over the 256-element block (= 64 consecutive threads = 2 warps)
            unsigned long long key = 0ull;
            if (on) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const unsigned long long kk = ((unsigned long long) __float_as_uint(fabsf(v[i])) << 32) | (unsigned) (255 - (4 * (t & 63) + i));
                    key = kk > key ? kk : key;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
                key = other > key ? other : key;
            }
            __syncthreads();
            if (lane == 0) sm.keys[warp] = key;
            __syncthreads();
            const unsigned long long k2 = sm.keys[warp ^ 1];
            key = k2 > key ? k2 : key;
            const int idx = 255 - (int) (key & 0xffffffffu);
            if (on && (idx >> 2) == (t & 63)) sm.bmax[t >> 6] = v[idx & 3];
            __syncthreads();
            if (on) {
                const float mx = sm.bmax[t >> 6];
                int q[4] = {0, 0, 0, 0};
                float dv = 0.0f;
                if (mx != 0.0f) {
                    const float iscale = __fdiv_rn(-127.f, mx);
#pragma unroll
                    for (int i = 0; i < 4; ++i) q[i] = min(127, __float2int_rn(__fmul_rn(iscale, v[i])));
                    dv = __fdiv_rn(1.0f, iscale);
                }
                const uint32_t packed = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
                *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e)) = packed;
                int s = q[0] + q[1] + q[2] + q[3];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                if ((t & 7) == 0) bs[e >> 5] = (int16_t) s;
                if ((t & 63) == 0) dd[e >> 8] = dv;
            }
        } else {
            float * dd = (float *) (base + L.d_off);
            int * bs = (int *) (base + L.bs_off);
            float amax = on ? fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3]))) : 0.0f;
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float dv = __fdiv_rn(amax, 127.f);
            const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
            int q[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) q[i] = __float2int_rn(__fmul_rn(v[i], id));
            int s = q[0] + q[1] + q[2] + q[3];
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (on) {
                const uint32_t packed = (uint32_t) (q[0] & 0xff) | ((uint32_t) (q[1] & 0xff) << 8) | ((uint32_t) (q[2] & 0xff) << 16) | ((uint32_t) (q[3] & 0xff) << 24);
                *reinterpret_cast<uint32_t *>(base + act_qs_off_q80(e)) = packed;
                if ((t & 7) == 0) { dd[e >> 5] = __half2float(__float2half_rn(dv)); bs[e >> 5] = s; }
            }
        }
    }
    if (Q8K) {  // zero the padding of the last (partial) 1024-element group so the main loop can read it blindly
        for (int64_t e = k + 4 * (int64_t) t; e < L.qs_bytes; e += step) *reinterpret_cast<uint32_t *>(base + act_qs_off_q8k(e)) = 0u;
    }
}

}  // namespace b200
