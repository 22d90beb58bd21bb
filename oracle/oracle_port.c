/*
 * oracle_port.c — CPU restatement ("port") of the reference's arithmetic on the
 * quantized-matmul decode hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared object.  The product
 * (chatllm.cpp_b200/csrc) never links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_pin.py checks every function here
 *   (a) against the reference itself compiled from /root/reference into
 *       oracle/_ref/lib (libggml-base.so: *_ref quantizers / dequantizers;
 *       libggml-cpu-*.so: the x86 quantizers and vec_dot kernels the CPU
 *       backend really runs), and
 *   (b) against committed fixtures in tests/golden/ produced by those same
 *       reference libraries (tests/golden/make_golden.py).
 * The reference ships no golden vectors of its own.
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Plain scalar C, compiled with -ffp-contract=off so no FMA
 * contraction changes the fp32 results.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define QK4_0 32
#define QK8_0 32
#define QK_K 256

/* ---- block layouts: ggml/src/ggml-common.h:170-176, :219-224, :288-306, :338-344 */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[QK4_0 / 2]; } oq_block_q4_0;                    /* 18 B */
typedef struct { uint16_t d; int8_t qs[QK8_0]; } oq_block_q8_0;                         /* 34 B */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qs[QK_K / 2]; } oq_block_q4_K; /* 144 B */
typedef struct { float d; int8_t qs[QK_K]; int16_t bsums[QK_K / 16]; } oq_block_q8_K;  /* 292 B */
#pragma pack(pop)

enum { OQ_TYPE_F32 = 0, OQ_TYPE_F16 = 1, OQ_TYPE_Q4_0 = 2, OQ_TYPE_Q8_0 = 8, OQ_TYPE_Q4_K = 12 }; /* ggml/include/ggml.h:389-405 */

/* ---- fp16 <-> fp32, IEEE round-to-nearest-even (ggml/src/ggml-impl.h ggml_compute_fp16_to_fp32 /
 *      ggml_compute_fp32_to_fp16; on x86 the reference uses F16C, which is the same IEEE conversion) */
float oq_fp16_to_fp32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1f;
    uint32_t man = h & 0x3ffu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            man &= 0x3ffu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112) << 23) | (man << 13);
    }
    float f; memcpy(&f, &bits, 4); return f;
}

uint16_t oq_fp32_to_fp16(float f) {
    uint32_t x; memcpy(&x, &f, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) { /* inf / nan */
        return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? (0x200u | ((ax >> 13) & 0x3ffu)) : 0));
    }
    if (ax >= 0x477ff000u) { /* rounds to >= 65520 -> inf */
        return (uint16_t)(sign | 0x7c00u);
    }
    if (ax < 0x38800000u) { /* subnormal half or zero */
        if (ax < 0x33000000u) return (uint16_t)sign; /* < 2^-25 -> 0 (2^-25 exactly ties to even = 0) */
        int e = (int)(ax >> 23);                   /* biased exponent, 102..112 */
        uint32_t man = (ax & 0x7fffffu) | 0x800000u;
        int shift = 126 - e;                        /* 14..24 */
        uint32_t q = man >> shift;
        uint32_t rem = man & ((1u << shift) - 1);
        uint32_t half = 1u << (shift - 1);
        if (rem > half || (rem == half && (q & 1))) q++;
        return (uint16_t)(sign | q);
    }
    uint32_t e = (ax >> 23) - 112;
    uint32_t man = ax & 0x7fffffu;
    uint32_t q = (e << 10) | (man >> 13);
    uint32_t rem = man & 0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (q & 1))) q++;
    return (uint16_t)(sign | q);
}

/* ============================ quantizers =============================== */

/* ggml/src/ggml-quants.c:36-70  quantize_row_q4_0_ref */
void oq_quantize_row_q4_0_ref(const float *x, void *vy, int64_t k) {
    oq_block_q4_0 *y = (oq_block_q4_0 *)vy;
    const int nb = (int)(k / QK4_0);
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f, max = 0.0f;
        for (int j = 0; j < QK4_0; j++) {
            const float v = x[i * QK4_0 + j];
            if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
        }
        const float d = max / -8;
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = oq_fp32_to_fp16(d);
        for (int j = 0; j < QK4_0 / 2; ++j) {
            const float x0 = x[i * QK4_0 + 0 + j] * id;
            const float x1 = x[i * QK4_0 + QK4_0 / 2 + j] * id;
            int8_t a0 = (int8_t)(x0 + 8.5f), a1 = (int8_t)(x1 + 8.5f);
            const uint8_t xi0 = a0 < 15 ? (uint8_t)a0 : 15;
            const uint8_t xi1 = a1 < 15 ? (uint8_t)a1 : 15;
            y[i].qs[j] = xi0 | (uint8_t)(xi1 << 4);
        }
    }
}

/* ggml/src/ggml-quants.c:199-222  quantize_row_q8_0_ref  (roundf = half away from zero, id = 1/d) */
void oq_quantize_row_q8_0_ref(const float *x, void *vy, int64_t k) {
    oq_block_q8_0 *y = (oq_block_q8_0 *)vy;
    const int nb = (int)(k / QK8_0);
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int j = 0; j < QK8_0; j++) { const float v = fabsf(x[i * QK8_0 + j]); if (v > amax) amax = v; }
        const float d = amax / ((1 << 7) - 1);
        const float id = d ? 1.0f / d : 0.0f;
        y[i].d = oq_fp32_to_fp16(d);
        for (int j = 0; j < QK8_0; ++j) y[i].qs[j] = (int8_t)roundf(x[i * QK8_0 + j] * id);
    }
}

/* ggml/src/ggml-cpu/arch/x86/quants.c:290-384  quantize_row_q8_0 (AVX/AVX2 build: what the CPU
 * backend runs on x86).  Differs from the _ref version in two places: id = 127/amax (not 1/d), and
 * round-to-nearest-EVEN (_mm256_round_ps NEAREST) instead of roundf. */
void oq_quantize_row_q8_0_x86(const float *x, void *vy, int64_t k) {
    oq_block_q8_0 *y = (oq_block_q8_0 *)vy;
    const int nb = (int)(k / QK8_0);
    for (int i = 0; i < nb; i++) {
        float amax = 0.0f;
        for (int j = 0; j < QK8_0; j++) { const float v = fabsf(x[i * QK8_0 + j]); if (v > amax) amax = v; }
        const float d = amax / 127.f;
        y[i].d = oq_fp32_to_fp16(d);
        const float id = (amax != 0.0f) ? 127.f / amax : 0.0f;
        for (int j = 0; j < QK8_0; ++j) y[i].qs[j] = (int8_t)(int)nearbyintf(x[i * QK8_0 + j] * id);
    }
}

/* ggml/src/ggml-quants.c:444-449  nearest_int (RNE via the 1.5*2^23 magic constant) */
static inline int oq_nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i; memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* ggml/src/ggml-quants.c:2555-2592  quantize_row_q8_K_ref  (x86 quantize_row_q8_K forwards to it:
 * ggml/src/ggml-cpu/arch/x86/quants.c:493-495).  bsums are left untouched for an all-zero block, as
 * in the reference; we zero them so the port is deterministic (d == 0 makes them irrelevant). */
void oq_quantize_row_q8_K_ref(const float *x, void *vy, int64_t k) {
    oq_block_q8_K *y = (oq_block_q8_K *)vy;
    const int64_t nb = k / QK_K;
    for (int64_t i = 0; i < nb; i++) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; ++j) {
            float ax = fabsf(x[j]);
            if (ax > amax) { amax = ax; max = x[j]; }
        }
        if (!amax) {
            y[i].d = 0;
            memset(y[i].qs, 0, QK_K);
            memset(y[i].bsums, 0, sizeof(y[i].bsums));
            x += QK_K;
            continue;
        }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; ++j) {
            int v = oq_nearest_int(iscale * x[j]);
            y[i].qs[j] = (int8_t)(v < 127 ? v : 127);
        }
        for (int j = 0; j < QK_K / 16; ++j) {
            int sum = 0;
            for (int ii = 0; ii < 16; ++ii) sum += y[i].qs[j * 16 + ii];
            y[i].bsums[j] = (int16_t)sum;
        }
        y[i].d = 1 / iscale;
        x += QK_K;
    }
}

/* ============================ dequantizers ============================= */

/* ggml/src/ggml-quants.c:307-325 */
void oq_dequantize_row_q4_0(const void *vx, float *y, int64_t k) {
    const oq_block_q4_0 *x = (const oq_block_q4_0 *)vx;
    const int nb = (int)(k / QK4_0);
    for (int i = 0; i < nb; i++) {
        const float d = oq_fp16_to_fp32(x[i].d);
        for (int j = 0; j < QK4_0 / 2; ++j) {
            const int x0 = (x[i].qs[j] & 0x0F) - 8;
            const int x1 = (x[i].qs[j] >> 4) - 8;
            y[i * QK4_0 + j + 0] = x0 * d;
            y[i * QK4_0 + j + QK4_0 / 2] = x1 * d;
        }
    }
}

/* ggml/src/ggml-quants.c:401-414 */
void oq_dequantize_row_q8_0(const void *vx, float *y, int64_t k) {
    const oq_block_q8_0 *x = (const oq_block_q8_0 *)vx;
    const int nb = (int)(k / QK8_0);
    for (int i = 0; i < nb; i++) {
        const float d = oq_fp16_to_fp32(x[i].d);
        for (int j = 0; j < QK8_0; ++j) y[i * QK8_0 + j] = x[i].qs[j] * d;
    }
}

/* ggml/src/ggml-quants.c:703-711 */
static inline void oq_get_scale_min_k4(int j, const uint8_t *q, uint8_t *d, uint8_t *m) {
    if (j < 4) {
        *d = q[j] & 63; *m = q[j + 4] & 63;
    } else {
        *d = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4);
        *m = (q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4);
    }
}

/* ggml/src/ggml-quants.c:1352-1373 */
void oq_dequantize_row_q4_K(const void *vx, float *y, int64_t k) {
    const oq_block_q4_K *x = (const oq_block_q4_K *)vx;
    const int nb = (int)(k / QK_K);
    for (int i = 0; i < nb; i++) {
        const uint8_t *q = x[i].qs;
        const float d = oq_fp16_to_fp32(x[i].d);
        const float min = oq_fp16_to_fp32(x[i].dmin);
        int is = 0;
        uint8_t sc, m;
        for (int j = 0; j < QK_K; j += 64) {
            oq_get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
            const float d1 = d * sc; const float m1 = min * m;
            oq_get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
            const float d2 = d * sc; const float m2 = min * m;
            for (int l = 0; l < 32; ++l) *y++ = d1 * (q[l] & 0xF) - m1;
            for (int l = 0; l < 32; ++l) *y++ = d2 * (q[l] >> 4) - m2;
            q += 32; is += 2;
        }
    }
}

/* ============================ dot products ============================= */

/* ggml/src/ggml-cpu/quants.c:115-148  ggml_vec_dot_q4_0_q8_0_generic */
float oq_vec_dot_q4_0_q8_0(int n, const void *vx, const void *vy) {
    const oq_block_q4_0 *x = (const oq_block_q4_0 *)vx;
    const oq_block_q8_0 *y = (const oq_block_q8_0 *)vy;
    const int nb = n / QK8_0;
    float sumf = 0;
    for (int ib = 0; ib < nb; ++ib) {
        int sumi0 = 0, sumi1 = 0;
        for (int j = 0; j < QK8_0 / 2; ++j) {
            const int v0 = (x[ib].qs[j] & 0x0F) - 8;
            const int v1 = (x[ib].qs[j] >> 4) - 8;
            sumi0 += v0 * y[ib].qs[j];
            sumi1 += v1 * y[ib].qs[j + QK8_0 / 2];
        }
        int sumi = sumi0 + sumi1;
        sumf += sumi * oq_fp16_to_fp32(x[ib].d) * oq_fp16_to_fp32(y[ib].d);
    }
    return sumf;
}

/* ggml/src/ggml-cpu/quants.c:305-333  ggml_vec_dot_q8_0_q8_0_generic */
float oq_vec_dot_q8_0_q8_0(int n, const void *vx, const void *vy) {
    const oq_block_q8_0 *x = (const oq_block_q8_0 *)vx;
    const oq_block_q8_0 *y = (const oq_block_q8_0 *)vy;
    const int nb = n / QK8_0;
    float sumf = 0;
    for (int ib = 0; ib < nb; ++ib) {
        int sumi = 0;
        for (int j = 0; j < QK8_0; j++) sumi += x[ib].qs[j] * y[ib].qs[j];
        sumf += sumi * (oq_fp16_to_fp32(x[ib].d) * oq_fp16_to_fp32(y[ib].d));
    }
    return sumf;
}

/* ggml/src/ggml-cpu/quants.c:550-625  ggml_vec_dot_q4_K_q8_K_generic (same 8-lane fp32 association) */
float oq_vec_dot_q4_K_q8_K(int n, const void *vx, const void *vy) {
    const oq_block_q4_K *x = (const oq_block_q4_K *)vx;
    const oq_block_q8_K *y = (const oq_block_q8_K *)vy;
    const int nb = n / QK_K;
    static const uint32_t kmask1 = 0x3f3f3f3f, kmask2 = 0x0f0f0f0f, kmask3 = 0x03030303;
    uint32_t utmp[4];
    const uint8_t *scales = (const uint8_t *)&utmp[0];
    const uint8_t *mins = (const uint8_t *)&utmp[2];
    int8_t aux8[QK_K];
    int16_t aux16[8];
    float sums[8];
    int32_t aux32[8];
    memset(sums, 0, sizeof(sums));
    float sumf = 0;
    for (int i = 0; i < nb; ++i) {
        const uint8_t *q4 = x[i].qs;
        const int8_t *q8 = y[i].qs;
        memset(aux32, 0, sizeof(aux32));
        int8_t *a = aux8;
        for (int j = 0; j < QK_K / 64; ++j) {
            for (int l = 0; l < 32; ++l) a[l] = (int8_t)(q4[l] & 0xF);
            a += 32;
            for (int l = 0; l < 32; ++l) a[l] = (int8_t)(q4[l] >> 4);
            a += 32; q4 += 32;
        }
        memcpy(utmp, x[i].scales, 12);
        utmp[3] = ((utmp[2] >> 4) & kmask2) | (((utmp[1] >> 6) & kmask3) << 4);
        const uint32_t uaux = utmp[1] & kmask1;
        utmp[1] = (utmp[2] & kmask2) | (((utmp[0] >> 6) & kmask3) << 4);
        utmp[2] = uaux;
        utmp[0] &= kmask1;
        int sumi = 0;
        for (int j = 0; j < QK_K / 16; ++j) sumi += y[i].bsums[j] * mins[j / 2];
        a = aux8;
        int is = 0;
        for (int j = 0; j < QK_K / 32; ++j) {
            int32_t scale = scales[is++];
            for (int r = 0; r < 4; ++r) {
                for (int l = 0; l < 8; ++l) aux16[l] = (int16_t)(q8[l] * a[l]);
                for (int l = 0; l < 8; ++l) aux32[l] += scale * aux16[l];
                q8 += 8; a += 8;
            }
        }
        const float d = oq_fp16_to_fp32(x[i].d) * y[i].d;
        for (int l = 0; l < 8; ++l) sums[l] += d * aux32[l];
        const float dmin = oq_fp16_to_fp32(x[i].dmin) * y[i].d;
        sumf -= dmin * sumi;
    }
    for (int l = 0; l < 8; ++l) sumf += sums[l];
    return sumf;
}

/* ============================ matmul =================================== */

static size_t oq_row_size(int type, int64_t k) {
    switch (type) {
        case OQ_TYPE_Q4_0: return (size_t)(k / 32) * 18;
        case OQ_TYPE_Q8_0: return (size_t)(k / 32) * 34;
        case OQ_TYPE_Q4_K: return (size_t)(k / 256) * 144;
        case OQ_TYPE_F16:  return (size_t)k * 2;
        case OQ_TYPE_F32:  return (size_t)k * 4;
    }
    return 0;
}
size_t oq_type_row_size(int type, int64_t k) { return oq_row_size(type, k); }

/* ggml/src/ggml-cpu/ggml-cpu.c:1229-1421 ggml_compute_forward_mul_mat + :1139-1227 one_chunk:
 * dst[m, n] = W[m, k] . X[k, n]; each X column is first converted to the weight type's vec_dot_type
 * (:1291-1326; table :207-308: Q4_0/Q8_0 -> Q8_0, Q4_K -> Q8_K, F16 -> F16) and every dst element is
 * one vec_dot.  q8_0_variant: 0 = *_ref quantizer, 1 = the x86 quantizer the CPU backend runs.
 * W rows are contiguous with stride w_row_stride bytes; X columns contiguous, stride k floats;
 * Y column stride m floats. */
int oq_mul_mat(int type, const void *w, size_t w_row_stride, int64_t k, int64_t m,
               const float *x, int64_t n, float *y, int q8_0_variant) {
    size_t qrow;
    if (type == OQ_TYPE_Q4_0 || type == OQ_TYPE_Q8_0) qrow = (size_t)(k / 32) * sizeof(oq_block_q8_0);
    else if (type == OQ_TYPE_Q4_K) qrow = (size_t)(k / 256) * sizeof(oq_block_q8_K);
    else if (type == OQ_TYPE_F16) qrow = (size_t)k * 2;
    else return -1;
    uint8_t *wdata = (uint8_t *)malloc(qrow * (size_t)n + 64);
    if (!wdata) return -2;
    for (int64_t j = 0; j < n; ++j) {
        void *q = wdata + qrow * j;
        if (type == OQ_TYPE_Q4_K) oq_quantize_row_q8_K_ref(x + j * k, q, k);
        else if (type == OQ_TYPE_F16) { uint16_t *h = (uint16_t *)q; for (int64_t i = 0; i < k; ++i) h[i] = oq_fp32_to_fp16(x[j * k + i]); }
        else if (q8_0_variant) oq_quantize_row_q8_0_x86(x + j * k, q, k);
        else oq_quantize_row_q8_0_ref(x + j * k, q, k);
    }
    for (int64_t j = 0; j < n; ++j) {
        const void *q = wdata + qrow * j;
        for (int64_t i = 0; i < m; ++i) {
            const void *row = (const uint8_t *)w + w_row_stride * i;
            float s;
            if (type == OQ_TYPE_Q4_0) s = oq_vec_dot_q4_0_q8_0((int)k, row, q);
            else if (type == OQ_TYPE_Q8_0) s = oq_vec_dot_q8_0_q8_0((int)k, row, q);
            else if (type == OQ_TYPE_Q4_K) s = oq_vec_dot_q4_K_q8_K((int)k, row, q);
            else { /* ggml/src/ggml-cpu/vec.cpp:264 ggml_vec_dot_f16 (scalar branch: double accumulate) */
                const uint16_t *a = (const uint16_t *)row, *b = (const uint16_t *)q;
                double sum = 0.0;
                for (int64_t t = 0; t < k; ++t) sum += (double)(oq_fp16_to_fp32(a[t]) * oq_fp16_to_fp32(b[t]));
                s = (float)sum;
            }
            y[j * m + i] = s;
        }
    }
    free(wdata);
    return 0;
}

/* ggml/src/ggml.c:3225-3240 (semantics) + ggml/src/ggml-cpu/ggml-cpu.c:1503-1700 ggml_compute_forward_mul_mat_id:
 *   as  [k, m, n_expert]  (expert stride = m rows), b [k, nb1, n_tokens] with nb1 == n_used or 1 (broadcast),
 *   ids [n_used, n_tokens] i32,  dst [m, n_used, n_tokens]:
 *   dst[:, e, t] = as[:, :, ids[e, t]] . b[:, e % nb1, t]
 * The CPU converts every src1 row to the weight type's vec_dot_type first (:1547-1580) and each dst element is one
 * vec_dot (one_chunk :1432-1501) — the same per-column arithmetic as oq_mul_mat, so this just routes columns to experts.
 * Returns -3 for an expert index outside [0, n_expert) (the reference asserts, :1611). */
int oq_mul_mat_id(int type, const void *as, size_t w_row_stride, int64_t k, int64_t m, int64_t n_expert,
                  const float *b, int64_t nb1, const int32_t *ids, int64_t n_used, int64_t n_tokens,
                  float *dst, int q8_0_variant) {
    for (int64_t t = 0; t < n_tokens; ++t) {
        for (int64_t e = 0; e < n_used; ++e) {
            const int32_t ex = ids[t * n_used + e];
            if (ex < 0 || ex >= n_expert) return -3;
            const uint8_t *w = (const uint8_t *)as + (size_t)ex * (size_t)m * w_row_stride;
            const float *x = b + (t * nb1 + (e % nb1)) * k;
            int rc = oq_mul_mat(type, w, w_row_stride, k, m, x, 1, dst + (t * n_used + e) * m, q8_0_variant);
            if (rc) return rc;
        }
    }
    return 0;
}

/* ============================ glue ops ================================= */

/* ggml/src/ggml-cpu/ops.cpp:3710-3758  rms_norm (double sum, float mean), then the `mul` by weight that
 * RMSNorm::forward appends (src/layers.cpp:2216-2225).  w may be NULL (plain rms_norm). */
void oq_rms_norm(const float *x, const float *w, float *y, int64_t ne0, int64_t nrows, float eps) {
    for (int64_t r = 0; r < nrows; ++r) {
        const float *xr = x + r * ne0; float *yr = y + r * ne0;
        double sum = 0.0;
        for (int64_t i = 0; i < ne0; i++) sum += (double)(xr[i] * xr[i]);
        const float mean = (float)(sum / ne0);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < ne0; i++) { float v = xr[i] * scale; yr[i] = w ? v * w[i] : v; }
    }
}

/* ggml/src/ggml-cpu/ops.cpp:5225-5335  soft_max (scale, optional f32 mask, no ALiBi), scalar tail of
 * ggml_vec_soft_max_f32 (ggml/src/ggml-cpu/vec.cpp:547-612): expf, double sum. mask: [ne0] per row or NULL */
void oq_soft_max(const float *x, const float *mask, float *y, int64_t ne0, int64_t nrows, float scale) {
    float *wp = (float *)malloc(sizeof(float) * (size_t)ne0);
    for (int64_t r = 0; r < nrows; ++r) {
        const float *xr = x + r * ne0; float *yr = y + r * ne0;
        for (int64_t i = 0; i < ne0; ++i) { wp[i] = xr[i] * scale; if (mask) wp[i] += mask[r * ne0 + i]; }
        float max = -INFINITY;
        for (int64_t i = 0; i < ne0; ++i) if (wp[i] > max) max = wp[i];
        double sum = 0.0;
        for (int64_t i = 0; i < ne0; ++i) { float v = expf(wp[i] - max); sum += (double)v; yr[i] = v; }
        sum = 1.0 / sum;
        for (int64_t i = 0; i < ne0; ++i) yr[i] *= (float)sum;
    }
    free(wp);
}

/* ggml/src/ggml-cpu/ops.cpp diag_mask_inf: x[i0, i1] = -inf for i0 > n_past + i1 */
void oq_diag_mask_inf(float *x, int64_t ne0, int64_t ne1, int64_t nz, int n_past) {
    for (int64_t z = 0; z < nz; ++z)
        for (int64_t j = 0; j < ne1; ++j)
            for (int64_t i = n_past; i < ne0; ++i)
                if (i > n_past + j) x[(z * ne1 + j) * ne0 + i] = -INFINITY;
}

/* ggml/src/ggml.c ggml_rope_yarn_corr_dim / ggml_rope_yarn_corr_dims */
static float oq_rope_yarn_corr_dim(int n_dims, int n_ctx_orig, float n_rot, float base) {
    return n_dims * logf(n_ctx_orig / (n_rot * 2 * (float)M_PI)) / (2 * logf(base));
}
void oq_rope_yarn_corr_dims(int n_dims, int n_ctx_orig, float freq_base, float beta_fast, float beta_slow, float dims[2]) {
    float start = floorf(oq_rope_yarn_corr_dim(n_dims, n_ctx_orig, beta_fast, freq_base));
    float end = ceilf(oq_rope_yarn_corr_dim(n_dims, n_ctx_orig, beta_slow, freq_base));
    dims[0] = start > 0 ? start : 0;
    dims[1] = end < n_dims - 1 ? end : n_dims - 1;
}

/* ggml/src/ggml-cpu/ops.cpp:5587-5611 rope_yarn_ramp / rope_yarn */
static float oq_rope_yarn_ramp(const float low, const float high, const int i0) {
    float den = high - low; if (den < 0.001f) den = 0.001f;
    const float y = (i0 / 2 - low) / den;
    float c = y < 0 ? 0 : y; c = c > 1 ? 1 : c;
    return 1 - c;
}
static void oq_rope_yarn(float theta_extrap, float freq_scale, const float corr_dims[2], int64_t i0, float ext_factor,
                         float mscale, float *cos_theta, float *sin_theta) {
    float theta_interp = freq_scale * theta_extrap;
    float theta = theta_interp;
    if (ext_factor != 0.0f) {
        float ramp_mix = oq_rope_yarn_ramp(corr_dims[0], corr_dims[1], (int)i0) * ext_factor;
        theta = theta_interp * (1 - ramp_mix) + theta_extrap * ramp_mix;
        mscale *= 1.0f + 0.1f * logf(1.0f / freq_scale);
    }
    *cos_theta = cosf(theta) * mscale;
    *sin_theta = sinf(theta) * mscale;
}

/* ggml/src/ggml-cpu/ops.cpp:5613-5628 cache init (theta recurrence seeded with the position),
 * :5705-5718 rotate_pairs, :5720-5865 rope_flt.  x: [ne0, n_heads, n_tokens] f32 contiguous, in -> out.
 * mode 0 = NORMAL (adjacent pairs), 2 = NEOX (pairs i, i + n_dims/2).  freq_factors may be NULL. */
void oq_rope(const float *x, float *y, const int32_t *pos, const float *freq_factors, int64_t ne0, int64_t n_heads,
             int64_t n_tokens, int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale, float ext_factor,
             float attn_factor, float beta_fast, float beta_slow) {
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    float corr_dims[2];
    oq_rope_yarn_corr_dims(n_dims, n_ctx_orig, freq_base, beta_fast, beta_slow, corr_dims);
    float *cache = (float *)malloc(sizeof(float) * (size_t)ne0);
    for (int64_t t = 0; t < n_tokens; ++t) {
        float theta = (float)pos[t];
        for (int64_t i0 = 0; i0 < ne0; i0 += 2) {
            const float ff = freq_factors ? freq_factors[i0 / 2] : 1.0f;
            oq_rope_yarn(theta / ff, freq_scale, corr_dims, i0, ext_factor, attn_factor, &cache[i0], &cache[i0 + 1]);
            theta *= theta_scale;
        }
        for (int64_t h = 0; h < n_heads; ++h) {
            const float *src = x + (t * n_heads + h) * ne0;
            float *dst = y + (t * n_heads + h) * ne0;
            const int64_t n_offset = (mode == 0) ? 1 : n_dims / 2;
            const int scale = (mode == 0) ? 1 : 2;
            for (int64_t i0 = 0; i0 < n_dims; i0 += 2) {
                const int64_t ic = i0 / scale;
                const float c = cache[i0], s = cache[i0 + 1];
                const float x0 = src[ic], x1 = src[ic + n_offset];
                dst[ic] = x0 * c - x1 * s;
                dst[ic + n_offset] = x0 * s + x1 * c;
            }
            for (int64_t i0 = n_dims; i0 < ne0; ++i0) dst[i0] = src[i0];
        }
    }
    free(cache);
}

/* ggml/src/ggml-cpu/vec.h:1061 ggml_silu_f32 ; SwiGLU of BaseMLP::forward (src/layers.cpp:2475-2483):
 * y = silu(gate) * up */
void oq_silu_mul(const float *gate, const float *up, float *y, int64_t n) {
    for (int64_t i = 0; i < n; ++i) { float g = gate[i]; y[i] = (g / (1.0f + expf(-g))) * up[i]; }
}

/* ggml/src/ggml-cpu/ops.cpp:4820 get_rows on a quantized table: dequantize row ids[i] */
int oq_get_rows(int type, const void *table, int64_t k, const int32_t *ids, int64_t n, float *y) {
    const size_t rs = oq_row_size(type, k);
    for (int64_t i = 0; i < n; ++i) {
        const void *row = (const uint8_t *)table + rs * (size_t)ids[i];
        if (type == OQ_TYPE_Q4_0) oq_dequantize_row_q4_0(row, y + i * k, k);
        else if (type == OQ_TYPE_Q8_0) oq_dequantize_row_q8_0(row, y + i * k, k);
        else if (type == OQ_TYPE_Q4_K) oq_dequantize_row_q4_K(row, y + i * k, k);
        else if (type == OQ_TYPE_F32) memcpy(y + i * k, row, (size_t)k * 4);
        else if (type == OQ_TYPE_F16) { const uint16_t *h = (const uint16_t *)row; for (int64_t t = 0; t < k; ++t) y[i * k + t] = oq_fp16_to_fp32(h[t]); }
        else return -1;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * One decoder step of the Llama family as the reference graph computes it:
 * LMBlock1Forward (src/layers.cpp:2719-2761) = RMSNorm -> attention -> add -> RMSNorm -> SwiGLU MLP -> add,
 * attention = q/k/v Linear, RoPE on q,k, K/V appended to an F16 cache, scores = K.Q (F16 operands,
 * ggml-cpu.c:213-219), scale, causal mask, softmax, V.P (F16 operands), o_proj
 * (src/layers.cpp:2541-2561, :2681-2698, :3044-3123).  Used by tests as the whole-step checker.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int type;              /* weight quant type */
    int hidden, n_heads, n_kv_heads, head_dim, ffn, max_len;
    int rope_mode;         /* 0 normal, 2 neox */
    float rope_theta, eps;
    const float *attn_norm, *ffn_norm;              /* [hidden] */
    const void *wq, *wk, *wv, *wo, *wgate, *wup, *wdown;
    const float *bq, *bk, *bv;                       /* optional biases (Qwen2) */
    uint16_t *k_cache;     /* [max_len][kv_hidden] f16 */
    uint16_t *v_cache;     /* [kv_hidden][max_len] f16 (transposed, src/layers.cpp:2937) */
} oq_layer;

void oq_layer_step(const oq_layer *L, float *h /* [hidden] in/out */, int pos, int q8_0_variant) {
    const int H = L->hidden, nh = L->n_heads, nkv = L->n_kv_heads, hd = L->head_dim, F = L->ffn;
    const int kvh = nkv * hd, qh = nh * hd;
    float *xn = (float *)malloc(sizeof(float) * (size_t)(H > F ? H : F));
    float *q = (float *)malloc(sizeof(float) * qh), *k = (float *)malloc(sizeof(float) * kvh), *v = (float *)malloc(sizeof(float) * kvh);
    float *att = (float *)malloc(sizeof(float) * qh);
    float *sc = (float *)malloc(sizeof(float) * (size_t)(pos + 1));
    float *g = (float *)malloc(sizeof(float) * F), *u = (float *)malloc(sizeof(float) * F), *o = (float *)malloc(sizeof(float) * H);
    const int32_t p32 = pos;

    oq_rms_norm(h, L->attn_norm, xn, H, 1, L->eps);
    oq_mul_mat(L->type, L->wq, oq_row_size(L->type, H), H, qh, xn, 1, q, q8_0_variant);
    oq_mul_mat(L->type, L->wk, oq_row_size(L->type, H), H, kvh, xn, 1, k, q8_0_variant);
    oq_mul_mat(L->type, L->wv, oq_row_size(L->type, H), H, kvh, xn, 1, v, q8_0_variant);
    if (L->bq) for (int i = 0; i < qh; ++i) q[i] += L->bq[i];
    if (L->bk) for (int i = 0; i < kvh; ++i) k[i] += L->bk[i];
    if (L->bv) for (int i = 0; i < kvh; ++i) v[i] += L->bv[i];
    oq_rope(q, q, &p32, NULL, hd, nh, 1, hd, L->rope_mode, 0, L->rope_theta, 1.0f, 0.0f, 1.0f, 32.0f, 1.0f);
    oq_rope(k, k, &p32, NULL, hd, nkv, 1, hd, L->rope_mode, 0, L->rope_theta, 1.0f, 0.0f, 1.0f, 32.0f, 1.0f);
    for (int i = 0; i < kvh; ++i) {
        L->k_cache[(size_t)pos * kvh + i] = oq_fp32_to_fp16(k[i]);
        L->v_cache[(size_t)i * L->max_len + pos] = oq_fp32_to_fp16(v[i]);
    }
    const int n_kv = pos + 1;
    const float scale = 1.0f / sqrtf((float)hd);
    for (int hh = 0; hh < nh; ++hh) {
        const int kvi = hh / (nh / nkv);
        /* scores = K . Q with both operands in f16 */
        for (int t = 0; t < n_kv; ++t) {
            double sum = 0.0;
            for (int d = 0; d < hd; ++d) {
                float a = oq_fp16_to_fp32(L->k_cache[(size_t)t * kvh + kvi * hd + d]);
                float b = oq_fp16_to_fp32(oq_fp32_to_fp16(q[hh * hd + d]));
                sum += (double)(a * b);
            }
            sc[t] = (float)sum;
        }
        oq_soft_max(sc, NULL, sc, n_kv, 1, scale);
        for (int d = 0; d < hd; ++d) {
            double sum = 0.0;
            const uint16_t *vr = L->v_cache + (size_t)(kvi * hd + d) * L->max_len;
            for (int t = 0; t < n_kv; ++t) sum += (double)(oq_fp16_to_fp32(vr[t]) * oq_fp16_to_fp32(oq_fp32_to_fp16(sc[t])));
            att[hh * hd + d] = (float)sum;
        }
    }
    oq_mul_mat(L->type, L->wo, oq_row_size(L->type, qh), qh, H, att, 1, o, q8_0_variant);
    for (int i = 0; i < H; ++i) h[i] += o[i];

    oq_rms_norm(h, L->ffn_norm, xn, H, 1, L->eps);
    oq_mul_mat(L->type, L->wgate, oq_row_size(L->type, H), H, F, xn, 1, g, q8_0_variant);
    oq_mul_mat(L->type, L->wup, oq_row_size(L->type, H), H, F, xn, 1, u, q8_0_variant);
    oq_silu_mul(g, u, g, F);
    oq_mul_mat(L->type, L->wdown, oq_row_size(L->type, F), F, H, g, 1, o, q8_0_variant);
    for (int i = 0; i < H; ++i) h[i] += o[i];

    free(xn); free(q); free(k); free(v); free(att); free(sc); free(g); free(u); free(o);
}
