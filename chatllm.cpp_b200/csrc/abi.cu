// abi.cu — extern "C" entry points declared in include/chatllm_b200.h
#include "../../include/chatllm_b200.h"
#include "common.cuh"
#include "kernels.h"

#include <cstring>
#include <mutex>

using namespace b200;

static GemvTuning g_tune = {0, 0, 0, 0, 0};

// per-device scratch for the convenience entry point (grown on demand; serialised by the stream it is used on)
static void * g_scratch[16] = {nullptr};
static size_t g_scratch_bytes[16] = {0};
static std::mutex g_mu;

static void * scratch(size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_scratch_bytes[dev] < bytes) {
        if (g_scratch[dev]) { cudaDeviceSynchronize(); cudaFree(g_scratch[dev]); }
        size_t nb = bytes < (1u << 20) ? (1u << 20) : bytes;
        if (cudaMalloc(&g_scratch[dev], nb) != cudaSuccess) { g_scratch[dev] = nullptr; g_scratch_bytes[dev] = 0; return nullptr; }
        g_scratch_bytes[dev] = nb;
    }
    return g_scratch[dev];
}

extern "C" {

int b200_abi_version(void) { return 1; }
int b200_device_sm_count(void) { return sm_count(); }

int b200_repack_weights(int wtype, void * native, void * dev_tensor, int64_t tensor_off, int64_t nbytes, int64_t k, int inverse, void * stream) {
    cudaStream_t st = (cudaStream_t) stream;
    if (wtype == B200_TYPE_Q4_0 || wtype == B200_TYPE_Q8_0) {
        if (k % 256) return B200_ERR_UNSUPPORTED;
        return repack_window(wtype, native, dev_tensor, tensor_off, nbytes, k, inverse != 0, st);
    }
    if (inverse) return (int) cudaMemcpyAsync(native, (uint8_t *) dev_tensor + tensor_off, (size_t) nbytes, cudaMemcpyDeviceToDevice, st);
    return (int) cudaMemcpyAsync((uint8_t *) dev_tensor + tensor_off, native, (size_t) nbytes, cudaMemcpyDeviceToDevice, st);
}

size_t b200_qact_col_bytes(int wtype, int64_t k) { return qact_col_bytes(wtype, k); }

int b200_quantize_act(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * qact, void * stream) {
    return quantize_act(wtype, x, x_col_stride, k, n, qact, (cudaStream_t) stream);
}

int b200_mul_mat_q(int wtype, const void * W, int64_t k, int64_t m, const void * qact, int64_t n, float * y, int64_t ldy, const float * bias,
                   void * stream) {
    return mul_mat_q(wtype, W, k, m, qact, n, y, ldy, bias, &g_tune, (cudaStream_t) stream);
}

int b200_mul_mat(int wtype, const void * W, int64_t k, int64_t m, const float * x, int64_t x_col_stride, int64_t n, float * y, int64_t ldy,
                 const float * bias, void * stream) {
    if (n <= 0) return B200_OK;
    if (n > B200_GEMV_MAX_COLS) {  // prompt-sized batch: int8 tensor-core path (prefill.cu)
        void * q = scratch(pact_col_bytes(wtype, k) * (size_t) n);
        if (!q) return (int) cudaErrorMemoryAllocation;
        int rc = quantize_plain(wtype, x, x_col_stride, k, n, q, (cudaStream_t) stream);
        if (rc) return rc;
        return mul_mat_q_batched(wtype, W, k, m, q, n, y, ldy, bias, (cudaStream_t) stream);
    }
    void * q = scratch(qact_col_bytes(wtype, k) * (size_t) n);
    if (!q) return (int) cudaErrorMemoryAllocation;
    int rc = quantize_act(wtype, x, x_col_stride, k, n, q, (cudaStream_t) stream);
    if (rc) return rc;
    return mul_mat_q(wtype, W, k, m, q, n, y, ldy, bias, &g_tune, (cudaStream_t) stream);
}

size_t b200_pact_col_bytes(int wtype, int64_t k) { return pact_col_bytes(wtype, k); }
int b200_quantize_plain(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * pact, void * stream) {
    return quantize_plain(wtype, x, x_col_stride, k, n, pact, (cudaStream_t) stream);
}
int b200_mul_mat_q_batched(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                           void * stream) {
    return mul_mat_q_batched(wtype, W, k, m, pact, n, y, ldy, bias, (cudaStream_t) stream);
}

int b200_mul_mat_q_batched_tc(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                              void * stream) {
    return mul_mat_q_batched_tc(wtype, W, k, m, pact, n, y, ldy, bias, pact_col_bytes(wtype, k), (cudaStream_t) stream);
}

int b200_gemv_set_tuning(int ks, int stages, int warps, int rg, int grid) {
    g_tune.ks = ks; g_tune.stages = stages; g_tune.warps = warps; g_tune.rg = rg; g_tune.grid = grid;
    return B200_OK;
}

int b200_rms_norm(const float * x, const float * w, float * y, int64_t ne0, int64_t nrows, float eps, void * stream) {
    return rms_norm_mul(x, w, y, ne0, nrows, eps, (cudaStream_t) stream);
}
int b200_add(const float * a, const float * b, float * y, int64_t n, void * stream) { return add_f32(a, b, y, n, (cudaStream_t) stream); }
int b200_silu_mul(const float * gate, const float * up, float * y, int64_t n, void * stream) { return silu_mul(gate, up, y, n, (cudaStream_t) stream); }
int b200_rope(const float * x, float * y, const int32_t * pos, const float * freq_factors, int64_t ne0, int64_t n_heads, int64_t n_tokens,
              int64_t xs_h, int64_t xs_t, int64_t ys_h, int64_t ys_t, int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale,
              float ext_factor, float attn_factor, float beta_fast, float beta_slow, void * stream) {
    return rope_f32(x, y, pos, freq_factors, ne0, n_heads, n_tokens, xs_h, xs_t, ys_h, ys_t, n_dims, mode, n_ctx_orig, freq_base, freq_scale,
                    ext_factor, attn_factor, beta_fast, beta_slow, (cudaStream_t) stream);
}
int b200_soft_max(const float * x, const float * mask, float * y, int64_t ne0, int64_t nrows, float scale, void * stream) {
    return soft_max_f32(x, mask, y, ne0, nrows, scale, (cudaStream_t) stream);
}
int b200_get_rows(int type, const void * table, int64_t k, const int32_t * ids, int64_t n, float * y, void * stream) {
    return get_rows_q(type, table, k, ids, n, y, (cudaStream_t) stream);
}

size_t b200_attn_decode_scratch_bytes(int n_heads, int n_kv) {
    const size_t a = attn_decode_scratch_bytes(n_heads, n_kv), b = attn_decode2_scratch_bytes(n_heads, n_kv);
    return a > b ? a : b;
}
int b200_attn_decode(const float * q, const void * k_cache, const void * v_cache_t, float * out, float * scratch, int n_heads, int kv_heads, int head_dim,
                     int n_kv, int64_t k_row_stride, int64_t v_row_stride, float scale, void * stream) {
    // two-launch version (16-byte row loads) when the cache strides allow it, else the three-launch one
    if (k_row_stride % 8 == 0 && v_row_stride % 8 == 0) {
        const int rc = attn_decode2(q, k_cache, v_cache_t, out, scratch, n_heads, kv_heads, head_dim, n_kv, k_row_stride, v_row_stride, scale, (cudaStream_t) stream);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    return attn_decode(q, k_cache, v_cache_t, out, scratch, n_heads, kv_heads, head_dim, n_kv, k_row_stride, v_row_stride, scale, (cudaStream_t) stream);
}
int b200_attn_decode_quant(const float * q, const void * k_cache, const void * v_cache_t, float * out, float * scratch, int n_heads, int kv_heads,
                           int head_dim, int n_kv, int64_t k_row_stride, int64_t v_row_stride, float scale, int wtype, void * qact, void * stream) {
    if (qact && ((int64_t) n_heads * head_dim) % 256) return B200_ERR_UNSUPPORTED;
    if (k_row_stride % 8 == 0 && v_row_stride % 8 == 0) {
        const int rc = attn_decode3(q, k_cache, v_cache_t, out, scratch, n_heads, kv_heads, head_dim, n_kv, k_row_stride, v_row_stride, scale, wtype, qact,
                                    (cudaStream_t) stream);
        if (rc != B200_ERR_UNSUPPORTED) return rc;
    }
    const int rc = attn_decode(q, k_cache, v_cache_t, out, scratch, n_heads, kv_heads, head_dim, n_kv, k_row_stride, v_row_stride, scale, (cudaStream_t) stream);
    if (rc || !qact) return rc;
    return quantize_act(wtype, out, (int64_t) n_heads * head_dim, (int64_t) n_heads * head_dim, 1, qact, (cudaStream_t) stream);
}
int b200_add_rmsnorm_quant(int wtype, const float * x, const float * r, const float * w, float * x_out, float * y_out, void * qact, int64_t ne0, int64_t nrows,
                           float eps, void * stream) {
    return add_rmsnorm_quant(wtype, x, r, w, x_out, y_out, qact, ne0, nrows, eps, (cudaStream_t) stream);
}
int b200_rope_kv_store(float * q, const float * k, const float * v, const int32_t * pos, const float * freq_factors, void * k_cache, void * v_cache_t,
                       int n_heads, int kv_heads, int head_dim, int mode, float freq_base, int64_t k_row_stride, int64_t v_row_stride, void * stream) {
    return rope_kv_store(q, k, v, pos, freq_factors, k_cache, v_cache_t, n_heads, kv_heads, head_dim, mode, freq_base, k_row_stride, v_row_stride,
                         (cudaStream_t) stream);
}
int b200_mul_mat_q_multi(int wtype, int mode, int nmat, const void * const * W, const int64_t * m, float * const * y, const int64_t * ldy,
                         const float * const * bias, int64_t k, const void * qact, int64_t n, void * stream) {
    return mul_mat_q_multi(wtype, mode, nmat, W, m, y, ldy, bias, k, qact, n, &g_tune, (cudaStream_t) stream);
}
int b200_mul_mat_q_id(int wtype, int paired, const void * W0, const void * W1, int64_t k, int64_t m, int n_expert, const int32_t * ids, int n_ids,
                      const void * qact, int act_cols, float * y, int64_t ldy, void * stream) {
    return mul_mat_q_id(wtype, paired, W0, W1, k, m, n_expert, ids, n_ids, qact, act_cols, y, ldy, &g_tune, (cudaStream_t) stream);
}
int b200_kv_store(const float * k, const float * v, void * k_cache, void * v_cache_t, int kv_hidden, int64_t k_row_stride, int64_t v_row_stride, int pos,
                  void * stream) {
    return kv_store(k, v, k_cache, v_cache_t, kv_hidden, k_row_stride, v_row_stride, pos, (cudaStream_t) stream);
}

// ---- persistent decode kernel (decode_mk.cu).  The C structs of the header and the C++ structs of kernels.h have the same layout.
static_assert(sizeof(b200_decode_layer) == sizeof(DecodeLayer) && sizeof(b200_decode_model) == sizeof(DecodeModel) && sizeof(b200_decode_io) == sizeof(DecodeIO),
              "C ABI structs and kernels.h structs must match");
void * b200_decode_plan_create(const b200_decode_model * model, int max_ctx, int * err) {
    if (!model) { if (err) *err = B200_ERR_ARG; return nullptr; }
    return decode_plan_create(*reinterpret_cast<const DecodeModel *>(model), max_ctx, err);
}
void b200_decode_plan_destroy(void * plan) { decode_plan_destroy(plan); }
int b200_decode_plan_set_kv(void * plan, int layer, void * k_cache, void * v_cache) { return decode_plan_set_kv(plan, layer, k_cache, v_cache); }
int b200_decode_plan_status(void * plan, void * stream) { return decode_plan_status(plan, (cudaStream_t) stream); }
int b200_decode_plan_info(void * plan, int * grid, int * smem_bytes, int * n_steps, int * stages, int * ks) { return decode_plan_info(plan, grid, smem_bytes, n_steps, stages, ks); }
int b200_decode_plan_times(void * plan, long long * out, int cap, void * stream) { return decode_plan_times(plan, out, cap, (cudaStream_t) stream); }
int b200_decode_step(void * plan, const b200_decode_io * io, void * stream) {
    if (!io) return B200_ERR_ARG;
    return decode_step(plan, *reinterpret_cast<const DecodeIO *>(io), (cudaStream_t) stream);
}

int b200_peer_wait(const void * flag, const void * seq, int offset, void * status, void * stream) { return peer_wait(flag, seq, offset, status, (cudaStream_t) stream); }
int b200_peer_send(const float * x, float * peer_x, int64_t n, const int32_t * tok, int32_t * peer_tok, void * peer_flag, void * seq, int32_t * pos, void * stream) {
    return peer_send(x, peer_x, n, tok, peer_tok, peer_flag, seq, pos, (cudaStream_t) stream);
}
int b200_argmax(const float * x, int64_t n, int32_t * out, void * stream) { return argmax_f32(x, n, out, (cudaStream_t) stream); }
int b200_ipc_alloc(size_t bytes, void ** dptr, void * handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dptr || !handle64 || bytes == 0) return B200_ERR_ARG;
    void * p = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&p, bytes));
    B200_CUDA_CHECK(cudaMemset(p, 0, bytes));
    B200_CUDA_CHECK(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return (int) e; }
    memcpy(handle64, &h, 64);
    *dptr = p;
    return B200_OK;
}
int b200_ipc_open(const void * handle64, void ** dptr) {
    if (!dptr || !handle64) return B200_ERR_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int) cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess);
}
int b200_ipc_close(void * dptr) { return (int) cudaIpcCloseMemHandle(dptr); }
int b200_ipc_free(void * dptr) { return (int) cudaFree(dptr); }

}  // extern "C"
