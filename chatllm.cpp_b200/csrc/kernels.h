// kernels.h — internal C++ interface between the .cu translation units and the C ABI (abi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace b200 {

// ---- quantize.cu
size_t qact_col_bytes(int wtype, int64_t k);
int quantize_act(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * qact, cudaStream_t st);
int repack_window(int wtype, const void * native_window, void * dev_tensor, int64_t tensor_off, int64_t nbytes, int64_t k, bool inverse,
                  cudaStream_t st);

// ---- gemv.cu
struct GemvTuning {
    int ks;      // k-units (256 elements) per pipeline stage
    int stages;  // pipeline depth per warp
    int warps;   // consumer warps per CTA
    int rg;      // rows per row-group (1, 2 or 4)
    int grid;    // CTAs (0 = auto)
};
// epilogue selectors for mul_mat_q
enum : int { EPI_NONE = 0, EPI_BIAS = 1 };
int mul_mat_q(int wtype, const void * W, int64_t k, int64_t m, const void * qact, int64_t n, float * y, int64_t ldy, const float * bias,
              const GemvTuning * tune, cudaStream_t st);
int mul_mat_q_multi(int wtype, int mode, int nmat, const void * const * W, const int64_t * m, float * const * y, const int64_t * ldy,
                    const float * const * bias, int64_t k, const void * qact, int64_t n, const GemvTuning * tune, cudaStream_t st);
// expert-indexed GEMV (ggml_mul_mat_id, one token): ids on the device; act_cols = 1 (shared column) or n_ids (one per slot)
int mul_mat_q_id(int wtype, int paired, const void * W0, const void * W1, int64_t k, int64_t m, int n_expert, const int32_t * ids, int n_ids,
                 const void * qact, int act_cols, float * y, int64_t ldy, const GemvTuning * tune, cudaStream_t st);
int sm_count();

// ---- decode_mk.cu: the whole decode token of a dense Llama-family model as one persistent kernel
struct DecodeLayer {
    const void * wq, * wk, * wv, * wo, * wgate, * wup, * wdown;  // quantized matrices in the device layout
    const float * bq, * bk, * bv;                                  // optional q/k/v biases (NULL)
    const float * attn_norm, * ffn_norm;
    void * k_cache;  // F16 [n_ctx][k_row_stride]
    void * v_cache;  // F16 transposed [kv_hidden][v_row_stride]
};
struct DecodeModel {
    int wtype, n_layers, hidden, heads, kv_heads, head_dim, ffn, vocab, rope_mode, embed_type;
    float rope_theta, eps, attn_scale;
    int64_t k_row_stride, v_row_stride;  // halves
    const DecodeLayer * layers;           // HOST array of n_layers entries
    const void * embed;                   // quantized embedding table (NULL: the caller always passes the hidden state)
    const float * final_norm;
    const void * lm_head;                 // NULL: the step ends with the hidden state (a layer shard that is not the last)
    const float * rope_freq_factors;
};
struct DecodeIO {
    const int32_t * tok;   // device; NULL: x holds the incoming hidden state
    const int32_t * pos;   // device: position of this token (RoPE angle, K-cache row)
    int n_kv;              // positions attended to, < 0: pos[0] + 1
    int v_col;             // V-cache column of this token, < 0: pos[0]
    float * x;             // residual stream [hidden] in/out (NULL: plan-owned)
    float * logits;        // [vocab]
    int32_t * next_tok;    // optional: argmax of the logits (first maximum)
    int flags;             // bit 0: tok[0] = next_tok, pos[0] += 1 at the end (greedy decoding without the host)
    int step_begin, step_end;  // debug: run only steps [begin, end) (0, 0 = all)
    // layer-sharded multi-GPU hand-off through NVLink peer memory (see include/chatllm_b200.h); all NULL / 0 on one GPU
    const void * wait_flag;
    float * send_x;
    void * send_flag;
    int32_t * send_tok;
    int wait_offset, reserved;
};
void * decode_plan_create(const DecodeModel & m, int max_ctx, int * err);
void decode_plan_destroy(void * plan);
int decode_plan_set_kv(void * plan, int layer, void * k_cache, void * v_cache);
int decode_plan_status(void * plan, cudaStream_t st);
int decode_plan_info(void * plan, int * grid, int * smem_bytes, int * n_steps, int * stages, int * ks);
int decode_step(void * plan, const DecodeIO & io, cudaStream_t st);
int decode_plan_times(void * plan, long long * out, int cap, cudaStream_t st);

// ---- decode_graph.cu: the same token as ONE replayed CUDA graph of the per-op kernels (the plugin's default decode path)
void * decode_graph_create(const DecodeModel & m, int max_ctx, int attn_cluster, int * err);
void decode_graph_destroy(void * plan);
int decode_graph_set_kv(void * plan, int layer, void * k_cache, void * v_cache);
int decode_graph_step(void * plan, const int32_t * tok, const int32_t * pos, const float * x_in, float * x_out, float * logits, int n_kv, cudaStream_t st);
bool decode_graph_ready(void * plan, int n_kv);
int decode_graph_prepare(void * plan, int n_kv_next, cudaStream_t live);
void decode_graph_stats(void * plan, long long * replays, long long * recaptures, long long * reinstantiations);

// ---- prefill.cu: batched (n > 8) quantized matmul on the int8 tensor cores, plain activation layout
size_t pact_col_bytes(int wtype, int64_t k);
int quantize_plain(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * pact, cudaStream_t st);
int mul_mat_q_batched(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                      cudaStream_t st);

// ---- prefill_tc.cu: OPT-IN (B200_MMQ_TCGEN05=1) tcgen05.mma kind::i8 version of the Q4_K prompt matmul, bit-identical to mmq_kernel by construction
bool mmq_tc_enabled();
int mul_mat_q_batched_tc(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy, const float * bias,
                         size_t col_bytes, cudaStream_t st);

// ---- ops.cu
int rms_norm_mul(const float * x, const float * w, float * y, int64_t ne0, int64_t nrows, float eps, cudaStream_t st);
int add_f32(const float * a, const float * b, float * y, int64_t n, cudaStream_t st);
int silu_mul(const float * gate, const float * up, float * y, int64_t n, cudaStream_t st);
int rope_f32(const float * x, float * y, const int32_t * pos, const float * freq_factors, int64_t ne0, int64_t n_heads, int64_t n_tokens,
             int64_t x_head_stride, int64_t x_tok_stride, int64_t y_head_stride, int64_t y_tok_stride, int n_dims, int mode, int n_ctx_orig,
             float freq_base, float freq_scale, float ext_factor, float attn_factor, float beta_fast, float beta_slow, cudaStream_t st);
int soft_max_f32(const float * x, const float * mask, float * y, int64_t ne0, int64_t nrows, float scale, cudaStream_t st);
// n_rows > 0: ids outside [0, n_rows) produce a zero row instead of an out-of-bounds read
int get_rows_q(int type, const void * table, int64_t k, const int32_t * ids, int64_t n, float * y, cudaStream_t st, int64_t n_rows = 0);
size_t attn_decode_scratch_bytes(int n_heads, int n_kv);
int attn_decode(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                int64_t k_row_stride, int64_t v_row_stride, float scale, cudaStream_t st);
int kv_store(const float * k, const float * v, void * kc, void * vc, int kv_hidden, int64_t k_row_stride, int64_t v_row_stride, int pos, cudaStream_t st);

int peer_wait(const void * flag, const void * seq, int offset, void * status, cudaStream_t st);
int peer_send(const float * x, float * peer_x, int64_t n, const int32_t * tok, int32_t * peer_tok, void * peer_flag, void * seq, int32_t * pos, cudaStream_t st);
int argmax_f32(const float * x, int64_t n, int32_t * out, cudaStream_t st);

// ---- fused.cu
int add_rmsnorm_quant(int wtype, const float * x, const float * r, const float * w, float * x_out, float * y_out, void * qact, int64_t ne0, int64_t nrows,
                      float eps, cudaStream_t st);
int rope_kv_store(float * q, const float * k, const float * v, const int32_t * pos, const float * ff, void * kc, void * vc, int n_heads, int kv_heads,
                  int head_dim, int mode, float freq_base, int64_t k_row_stride, int64_t v_row_stride, cudaStream_t st);
// q_out may differ from q; v_col >= 0: vc already points at the token's column (the host's V-cache view), else column = pos[0]
int rope_kv_store2(const float * q, float * q_out, const float * k, const float * v, const int32_t * pos, const float * ff, void * kc, void * vc, int n_heads,
                   int kv_heads, int head_dim, int mode, float freq_base, int64_t k_row_stride, int64_t v_row_stride, int v_col, cudaStream_t st);
size_t attn_decode2_scratch_bytes(int n_heads, int n_kv);
int attn_decode2(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                 int64_t k_row_stride, int64_t v_row_stride, float scale, cudaStream_t st);
// + qact (may be NULL) = out quantized as the activations of a following matmul with weight type wtype (n_heads*head_dim % 256 == 0)
// preload: 1 = K / V rows older than n_kv - 1 may be streamed before the programmatic-dependent-launch wait (they were written by
// EARLIER decode steps), 0 = a launch just before this call may have rewritten them (cache shift), -1 = library default
int attn_decode3(const float * q, const void * kc, const void * vc, float * out, float * scratch, int n_heads, int kv_heads, int head_dim, int n_kv,
                 int64_t k_row_stride, int64_t v_row_stride, float scale, int wtype, void * qact, cudaStream_t st, int preload = -1, int cluster = -1);
bool attn_decode3_supported(int n_heads, int kv_heads, int head_dim, int64_t k_row_stride, int64_t v_row_stride);
// cluster: 1 = thread-block-cluster V.P (2 launches), 0 = split V.P + tail launch (3 launches), -1 = library default (B200_ATTN_CLUSTER, on)

// ---- ops_generic.cu: strided / broadcasting ops as ggml graphs present them (ne[0] fastest, nb in bytes)
struct TV {
    void * data;
    int type;  // ggml type id
    int64_t ne[4];
    int64_t nb[4];
};
int op_bin(int op /*0 add, 1 mul, 2 div*/, const TV & a, const TV & b, const TV & d, cudaStream_t st);
int op_cpy(const TV & s, const TV & d, cudaStream_t st);
int op_set_rows(const TV & s, const TV & ids, const TV & d, cudaStream_t st);
int op_scale(const float * x, float * y, int64_t n, float s, float b, cudaStream_t st);
int op_clamp(const float * x, float * y, int64_t n, float lo, float hi, cudaStream_t st);
int op_silu(const float * x, float * y, int64_t n, cudaStream_t st);
int op_diag_mask_inf(const float * x, float * y, int64_t ne0, int64_t ne1, int64_t n, int n_past, cudaStream_t st);
int op_soft_max(const TV & x, const TV * mask, const TV & y, float scale, cudaStream_t st);
int op_rms_norm(const TV & x, const TV & y, float eps, cudaStream_t st);
int op_mul_mat_f(const TV & a, const TV & b, const TV & d, cudaStream_t st);
int op_sum_rows(const TV & x, const TV & y, cudaStream_t st);
int op_repeat(const TV & s, const TV & d, cudaStream_t st);
int op_argsort(const TV & x, const TV & y, int k_out, bool ascending, bool swap01, cudaStream_t st);
int op_get_rows_f(const TV & a, const TV & ids, const TV & d, cudaStream_t st);

}  // namespace b200
