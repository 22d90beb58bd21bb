/* chatllm_b200.h — kernel-level C ABI of libchatllm_b200.so (sm_100a).
 *
 * Plain C, device pointers and sizes only.  Every entry point returns 0 on success, a positive cudaError_t
 * value on a CUDA failure, or a negative B200_ERR_* code; nothing here falls back to the CPU.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).  All calls are asynchronous on it.
 *
 * This is the layer the ggml backend plugin (include/ggml_b200_backend.h, libggml-cuda.so) is built on; each
 * function names the reference routine it replaces (paths relative to the reference tree).
 *
 * Weight layouts in device memory
 *   Q4_K : native ggml block_q4_K stream (ggml/src/ggml-common.h:288-306), rows contiguous.
 *   Q4_0 / Q8_0 : per-row SoA  "qs[nb][16|32] then fp16 d[nb]"  (same byte count and row stride as the native
 *          AoS blocks, ggml-common.h:170-176 / :219-224).  Convert with b200_repack_weights(); the plugin does this
 *          inside buffer.set_tensor / get_tensor so the host application never sees it.
 *   k (= ne00) must be a multiple of 256 for the quantized matmul entry points.
 */
#ifndef CHATLLM_B200_H
#define CHATLLM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_TYPE_F32_ 0
#define B200_TYPE_F16_ 1
#define B200_TYPE_Q4_0_ 2
#define B200_TYPE_Q8_0_ 8
#define B200_TYPE_Q4_K_ 12

/* library / device info */
int b200_abi_version(void);                 /* = 1 */
int b200_device_sm_count(void);

/* ---- weight layout ------------------------------------------------------------------------------------ */
/* Convert a window [tensor_off, tensor_off+nbytes) of a tensor given in the NATIVE ggml layout (`native`, device
 * memory) into the device layout inside `dev_tensor` (inverse=0), or back (inverse=1: `native` is written).
 * Q4_K: plain copy.  Replaces nothing in the reference (the reference CUDA backend keeps AoS blocks); it is what
 * ggml_backend_buffer_i.set_tensor / get_tensor (ggml/src/ggml-backend-impl.h:46-51) do in our plugin. */
int b200_repack_weights(int wtype, void * native, void * dev_tensor, int64_t tensor_off, int64_t nbytes, int64_t k,
                        int inverse, void * stream);

/* ---- activation quantization ---------------------------------------------------------------------------- */
/* bytes of one quantized activation column for a weight type (Q4_K -> Q8_K codes, Q4_0/Q8_0 -> Q8_0 codes) */
size_t b200_qact_col_bytes(int wtype, int64_t k);
/* x: n columns of k floats, column stride x_col_stride floats -> qact (n * b200_qact_col_bytes).
 * Reproduces bit-for-bit quantize_row_q8_K_ref (ggml/src/ggml-quants.c:2555-2592) resp. the x86
 * quantize_row_q8_0 (ggml/src/ggml-cpu/arch/x86/quants.c:290-384) the CPU backend applies to src1
 * (ggml/src/ggml-cpu/ggml-cpu.c:1291-1326). */
int b200_quantize_act(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * qact, void * stream);

/* ---- quantized matmul (decode GEMV / skinny GEMM, n <= 8 fast; larger n loops column groups) -------------- */
/* y[c*ldy + i] = sum_k W[i,k] * x_c[k]  (+ bias[i]),  i < m, c < n.  Replaces ggml_compute_forward_mul_mat
 * (ggml/src/ggml-cpu/ggml-cpu.c:1229-1421) / mul_mat_vec_q (ggml/src/ggml-cuda/mmvq.cu:140-356) for
 * src0 in {Q4_K, Q4_0, Q8_0}. */
int b200_mul_mat_q(int wtype, const void * W, int64_t k, int64_t m, const void * qact, int64_t n, float * y, int64_t ldy,
                   const float * bias, void * stream);
/* convenience: quantize x into an internal per-device scratch, then b200_mul_mat_q */
int b200_mul_mat(int wtype, const void * W, int64_t k, int64_t m, const float * x, int64_t x_col_stride, int64_t n, float * y,
                 int64_t ldy, const float * bias, void * stream);
/* ---- prompt-sized batches (n > B200_GEMV_MAX_COLS): exact int8 tensor-core GEMM -------------------------------------
 * Same arithmetic as above (the CPU reference quantizes every src1 column to Q8_K / Q8_0, ggml-cpu.c:1291-1326, and
 * sums exact int8 block dots); activations use a plain k-contiguous layout: qs[k] | d | block sums, 16-byte padded.
 * Replaces the prompt branch of ggml_compute_forward_mul_mat (ggml/src/ggml-cpu/ggml-cpu.c:1229-1421; llamafile
 * tinyBLAS ggml/src/ggml-cpu/llamafile/sgemm.cpp) and mul_mat_q (ggml/src/ggml-cuda/mmq.cuh). b200_mul_mat picks it. */
#define B200_GEMV_MAX_COLS 8
size_t b200_pact_col_bytes(int wtype, int64_t k);
int b200_quantize_plain(int wtype, const float * x, int64_t x_col_stride, int64_t k, int64_t n, void * pact, void * stream);
int b200_mul_mat_q_batched(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy,
                           const float * bias, void * stream);
/* OPT-IN tcgen05 version of b200_mul_mat_q_batched (csrc/prefill_tc.cu: tcgen05.mma kind::i8, accumulators in TMEM).  Q4_K: the 6-bit
 * sub-block scales are folded into three int8 weight planes so the super-block sums stay exact in int32; Q4_0 / Q8_0: one TMEM
 * accumulator per 32-element block.  Same integers and the same fp32 rescale expression as b200_mul_mat_q_batched; not yet run on a
 * GPU (round 1) — the plugin uses it only with B200_MMQ_TCGEN05=1. */
int b200_mul_mat_q_batched_tc(int wtype, const void * W, int64_t k, int64_t m, const void * pact, int64_t n, float * y, int64_t ldy,
                              const float * bias, void * stream);
/* Several matrices that share the activation vector in ONE launch.
 *   mode 0 (concat, nmat <= 3): y_i[c*ldy_i + r] = W_i[r,:].x_c (+ bias_i[r])       — q/k/v projections of a layer
 *   mode 1 (paired, nmat == 2, m_0 == m_1): y_0[c*ldy_0 + r] = silu(W_0[r,:].x_c) * (W_1[r,:].x_c)
 *                                            — gate/up projections + SwiGLU of BaseMLP::forward (src/layers.cpp:2475-2483) */
int b200_mul_mat_q_multi(int wtype, int mode, int nmat, const void * const * W, const int64_t * m, float * const * y, const int64_t * ldy,
                         const float * const * bias, int64_t k, const void * qact, int64_t n, void * stream);
/* Expert-indexed matmul for ONE token = ggml_mul_mat_id (ggml/src/ggml.c:3225-3240; CPU ggml_compute_forward_mul_mat_id
 * ggml/src/ggml-cpu/ggml-cpu.c:1503-1700; caller MultiLinear::forward src/layers.cpp:2145-2151, MultiMLP::forward :3674-3688).
 *   W0 (and W1): stacks of n_expert matrices [m, k] in the device layout, expert e at byte offset e * m * row_bytes.
 *   ids: n_ids expert indices in DEVICE memory (the router's top_k output).  qact: act_cols quantized columns;
 *   act_cols == 1: every slot uses column 0 (src1 broadcast: gate / up), act_cols == n_ids: slot s uses column s (down).
 *   paired = 0:  y[s*ldy + r] = W0[ids[s]][r,:] . x_c(s)
 *   paired = 1:  y[s*ldy + r] = silu(W0[ids[s]][r,:] . x) * (W1[ids[s]][r,:] . x)        (experts' SwiGLU in one launch) */
int b200_mul_mat_q_id(int wtype, int paired, const void * W0, const void * W1, int64_t k, int64_t m, int n_expert, const int32_t * ids,
                      int n_ids, const void * qact, int act_cols, float * y, int64_t ldy, void * stream);
/* override the pipeline shape of the GEMV kernel (0 = keep default): units(256 elts)/stage, stages, warps/CTA,
 * rows/group, grid */
int b200_gemv_set_tuning(int ks, int stages, int warps, int rg, int grid);

/* ---- fp32 glue ops ----------------------------------------------------------------------------------------- */
/* y = rms_norm(x) * w   (w may be NULL).  ggml/src/ggml-cpu/ops.cpp:3710-3758 + src/layers.cpp:2216-2225 */
int b200_rms_norm(const float * x, const float * w, float * y, int64_t ne0, int64_t nrows, float eps, void * stream);
int b200_add(const float * a, const float * b, float * y, int64_t n, void * stream);
/* y = silu(gate) * up.  src/layers.cpp:2475-2483 */
int b200_silu_mul(const float * gate, const float * up, float * y, int64_t n, void * stream);
/* RoPE on x[ne0, n_heads, n_tokens] (strides in floats), mode 0 = NORMAL, 2 = NEOX.
 * ggml/src/ggml-cpu/ops.cpp:5587-5865 */
int b200_rope(const float * x, float * y, const int32_t * pos, const float * freq_factors, int64_t ne0, int64_t n_heads,
              int64_t n_tokens, int64_t x_head_stride, int64_t x_tok_stride, int64_t y_head_stride, int64_t y_tok_stride,
              int n_dims, int mode, int n_ctx_orig, float freq_base, float freq_scale, float ext_factor, float attn_factor,
              float beta_fast, float beta_slow, void * stream);
/* y = softmax(x*scale + mask) over rows of ne0.  ggml/src/ggml-cpu/ops.cpp:5225-5335 */
int b200_soft_max(const float * x, const float * mask, float * y, int64_t ne0, int64_t nrows, float scale, void * stream);
/* y[i, :] = dequant(table[ids[i], :]).  ggml/src/ggml-cpu/ops.cpp:4820 */
int b200_get_rows(int type, const void * table, int64_t k, const int32_t * ids, int64_t n, float * y, void * stream);

/* ---- fused decode-step kernels --------------------------------------------------------------------------------- */
/* [x_out = x + r] ; y = rms_norm(x_out) * w ; qact = quantized y for weight type `wtype` (r, x_out, y_out, qact may be NULL).
 * Replaces ADD + RMS_NORM + MUL (src/layers.cpp:2719-2761, :2216-2225) + the src1 conversion of the following matmuls
 * (ggml-cpu.c:1291-1326).  ne0 % 256 == 0, ne0 <= 20480. */
int b200_add_rmsnorm_quant(int wtype, const float * x, const float * r, const float * w, float * x_out, float * y_out, void * qact,
                           int64_t ne0, int64_t nrows, float eps, void * stream);
/* one token: RoPE(q) in place, RoPE(k) -> K cache row pos[0], v -> V cache column pos[0] (F16, RNE).  Replaces the two
 * rope_ext_inplace nodes + SET_ROWS + TRANSPOSE/CPY of src/layers.h:2103-2114, src/layers.cpp:3044-3123. */
int b200_rope_kv_store(float * q, const float * k, const float * v, const int32_t * pos, const float * freq_factors, void * k_cache,
                       void * v_cache_t, int n_heads, int kv_heads, int head_dim, int mode, float freq_base, int64_t k_row_stride,
                       int64_t v_row_stride, void * stream);

/* ---- decode attention over the reference's F16 KV-cache layouts ---------------------------------------------- */
/* K cache [max_len][kv_heads*head_dim] (row stride k_row_stride halves), V cache TRANSPOSED [kv_heads*head_dim][max_len]
 * (row stride v_row_stride halves) — src/layers.cpp:2933-2941.  out[h*head_dim+d] for one query token attending to
 * positions [0, n_kv).  Replaces the mul_mat / scale / diag_mask_inf / soft_max / mul_mat / permute / cont chain of
 * CoreAttention::calc_attn_scores (src/layers.cpp:2541-2561) with the same f16-operand, fp32-accumulate arithmetic.
 * Ordering contract: only position n_kv-1 (the token just appended by b200_rope_kv_store / b200_kv_store on the same stream) may have
 * been written by the launches immediately preceding this call; older positions must come from earlier decode steps (or from work the
 * host has synchronized with) — the kernels start streaming them before their programmatic-dependent-launch wait. */
size_t b200_attn_decode_scratch_bytes(int n_heads, int n_kv);
int b200_attn_decode(const float * q, const void * k_cache, const void * v_cache_t, float * out, float * scratch, int n_heads,
                     int kv_heads, int head_dim, int n_kv, int64_t k_row_stride, int64_t v_row_stride, float scale, void * stream);
/* Same, and additionally qact (may be NULL) = `out` quantized as the activations of the following matmul with weight type
 * wtype (the o-projection: src1 conversion of ggml-cpu.c:1291-1326), written by the kernel that sums the split V.P partials.
 * Requires n_heads*head_dim % 256 == 0 when qact != NULL. */
int b200_attn_decode_quant(const float * q, const void * k_cache, const void * v_cache_t, float * out, float * scratch, int n_heads,
                           int kv_heads, int head_dim, int n_kv, int64_t k_row_stride, int64_t v_row_stride, float scale, int wtype,
                           void * qact, void * stream);
/* append one token: K row `pos`, V column `pos` (F32 -> F16, RNE).  KVCacheAttention::save_to_cache src/layers.cpp:3044-3123 */
int b200_kv_store(const float * k, const float * v, void * k_cache, void * v_cache_t, int kv_hidden, int64_t k_row_stride,
                  int64_t v_row_stride, int pos, void * stream);

/* ---- the whole decode token as ONE persistent kernel (csrc/decode_mk.cu) ------------------------------------------------------
 * Replaces, for one token of a dense Llama-family model (RMSNorm -> q/k/v -> RoPE -> KV append -> attention -> o -> residual ->
 * RMSNorm -> SwiGLU MLP -> residual, then final norm + lm_head), the ~1000-node graph the reference builds and schedules per token:
 * HeterogeneousModel::forward (src/models.cpp:1399-1424) -> LMBlock1Forward::forward (src/layers.cpp:2719-2761) ->
 * LMFinalSteps::forward (src/models.cpp:1736-1785).  One cooperative launch of one CTA per SM: every warp streams ITS rows of every
 * quantized matmul of the token through a private bulk-copy ring that runs ahead across the grid barriers between steps, so HBM stays
 * busy while activations are normalised / quantized and attention runs.  Same arithmetic as the per-op entry points above.
 * All pointers inside the structs are DEVICE pointers (weights in the device layout); `layers` itself is a host array. */
typedef struct b200_decode_layer {
    const void * wq, * wk, * wv, * wo, * wgate, * wup, * wdown;
    const float * bq, * bk, * bv;          /* optional q/k/v biases (NULL) */
    const float * attn_norm, * ffn_norm;
    void * k_cache;                        /* F16 [n_ctx][k_row_stride]            (src/layers.cpp:2933) */
    void * v_cache;                        /* F16 transposed [kv_hidden][v_row_stride] (src/layers.cpp:2937) */
} b200_decode_layer;
typedef struct b200_decode_model {
    int32_t wtype, n_layers, hidden, heads, kv_heads, head_dim, ffn, vocab, rope_mode, embed_type; /* embed_type 0: = wtype */
    float rope_theta, eps, attn_scale;
    int64_t k_row_stride, v_row_stride;    /* in halves, multiples of 8 */
    const b200_decode_layer * layers;      /* HOST array [n_layers] */
    const void * embed;                    /* quantized embedding table, NULL: callers always pass the hidden state */
    const float * final_norm;
    const void * lm_head;                  /* NULL: the step ends with the hidden state (a layer shard that is not the last) */
    const float * rope_freq_factors;       /* optional [head_dim/2] */
} b200_decode_model;
typedef struct b200_decode_io {
    const int32_t * tok;                   /* device; NULL: x holds the incoming hidden state */
    const int32_t * pos;                   /* device: position of this token (RoPE angle, K-cache row) */
    int32_t n_kv;                          /* positions attended to; < 0: pos[0] + 1 (read on the device) */
    int32_t v_col;                         /* V-cache column of this token; < 0: pos[0] */
    float * x;                             /* residual stream [hidden], in/out; NULL: plan-owned */
    float * logits;                        /* [vocab] (required when the model has an lm_head) */
    int32_t * next_tok;                    /* optional: argmax of the logits, first maximum */
    int32_t flags;                         /* bit 0: tok[0] = next_tok, pos[0] += 1 when the step ends (greedy decoding without the host);
                                              bit 1: pos[0] += 1 only (a shard that does not own the token) */
    int32_t step_begin, step_end;          /* debug: run only steps [begin, end); 0, 0 = the whole token */
    /* Layer-sharded multi-GPU (the reference copies the hidden row with cpy_tensor_async + events,
     * ggml/src/ggml-cuda/ggml-cuda.cu:2806-2866): one process per GPU, each running its layer range with this entry point.  The kernel
     * waits (after starting its weight streams) until the 64-bit LOCAL flag `wait_flag` reaches <launches of this plan completed so far> +
     * wait_offset, and when the token is done stores the hidden row into `send_x` / the next token into `send_tok` — PEER memory of the
     * next shard, mapped with b200_ipc_open — and raises the peer's flag `send_flag` to <launches completed> + 1 (system-scope release).
     * First shard: wait_flag = its token flag, wait_offset 0;  later shards: wait_flag = their x flag, wait_offset 1. */
    const void * wait_flag;
    float * send_x;
    void * send_flag;
    int32_t * send_tok;
    int32_t wait_offset, reserved;
} b200_decode_io;
/* returns NULL and *err (B200_ERR_* / cudaError_t) when the shape is not supported; max_ctx sizes the attention workspace */
void * b200_decode_plan_create(const b200_decode_model * model, int max_ctx, int * err);
void b200_decode_plan_destroy(void * plan);
/* the KV cache tensors of a layer moved (they are owned by the host application) */
int b200_decode_plan_set_kv(void * plan, int layer, void * k_cache, void * v_cache);
/* 0 = fine, 1 = a grid barrier timed out in an earlier step (synchronizes `stream`) */
int b200_decode_plan_status(void * plan, void * stream);
int b200_decode_plan_info(void * plan, int * grid, int * smem_bytes, int * n_steps, int * stages, int * ks);
/* profiling aid (B200_MK_TIMES=1 when the plan is created): SM-clock stamps at kernel start and after every step of the last launch */
int b200_decode_plan_times(void * plan, long long * out, int cap, void * stream);
int b200_decode_step(void * plan, const b200_decode_io * io, void * stream);
/* out[0] = index of the first maximum of x[0..n): greedy sampling on the device (the reference reads the logits back and scans them on the
 * host every token, src/models.cpp:1026-1031). */
int b200_argmax(const float * x, int64_t n, int32_t * out, void * stream);
/* The two ends of a layer shard's CUDA graph (reference: cpy_tensor_async + events, ggml/src/ggml-cuda/ggml-cuda.cu:2806-2866).
 * b200_peer_wait: spin (on the device) until the LOCAL 64-bit flag reaches *seq + offset (seq = tokens this shard has completed; offset 0 for
 *   the first shard, which waits for the previous step's token, 1 for later shards, which wait for this step's hidden row); status (optional)
 *   is set to 1 after a 20 s timeout instead of hanging the GPU.
 * b200_peer_send: store x[0..n) into peer_x and / or tok[0] into peer_tok (PEER memory mapped with b200_ipc_open), raise *peer_flag to
 *   *seq + 1 with a system-scope release, then *seq += 1 and (optionally) pos[0] += 1. */
int b200_peer_wait(const void * flag, const void * seq, int offset, void * status, void * stream);
int b200_peer_send(const float * x, float * peer_x, int64_t n, const int32_t * tok, int32_t * peer_tok, void * peer_flag, void * seq, int32_t * pos,
                   void * stream);
/* Device memory that another PROCESS on the same node can map (cudaIpc*): hidden-row / flag / token mailboxes of the sharded decode.
 * b200_ipc_alloc: cudaMalloc + zero-fill + export a 64-byte handle;  b200_ipc_open: map a peer's handle (peer access enabled lazily). */
int b200_ipc_alloc(size_t bytes, void ** dptr, void * handle64);
int b200_ipc_open(const void * handle64, void ** dptr);
int b200_ipc_close(void * dptr);
int b200_ipc_free(void * dptr);

#ifdef __cplusplus
}
#endif
#endif /* CHATLLM_B200_H */
