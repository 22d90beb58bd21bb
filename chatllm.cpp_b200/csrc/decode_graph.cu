// decode_graph.cu — one decode token of a dense Llama-family model as ONE replayed CUDA graph of the per-op kernels.
//
// the per-token host overhead of walking ~360 graph nodes.  The reference's CUDA backend answers it by capturing its node
// launches into a CUDA graph and re-launching it while the node properties stay equal (ggml/src/ggml-cuda/ggml-cuda.cu:2875-3071, :3993-4089) —
// which in a decode loop they never do (n_kv grows every token), so it patches kernel parameters each token.  Here the ggml graph is not
// replayed at all: the plugin's whole-token matcher (ggml_backend/ggml-b200.cu try_whole_token) reduces it to a DecodeModel (weights, caches,
// norms), and this file owns the launch sequence
//     [get_rows]  N x { add+RMSNorm+quantize, q/k/v GEMV, RoPE+KV append, attention (2 launches), o GEMV, add+RMSNorm+quantize,
//                       gate/up GEMV + SwiGLU, quantize, down GEMV }  [add+RMSNorm+quantize, lm_head GEMV]
// — the same kernels, in the same order, as the node-by-node path, so the logits are bit-identical.  Token t's sequence is captured into a
// cudaGraph on a side stream WHILE TOKEN t-1 EXECUTES (decode_graph_prepare: the only thing that changes is n_kv, which is known in advance)
// and applied to the live executable graph with cudaGraphExecUpdate (which only affects future launches), so a token costs the host three
// small copies and one cudaGraphLaunch, and the GPU runs the kernels back to back under programmatic dependent launch.
// Token ids, position and logits go through plan-owned buffers (the host's tensors may move between graphs).
#include "common.cuh"
#include "kernels.h"

#include <vector>

namespace b200 {

namespace {

struct DecodeGraph {
    DecodeModel M;
    std::vector<DecodeLayer> L;
    int max_ctx = 0;
    int device = 0;
    // plan-owned activations
    float * x = nullptr, * q = nullptr, * k = nullptr, * v = nullptr, * att = nullptr, * o = nullptr, * gate = nullptr, * logits = nullptr, * scratch = nullptr;
    uint8_t * qact = nullptr;
    int32_t * tok = nullptr, * pos = nullptr;
    cudaStream_t cap = nullptr;
    cudaGraphExec_t exec = nullptr;
    int exec_n_kv = -1;     // the n_kv the executable graph is currently parameterised for
    bool warmed = false;    // first token runs eagerly (first-use cudaFuncSetAttribute calls are not capturable)
    int attn_cluster = -1;
    long long replays = 0, recaptures = 0, reinstantiations = 0;
};

int enqueue(DecodeGraph & P, int n_kv, cudaStream_t s) {
    const DecodeModel & M = P.M;
    const int wt = M.wtype, hidden = M.hidden, kvd = M.kv_heads * M.head_dim;
    int rc = 0;
    if (M.embed) rc |= get_rows_q(M.embed_type, M.embed, hidden, P.tok, 1, P.x, s, M.lm_head ? M.vocab : 0);
    const float * pending = nullptr;
    for (size_t l = 0; l < P.L.size() && !rc; ++l) {
        const DecodeLayer & W = P.L[l];
        rc |= add_rmsnorm_quant(wt, P.x, pending, W.attn_norm, pending ? P.x : nullptr, nullptr, P.qact, hidden, 1, M.eps, s);
        {
            const void * Ws[3] = {W.wq, W.wk, W.wv};
            const int64_t ms[3] = {hidden, kvd, kvd};
            float * ys[3] = {P.q, P.k, P.v};
            const float * bs[3] = {W.bq, W.bk, W.bv};
            rc |= mul_mat_q_multi(wt, 0, 3, Ws, ms, ys, ms, bs, hidden, P.qact, 1, nullptr, s);
        }
        rc |= rope_kv_store(P.q, P.k, P.v, P.pos, M.rope_freq_factors, W.k_cache, W.v_cache, M.heads, M.kv_heads, M.head_dim, M.rope_mode, M.rope_theta,
                            M.k_row_stride, M.v_row_stride, s);
        rc |= attn_decode3(P.q, W.k_cache, W.v_cache, P.att, P.scratch, M.heads, M.kv_heads, M.head_dim, n_kv, M.k_row_stride, M.v_row_stride, M.attn_scale, wt,
                           P.qact, s, 1, P.attn_cluster);
        rc |= mul_mat_q(wt, W.wo, hidden, hidden, P.qact, 1, P.o, hidden, nullptr, nullptr, s);
        rc |= add_rmsnorm_quant(wt, P.x, P.o, W.ffn_norm, P.x, nullptr, P.qact, hidden, 1, M.eps, s);
        {
            const void * Ws[2] = {W.wgate, W.wup};
            const int64_t ms[2] = {M.ffn, M.ffn};
            float * ys[2] = {P.gate, nullptr};
            rc |= mul_mat_q_multi(wt, 1, 2, Ws, ms, ys, ms, nullptr, hidden, P.qact, 1, nullptr, s);
        }
        rc |= quantize_act(wt, P.gate, M.ffn, M.ffn, 1, P.qact, s);
        rc |= mul_mat_q(wt, W.wdown, M.ffn, hidden, P.qact, 1, P.o, hidden, nullptr, nullptr, s);
        pending = P.o;
    }
    if (rc) return rc;
    if (M.lm_head) {
        rc |= add_rmsnorm_quant(wt, P.x, pending, M.final_norm, pending ? P.x : nullptr, nullptr, P.qact, hidden, 1, M.eps, s);
        rc |= mul_mat_q(wt, M.lm_head, hidden, M.vocab, P.qact, 1, P.logits, M.vocab, nullptr, nullptr, s);
    } else if (pending) {
        rc |= add_f32(P.x, pending, P.x, hidden, s);
    }
    return rc;
}

// capture the sequence for n_kv and make it the executable graph (update in place when the topology allows it)
int capture(DecodeGraph & P, int n_kv, cudaStream_t live, bool in_flight) {
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamBeginCapture(P.cap, cudaStreamCaptureModeRelaxed);
    if (e != cudaSuccess) return (int) e;
    const int rc = enqueue(P, n_kv, P.cap);
    e = cudaStreamEndCapture(P.cap, &g);
    if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        cudaGetLastError();
        return rc ? rc : (int) (e != cudaSuccess ? e : cudaErrorUnknown);
    }
    P.recaptures++;
    if (P.exec) {
        cudaGraphExecUpdateResultInfo info;
        e = cudaGraphExecUpdate(P.exec, g, &info);
        if (e != cudaSuccess) {   // e.g. the attention's cluster width changed with n_kv: a new executable graph is needed
            cudaGetLastError();
            if (in_flight) {      // the old one may still be running: leave it alone, the next step re-captures after synchronizing
                cudaGraphDestroy(g);
                P.exec_n_kv = -1;
                return 0;
            }
            cudaStreamSynchronize(live);
            cudaGraphExecDestroy(P.exec);
            P.exec = nullptr;
        }
    }
    if (!P.exec) {
        e = cudaGraphInstantiate(&P.exec, g, 0);
        P.reinstantiations++;
        if (e != cudaSuccess) { cudaGraphDestroy(g); P.exec = nullptr; P.exec_n_kv = -1; return (int) e; }
    }
    cudaGraphDestroy(g);
    P.exec_n_kv = n_kv;
    return 0;
}

template <class T>
bool dev_alloc(T ** p, size_t n) { return cudaMalloc((void **) p, n * sizeof(T)) == cudaSuccess; }

}  // namespace

void * decode_graph_create(const DecodeModel & m, int max_ctx, int attn_cluster, int * err) {
    *err = 0;
    const int64_t hidden = m.hidden, kvd = (int64_t) m.kv_heads * m.head_dim;
    if (m.n_layers <= 0 || hidden % 256 || m.ffn % 256 || hidden > 20480 || (int64_t) m.heads * m.head_dim != hidden || max_ctx <= 0) { *err = -1; return nullptr; }
    if (m.wtype != B200_TYPE_Q4_K && m.wtype != B200_TYPE_Q4_0 && m.wtype != B200_TYPE_Q8_0) { *err = -2; return nullptr; }
    if (!attn_decode3_supported(m.heads, m.kv_heads, m.head_dim, m.k_row_stride, m.v_row_stride) || (m.rope_mode != 0 && m.rope_mode != 2) || m.head_dim % 2) {
        *err = -4;   // shapes the fused attention / RoPE kernels are not instantiated for: the node-by-node path handles them
        return nullptr;
    }
    DecodeGraph * P = new DecodeGraph();
    P->M = m;
    P->L.assign(m.layers, m.layers + m.n_layers);
    P->M.layers = nullptr;
    P->max_ctx = max_ctx;
    P->attn_cluster = attn_cluster;
    cudaGetDevice(&P->device);
    const size_t kmax = (size_t) (m.ffn > hidden ? m.ffn : hidden);
    const size_t scratch = attn_decode2_scratch_bytes(m.heads, max_ctx) + attn_decode_scratch_bytes(m.heads, max_ctx);
    bool ok = dev_alloc(&P->x, hidden) && dev_alloc(&P->q, hidden) && dev_alloc(&P->k, kvd) && dev_alloc(&P->v, kvd) && dev_alloc(&P->att, hidden) &&
              dev_alloc(&P->o, hidden) && dev_alloc(&P->gate, m.ffn) && dev_alloc(&P->qact, qact_col_bytes(m.wtype, (int64_t) kmax) + 1024) &&
              dev_alloc(&P->tok, 4) && dev_alloc(&P->pos, 4) && cudaMalloc((void **) &P->scratch, scratch) == cudaSuccess &&
              (!m.lm_head || dev_alloc(&P->logits, m.vocab)) && cudaStreamCreateWithFlags(&P->cap, cudaStreamNonBlocking) == cudaSuccess;
    if (!ok) { cudaGetLastError(); *err = -3; decode_graph_destroy(P); return nullptr; }
    cudaMemset(P->tok, 0, 16);
    cudaMemset(P->pos, 0, 16);
    return P;
}

void decode_graph_destroy(void * plan) {
    DecodeGraph * P = (DecodeGraph *) plan;
    if (!P) return;
    if (P->exec) cudaGraphExecDestroy(P->exec);
    if (P->cap) cudaStreamDestroy(P->cap);
    void * bufs[] = {P->x, P->q, P->k, P->v, P->att, P->o, P->gate, P->logits, P->scratch, P->qact, P->tok, P->pos};
    for (void * b : bufs) if (b) cudaFree(b);
    delete P;
}

int decode_graph_set_kv(void * plan, int layer, void * k_cache, void * v_cache) {
    DecodeGraph * P = (DecodeGraph *) plan;
    if (!P || layer < 0 || layer >= (int) P->L.size()) return B200_ERR_ARG;
    P->L[layer].k_cache = k_cache;
    P->L[layer].v_cache = v_cache;
    P->exec_n_kv = -1;  // the captured pointers are stale
    return 0;
}

// One token on stream st.  tok / pos / x_in / x_out / logits are the CALLER's device buffers (copied in / out): tok when the model has an
// embedding table, else x_in; logits when it has an lm_head, else x_out.  0 < n_kv <= max_ctx = pos[0] + 1.
int decode_graph_step(void * plan, const int32_t * tok, const int32_t * pos, const float * x_in, float * x_out, float * logits, int n_kv, cudaStream_t st) {
    DecodeGraph * P = (DecodeGraph *) plan;
    if (!P || n_kv <= 0 || n_kv > P->max_ctx || !pos) return B200_ERR_ARG;
    const DecodeModel & M = P->M;
    cudaError_t e = cudaSuccess;
    if (M.embed) {
        if (!tok) return B200_ERR_ARG;
        e = cudaMemcpyAsync(P->tok, tok, 4, cudaMemcpyDeviceToDevice, st);
    } else {
        if (!x_in) return B200_ERR_ARG;
        e = cudaMemcpyAsync(P->x, x_in, (size_t) M.hidden * 4, cudaMemcpyDeviceToDevice, st);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(P->pos, pos, 4, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int) e;
    int rc = 0;
    if (!P->warmed) {
        rc = enqueue(*P, n_kv, st);
        P->warmed = true;
    } else {
        if (P->exec_n_kv != n_kv) rc = capture(*P, n_kv, st, false);
        if (!rc) { rc = (int) cudaGraphLaunch(P->exec, st); P->replays++; }
    }
    if (rc) return rc;
    if (M.lm_head) {
        if (!logits) return B200_ERR_ARG;
        e = cudaMemcpyAsync(logits, P->logits, (size_t) M.vocab * 4, cudaMemcpyDeviceToDevice, st);
    } else {
        if (!x_out) return B200_ERR_ARG;
        e = cudaMemcpyAsync(x_out, P->x, (size_t) M.hidden * 4, cudaMemcpyDeviceToDevice, st);
    }
    return (int) e;
}

// Is the executable graph already parameterised for n_kv (or is this the eager first token)?  The plugin sends a token whose n_kv was not
// predicted down its node-by-node path when that keeps happening (interleaved sequences): a synchronous capture before the launch would leave
// the GPU idle, while node-by-node launches overlap their own enqueue.
bool decode_graph_ready(void * plan, int n_kv) {
    DecodeGraph * P = (DecodeGraph *) plan;
    return P && (!P->warmed || P->exec_n_kv == n_kv);
}

// Host-only: parameterise the executable graph for the NEXT token while the current one runs.
int decode_graph_prepare(void * plan, int n_kv_next, cudaStream_t live) {
    DecodeGraph * P = (DecodeGraph *) plan;
    if (!P || !P->warmed || n_kv_next <= 0 || n_kv_next > P->max_ctx || P->exec_n_kv == n_kv_next) return 0;
    return capture(*P, n_kv_next, live, true);
}

void decode_graph_stats(void * plan, long long * replays, long long * recaptures, long long * reinstantiations) {
    DecodeGraph * P = (DecodeGraph *) plan;
    *replays = P ? P->replays : 0;
    *recaptures = P ? P->recaptures : 0;
    *reinstantiations = P ? P->reinstantiations : 0;
}

}  // namespace b200
