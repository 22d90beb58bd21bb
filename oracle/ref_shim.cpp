// ref_shim.cpp — TEST INFRASTRUCTURE ONLY (lives under oracle/, built into oracle/_ref/lib/libref_shim.so).
//
// A thin C wrapper over the reference's *public* ggml API (ggml/include/ggml.h, ggml-backend.h) that lets
// the Python tests build a small ggml graph from a flat instruction list and execute it on any registered
// ggml backend device:  "CPU" (the reference's own ggml-cpu = the oracle) or "CUDA0" (our plugin, loaded
// through the very same ggml_backend_load_all_from_path() code path chatllm uses,
// src/backend.cpp:277-285 / ggml/src/ggml-backend-reg.cpp:549-570).
// It contains no arithmetic of its own and copies no reference code; it only calls the reference library.
//
// Build: see oracle/Makefile (target `shim`).
#include "ggml.h"
#include "ggml-alloc.h"
#include "ggml-backend.h"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" {

// ---- instruction encoding -------------------------------------------------------------------------
// Each instruction is 16 int64 + 8 float:  I[0]=opcode, I[1..3]=src tensor ids (or -1), I[4..15]=int params.
enum rs_opcode {
    RS_INPUT = 0,      // I[4]=ggml type, I[5..8]=ne0..ne3 ; data supplied by host (may be NULL = zeros)
    RS_MUL_MAT = 1,
    RS_ADD = 2,
    RS_MUL = 3,
    RS_RMS_NORM = 4,   // F[0]=eps
    RS_ROPE = 5,       // src a, pos b, freq_factors c(-1) ; I[4]=n_dims I[5]=mode I[6]=n_ctx_orig ; F[0..5]=freq_base,freq_scale,ext,attn,beta_fast,beta_slow ; I[7]=inplace
    RS_SOFT_MAX = 6,   // src a, mask b(-1) ; F[0]=scale F[1]=max_bias ; I[4]=inplace
    RS_SCALE = 7,      // F[0]=s ; I[4]=inplace
    RS_DIAG_MASK_INF = 8, // I[4]=n_past ; I[5]=inplace
    RS_SILU = 9,       // I[4]=inplace
    RS_GET_ROWS = 10,
    RS_SET_ROWS = 11,  // a=dst, b=src, c=ids
    RS_CPY = 12,       // a -> b
    RS_CONT = 13,
    RS_VIEW = 14,      // I[4..7]=ne0..3, I[8..10]=nb1..3 (bytes), I[11]=offset (bytes), I[12]=ndims
    RS_RESHAPE = 15,   // I[4..7]=ne
    RS_PERMUTE = 16,   // I[4..7]=axes
    RS_TRANSPOSE = 17,
    RS_MUL_MAT_ID = 18,// a=as, b=b, c=ids
    RS_TOP_K = 19,     // I[4]=k
    RS_SUM_ROWS = 20,
    RS_DIV = 21,
    RS_DUP = 22,
    RS_ADD_INPLACE = 23,
    RS_MUL_INPLACE = 24,
    RS_REPEAT = 25,    // a repeated to shape of b
    RS_CLAMP = 26,     // F[0]=min F[1]=max
    RS_ARGSORT = 27,   // I[4]=order
    RS_MUL_MAT_PREC_F32 = 28, // mul_mat + ggml_mul_mat_set_prec(F32)
};

// may be called once per directory; the order of the calls is the registration order of the devices
// NOTE: this fork registers devices lazily on the first device query (ggml-backend-reg.cpp register_devices), so no
// device query may happen before the last directory has been loaded.
int refshim_init(const char * backend_dir) {
    ggml_backend_load_all_from_path(backend_dir);
    return 0;
}

int refshim_n_devices(void) { return (int) ggml_backend_dev_count(); }

const char * refshim_device_name(int i) {
    if (i < 0 || i >= (int) ggml_backend_dev_count()) return "";
    return ggml_backend_dev_name(ggml_backend_dev_get(i));
}

const char * refshim_device_desc(int i) {
    if (i < 0 || i >= (int) ggml_backend_dev_count()) return "";
    return ggml_backend_dev_description(ggml_backend_dev_get(i));
}

size_t refshim_row_size(int type, int64_t ne) { return ggml_row_size((enum ggml_type) type, ne); }

// Run a graph.  Returns 0 on success, <0 on error, >0 = number of graph nodes the device reported it
// could NOT support (when strict != 0 nothing is executed in that case).
//   n_instr, I[n_instr*16], F[n_instr*8], in_data[n_instr] (host pointer for RS_INPUT rows, else ignored)
//   n_out, out_ids[n_out], out_data[n_out], out_bytes[n_out]
//   repeat: execute the graph this many times (timing loops / idempotence checks)
int refshim_run(const char * device, int n_instr, const int64_t * I, const float * F, const void * const * in_data,
                int n_out, const int * out_ids, void * const * out_data, const size_t * out_bytes,
                int n_threads, int strict, int repeat, double * elapsed_ms) {
    ggml_backend_dev_t dev = ggml_backend_dev_by_name(device);
    if (!dev) { fprintf(stderr, "refshim: no device named %s\n", device); return -1; }
    ggml_backend_t backend = ggml_backend_dev_init(dev, nullptr);
    if (!backend) return -2;
    if (n_threads > 0) {
        ggml_backend_reg_t reg = ggml_backend_dev_backend_reg(dev);
        auto fn = (ggml_backend_set_n_threads_t) ggml_backend_reg_get_proc_address(reg, "ggml_backend_set_n_threads");
        if (fn) fn(backend, n_threads);
    }

    ggml_init_params ip = { ggml_tensor_overhead() * (size_t)(n_instr + 8) + ggml_graph_overhead() + 4096, nullptr, true };
    ggml_context * ctx_in = ggml_init(ip);   // inputs (allocated statically)
    ggml_context * ctx_g  = ggml_init(ip);   // graph nodes (allocated by gallocr)
    std::vector<ggml_tensor *> t(n_instr, nullptr);
    int rc = 0;

    for (int i = 0; i < n_instr && rc == 0; ++i) {
        const int64_t * q = I + (size_t) i * 16;
        const float * f = F + (size_t) i * 8;
        auto S = [&](int k) -> ggml_tensor * { return q[k] >= 0 ? t[q[k]] : nullptr; };
        ggml_tensor * r = nullptr;
        switch ((rs_opcode) q[0]) {
            case RS_INPUT:  r = ggml_new_tensor_4d(ctx_in, (ggml_type) q[4], q[5], q[6], q[7], q[8]); break;
            case RS_MUL_MAT: r = ggml_mul_mat(ctx_g, S(1), S(2)); break;
            case RS_MUL_MAT_PREC_F32: r = ggml_mul_mat(ctx_g, S(1), S(2)); ggml_mul_mat_set_prec(r, GGML_PREC_F32); break;
            case RS_ADD: r = ggml_add(ctx_g, S(1), S(2)); break;
            case RS_ADD_INPLACE: r = ggml_add_inplace(ctx_g, S(1), S(2)); break;
            case RS_MUL: r = ggml_mul(ctx_g, S(1), S(2)); break;
            case RS_MUL_INPLACE: r = ggml_mul_inplace(ctx_g, S(1), S(2)); break;
            case RS_DIV: r = ggml_div(ctx_g, S(1), S(2)); break;
            case RS_RMS_NORM: r = ggml_rms_norm(ctx_g, S(1), f[0]); break;
            case RS_ROPE:
                r = q[7] ? ggml_rope_ext_inplace(ctx_g, S(1), S(2), S(3), (int) q[4], (int) q[5], (int) q[6], f[0], f[1], f[2], f[3], f[4], f[5])
                         : ggml_rope_ext(ctx_g, S(1), S(2), S(3), (int) q[4], (int) q[5], (int) q[6], f[0], f[1], f[2], f[3], f[4], f[5]);
                break;
            case RS_SOFT_MAX:
                r = q[4] ? ggml_soft_max_ext_inplace(ctx_g, S(1), S(2), f[0], f[1]) : ggml_soft_max_ext(ctx_g, S(1), S(2), f[0], f[1]);
                break;
            case RS_SCALE: r = q[4] ? ggml_scale_inplace(ctx_g, S(1), f[0]) : ggml_scale(ctx_g, S(1), f[0]); break;
            case RS_DIAG_MASK_INF: r = q[5] ? ggml_diag_mask_inf_inplace(ctx_g, S(1), (int) q[4]) : ggml_diag_mask_inf(ctx_g, S(1), (int) q[4]); break;
            case RS_SILU: r = q[4] ? ggml_silu_inplace(ctx_g, S(1)) : ggml_silu(ctx_g, S(1)); break;
            case RS_GET_ROWS: r = ggml_get_rows(ctx_g, S(1), S(2)); break;
            case RS_SET_ROWS: r = ggml_set_rows(ctx_g, S(1), S(2), S(3)); break;
            case RS_CPY: r = ggml_cpy(ctx_g, S(1), S(2)); break;
            case RS_CONT: r = ggml_cont(ctx_g, S(1)); break;
            case RS_DUP: r = ggml_dup(ctx_g, S(1)); break;
            case RS_VIEW:
                switch ((int) q[12]) {
                    case 1: r = ggml_view_1d(ctx_g, S(1), q[4], (size_t) q[11]); break;
                    case 2: r = ggml_view_2d(ctx_g, S(1), q[4], q[5], (size_t) q[8], (size_t) q[11]); break;
                    case 3: r = ggml_view_3d(ctx_g, S(1), q[4], q[5], q[6], (size_t) q[8], (size_t) q[9], (size_t) q[11]); break;
                    default: r = ggml_view_4d(ctx_g, S(1), q[4], q[5], q[6], q[7], (size_t) q[8], (size_t) q[9], (size_t) q[10], (size_t) q[11]); break;
                }
                break;
            case RS_RESHAPE: r = ggml_reshape_4d(ctx_g, S(1), q[4], q[5], q[6], q[7]); break;
            case RS_PERMUTE: r = ggml_permute(ctx_g, S(1), (int) q[4], (int) q[5], (int) q[6], (int) q[7]); break;
            case RS_TRANSPOSE: r = ggml_transpose(ctx_g, S(1)); break;
            case RS_MUL_MAT_ID: r = ggml_mul_mat_id(ctx_g, S(1), S(2), S(3)); break;
            case RS_TOP_K: r = ggml_top_k(ctx_g, S(1), (int) q[4]); break;
            case RS_ARGSORT: r = ggml_argsort(ctx_g, S(1), (ggml_sort_order) q[4]); break;
            case RS_SUM_ROWS: r = ggml_sum_rows(ctx_g, S(1)); break;
            case RS_REPEAT: r = ggml_repeat(ctx_g, S(1), S(2)); break;
            case RS_CLAMP: r = ggml_clamp(ctx_g, S(1), f[0], f[1]); break;
            default: rc = -3; break;
        }
        if (!r && rc == 0) rc = -4;
        t[i] = r;
    }

    ggml_backend_buffer_t buf_in = nullptr;
    ggml_gallocr_t galloc = nullptr;
    if (rc == 0) {
        buf_in = ggml_backend_alloc_ctx_tensors(ctx_in, backend);
        if (!buf_in) rc = -5;
    }
    if (rc == 0) {
        for (int i = 0; i < n_instr; ++i) {
            if (I[(size_t) i * 16] != RS_INPUT) continue;
            if (in_data[i]) ggml_backend_tensor_set(t[i], in_data[i], 0, ggml_nbytes(t[i]));
            else ggml_backend_tensor_memset(t[i], 0, 0, ggml_nbytes(t[i]));
        }
        ggml_cgraph * gf = ggml_new_graph_custom(ctx_g, (size_t) n_instr + 8, false);
        for (int o = 0; o < n_out; ++o) {
            ggml_set_output(t[out_ids[o]]);
            ggml_build_forward_expand(gf, t[out_ids[o]]);
        }
        int unsupported = 0;
        for (int i = 0; i < ggml_graph_n_nodes(gf); ++i) {
            ggml_tensor * nd = ggml_graph_node(gf, i);
            if (!ggml_backend_dev_supports_op(dev, nd)) {
                unsupported++;
                fprintf(stderr, "refshim: %s does not support node %d op %s (%s)\n", device, i, ggml_op_name(nd->op), nd->name);
            }
        }
        if (unsupported && strict) {
            rc = unsupported;
        } else {
            galloc = ggml_gallocr_new(ggml_backend_get_default_buffer_type(backend));
            if (!ggml_gallocr_alloc_graph(galloc, gf)) rc = -6;
            if (rc == 0) {
                int64_t t0 = ggml_time_us();
                for (int it = 0; it < (repeat > 0 ? repeat : 1) && rc == 0; ++it) {
                    if (ggml_backend_graph_compute(backend, gf) != GGML_STATUS_SUCCESS) rc = -7;
                }
                ggml_backend_synchronize(backend);
                if (elapsed_ms) *elapsed_ms = (ggml_time_us() - t0) / 1000.0;
            }
            if (rc == 0) {
                for (int o = 0; o < n_out; ++o) {
                    ggml_tensor * x = t[out_ids[o]];
                    size_t nb = ggml_nbytes(x);
                    if (nb > out_bytes[o]) { rc = -8; break; }
                    ggml_backend_tensor_get(x, out_data[o], 0, nb);
                }
            }
        }
    }
    if (galloc) ggml_gallocr_free(galloc);
    if (buf_in) ggml_backend_buffer_free(buf_in);
    ggml_free(ctx_g);
    ggml_free(ctx_in);
    ggml_backend_free(backend);
    return rc;
}

} // extern "C"
