"""chatllm.cpp_b200 — B200-native (sm_100a) implementation of chatllm.cpp's quantized-matmul decode hot path.

This Python package is only the host-side mirror used by tests and bench.py: it loads the in-tree C-ABI library
(lib/libchatllm_b200.so, built from csrc/ by csrc/Makefile) with ctypes and wraps its entry points over torch
tensors (torch is used for device memory / streams only).  The product is the CUDA code in csrc/ and the ggml
backend plugin lib/libggml-cuda.so.  There is NO CPU fallback: if the extension is missing or no GPU is present
the calls raise.
"""
import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(PKG_DIR, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libchatllm_b200.so")
PLUGIN_PATH = os.path.join(LIB_DIR, "libggml-cuda.so")

# ggml type ids (reference ggml/include/ggml.h:389-405)
F32, F16, Q4_0, Q8_0, Q4_K = 0, 1, 2, 8, 12

_lib = None


class DecodeLayer(C.Structure):
    """b200_decode_layer (include/chatllm_b200.h)"""
    _fields_ = [(n, C.c_void_p) for n in ("wq", "wk", "wv", "wo", "wgate", "wup", "wdown", "bq", "bk", "bv", "attn_norm", "ffn_norm", "k_cache", "v_cache")]


class DecodeModel(C.Structure):
    """b200_decode_model"""
    _fields_ = [(n, C.c_int32) for n in ("wtype", "n_layers", "hidden", "heads", "kv_heads", "head_dim", "ffn", "vocab", "rope_mode", "embed_type")] + \
               [(n, C.c_float) for n in ("rope_theta", "eps", "attn_scale")] + [("k_row_stride", C.c_int64), ("v_row_stride", C.c_int64)] + \
               [("layers", C.POINTER(DecodeLayer)), ("embed", C.c_void_p), ("final_norm", C.c_void_p), ("lm_head", C.c_void_p), ("rope_freq_factors", C.c_void_p)]


class DecodeIO(C.Structure):
    """b200_decode_io"""
    _fields_ = [("tok", C.c_void_p), ("pos", C.c_void_p), ("n_kv", C.c_int32), ("v_col", C.c_int32), ("x", C.c_void_p), ("logits", C.c_void_p),
                ("next_tok", C.c_void_p), ("flags", C.c_int32), ("step_begin", C.c_int32), ("step_end", C.c_int32),
                ("wait_flag", C.c_void_p), ("send_x", C.c_void_p), ("send_flag", C.c_void_p), ("send_tok", C.c_void_p), ("wait_offset", C.c_int32),
                ("reserved", C.c_int32)]


class B200Error(RuntimeError):
    pass


def lib():
    """ctypes handle to libchatllm_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "or `make -C chatllm.cpp_b200/csrc`")
        L = C.CDLL(LIB_PATH)
        vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int, C.c_float
        L.b200_abi_version.restype = i32
        L.b200_device_sm_count.restype = i32
        L.b200_repack_weights.argtypes = [i32, vp, vp, i64, i64, i64, i32, vp]
        L.b200_qact_col_bytes.argtypes = [i32, i64]; L.b200_qact_col_bytes.restype = C.c_size_t
        L.b200_quantize_act.argtypes = [i32, vp, i64, i64, i64, vp, vp]
        L.b200_mul_mat_q.argtypes = [i32, vp, i64, i64, vp, i64, vp, i64, vp, vp]
        L.b200_mul_mat.argtypes = [i32, vp, i64, i64, vp, i64, i64, vp, i64, vp, vp]
        L.b200_gemv_set_tuning.argtypes = [i32] * 5
        L.b200_rms_norm.argtypes = [vp, vp, vp, i64, i64, f32, vp]
        L.b200_add.argtypes = [vp, vp, vp, i64, vp]
        L.b200_silu_mul.argtypes = [vp, vp, vp, i64, vp]
        L.b200_rope.argtypes = [vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, i64, i32, i32, i32] + [f32] * 6 + [vp]
        L.b200_soft_max.argtypes = [vp, vp, vp, i64, i64, f32, vp]
        L.b200_get_rows.argtypes = [i32, vp, i64, vp, i64, vp, vp]
        L.b200_attn_decode_scratch_bytes.argtypes = [i32, i32]; L.b200_attn_decode_scratch_bytes.restype = C.c_size_t
        L.b200_attn_decode.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, f32, vp]
        L.b200_attn_decode_quant.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i64, i64, f32, i32, vp, vp]
        L.b200_kv_store.argtypes = [vp, vp, vp, vp, i32, i64, i64, i32, vp]
        L.b200_add_rmsnorm_quant.argtypes = [i32, vp, vp, vp, vp, vp, vp, i64, i64, f32, vp]
        L.b200_rope_kv_store.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, f32, i64, i64, vp]
        L.b200_mul_mat_q_multi.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, i64, vp, i64, vp]
        L.b200_mul_mat_q_id.argtypes = [i32, i32, vp, vp, i64, i64, i32, vp, i32, vp, i32, vp, i64, vp]
        L.b200_pact_col_bytes.argtypes = [i32, i64]; L.b200_pact_col_bytes.restype = C.c_size_t
        L.b200_quantize_plain.argtypes = [i32, vp, i64, i64, i64, vp, vp]
        L.b200_mul_mat_q_batched.argtypes = [i32, vp, i64, i64, vp, i64, vp, i64, vp, vp]
        L.b200_mul_mat_q_batched_tc.argtypes = [i32, vp, i64, i64, vp, i64, vp, i64, vp, vp]
        L.b200_decode_plan_create.argtypes = [C.POINTER(DecodeModel), i32, C.POINTER(i32)]; L.b200_decode_plan_create.restype = vp
        L.b200_decode_plan_destroy.argtypes = [vp]; L.b200_decode_plan_destroy.restype = None
        L.b200_decode_plan_set_kv.argtypes = [vp, i32, vp, vp]
        L.b200_decode_plan_status.argtypes = [vp, vp]
        L.b200_decode_plan_info.argtypes = [vp] + [C.POINTER(i32)] * 5
        L.b200_decode_step.argtypes = [vp, C.POINTER(DecodeIO), vp]
        L.b200_decode_plan_times.argtypes = [vp, vp, i32, vp]
        L.b200_argmax.argtypes = [vp, i64, vp, vp]
        L.b200_peer_wait.argtypes = [vp, vp, i32, vp, vp]
        L.b200_peer_send.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, vp]
        L.b200_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(vp), vp]
        L.b200_ipc_open.argtypes = [vp, C.POINTER(vp)]
        L.b200_ipc_close.argtypes = [vp]
        L.b200_ipc_free.argtypes = [vp]
        _lib = L
    return _lib


EXPORTS = ["b200_abi_version", "b200_device_sm_count", "b200_repack_weights", "b200_qact_col_bytes", "b200_quantize_act",
           "b200_mul_mat_q", "b200_mul_mat", "b200_gemv_set_tuning", "b200_rms_norm", "b200_add", "b200_silu_mul", "b200_rope",
           "b200_soft_max", "b200_get_rows", "b200_attn_decode_scratch_bytes", "b200_attn_decode", "b200_attn_decode_quant", "b200_kv_store", "b200_add_rmsnorm_quant", "b200_rope_kv_store",
           "b200_mul_mat_q_multi", "b200_mul_mat_q_id", "b200_pact_col_bytes", "b200_quantize_plain", "b200_mul_mat_q_batched", "b200_mul_mat_q_batched_tc",
           "b200_decode_plan_create", "b200_decode_plan_destroy", "b200_decode_plan_set_kv", "b200_decode_plan_status", "b200_decode_plan_info", "b200_decode_plan_times", "b200_decode_step", "b200_argmax", "b200_peer_wait", "b200_peer_send", "b200_ipc_alloc", "b200_ipc_open", "b200_ipc_close", "b200_ipc_free"]
