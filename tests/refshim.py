"""Python side of oracle/ref_shim.cpp — TEST INFRASTRUCTURE.

Builds small ggml graphs from a flat instruction list and runs them on a named ggml backend device of the
reference build in oracle/_ref ("CPU" = the reference's own ggml-cpu backend = the oracle;
"CUDA0" = our plugin libggml-cuda.so loaded through the reference's registry).
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "lib")
PLUGIN_DIR = os.path.join(ROOT, "chatllm.cpp_b200", "lib")

# ggml type ids (ggml/include/ggml.h:389-405)
F32, F16, Q4_0, Q8_0, Q4_K, I32 = 0, 1, 2, 8, 12, 26

(INPUT, MUL_MAT, ADD, MUL, RMS_NORM, ROPE, SOFT_MAX, SCALE, DIAG_MASK_INF, SILU, GET_ROWS, SET_ROWS, CPY, CONT, VIEW,
 RESHAPE, PERMUTE, TRANSPOSE, MUL_MAT_ID, TOP_K, SUM_ROWS, DIV, DUP, ADD_INPLACE, MUL_INPLACE, REPEAT, CLAMP, ARGSORT,
 MUL_MAT_PREC_F32) = range(29)

_TYPE_NP = {F32: np.float32, F16: np.float16, I32: np.int32}
_BLK = {Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), F32: (1, 4), F16: (1, 2), I32: (1, 4)}


def row_size(t, ne0):
    b, s = _BLK[t]
    assert ne0 % b == 0
    return ne0 // b * s


def available():
    return os.path.exists(os.path.join(REF_LIB, "libref_shim.so"))


_lib = None


def lib(backend_dirs=None):
    """Load the shim and register backends.  CPU variants live in oracle/_ref/lib; our plugin (if built)
    in chatllm.cpp_b200/lib.  The plugin must be registered BEFORE the cpu backend (src/backend.cpp:727-733),
    so its directory is loaded first."""
    global _lib
    if _lib is None:
        C.CDLL(os.path.join(REF_LIB, "libggml-base.so"), mode=C.RTLD_GLOBAL)
        C.CDLL(os.path.join(REF_LIB, "libggml.so"), mode=C.RTLD_GLOBAL)
        L = C.CDLL(os.path.join(REF_LIB, "libref_shim.so"))
        L.refshim_init.argtypes = [C.c_char_p]
        L.refshim_device_name.restype = C.c_char_p
        L.refshim_device_desc.restype = C.c_char_p
        L.refshim_run.argtypes = [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
        dirs = backend_dirs
        if dirs is None:
            dirs = []
            if os.path.exists(os.path.join(PLUGIN_DIR, "libggml-cuda.so")) and _have_gpu():
                dirs.append(PLUGIN_DIR)
            dirs.append(REF_LIB)
        for d in dirs:
            L.refshim_init(d.encode())
        _lib = L
    return _lib


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def devices():
    L = lib()
    return [L.refshim_device_name(i).decode() for i in range(L.refshim_n_devices())]


class T:
    """Handle to a tensor in a Graph (index + shape bookkeeping done by the caller)."""
    def __init__(self, g, idx):
        self.g, self.idx = g, idx


class Graph:
    def __init__(self):
        self.I, self.F, self.data, self.keep = [], [], [], []

    def _emit(self, op, srcs=(), ip=(), fp=(), data=None):
        row = [op] + [(-1 if s is None else s.idx) for s in srcs] + [-1] * (3 - len(srcs)) + list(ip)
        row += [0] * (16 - len(row))
        f = list(fp) + [0.0] * (8 - len(fp))
        self.I.append(row); self.F.append(f); self.data.append(data)
        return T(self, len(self.I) - 1)

    # ---- leaves
    def input(self, arr_or_bytes, gtype=None, ne=None):
        """numpy array (F32/F16/I32; shape reversed = ggml ne order) or raw bytes for quantized types."""
        if isinstance(arr_or_bytes, np.ndarray) and gtype is None:
            a = np.ascontiguousarray(arr_or_bytes)
            gtype = {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.int32): I32}[a.dtype]
            ne = list(a.shape[::-1]) + [1] * (4 - a.ndim)
            buf = a
        else:
            buf = arr_or_bytes if arr_or_bytes is None else np.frombuffer(bytes(arr_or_bytes), dtype=np.uint8).copy() \
                if not isinstance(arr_or_bytes, np.ndarray) else np.ascontiguousarray(arr_or_bytes)
            ne = list(ne) + [1] * (4 - len(ne))
        self.keep.append(buf)
        return self._emit(INPUT, (), [gtype] + ne, (), buf)

    # ---- ops
    def mul_mat(self, a, b, prec_f32=False): return self._emit(MUL_MAT_PREC_F32 if prec_f32 else MUL_MAT, (a, b))
    def add(self, a, b, inplace=False): return self._emit(ADD_INPLACE if inplace else ADD, (a, b))
    def mul(self, a, b, inplace=False): return self._emit(MUL_INPLACE if inplace else MUL, (a, b))
    def div(self, a, b): return self._emit(DIV, (a, b))
    def rms_norm(self, a, eps): return self._emit(RMS_NORM, (a,), (), (eps,))
    def rope(self, a, pos, n_dims, mode, freq_base, ff=None, n_ctx_orig=0, freq_scale=1.0, ext=0.0, attn=1.0,
             beta_fast=32.0, beta_slow=1.0, inplace=False):
        return self._emit(ROPE, (a, pos, ff), (n_dims, mode, n_ctx_orig, int(inplace)),
                          (freq_base, freq_scale, ext, attn, beta_fast, beta_slow))
    def soft_max(self, a, mask=None, scale=1.0, max_bias=0.0, inplace=False):
        return self._emit(SOFT_MAX, (a, mask), (int(inplace),), (scale, max_bias))
    def scale(self, a, s, inplace=False): return self._emit(SCALE, (a,), (int(inplace),), (s,))
    def diag_mask_inf(self, a, n_past, inplace=False): return self._emit(DIAG_MASK_INF, (a,), (n_past, int(inplace)))
    def silu(self, a, inplace=False): return self._emit(SILU, (a,), (int(inplace),))
    def get_rows(self, a, ids): return self._emit(GET_ROWS, (a, ids))
    def set_rows(self, dst, src, ids): return self._emit(SET_ROWS, (dst, src, ids))
    def cpy(self, a, b): return self._emit(CPY, (a, b))
    def cont(self, a): return self._emit(CONT, (a,))
    def dup(self, a): return self._emit(DUP, (a,))
    def view(self, a, ne, nb=(0, 0, 0), offset=0):
        ne4 = list(ne) + [1] * (4 - len(ne)); nb3 = list(nb) + [0] * (3 - len(nb))
        return self._emit(VIEW, (a,), ne4 + nb3 + [offset, len(ne)])
    def reshape(self, a, ne): return self._emit(RESHAPE, (a,), list(ne) + [1] * (4 - len(ne)))
    def permute(self, a, axes): return self._emit(PERMUTE, (a,), list(axes))
    def transpose(self, a): return self._emit(TRANSPOSE, (a,))
    def mul_mat_id(self, as_, b, ids): return self._emit(MUL_MAT_ID, (as_, b, ids))
    def top_k(self, a, k): return self._emit(TOP_K, (a,), (k,))
    def argsort(self, a, order=1): return self._emit(ARGSORT, (a,), (order,))
    def sum_rows(self, a): return self._emit(SUM_ROWS, (a,))
    def repeat(self, a, b): return self._emit(REPEAT, (a, b))
    def clamp(self, a, lo, hi): return self._emit(CLAMP, (a,), (), (lo, hi))

    def run(self, device, outputs, n_threads=4, strict=True, repeat=1):
        """outputs: list of (T, numpy dtype, shape).  Returns list of numpy arrays (+ elapsed ms)."""
        L = lib()
        n = len(self.I)
        I = np.asarray(self.I, dtype=np.int64)
        F = np.asarray(self.F, dtype=np.float32)
        ptrs = (C.c_void_p * n)()
        for i, d in enumerate(self.data):
            ptrs[i] = d.ctypes.data if d is not None else None
        outs = [np.zeros(shape, dtype=dt) for (_, dt, shape) in outputs]
        out_ids = np.asarray([t.idx for (t, _, _) in outputs], dtype=np.int32)
        out_ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        out_bytes = (C.c_size_t * len(outs))(*[o.nbytes for o in outs])
        ms = C.c_double(0.0)
        rc = L.refshim_run(device.encode(), n, I.ctypes.data, F.ctypes.data, C.cast(ptrs, C.c_void_p), len(outs),
                           out_ids.ctypes.data, C.cast(out_ptrs, C.c_void_p), C.cast(out_bytes, C.c_void_p),
                           n_threads, int(strict), repeat, C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"refshim_run({device}) failed rc={rc}")
        self.elapsed_ms = ms.value
        return outs
