"""Quantized block formats as numpy byte arrays + ctypes loaders for the oracle — TEST INFRASTRUCTURE.

Block layouts: ggml/src/ggml-common.h:170-176 (q4_0), :219-224 (q8_0), :288-306 (q4_K), :338-344 (q8_K).
`random_blocks` writes random *valid* blocks directly (SURVEY.md §8d route (i)): parity only needs both sides to
read the same bytes; scales are chosen so dequantized weights are ~zero-mean with std ~0.02.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")

F32, F16, Q4_0, Q8_0, Q4_K = 0, 1, 2, 8, 12
BLK = {Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), F16: (1, 2), F32: (1, 4)}
NAMES = {Q4_0: "q4_0", Q8_0: "q8_0", Q4_K: "q4_K", F16: "f16", F32: "f32"}


def row_size(t, k):
    b, s = BLK[t]
    assert k % b == 0, (t, k)
    return k // b * s


def random_blocks(t, rows, k, seed=0, rng=None):
    """uint8 array [rows, row_size] of random valid blocks of type t."""
    rng = rng or np.random.default_rng(seed)
    nb = k // BLK[t][0]
    if t == Q4_0:
        out = np.empty((rows, nb, 18), dtype=np.uint8)
        d = (rng.uniform(0.5, 1.5, size=(rows, nb)) * 4.3e-3 * rng.choice([-1.0, 1.0], size=(rows, nb))).astype(np.float16)
        out[:, :, 0:2] = d.view(np.uint8).reshape(rows, nb, 2)
        out[:, :, 2:] = rng.integers(0, 256, size=(rows, nb, 16), dtype=np.uint8)
    elif t == Q8_0:
        out = np.empty((rows, nb, 34), dtype=np.uint8)
        d = (rng.uniform(0.5, 1.5, size=(rows, nb)) * 2.7e-4).astype(np.float16)
        out[:, :, 0:2] = d.view(np.uint8).reshape(rows, nb, 2)
        out[:, :, 2:] = rng.integers(-127, 128, size=(rows, nb, 32), dtype=np.int8).view(np.uint8)
    elif t == Q4_K:
        out = np.empty((rows, nb, 144), dtype=np.uint8)
        d = (rng.uniform(0.5, 1.5, size=(rows, nb)) * 1.2e-4).astype(np.float16)
        dmin = (d.astype(np.float32) * 7.5).astype(np.float16)
        out[:, :, 0:2] = d.view(np.uint8).reshape(rows, nb, 2)
        out[:, :, 2:4] = dmin.view(np.uint8).reshape(rows, nb, 2)
        # 8 six-bit scales; mins = scales so each sub-block is ~zero-mean (w = d*sc*(q-7.5)); packed per
        # get_scale_min_k4 (ggml-quants.c:703-711)
        sc = rng.integers(1, 64, size=(rows, nb, 8), dtype=np.uint8)
        mn = sc.copy()
        s = np.zeros((rows, nb, 12), dtype=np.uint8)
        s[:, :, 0:4] = (sc[:, :, 0:4] & 63) | ((sc[:, :, 4:8] >> 4) << 6)
        s[:, :, 4:8] = (mn[:, :, 0:4] & 63) | ((mn[:, :, 4:8] >> 4) << 6)
        s[:, :, 8:12] = (sc[:, :, 4:8] & 0xF) | ((mn[:, :, 4:8] & 0xF) << 4)
        out[:, :, 4:16] = s
        out[:, :, 16:] = rng.integers(0, 256, size=(rows, nb, 128), dtype=np.uint8)
    elif t == F16:
        return (rng.standard_normal((rows, k)) * 0.02).astype(np.float16).view(np.uint8).reshape(rows, k * 2)
    else:
        raise ValueError(t)
    return out.reshape(rows, nb * BLK[t][1])


def dequant_np(t, blocks, k):
    """Vectorized numpy restatement of dequantize_row_* (ggml-quants.c:307-325, :401-414, :1352-1373).
    blocks: uint8 [rows, row_size] -> float32 [rows, k]."""
    rows = blocks.shape[0]
    nb = k // BLK[t][0]
    b = blocks.reshape(rows, nb, BLK[t][1])
    if t == Q4_0:
        d = b[:, :, 0:2].copy().view(np.float16).astype(np.float32)
        qs = b[:, :, 2:]
        lo = (qs & 0xF).astype(np.int32) - 8
        hi = (qs >> 4).astype(np.int32) - 8
        return (np.concatenate([lo, hi], axis=2).astype(np.float32) * d).reshape(rows, k)
    if t == Q8_0:
        d = b[:, :, 0:2].copy().view(np.float16).astype(np.float32)
        return (b[:, :, 2:].view(np.int8).astype(np.float32) * d).reshape(rows, k)
    if t == Q4_K:
        d = b[:, :, 0:2].copy().view(np.float16).astype(np.float32)
        dmin = b[:, :, 2:4].copy().view(np.float16).astype(np.float32)
        s = b[:, :, 4:16]
        sc = np.empty((rows, nb, 8), dtype=np.uint8); mn = np.empty_like(sc)
        sc[:, :, 0:4] = s[:, :, 0:4] & 63
        mn[:, :, 0:4] = s[:, :, 4:8] & 63
        sc[:, :, 4:8] = (s[:, :, 8:12] & 0xF) | ((s[:, :, 0:4] >> 6) << 4)
        mn[:, :, 4:8] = (s[:, :, 8:12] >> 4) | ((s[:, :, 4:8] >> 6) << 4)
        qs = b[:, :, 16:].reshape(rows, nb, 4, 32)
        q = np.stack([qs & 0xF, qs >> 4], axis=3).reshape(rows, nb, 8, 32).astype(np.float32)
        d1 = (d * sc.astype(np.float32))[..., None]
        m1 = (dmin * mn.astype(np.float32))[..., None]
        return (d1 * q - m1).reshape(rows, k)
    if t == F16:
        return blocks.view(np.float16).astype(np.float32).reshape(rows, k)
    raise ValueError(t)


# ---------------------------------------------------------------------------------------------------
_port = None


def port():
    """ctypes handle to oracle/_ref/liboracle_port.so (our C restatement)."""
    global _port
    if _port is None:
        L = C.CDLL(os.path.join(REF, "liboracle_port.so"))
        fp, vp, i64, i32 = C.POINTER(C.c_float), C.c_void_p, C.c_int64, C.c_int
        for n in ("oq_quantize_row_q4_0_ref", "oq_quantize_row_q8_0_ref", "oq_quantize_row_q8_0_x86", "oq_quantize_row_q8_K_ref"):
            getattr(L, n).argtypes = [vp, vp, i64]
        for n in ("oq_dequantize_row_q4_0", "oq_dequantize_row_q8_0", "oq_dequantize_row_q4_K"):
            getattr(L, n).argtypes = [vp, vp, i64]
        for n in ("oq_vec_dot_q4_0_q8_0", "oq_vec_dot_q8_0_q8_0", "oq_vec_dot_q4_K_q8_K"):
            getattr(L, n).argtypes = [i32, vp, vp]; getattr(L, n).restype = C.c_float
        L.oq_mul_mat.argtypes = [i32, vp, C.c_size_t, i64, i64, vp, i64, vp, i32]
        L.oq_mul_mat_id.argtypes = [i32, vp, C.c_size_t, i64, i64, i64, vp, i64, vp, i64, i64, vp, i32]
        L.oq_rms_norm.argtypes = [vp, vp, vp, i64, i64, C.c_float]
        L.oq_soft_max.argtypes = [vp, vp, vp, i64, i64, C.c_float]
        L.oq_rope.argtypes = [vp, vp, vp, vp, i64, i64, i64, i32, i32, i32] + [C.c_float] * 6
        L.oq_silu_mul.argtypes = [vp, vp, vp, i64]
        L.oq_get_rows.argtypes = [i32, vp, i64, vp, i64, vp]
        L.oq_fp16_to_fp32.argtypes = [C.c_uint16]; L.oq_fp16_to_fp32.restype = C.c_float
        L.oq_fp32_to_fp16.argtypes = [C.c_float]; L.oq_fp32_to_fp16.restype = C.c_uint16
        _port = L
    return _port


def p(a):
    return a.ctypes.data if a is not None else None


def port_mul_mat(t, w_bytes, k, m, x, variant=1):
    """x: float32 [n, k] -> y float32 [n, m] via the oracle port."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    n = x.shape[0]
    y = np.empty((n, m), dtype=np.float32)
    w = np.ascontiguousarray(w_bytes)
    rc = port().oq_mul_mat(t, p(w), row_size(t, k), k, m, p(x), n, p(y), variant)
    assert rc == 0
    return y


def port_mul_mat_id(t, as_bytes, k, m, n_expert, b, ids, variant=1):
    """as_bytes: n_expert stacked [m, row_size] matrices; b: float32 [n_tokens, nb1, k]; ids: int32 [n_tokens, n_used]
    -> float32 [n_tokens, n_used, m] via the oracle port."""
    b = np.ascontiguousarray(b, dtype=np.float32); ids = np.ascontiguousarray(ids, dtype=np.int32)
    n_tokens, nb1, _ = b.shape
    n_used = ids.shape[1]
    y = np.empty((n_tokens, n_used, m), dtype=np.float32)
    w = np.ascontiguousarray(as_bytes)
    rc = port().oq_mul_mat_id(t, p(w), row_size(t, k), k, m, n_expert, p(b), nb1, p(ids), n_used, n_tokens, p(y), variant)
    assert rc == 0, rc
    return y


_ref = {}


def ref_lib(name):
    """ctypes handle to a reference library in oracle/_ref/lib ('base' or 'cpu')."""
    if name not in _ref:
        C.CDLL(os.path.join(REF, "lib", "libggml-base.so"), mode=C.RTLD_GLOBAL)
        if name == "base":
            _ref[name] = C.CDLL(os.path.join(REF, "lib", "libggml-base.so"))
        else:
            flags = open("/proc/cpuinfo").read()
            v = "avx512" if all(f in flags for f in ("avx512f", "avx512vnni", "avx512vbmi", "avx512bw")) else "avx2"
            _ref[name] = C.CDLL(os.path.join(REF, "lib", f"libggml-cpu-{v}.so"))
            _ref[name].ggml_cpu_init()   # fills the fp16->fp32 lookup table the dot kernels use (ggml-cpu.c:3677)
    return _ref[name]


def have_ref():
    return os.path.exists(os.path.join(REF, "lib", "libggml-base.so"))


def have_port():
    return os.path.exists(os.path.join(REF, "liboracle_port.so"))
