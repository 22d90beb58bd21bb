"""Thin torch-tensor wrappers over the C ABI (include/chatllm_b200.h).  Device memory comes from torch; every call
runs on torch's current CUDA stream so torch.cuda.Event timing sees the kernels."""
import torch

from . import F16, F32, Q4_0, Q4_K, Q8_0, B200Error, lib

BLK = {Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144), F16: (1, 2), F32: (1, 4)}


def _chk(rc, what):
    if rc != 0:
        raise B200Error(f"{what} failed rc={rc}")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def row_size(t, k):
    b, s = BLK[t]
    return k // b * s


def upload_weights(wtype, native_bytes, k, m):
    """native ggml-layout bytes (numpy uint8 / torch uint8 CPU, m rows) -> device tensor in the device layout."""
    src = torch.as_tensor(native_bytes).reshape(-1).cuda()
    assert src.numel() == m * row_size(wtype, k)
    dst = torch.empty_like(src)
    _chk(lib().b200_repack_weights(wtype, src.data_ptr(), dst.data_ptr(), 0, src.numel(), k, 0, _stream()), "repack")
    torch.cuda.current_stream().synchronize()
    return dst


def download_weights(wtype, dev, k):
    out = torch.empty_like(dev)
    _chk(lib().b200_repack_weights(wtype, out.data_ptr(), dev.data_ptr(), 0, dev.numel(), k, 1, _stream()), "unpack")
    return out


def quantize_act(wtype, x):
    """x: float32 cuda [n, k] -> uint8 tensor [n, col_bytes] of quantized activation columns."""
    n, k = x.shape
    cb = lib().b200_qact_col_bytes(wtype, k)
    q = torch.empty((n, cb), dtype=torch.uint8, device=x.device)
    _chk(lib().b200_quantize_act(wtype, x.data_ptr(), x.stride(0), k, n, q.data_ptr(), _stream()), "quantize_act")
    return q


def mul_mat_q(wtype, w_dev, k, m, qact, n, bias=None, out=None):
    y = out if out is not None else torch.empty((n, m), dtype=torch.float32, device=w_dev.device)
    _chk(lib().b200_mul_mat_q(wtype, w_dev.data_ptr(), k, m, qact.data_ptr(), n, y.data_ptr(), y.stride(0), _p(bias), _stream()),
         "mul_mat_q")
    return y


def mul_mat_id(wtype, w0_dev, k, m, n_expert, ids, x, w1_dev=None):
    """ggml_mul_mat_id for one token.  w0_dev (w1_dev): expert stacks [n_expert, m, k] in the device layout; ids: int32 cuda
    [n_ids]; x: float32 cuda [1, k] (shared by every slot) or [n_ids, k] (one column per slot).  w1_dev given -> the experts'
    SwiGLU silu(w0 x) * (w1 x).  Returns [n_ids, m]."""
    n_ids = ids.numel()
    assert x.shape[0] in (1, n_ids)
    q = quantize_act(wtype, x)
    y = torch.empty((n_ids, m), dtype=torch.float32, device=x.device)
    _chk(lib().b200_mul_mat_q_id(wtype, 1 if w1_dev is not None else 0, w0_dev.data_ptr(), _p(w1_dev), k, m, n_expert, ids.data_ptr(), n_ids,
                                 q.data_ptr(), x.shape[0], y.data_ptr(), y.stride(0), _stream()), "mul_mat_q_id")
    return y


def mul_mat(wtype, w_dev, k, m, x, bias=None, out=None):
    n = x.shape[0]
    y = out if out is not None else torch.empty((n, m), dtype=torch.float32, device=x.device)
    _chk(lib().b200_mul_mat(wtype, w_dev.data_ptr(), k, m, x.data_ptr(), x.stride(0), n, y.data_ptr(), y.stride(0), _p(bias), _stream()),
         "mul_mat")
    return y


def rms_norm(x, w, eps):
    y = torch.empty_like(x)
    _chk(lib().b200_rms_norm(x.data_ptr(), _p(w), y.data_ptr(), x.shape[-1], x.numel() // x.shape[-1], eps, _stream()), "rms_norm")
    return y


def add(a, b):
    y = torch.empty_like(a)
    _chk(lib().b200_add(a.data_ptr(), b.data_ptr(), y.data_ptr(), a.numel(), _stream()), "add")
    return y


def silu_mul(g, u):
    y = torch.empty_like(g)
    _chk(lib().b200_silu_mul(g.data_ptr(), u.data_ptr(), y.data_ptr(), g.numel(), _stream()), "silu_mul")
    return y


def rope(x, pos, n_dims, mode, freq_base, ff=None, n_ctx_orig=0, freq_scale=1.0, ext=0.0, attn=1.0, beta_fast=32.0, beta_slow=1.0,
         inplace=False):
    """x: [n_tokens, n_heads, ne0] float32 contiguous."""
    nt, nh, ne0 = x.shape
    y = x if inplace else torch.empty_like(x)
    _chk(lib().b200_rope(x.data_ptr(), y.data_ptr(), pos.data_ptr(), _p(ff), ne0, nh, nt, x.stride(1), x.stride(0), y.stride(1), y.stride(0),
                         n_dims, mode, n_ctx_orig, freq_base, freq_scale, ext, attn, beta_fast, beta_slow, _stream()), "rope")
    return y


def soft_max(x, scale, mask=None):
    y = torch.empty_like(x)
    _chk(lib().b200_soft_max(x.data_ptr(), _p(mask), y.data_ptr(), x.shape[-1], x.numel() // x.shape[-1], scale, _stream()), "soft_max")
    return y


def get_rows(wtype, table_dev, k, ids):
    y = torch.empty((ids.numel(), k), dtype=torch.float32, device=ids.device)
    _chk(lib().b200_get_rows(wtype, table_dev.data_ptr(), k, ids.data_ptr(), ids.numel(), y.data_ptr(), _stream()), "get_rows")
    return y
