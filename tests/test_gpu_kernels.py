"""GPU parity tests (run with -m gpu on the B200): the CUDA path, called through the C ABI
(include/chatllm_b200.h), against the oracle port (pinned in test_oracle_pin.py) and the committed golden vectors
produced by executing the reference.  Integer results bit-exact; fp32 results within the stated tolerance
(north star: logits within 1e-3 relative — per-op we hold 2e-5)."""
import os

import numpy as np
import pytest

import qformats as qf

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))


@pytest.fixture(scope="module")
def K():
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    pkg.lib()  # raises if the extension is missing — no fallback
    from chatllm_cpp_b200 import kernels
    assert torch.cuda.is_available()
    return kernels


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _acts(rng, n, k, heavy=True):
    x = rng.standard_normal((n, k)).astype(np.float32)
    if heavy:
        x[rng.random((n, k)) < 1e-3] *= 20.0
    return x


def _decode_qact(q, wtype, k):
    """device qact column bytes (layout: chatllm.cpp_b200/csrc/actlayout.cuh) -> (qs int8[k], d float32[k/G], bs int[k/32])"""
    a16 = lambda v: (v + 15) & ~15
    e = np.arange(k)
    if wtype == qf.Q4_K:
        qs_bytes = (k + 1023) // 1024 * 1024
        u, h, j, b = e >> 8, (e >> 7) & 1, (e >> 4) & 7, e & 15
        off = (u >> 2) * 1024 + j * 128 + (u & 3) * 32 + h * 16 + b
        o1 = qs_bytes; o2 = o1 + a16(k // 256 * 4)
        d = q[:, o1:o1 + k // 256 * 4].copy().view(np.float32)
        bs = q[:, o2:o2 + k // 32 * 2].copy().view(np.int16).astype(np.int32)
    else:
        blk, j, b = e >> 5, (e >> 4) & 1, e & 15
        off = (blk >> 3) * 256 + j * 128 + (blk & 7) * 16 + b
        o1 = a16(k); o2 = o1 + a16(k // 32 * 4)
        d = q[:, o1:o1 + k // 32 * 4].copy().view(np.float32)
        bs = q[:, o2:o2 + k // 32 * 4].copy().view(np.int32)
    qs = q[:, off].view(np.int8)
    return qs, d, bs


@pytest.mark.parametrize("k", [256, 4096, 14336])
def test_quantize_q8_K_bit_exact(K, k):
    rng = np.random.default_rng(k)
    x = _acts(rng, 5, k)
    x[1, :256] = 0.0                      # all-zero block
    x[2, 3] = 7.5; x[2, 200] = -7.5       # +/- tie on |max|: the FIRST occurrence decides the sign of the scale
    x[3, 10] = -9.25; x[3, 11] = 9.25
    q = K.quantize_act(qf.Q4_K, _t(x)).cpu().numpy()
    qs, d, bs = _decode_qact(q, qf.Q4_K, k)
    ref = np.zeros((5, k // 256 * 292), dtype=np.uint8)
    for r in range(5):
        qf.port().oq_quantize_row_q8_K_ref(x[r].ctypes.data, ref[r].ctypes.data, k)
    rb = ref.reshape(5, k // 256, 292)
    assert np.array_equal(d, rb[:, :, 0:4].copy().view(np.float32).reshape(5, -1))
    assert np.array_equal(qs, rb[:, :, 4:260].view(np.int8).reshape(5, k))
    bsum16 = rb[:, :, 260:].copy().view(np.int16).reshape(5, k // 256, 8, 2).astype(np.int32).sum(-1).reshape(5, -1)
    assert np.array_equal(bs, bsum16)


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("k", [256, 2048, 18944])
def test_quantize_q8_0_bit_exact(K, wtype, k):
    rng = np.random.default_rng(k + wtype)
    x = _acts(rng, 3, k)
    x[0, :32] = 0.0
    x[1, 0:32] = (np.arange(32, dtype=np.float32) + 0.5) * np.where(np.arange(32) % 2 == 0, 1.0, -1.0); x[1, 0] = 127.0  # RNE ties
    q = K.quantize_act(wtype, _t(x)).cpu().numpy()
    qs, d, bs = _decode_qact(q, wtype, k)
    ref = np.zeros((3, k // 32 * 34), dtype=np.uint8)
    for r in range(3):
        qf.port().oq_quantize_row_q8_0_x86(x[r].ctypes.data, ref[r].ctypes.data, k)
    rb = ref.reshape(3, k // 32, 34)
    assert np.array_equal(qs, rb[:, :, 2:].view(np.int8).reshape(3, k))
    assert np.array_equal(d, rb[:, :, 0:2].copy().view(np.float16).astype(np.float32).reshape(3, -1))
    assert np.array_equal(bs, rb[:, :, 2:].view(np.int8).astype(np.int32).sum(-1))


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
def test_repack_round_trip_and_windows(K, wtype):
    import torch
    rng = np.random.default_rng(3)
    k, m = 1024, 37
    w = qf.random_blocks(wtype, m, k, rng=rng)
    dev = K.upload_weights(wtype, w, k, m)
    back = K.download_weights(wtype, dev, k).cpu().numpy().reshape(w.shape)
    assert np.array_equal(back, w)
    # windowed upload at odd offsets (ggml_backend_tensor_set writes 1 MiB chunks, src/chat.cpp:1322-1338)
    import chatllm_cpp_b200 as pkg
    flat = torch.from_numpy(w.reshape(-1)).cuda()
    dst = torch.zeros_like(flat)
    cuts = [0, 7, 1000, 1001, 20011, flat.numel()]
    for a, b in zip(cuts[:-1], cuts[1:]):
        chunk = flat[a:b].clone()
        assert pkg.lib().b200_repack_weights(wtype, chunk.data_ptr(), dst.data_ptr(), a, b - a, k, 0, 0) == 0
    torch.cuda.synchronize()
    assert torch.equal(dst, dev)


SHAPES = [(256, 1), (512, 33), (1024, 48), (4096, 300), (5632, 64), (2048, 257)]


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
@pytest.mark.parametrize("k,m", SHAPES)
@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 8])
def test_mul_mat_vs_oracle(K, wtype, k, m, n):
    rng = np.random.default_rng(k * 7 + m + n)
    w = qf.random_blocks(wtype, m, k, rng=rng)
    x = _acts(rng, n, k)
    bias = rng.standard_normal(m).astype(np.float32) if (m + n) % 2 else None
    y = K.mul_mat(wtype, K.upload_weights(wtype, w, k, m), k, m, _t(x), bias=None if bias is None else _t(bias)).cpu().numpy()
    ref = qf.port_mul_mat(wtype, w, k, m, x, variant=1)
    if bias is not None:
        ref = ref + bias[None, :]
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
@pytest.mark.parametrize("n", [1, 3, 8])
def test_mul_mat_vs_reference_golden(K, wtype, n):
    """fixtures produced by the reference CPU backend itself (tests/golden/make_golden.py)"""
    w = G[f"mm_{qf.NAMES[wtype]}_w"]; x = G[f"mm_{qf.NAMES[wtype]}_n{n}_x"]; ref = G[f"mm_{qf.NAMES[wtype]}_n{n}_y"]
    k, m = x.shape[1], w.shape[0]
    y = K.mul_mat(wtype, K.upload_weights(wtype, w, k, m), k, m, _t(x)).cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
@pytest.mark.parametrize("k,m,n", [(256, 1, 9), (512, 33, 16), (1024, 48, 37), (2048, 257, 64), (4096, 300, 200), (5632, 130, 65)])
def test_mul_mat_batched_vs_oracle(K, wtype, k, m, n):
    """prompt-sized batches take the int8 tensor-core path (prefill.cu): same quantized arithmetic as the oracle"""
    rng = np.random.default_rng(k * 3 + m + n)
    w = qf.random_blocks(wtype, m, k, rng=rng)
    x = _acts(rng, n, k)
    bias = rng.standard_normal(m).astype(np.float32) if (m + n) % 2 else None
    y = K.mul_mat(wtype, K.upload_weights(wtype, w, k, m), k, m, _t(x), bias=None if bias is None else _t(bias)).cpu().numpy()
    ref = qf.port_mul_mat(wtype, w, k, m, x, variant=1)
    if bias is not None:
        ref = ref + bias[None, :]
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q4_K])
def test_mul_mat_batched_matches_gemv(K, wtype):
    """the prompt path and the decode path must agree column by column (a token's logits may not depend on batch size
    beyond fp32 summation order)"""
    rng = np.random.default_rng(11)
    k, m, n = 4096, 1024, 24
    w = qf.random_blocks(wtype, m, k, rng=rng)
    x = _acts(rng, n, k)
    wd = K.upload_weights(wtype, w, k, m)
    yb = K.mul_mat(wtype, wd, k, m, _t(x)).cpu().numpy()
    yg = np.concatenate([K.mul_mat(wtype, wd, k, m, _t(x[c:c + 8])).cpu().numpy() for c in range(0, n, 8)])
    assert np.abs(yb - yg).max() <= 2e-5 * np.abs(yg).max()


@pytest.mark.parametrize("wtype,k,m", [(qf.Q4_K, 4096, 14336), (qf.Q4_K, 14336, 4096), (qf.Q4_0, 3584, 18944), (qf.Q8_0, 2048, 5632),
                                       (qf.Q4_K, 4096, 128256)])
def test_mul_mat_full_size_semantics(K, wtype, k, m):
    """BASELINE.json's real shapes: result == dequant(W) . dequant(Q(x)) (the oracle's definition, pinned in
    test_oracle_pin.py) evaluated in float64 with the vectorized numpy dequantizers."""
    rng = np.random.default_rng(k + m)
    w = qf.random_blocks(wtype, m, k, rng=rng)
    x = _acts(rng, 1, k)
    y = K.mul_mat(wtype, K.upload_weights(wtype, w, k, m), k, m, _t(x)).cpu().numpy()[0]
    if wtype == qf.Q4_K:
        q = np.zeros((1, k // 256 * 292), dtype=np.uint8)
        qf.port().oq_quantize_row_q8_K_ref(x.ctypes.data, q.ctypes.data, k)
        qb = q.reshape(k // 256, 292)
        xd = (qb[:, 4:260].view(np.int8).astype(np.float64) * qb[:, 0:4].copy().view(np.float32).astype(np.float64)).reshape(k)
    else:
        q = np.zeros((1, k // 32 * 34), dtype=np.uint8)
        qf.port().oq_quantize_row_q8_0_x86(x.ctypes.data, q.ctypes.data, k)
        xd = qf.dequant_np(qf.Q8_0, q, k)[0].astype(np.float64)
    ref = np.empty(m)
    for r0 in range(0, m, 8192):
        ref[r0:r0 + 8192] = qf.dequant_np(wtype, w[r0:r0 + 8192], k).astype(np.float64) @ xd
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()
    # spot-check 64 rows against the scalar oracle port as well
    idx = rng.choice(m, 64, replace=False)
    ref2 = qf.port_mul_mat(wtype, w[idx], k, 64, x, variant=1)[0]
    assert np.abs(y[idx] - ref2).max() <= 2e-5 * np.abs(ref).max()


def test_rms_norm(K):
    x, w, ref = G["rms_x"], G["rms_w"], G["rms_y"]
    y = K.rms_norm(_t(x), _t(w), 1e-5).cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()
    rng = np.random.default_rng(0)
    for ne0 in (64, 4096, 5120):
        x = _acts(rng, 7, ne0); w = rng.standard_normal(ne0).astype(np.float32)
        y = K.rms_norm(_t(x), _t(w), 1e-6).cpu().numpy()
        r = np.zeros_like(x); qf.port().oq_rms_norm(x.ctypes.data, w.ctypes.data, r.ctypes.data, ne0, 7, 1e-6)
        assert np.abs(y - r).max() <= 2e-6 * np.abs(r).max()


def test_soft_max(K):
    x, ref = G["sm_x"], G["sm_y"]
    y = K.soft_max(_t(x), 0.088388).cpu().numpy()
    assert np.abs(y - ref).max() <= 1e-6
    rng = np.random.default_rng(1)
    x = (_acts(rng, 32, 4097) * 4).astype(np.float32)
    mask = np.where(rng.random((32, 4097)) < 0.1, -np.inf, 0.0).astype(np.float32)
    y = K.soft_max(_t(x), 0.125, mask=_t(mask)).cpu().numpy()
    r = np.zeros_like(x); qf.port().oq_soft_max(x.ctypes.data, mask.ctypes.data, r.ctypes.data, 4097, 32, 0.125)
    assert np.abs(y - r).max() <= 1e-6
    assert np.all(y[mask == -np.inf] == 0)


@pytest.mark.parametrize("mode,nm", [(0, "norm"), (2, "neox")])
@pytest.mark.parametrize("use_ff", [0, 1])
def test_rope(K, mode, nm, use_ff):
    x, pos, ff, ref = G[f"rope_{nm}_x"], G[f"rope_{nm}_pos"], G[f"rope_{nm}_ff"], G[f"rope_{nm}_ff{use_ff}_y"]
    y = K.rope(_t(x), _t(pos), 128, mode, 500000.0, ff=_t(ff) if use_ff else None).cpu().numpy()
    # the angle is the same fp32 recurrence as the CPU's; device cosf/sinf differ from glibc by <= 2 ulp
    assert np.abs(y - ref).max() <= 4e-6 * np.abs(ref).max()
    y2 = K.rope(_t(x).clone(), _t(pos), 128, mode, 500000.0, ff=_t(ff) if use_ff else None, inplace=True).cpu().numpy()
    assert np.array_equal(y, y2)


def test_rope_partial_dims_and_yarn(K):
    rng = np.random.default_rng(2)
    x = _acts(rng, 5 * 3, 96, heavy=False).reshape(5, 3, 96)
    pos = np.array([0, 1, 100, 4095, 8000], dtype=np.int32)
    for mode in (0, 2):
        y = K.rope(_t(x), _t(pos), 64, mode, 10000.0, n_ctx_orig=4096, freq_scale=0.25, ext=1.0, attn=1.1).cpu().numpy()
        r = np.zeros_like(x)
        qf.port().oq_rope(x.ctypes.data, r.ctypes.data, pos.ctypes.data, None, 96, 3, 5, 64, mode, 4096, 10000.0, 0.25, 1.0, 1.1, 32.0, 1.0)
        # the YaRN mix theta_interp*(1-ramp) + theta_extrap*ramp is FMA-contracted by nvcc (and by gcc in the reference
        # build) but not in the -ffp-contract=off port: 1 ulp of a ~1e3 rad angle is ~6e-5 rad -> tolerance 1e-4
        assert np.abs(y - r).max() <= 1e-4 * np.abs(r).max()


def test_silu_mul_and_add(K):
    g, u, ref = G["silu_g"], G["silu_u"], G["silu_y"]
    y = K.silu_mul(_t(g), _t(u)).cpu().numpy()
    assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.array_equal(K.add(_t(g), _t(u)).cpu().numpy(), g + u)


@pytest.mark.parametrize("wtype", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
def test_get_rows(K, wtype):
    rng = np.random.default_rng(4)
    k, rows = 1024, 50
    w = qf.random_blocks(wtype, rows, k, rng=rng)
    ids = np.array([0, 49, 7, 7, 13], dtype=np.int32)
    y = K.get_rows(wtype, K.upload_weights(wtype, w, k, rows), k, _t(ids)).cpu().numpy()
    ref = qf.dequant_np(wtype, w[ids], k)
    assert np.abs(y - ref).max() <= 1e-6 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("ne0", [256, 2048, 4096, 5632, 14336, 18944])
def test_add_rmsnorm_quant_matches_unfused(K, wtype, ne0):
    """fused ADD+RMS_NORM+MUL+quantize == the separate kernels (codes bit-exact, floats exact)"""
    import torch
    import chatllm_cpp_b200 as pkg
    rng = np.random.default_rng(ne0 + wtype)
    x, r = _acts(rng, 3, ne0), _acts(rng, 3, ne0)
    r[1, :256] = -x[1, :256]            # an all-zero block after the add
    w = (1 + 0.1 * rng.standard_normal(ne0)).astype(np.float32)
    xd, rd, wd = _t(x), _t(r), _t(w)
    xs = K.add(xd, rd)
    y_ref = K.rms_norm(xs, wd, 1e-5)
    q_ref = K.quantize_act(wtype, y_ref)
    x_out = torch.empty_like(xd); y_out = torch.empty_like(xd); q = torch.zeros_like(q_ref)
    rc = pkg.lib().b200_add_rmsnorm_quant(wtype, xd.data_ptr(), rd.data_ptr(), wd.data_ptr(), x_out.data_ptr(), y_out.data_ptr(), q.data_ptr(), ne0, 3, 1e-5, 0)
    assert rc == 0
    torch.cuda.synchronize()
    assert torch.equal(x_out, xs)
    # the block-sum association differs between the two norm kernels: y may differ by 1 ulp, codes by <= 1 on a tie
    assert (y_out - y_ref).abs().max().item() <= 2e-6 * y_ref.abs().max().item()
    a, b = _decode_qact(q.cpu().numpy(), wtype, ne0), _decode_qact(q_ref.cpu().numpy(), wtype, ne0)
    assert np.abs(a[0].astype(int) - b[0].astype(int)).max() <= 1
    assert (a[0] == b[0]).mean() > 0.999
    assert np.allclose(a[1], b[1], rtol=1e-6)
    # and bit-exact against the oracle quantizer applied to the fused kernel's own float output
    yo = y_out.cpu().numpy()
    if wtype == qf.Q4_K:
        ref = np.zeros((3, ne0 // 256 * 292), dtype=np.uint8)
        for i in range(3):
            qf.port().oq_quantize_row_q8_K_ref(yo[i].ctypes.data, ref[i].ctypes.data, ne0)
        assert np.array_equal(a[0], ref.reshape(3, -1, 292)[:, :, 4:260].view(np.int8).reshape(3, ne0))
    else:
        ref = np.zeros((3, ne0 // 32 * 34), dtype=np.uint8)
        for i in range(3):
            qf.port().oq_quantize_row_q8_0_x86(yo[i].ctypes.data, ref[i].ctypes.data, ne0)
        assert np.array_equal(a[0], ref.reshape(3, -1, 34)[:, :, 2:].view(np.int8).reshape(3, ne0))


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
def test_mul_mat_multi_concat_and_paired(K, wtype):
    import ctypes as C
    import torch
    import chatllm_cpp_b200 as pkg
    rng = np.random.default_rng(5 + wtype)
    k = 1024
    ms = [512, 128, 128]
    ws = [qf.random_blocks(wtype, m, k, rng=rng) for m in ms]
    x = _acts(rng, 1, k)
    biases = [rng.standard_normal(m).astype(np.float32) for m in ms]
    wd = [K.upload_weights(wtype, w, k, m) for w, m in zip(ws, ms)]
    bd = [_t(b) for b in biases]
    q = K.quantize_act(wtype, _t(x))
    ys = [torch.zeros((1, m), device="cuda") for m in ms]
    arr = lambda ct, v: (ct * len(v))(*v)
    rc = pkg.lib().b200_mul_mat_q_multi(wtype, 0, 3, arr(C.c_void_p, [w.data_ptr() for w in wd]), arr(C.c_int64, ms), arr(C.c_void_p, [y.data_ptr() for y in ys]),
                                        arr(C.c_int64, ms), arr(C.c_void_p, [b.data_ptr() for b in bd]), k, q.data_ptr(), 1, 0)
    assert rc == 0
    for w, m, b, y in zip(ws, ms, biases, ys):
        ref = qf.port_mul_mat(wtype, w, k, m, x)[0] + b
        assert np.abs(y.cpu().numpy()[0] - ref).max() <= 2e-5 * np.abs(ref).max()
    # paired SwiGLU: silu(Wg x) * (Wu x)
    m = 768
    wg, wu = qf.random_blocks(wtype, m, k, rng=rng), qf.random_blocks(wtype, m, k, rng=rng)
    dg, du = K.upload_weights(wtype, wg, k, m), K.upload_weights(wtype, wu, k, m)
    y = torch.zeros((1, m), device="cuda")
    rc = pkg.lib().b200_mul_mat_q_multi(wtype, 1, 2, arr(C.c_void_p, [dg.data_ptr(), du.data_ptr()]), arr(C.c_int64, [m, m]), arr(C.c_void_p, [y.data_ptr(), 0]),
                                        arr(C.c_int64, [m, m]), arr(C.c_void_p, [0, 0]), k, q.data_ptr(), 1, 0)
    assert rc == 0
    g = qf.port_mul_mat(wtype, wg, k, m, x); u = qf.port_mul_mat(wtype, wu, k, m, x)
    ref = np.zeros_like(g); qf.port().oq_silu_mul(g.ctypes.data, u.ctypes.data, ref.ctypes.data, m)
    assert np.abs(y.cpu().numpy() - ref).max() <= 3e-5 * np.abs(ref).max()


@pytest.mark.parametrize("heads,kvh,hd", [(32, 8, 128), (32, 4, 64), (28, 4, 128), (4, 2, 64)])
@pytest.mark.parametrize("n_kv", [1, 7, 128, 129, 1000, 4097])
def test_attn_decode_vs_reference_semantics(K, heads, kvh, hd, n_kv):
    """scores = K.f16(q) (fp32 acc) * scale -> softmax -> f16(P) -> V.P, on the reference's cache layouts
    (src/layers.cpp:2541-2561, ggml-cpu.c:213-219), evaluated in float64 with explicit f16 roundings."""
    import torch
    import chatllm_cpp_b200 as pkg
    rng = np.random.default_rng(n_kv + heads)
    max_len = 4352
    kv_hidden = kvh * hd
    kc = rng.standard_normal((max_len, kv_hidden)).astype(np.float16)
    vc = rng.standard_normal((kv_hidden, max_len)).astype(np.float16)
    q = rng.standard_normal((heads, hd)).astype(np.float32) * 2
    scale = 1.0 / np.sqrt(hd)
    L = pkg.lib()
    out = torch.zeros((heads, hd), device="cuda")
    scratch = torch.empty(L.b200_attn_decode_scratch_bytes(heads, max_len) // 4 + 16, dtype=torch.float32, device="cuda")
    qd, kd, vd = _t(q), _t(kc), _t(vc)
    rc = L.b200_attn_decode(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out.data_ptr(), scratch.data_ptr(), heads, kvh, hd, n_kv,
                            kv_hidden, max_len, scale, 0)
    assert rc == 0
    torch.cuda.synchronize()
    g = heads // kvh
    q16 = q.astype(np.float16).astype(np.float64)
    ref = np.zeros((heads, hd))
    for h in range(heads):
        kk = kc[:n_kv, (h // g) * hd:(h // g + 1) * hd].astype(np.float64)
        s = (kk @ q16[h]).astype(np.float32) * np.float32(scale)
        e = np.exp((s - s.max()).astype(np.float32)).astype(np.float32)
        p = (e / np.float32(e.astype(np.float64).sum())).astype(np.float16).astype(np.float64)
        vv = vc[(h // g) * hd:(h // g + 1) * hd, :n_kv].astype(np.float64)
        ref[h] = vv @ p
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() <= 2e-3 * np.abs(ref).max() + 1e-6   # f16 rounding of P can flip on an ulp: ~1e-3 of one term


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0])
@pytest.mark.parametrize("heads,kvh,hd,n_kv", [(32, 8, 128, 4097), (32, 8, 128, 300), (28, 4, 128, 1000), (32, 4, 64, 777), (4, 2, 64, 9000)])
def test_attn_decode_quant_tail_emits_the_o_proj_activations(K, wtype, heads, kvh, hd, n_kv):
    """b200_attn_decode_quant: `out` equals b200_attn_decode's, and qact is bit-identical to quantizing that `out` with the
    stand-alone quantizer (which test_quantize_* pin to the reference bit for bit).  n_kv = 9000 exercises multi-round spans."""
    import torch
    import chatllm_cpp_b200 as pkg
    rng = np.random.default_rng(n_kv + heads + wtype)
    max_len = ((n_kv + 255) // 256) * 256 + 256
    kv_hidden = kvh * hd
    kd = _t(rng.standard_normal((max_len, kv_hidden)).astype(np.float16))
    vd = _t(rng.standard_normal((kv_hidden, max_len)).astype(np.float16))
    vd[:, n_kv:] = float("nan")   # beyond the context the cache may hold anything
    qd = _t(rng.standard_normal((heads, hd)).astype(np.float32) * 2)
    L = pkg.lib()
    scratch = torch.empty(L.b200_attn_decode_scratch_bytes(heads, max_len) // 4 + 16, dtype=torch.float32, device="cuda")
    out_a = torch.zeros((1, heads * hd), device="cuda"); out_b = torch.zeros_like(out_a)
    scale = 1.0 / np.sqrt(hd)
    assert L.b200_attn_decode(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out_a.data_ptr(), scratch.data_ptr(), heads, kvh, hd, n_kv, kv_hidden, max_len,
                              scale, 0) == 0
    q = torch.zeros(L.b200_qact_col_bytes(wtype, heads * hd), dtype=torch.uint8, device="cuda")
    assert L.b200_attn_decode_quant(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out_b.data_ptr(), scratch.data_ptr(), heads, kvh, hd, n_kv, kv_hidden,
                                    max_len, scale, wtype, q.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    assert torch.isfinite(out_b).all()
    assert torch.equal(out_a, out_b)
    k = heads * hd
    got = _decode_qact(q.cpu().numpy()[None, :], wtype, k)
    want = _decode_qact(K.quantize_act(wtype, out_b).cpu().numpy(), wtype, k)   # padding bytes of the layout are not compared
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("k,m,n_expert,n_used", [(512, 48, 8, 2), (4096, 14336, 8, 2), (14336, 4096, 8, 2), (2048, 1408, 60, 4), (1024, 6, 8, 3)])
def test_mul_mat_id_vs_oracle(K, wtype, k, m, n_expert, n_used):
    """expert-indexed GEMV (ggml_mul_mat_id, one token) against the oracle port (pinned to the reference's CPU backend in
    test_oracle_pin.py): shared activation column (gate / up) and one column per slot (down)."""
    import torch
    rng = np.random.default_rng(k + m + wtype)
    pool = qf.random_blocks(wtype, min(m, 512) * 2, k, rng=rng)
    as_ = pool[rng.integers(0, pool.shape[0], size=n_expert * m)]
    wd = K.upload_weights(wtype, as_, k, n_expert * m)
    ids = rng.choice(n_expert, size=n_used, replace=False).astype(np.int32)
    for nb1 in (1, n_used):
        x = _acts(rng, nb1, k)
        y = K.mul_mat_id(wtype, wd, k, m, n_expert, _t(ids), _t(x)).cpu().numpy()
        ref = qf.port_mul_mat_id(wtype, as_, k, m, n_expert, x[None], ids[None])[0]
        assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0])
@pytest.mark.parametrize("k,m", [(4096, 14336), (512, 96)])
def test_mul_mat_id_paired_swiglu_vs_oracle(K, wtype, k, m):
    """MultiMLP::forward's gate/act/up/mul for the selected experts in one launch (src/layers.cpp:3674-3688)."""
    import torch
    rng = np.random.default_rng(k + m + wtype + 1)
    n_expert, n_used = 8, 2
    pool = qf.random_blocks(wtype, 1024, k, rng=rng)
    wg = pool[rng.integers(0, 1024, size=n_expert * m)]; wu = pool[rng.integers(0, 1024, size=n_expert * m)]
    dg, du = K.upload_weights(wtype, wg, k, n_expert * m), K.upload_weights(wtype, wu, k, n_expert * m)
    ids = np.array([5, 2], dtype=np.int32)
    x = _acts(rng, 1, k)
    y = K.mul_mat_id(wtype, dg, k, m, n_expert, _t(ids), _t(x), w1_dev=du).cpu().numpy()
    g = qf.port_mul_mat_id(wtype, wg, k, m, n_expert, x[None], ids[None])[0]
    u = qf.port_mul_mat_id(wtype, wu, k, m, n_expert, x[None], ids[None])[0]
    ref = np.zeros_like(g); qf.port().oq_silu_mul(g.ctypes.data, u.ctypes.data, ref.ctypes.data, g.size)
    assert np.abs(y - ref).max() <= 3e-5 * np.abs(ref).max()


GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_moe_vectors.npz"))


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("tag", ["bcast", "slot"])
def test_mul_mat_id_vs_reference_golden(K, wtype, tag):
    """expert-indexed GEMV against outputs stored from the reference CPU backend's ggml_mul_mat_id (make_golden.py main_moe)"""
    nm = qf.NAMES[wtype]
    as_, b, ids, ref = GM[f"mmid_{nm}_as"], GM[f"mmid_{nm}_{tag}_b"], GM["mmid_ids"], GM[f"mmid_{nm}_{tag}_y"]
    k, m, n_expert = b.shape[2], ref.shape[2], 8
    wd = K.upload_weights(wtype, as_, k, n_expert * m)
    for t in range(b.shape[0]):
        y = K.mul_mat_id(wtype, wd, k, m, n_expert, _t(ids[t]), _t(b[t])).cpu().numpy()
        assert np.abs(y - ref[t]).max() <= 2e-5 * np.abs(ref).max()


def test_moe_experts_swiglu_vs_reference_golden(K):
    """the paired expert launch (gate / silu / up / mul of MultiMLP::forward) against the reference's `par` tensor, experts as
    chosen by the reference's own top_k"""
    wg, wu, x, sel, ref = GM["moe_wg"], GM["moe_wu"], GM["moe_x"], GM["moe_sel"][0], GM["moe_par"][0]
    hidden, ffn, n_expert = x.shape[1], ref.shape[1], 8
    dg, du = K.upload_weights(qf.Q4_K, wg, hidden, n_expert * ffn), K.upload_weights(qf.Q4_K, wu, hidden, n_expert * ffn)
    y = K.mul_mat_id(qf.Q4_K, dg, hidden, ffn, n_expert, _t(sel.astype(np.int32)), _t(x), w1_dev=du).cpu().numpy()
    assert np.abs(y - ref).max() <= 3e-5 * np.abs(ref).max()


@pytest.mark.parametrize("n", [1, 1000, 32000, 128256, 152064])
def test_argmax_first_maximum(n):
    """b200_argmax = the host's greedy pick (src/models.cpp:1026-1031: first maximum), ties and negative values included"""
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n).astype(np.float32)
    if n > 10:
        x[rng.integers(0, n, 3)] = x.max()          # ties: the first one must win
    out = torch.full((1,), -1, dtype=torch.int32, device="cuda")
    xd = torch.from_numpy(x).cuda()
    for _ in range(2):                                # twice: the scratch re-arms itself
        assert pkg.lib().b200_argmax(xd.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
        torch.cuda.synchronize()
        assert int(out.item()) == int(np.argmax(x))
    xn = -np.abs(x) - 1.0
    xnd = torch.from_numpy(xn).cuda()
    assert pkg.lib().b200_argmax(xnd.data_ptr(), n, out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert int(out.item()) == int(np.argmax(xn))
