"""Pins the oracle port (oracle/oracle_port.c) against (a) committed golden vectors produced by executing the
reference (tests/golden/make_golden.py) and (b) the reference libraries themselves when oracle/_ref/lib exists.
CPU only.  Bit-exact for integer/byte results; fp32 results may differ by summation order only (tolerances stated).
"""
import ctypes as C
import os

import numpy as np
import pytest

import qformats as qf

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_vectors.npz"))

pytestmark = pytest.mark.skipif(not qf.have_port(), reason="oracle port not built (make -C oracle port)")


def _quant(fn, x, bs, blk):
    y = np.zeros((x.shape[0], x.shape[1] // blk * bs), dtype=np.uint8)
    for r in range(x.shape[0]):
        fn(x[r].ctypes.data, y[r].ctypes.data, x.shape[1])
    return y


@pytest.mark.parametrize("name,fn,bs,blk", [
    ("q4_0_ref", "oq_quantize_row_q4_0_ref", 18, 32),
    ("q8_0_ref", "oq_quantize_row_q8_0_ref", 34, 32),
    ("q8_0_x86", "oq_quantize_row_q8_0_x86", 34, 32),
    ("q8_K_ref", "oq_quantize_row_q8_K_ref", 292, 256),
])
def test_quantizers_bit_exact_vs_golden(name, fn, bs, blk):
    x = np.ascontiguousarray(G["quant_x"])
    y = _quant(getattr(qf.port(), fn), x, bs, blk)
    ref = G["quant_" + name].copy()
    if name == "q8_K_ref":
        # the reference leaves bsums untouched for an all-zero block (ggml-quants.c:2570-2575); mask them
        yb, rb = y.reshape(x.shape[0], -1, 292), ref.reshape(x.shape[0], -1, 292)
        zero = (rb[:, :, 0:4].copy().view(np.float32) == 0).reshape(rb.shape[0], rb.shape[1])
        yb[zero, 260:] = 0; rb[zero, 260:] = 0
    assert np.array_equal(y, ref)


def test_q8_0_variants_differ_only_where_expected():
    # the x86 quantizer (RNE, id=127/amax) and the _ref quantizer (roundf, id=1/d) are different functions;
    # the fixture row 2 contains exact-.5 products so the two rounding modes must actually disagree there.
    a, b = G["quant_q8_0_ref"], G["quant_q8_0_x86"]
    assert not np.array_equal(a, b)
    diff = (a.view(np.int8).astype(int) - b.view(np.int8).astype(int))
    assert np.abs(diff).max() <= 1


@pytest.mark.parametrize("t,fn", [(qf.Q4_0, "oq_dequantize_row_q4_0"), (qf.Q8_0, "oq_dequantize_row_q8_0"),
                                  (qf.Q4_K, "oq_dequantize_row_q4_K")])
def test_dequantizers_bit_exact_vs_golden(t, fn):
    w = np.ascontiguousarray(G[f"deq_{qf.NAMES[t]}_w"]); ref = G[f"deq_{qf.NAMES[t]}_y"]
    y = np.zeros_like(ref)
    for r in range(w.shape[0]):
        getattr(qf.port(), fn)(w[r].ctypes.data, y[r].ctypes.data, ref.shape[1])
    assert np.array_equal(y, ref)
    # and the vectorized numpy restatement used for large-size checks
    assert np.array_equal(qf.dequant_np(t, w, ref.shape[1]), ref)


@pytest.mark.parametrize("t", [qf.Q4_0, qf.Q8_0, qf.Q4_K])
@pytest.mark.parametrize("n", [1, 3, 8])
def test_mul_mat_vs_golden(t, n):
    """Integer dots are exact; only fp32 association differs between the scalar port and the reference's SIMD
    kernels -> tolerance 2e-5 of the output scale (north star allows 1e-3)."""
    w = G[f"mm_{qf.NAMES[t]}_w"]; x = G[f"mm_{qf.NAMES[t]}_n{n}_x"]; ref = G[f"mm_{qf.NAMES[t]}_n{n}_y"]
    y = qf.port_mul_mat(t, w, x.shape[1], w.shape[0], x, variant=1)
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


def test_mul_mat_matches_dequant_times_dequant():
    """Semantics check (SURVEY.md §8a parity note): result == dequant(W) . dequant(Q8(x))."""
    rng = np.random.default_rng(5)
    k, m = 512, 16
    for t in (qf.Q4_0, qf.Q8_0):
        w = qf.random_blocks(t, m, k, rng=rng)
        x = rng.standard_normal((1, k)).astype(np.float32)
        q = np.zeros((1, k // 32 * 34), dtype=np.uint8)
        qf.port().oq_quantize_row_q8_0_x86(x.ctypes.data, q.ctypes.data, k)
        xd = qf.dequant_np(qf.Q8_0, q, k)
        ref = (qf.dequant_np(t, w, k).astype(np.float64) @ xd[0].astype(np.float64))
        y = qf.port_mul_mat(t, w, k, m, x, variant=1)[0]
        assert np.abs(y - ref).max() <= 1e-5 * np.abs(ref).max() + 1e-7


def test_rms_norm_vs_golden():
    x, w, ref = (np.ascontiguousarray(G[k]) for k in ("rms_x", "rms_w", "rms_y"))
    y = np.zeros_like(ref)
    qf.port().oq_rms_norm(x.ctypes.data, w.ctypes.data, y.ctypes.data, x.shape[1], x.shape[0], 1e-5)
    assert np.abs(y - ref).max() <= 1e-6 * np.abs(ref).max()


def test_soft_max_vs_golden():
    x, ref = np.ascontiguousarray(G["sm_x"]), G["sm_y"]
    y = np.zeros_like(ref)
    qf.port().oq_soft_max(x.ctypes.data, None, y.ctypes.data, x.shape[1], x.shape[0], 0.088388)
    # reference uses a vectorized expf (ggml_v_expf, ~2 ulp) — tolerance 1e-6 absolute on probabilities
    assert np.abs(y - ref).max() <= 1e-6


@pytest.mark.parametrize("mode,nm", [(0, "norm"), (2, "neox")])
@pytest.mark.parametrize("use_ff", [0, 1])
def test_rope_vs_golden(mode, nm, use_ff):
    x = np.ascontiguousarray(G[f"rope_{nm}_x"]); pos = np.ascontiguousarray(G[f"rope_{nm}_pos"])
    ff = np.ascontiguousarray(G[f"rope_{nm}_ff"]); ref = G[f"rope_{nm}_ff{use_ff}_y"]
    y = np.zeros_like(ref)
    qf.port().oq_rope(x.ctypes.data, y.ctypes.data, pos.ctypes.data, ff.ctypes.data if use_ff else None,
                      128, x.shape[1], x.shape[0], 128, mode, 0, 500000.0, 1.0, 0.0, 1.0, 32.0, 1.0)
    # same fp32 recurrence for theta (ops.cpp:5613-5628) and same libm cosf/sinf -> agreement to ~1 ulp
    assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()


def test_silu_mul_vs_golden():
    g, u, ref = (np.ascontiguousarray(G[k]) for k in ("silu_g", "silu_u", "silu_y"))
    y = np.zeros_like(ref)
    qf.port().oq_silu_mul(g.ctypes.data, u.ctypes.data, y.ctypes.data, g.size)
    assert np.abs(y - ref).max() <= 2e-6 * np.abs(ref).max()


def test_fp16_roundtrip_matches_numpy():
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.standard_normal(2000) * 10.0 ** rng.integers(-9, 5, 2000), [0.0, -0.0, 65504.0, 65520.0, 1e-8, 6e-8]]).astype(np.float32)
    for f in v:
        h = qf.port().oq_fp32_to_fp16(float(f))
        assert h == int(np.float32(f).astype(np.float16).view(np.uint16)), f
        assert qf.port().oq_fp16_to_fp32(h) == float(np.uint16(h).view(np.float16))


# ---- live checks against the compiled reference (only where oracle/_ref/lib travelled) ---------------------
needs_ref = pytest.mark.skipif(not qf.have_ref(), reason="oracle/_ref/lib not built")


@needs_ref
@pytest.mark.parametrize("t,wfn,cpu_fn,port_fn,qfn,qbs,qblk", [
    (qf.Q4_0, None, "ggml_vec_dot_q4_0_q8_0", "oq_vec_dot_q4_0_q8_0", "quantize_row_q8_0", 34, 32),
    (qf.Q8_0, None, "ggml_vec_dot_q8_0_q8_0", "oq_vec_dot_q8_0_q8_0", "quantize_row_q8_0", 34, 32),
    (qf.Q4_K, None, "ggml_vec_dot_q4_K_q8_K", "oq_vec_dot_q4_K_q8_K", "quantize_row_q8_K", 292, 256),
])
def test_vec_dot_live_vs_reference_simd_and_generic(t, wfn, cpu_fn, port_fn, qfn, qbs, qblk):
    rng = np.random.default_rng(11)
    cpu = qf.ref_lib("cpu")
    k = 4096
    w = qf.random_blocks(t, 8, k, rng=rng)
    x = rng.standard_normal(k).astype(np.float32); x[7] = 25.0
    q = np.zeros(k // qblk * qbs, dtype=np.uint8)
    f = getattr(cpu, qfn); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]; f(x.ctypes.data, q.ctypes.data, k)
    # our port's quantizer must produce the same bytes as the one the CPU backend runs
    q2 = np.zeros_like(q)
    pq = qf.port().oq_quantize_row_q8_K_ref if qblk == 256 else qf.port().oq_quantize_row_q8_0_x86
    pq(x.ctypes.data, q2.ctypes.data, k)
    assert np.array_equal(q, q2)
    for variant in (cpu_fn, cpu_fn + "_generic"):
        vd = getattr(cpu, variant)
        vd.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        ref, mine = [], []
        for r in range(w.shape[0]):
            s = C.c_float(0)
            vd(k, C.byref(s), 0, w[r].ctypes.data, 0, q.ctypes.data, 0, 1)
            ref.append(s.value)
            mine.append(getattr(qf.port(), port_fn)(k, w[r].ctypes.data, q.ctypes.data))
        ref, mine = np.array(ref), np.array(mine)
        # integer partial sums are exact on both sides; fp32 association (SIMD lanes; gcc FMA contraction in the
        # -mfma reference objects vs -ffp-contract=off here) differs -> a few ulp of the largest partial sum
        tol = (2e-6 if variant.endswith("_generic") else 2e-5) * np.abs(ref).max()
        assert np.abs(ref - mine).max() <= tol


import refshim as rs  # noqa: E402

needs_shim = pytest.mark.skipif(not rs.available(), reason="oracle/_ref/lib/libref_shim.so not built")


@needs_shim
@pytest.mark.parametrize("t", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("n_tokens,broadcast", [(1, True), (1, False), (5, True), (5, False)])
def test_mul_mat_id_live_vs_reference_cpu_backend(t, n_tokens, broadcast):
    """oq_mul_mat_id pinned against the reference executing ggml_mul_mat_id on its own CPU backend (ggml-cpu.c:1503-1700)
    through the public ggml API (oracle/ref_shim.cpp): expert routing, the src1 broadcast rule and the per-row arithmetic."""
    rng = np.random.default_rng(100 * t + n_tokens + broadcast)
    k, m, n_expert, n_used = 512, 48, 8, 2
    as_ = qf.random_blocks(t, n_expert * m, k, rng=rng)
    nb1 = 1 if broadcast else n_used
    b = rng.standard_normal((n_tokens, nb1, k)).astype(np.float32)
    ids = np.stack([rng.choice(n_expert, size=n_used, replace=False) for _ in range(n_tokens)]).astype(np.int32)
    g = rs.Graph()
    y = g.mul_mat_id(g.input(as_.reshape(-1), t, (k, m, n_expert)), g.input(b), g.input(ids))
    ref = g.run("CPU", [(y, np.float32, (n_tokens, n_used, m))], n_threads=4, strict=True)[0]
    mine = qf.port_mul_mat_id(t, as_, k, m, n_expert, b, ids)
    assert np.abs(ref - mine).max() <= 2e-5 * np.abs(ref).max()
    # routing sanity: swapping the two selected experts swaps the two output slots (only when src1 is shared)
    if broadcast:
        mine_sw = qf.port_mul_mat_id(t, as_, k, m, n_expert, b, ids[:, ::-1])
        assert np.array_equal(mine_sw, mine[:, ::-1])


GM = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_moe_vectors.npz"))


@pytest.mark.parametrize("t", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("tag", ["bcast", "slot"])
def test_mul_mat_id_vs_golden(t, tag):
    """oq_mul_mat_id against stored outputs of the reference CPU backend (tests/golden/make_golden.py main_moe): works where
    oracle/_ref is absent."""
    nm = qf.NAMES[t]
    as_, b, ids, ref = GM[f"mmid_{nm}_as"], GM[f"mmid_{nm}_{tag}_b"], GM["mmid_ids"], GM[f"mmid_{nm}_{tag}_y"]
    y = qf.port_mul_mat_id(t, as_, b.shape[2], ref.shape[2], 8, b, ids)
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()


def moe_block_port(wr, wg, wu, wd, x, n_expert=8, n_used=2, t=qf.Q4_K):
    """GenericSparseMLP::forward + MultiMLP::forward (src/layers.cpp:3755-3880, :3674-3688) for one token, composed from the
    port's primitives: router matmul -> softmax(40 * logits) -> top-k -> normalised weights -> SwiGLU experts -> down -> sum."""
    hidden, ffn = x.shape[1], wg.shape[0] // n_expert
    logits = qf.port_mul_mat(t, wr, hidden, n_expert, x) * np.float32(40.0)
    probs = np.zeros_like(logits)
    qf.port().oq_soft_max(logits.ctypes.data, None, probs.ctypes.data, n_expert, 1, 1.0)
    sel = np.argsort(-probs[0], kind="stable")[:n_used].astype(np.int32)
    w = probs[0, sel]; w = (w / np.float32(w.sum(dtype=np.float32))).astype(np.float32)
    g = qf.port_mul_mat_id(t, wg, hidden, ffn, n_expert, x[None], sel[None])[0]
    u = qf.port_mul_mat_id(t, wu, hidden, ffn, n_expert, x[None], sel[None])[0]
    par = np.zeros_like(g); qf.port().oq_silu_mul(g.ctypes.data, u.ctypes.data, par.ctypes.data, g.size)
    e = qf.port_mul_mat_id(t, wd, ffn, hidden, n_expert, par[None], sel[None])[0]        # [n_used, hidden], one column per slot
    out = (e[0] * w[0]).astype(np.float32)
    for i in range(1, n_used):
        out = (out + e[i] * w[i]).astype(np.float32)
    return sel, w, par, out


def test_sparse_moe_block_vs_golden():
    sel, w, par, out = moe_block_port(GM["moe_wr"], GM["moe_wg"], GM["moe_wu"], GM["moe_wd"], GM["moe_x"])
    assert set(sel.tolist()) == set(GM["moe_sel"][0].tolist())
    order = [sel.tolist().index(e) for e in GM["moe_sel"][0].tolist()]   # the reference's top_k order of the two slots
    assert np.abs(w[order] - GM["moe_w"].reshape(-1)).max() <= 1e-6
    assert np.abs(par[order] - GM["moe_par"][0]).max() <= 3e-5 * np.abs(GM["moe_par"]).max()
    assert np.abs(out - GM["moe_out"][0]).max() <= 5e-5 * np.abs(GM["moe_out"]).max()


def test_q4k_scale_split_planes_are_exact():
    """The algebra behind csrc/prefill_tc.cu: with sc = 8*hi + lo (hi, lo < 8) the three int8 planes hi*q, lo*q (<= 105) and m give the
    reference's per-super-block integers exactly — 8*(A_hi.X) + (A_lo.X) == sum_j sc_j sum q x and (A_m.X) == sum_j m_j bsum_j — so the
    tensor-core result, rescaled once per super-block, equals the oracle's mul_mat up to fp32 association."""
    rng = np.random.default_rng(99)
    k, m, n = 1024, 24, 5
    w = qf.random_blocks(qf.Q4_K, m, k, rng=rng)
    x = rng.standard_normal((n, k)).astype(np.float32); x[0, 5] = 30.0
    # activation codes exactly as the reference quantizes them (Q8_K)
    q8 = np.zeros((n, k // 256 * 292), dtype=np.uint8)
    for r in range(n):
        qf.port().oq_quantize_row_q8_K_ref(x[r].ctypes.data, q8[r].ctypes.data, k)
    qb = q8.reshape(n, k // 256, 292)
    dx = qb[:, :, 0:4].copy().view(np.float32).reshape(n, -1)
    xc = qb[:, :, 4:260].view(np.int8).astype(np.int32)                                   # [n, nb, 256]
    wb = w.reshape(m, k // 256, 144)
    d = wb[:, :, 0:2].copy().view(np.float16).astype(np.float32).reshape(m, -1)
    dmin = wb[:, :, 2:4].copy().view(np.float16).astype(np.float32).reshape(m, -1)
    sc8 = wb[:, :, 4:16].astype(np.int32)                                                  # 12 scale bytes (get_scale_min_k4)
    sc = np.zeros((m, k // 256, 8), dtype=np.int32); mn = np.zeros_like(sc)
    for j in range(8):
        if j < 4:
            sc[:, :, j] = sc8[:, :, j] & 63; mn[:, :, j] = sc8[:, :, j + 4] & 63
        else:
            sc[:, :, j] = (sc8[:, :, j + 4] & 0xF) | ((sc8[:, :, j - 4] >> 6) << 4)
            mn[:, :, j] = (sc8[:, :, j + 4] >> 4) | ((sc8[:, :, j] >> 6) << 4)
    qs = wb[:, :, 16:].astype(np.int32).reshape(m, k // 256, 4, 32)
    q = np.stack([qs & 0xF, qs >> 4], axis=3).reshape(m, k // 256, 8, 32)                 # sub-block j = (group j//2, nibble j%2)
    hi, lo = sc >> 3, sc & 7
    A_hi, A_lo = (hi[..., None] * q), (lo[..., None] * q)
    assert A_hi.max() <= 105 and A_lo.max() <= 105 and mn.max() <= 63                      # all three planes fit int8
    A_m = np.broadcast_to(mn[..., None], q.shape)
    X = xc.reshape(n, k // 256, 8, 32)
    g_hi = np.einsum("mbjk,nbjk->nmb", A_hi, X); g_lo = np.einsum("mbjk,nbjk->nmb", A_lo, X); g_m = np.einsum("mbjk,nbjk->nmb", A_m, X)
    isum = np.einsum("mbj,nmbj->nmb", sc, np.einsum("mbjk,nbjk->nmbj", q, X))             # the reference's integers
    msum = np.einsum("mbj,nbj->nmb", mn, X.sum(-1))
    assert np.array_equal(8 * g_hi + g_lo, isum) and np.array_equal(g_m, msum)
    y = np.zeros((n, m), dtype=np.float32)
    for b in range(k // 256):
        y += (dx[:, None, b] * d[None, :, b]) * (8 * g_hi + g_lo)[:, :, b].astype(np.float32) - (dx[:, None, b] * dmin[None, :, b]) * g_m[:, :, b].astype(np.float32)
    ref = qf.port_mul_mat(qf.Q4_K, w, k, m, x)
    assert np.abs(y - ref).max() <= 2e-5 * np.abs(ref).max()
