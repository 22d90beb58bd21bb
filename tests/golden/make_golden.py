"""Generates tests/golden/*.npz by EXECUTING THE REFERENCE (oracle/_ref libs compiled from /root/reference by
oracle/Makefile).  Run in the build container:  python tests/golden/make_golden.py
The reference ships no golden vectors; these fixtures are what pins the oracle port and the CUDA
kernels on machines where /root/reference is absent.  Inputs are seeded; outputs are the reference's.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import qformats as qf  # noqa: E402
import refshim as rs  # noqa: E402


def acts(rng, n, k, heavy=False):
    x = rng.standard_normal((n, k)).astype(np.float32)
    if heavy:  # 1-in-1000 outliers x20 to exercise amax-driven scales
        m = rng.random((n, k)) < 1e-3
        x[m] *= 20.0
    return x


def main():
    rng = np.random.default_rng(1234)
    base, cpu = qf.ref_lib("base"), qf.ref_lib("cpu")
    out = {}

    # --- quantizers (ggml-quants.c *_ref; ggml-cpu/arch/x86/quants.c quantize_row_q8_0) -------------
    x = acts(rng, 4, 512, heavy=True)
    x[1, :32] = 0.0          # all-zero block
    # rounding-mode probe: amax = 127 -> id = 1, the other values are exact .5 ties (roundf vs RNE disagree)
    x[2, 0:32] = (np.arange(32, dtype=np.float32) + 0.5) * np.where(np.arange(32) % 2 == 0, 1.0, -1.0)
    x[2, 0] = 127.0
    out["quant_x"] = x
    for name, libh, fn, t, bs in (("q4_0_ref", base, "quantize_row_q4_0_ref", qf.Q4_0, 18),
                                  ("q8_0_ref", base, "quantize_row_q8_0_ref", qf.Q8_0, 34),
                                  ("q8_0_x86", cpu, "quantize_row_q8_0", qf.Q8_0, 34),
                                  ("q8_K_ref", base, "quantize_row_q8_K_ref", None, 292),
                                  ("q4_K_ref", base, "quantize_row_q4_K_ref", qf.Q4_K, 144)):
        blk = 256 if "K" in name else 32
        y = np.zeros((x.shape[0], x.shape[1] // blk * bs), dtype=np.uint8)
        f = getattr(libh, fn); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        for r in range(x.shape[0]):
            f(x[r].ctypes.data, y[r].ctypes.data, x.shape[1])
        out["quant_" + name] = y

    # --- dequantizers ---------------------------------------------------------------------------------
    for t, fn in ((qf.Q4_0, "dequantize_row_q4_0"), (qf.Q8_0, "dequantize_row_q8_0"), (qf.Q4_K, "dequantize_row_q4_K")):
        w = qf.random_blocks(t, 3, 512, rng=rng)
        y = np.zeros((3, 512), dtype=np.float32)
        f = getattr(base, fn); f.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        for r in range(3):
            f(w[r].ctypes.data, y[r].ctypes.data, 512)
        out[f"deq_{qf.NAMES[t]}_w"] = w
        out[f"deq_{qf.NAMES[t]}_y"] = y

    # --- mul_mat through the reference CPU backend (ggml-cpu.c:1229-1421), n = 1, 3, 8 ---------------
    for t in (qf.Q4_0, qf.Q8_0, qf.Q4_K):
        k, m = 1024, 48
        w = qf.random_blocks(t, m, k, rng=rng)
        for n in (1, 3, 8):
            xx = acts(rng, n, k, heavy=(n == 3))
            g = rs.Graph()
            r = g.mul_mat(g.input(w.reshape(-1), t, (k, m)), g.input(xx))
            (y,) = g.run("CPU", [(r, np.float32, (n, m))])
            out[f"mm_{qf.NAMES[t]}_n{n}_x"] = xx
            out[f"mm_{qf.NAMES[t]}_n{n}_y"] = y
        out[f"mm_{qf.NAMES[t]}_w"] = w

    # --- rms_norm * weight, soft_max, rope (normal + neox + freq_factors), silu*mul --------------------
    xx = acts(rng, 3, 2048)
    wn = (1.0 + 0.1 * rng.standard_normal(2048)).astype(np.float32)
    g = rs.Graph(); r = g.mul(g.rms_norm(g.input(xx), 1e-5), g.input(wn))
    (y,) = g.run("CPU", [(r, np.float32, xx.shape)])
    out["rms_x"], out["rms_w"], out["rms_y"] = xx, wn, y

    sx = (acts(rng, 8, 777) * 3).astype(np.float32)
    g = rs.Graph(); r = g.soft_max(g.input(sx), None, scale=0.088388)
    (y,) = g.run("CPU", [(r, np.float32, sx.shape)])
    out["sm_x"], out["sm_y"] = sx, y

    for mode, nm in ((0, "norm"), (2, "neox")):
        q = acts(rng, 3 * 4, 128).reshape(3, 4, 128)          # [tokens, heads, head_dim]
        pos = np.array([0, 17, 4095], dtype=np.int32)
        ff = (1.0 + rng.random(64)).astype(np.float32)
        for use_ff in (False, True):
            g = rs.Graph()
            r = g.rope(g.input(q), g.input(pos), 128, mode, 500000.0, ff=g.input(ff) if use_ff else None)
            (y,) = g.run("CPU", [(r, np.float32, q.shape)])
            out[f"rope_{nm}_ff{int(use_ff)}_y"] = y
        out[f"rope_{nm}_x"], out[f"rope_{nm}_pos"], out[f"rope_{nm}_ff"] = q, pos, ff

    ga, up = acts(rng, 2, 512) * 2, acts(rng, 2, 512)
    g = rs.Graph(); r = g.mul(g.silu(g.input(ga)), g.input(up))
    (y,) = g.run("CPU", [(r, np.float32, ga.shape)])
    out["silu_g"], out["silu_u"], out["silu_y"] = ga, up, y

    # --- attention matmuls with F16 operands (ggml-cpu.c:213-219; src/layers.cpp:2541-2561) ------------
    kc = (rng.standard_normal((37, 2, 64))).astype(np.float16)   # [n_kv, kv_heads, head_dim]
    qv = acts(rng, 4, 64).reshape(1, 4, 64)                      # [qlen, heads, head_dim]
    g = rs.Graph()
    K = g.permute(g.input(kc), (0, 2, 1, 3))     # -> [head_dim, n_kv, kv_heads]
    Q = g.permute(g.input(qv), (0, 2, 1, 3))     # -> [head_dim, qlen, heads]
    r = g.mul_mat(K, Q, prec_f32=True)           # [n_kv, qlen, heads]
    (y,) = g.run("CPU", [(r, np.float32, (4, 1, 37))])
    out["att_k"], out["att_q"], out["att_s"] = kc, qv, y

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_vectors.npz"), {k: v.shape for k, v in out.items()})


def main_moe():
    """ref_moe_vectors.npz: ggml_mul_mat_id (ggml/src/ggml-cpu/ggml-cpu.c:1503-1700) and the whole sparse-MoE block of
    GenericSparseMLP::forward / MultiMLP::forward (src/layers.cpp:3755-3880, :3674-3688), executed by the reference CPU backend."""
    rng = np.random.default_rng(4321)
    out = {}
    k, m, n_expert, n_used, n_tok = 512, 32, 8, 2, 3
    ids = np.stack([rng.choice(n_expert, size=n_used, replace=False) for _ in range(n_tok)]).astype(np.int32)
    out["mmid_ids"] = ids
    for t in (qf.Q4_K, qf.Q4_0, qf.Q8_0):
        nm = qf.NAMES[t]
        as_ = qf.random_blocks(t, n_expert * m, k, rng=rng)
        out[f"mmid_{nm}_as"] = as_
        for tag, nb1 in (("bcast", 1), ("slot", n_used)):
            b = acts(rng, n_tok * nb1, k, heavy=True).reshape(n_tok, nb1, k)
            g = rs.Graph()
            y = g.mul_mat_id(g.input(as_.reshape(-1), t, (k, m, n_expert)), g.input(b), g.input(ids))
            (yv,) = g.run("CPU", [(y, np.float32, (n_tok, n_used, m))], n_threads=4)
            out[f"mmid_{nm}_{tag}_b"], out[f"mmid_{nm}_{tag}_y"] = b, yv

    # the block: router -> softmax -> top-2 -> normalised weights -> experts' SwiGLU -> down -> weighted sum over the slots
    hidden, ffn = 256, 512
    t = qf.Q4_K
    wr = qf.random_blocks(t, n_expert, hidden, rng=rng)
    wg = qf.random_blocks(t, n_expert * ffn, hidden, rng=rng); wu = qf.random_blocks(t, n_expert * ffn, hidden, rng=rng)
    wd = qf.random_blocks(t, n_expert * hidden, ffn, rng=rng)
    x = acts(rng, 1, hidden)
    g = rs.Graph()
    h = g.input(x)
    probs = g.soft_max(g.scale(g.mul_mat(g.input(wr.reshape(-1), t, (hidden, n_expert)), h), 40.0))
    sel = g.top_k(probs, n_used)
    w = g.reshape(g.get_rows(g.reshape(probs, (1, n_expert, 1)), sel), (n_used, 1))
    w = g.reshape(g.div(w, g.sum_rows(w)), (1, n_used, 1))
    h3 = g.reshape(h, (hidden, 1, 1))
    act = g.silu(g.mul_mat_id(g.input(wg.reshape(-1), t, (hidden, ffn, n_expert)), h3, sel))
    par = g.mul(g.mul_mat_id(g.input(wu.reshape(-1), t, (hidden, ffn, n_expert)), h3, sel), act, inplace=True)
    experts = g.mul(g.mul_mat_id(g.input(wd.reshape(-1), t, (ffn, hidden, n_expert)), par, sel), w)
    o = g.add(g.view(experts, (hidden, 1), nb=(hidden * n_used * 4,), offset=0), g.view(experts, (hidden, 1), nb=(hidden * n_used * 4,), offset=hidden * 4))
    # intermediates are views / in-place results whose storage the graph allocator recycles: read them through contiguous copies
    w_out, par_out = g.cont(w), g.cont(par)
    selv, wv, parv, ov = g.run("CPU", [(sel, np.int32, (1, n_used)), (w_out, np.float32, (1, n_used, 1)), (par_out, np.float32, (1, n_used, ffn)),
                                       (o, np.float32, (1, hidden))], n_threads=4)
    out.update({"moe_wr": wr, "moe_wg": wg, "moe_wu": wu, "moe_wd": wd, "moe_x": x, "moe_sel": selv, "moe_w": wv, "moe_par": parv, "moe_out": ov})
    np.savez_compressed(os.path.join(HERE, "ref_moe_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_moe_vectors.npz"), {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    if "--moe-only" not in sys.argv:
        main()
    main_moe()
