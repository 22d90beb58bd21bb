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


if __name__ == "__main__":
    main()
