"""GEMV bandwidth sweep on one B200: CUDA-event timing of b200_mul_mat_q over Llama-3-8B shapes, rotating over
enough distinct weight copies that the working set exceeds L2 (126 MB).  Usage:
    python tools/gemv_sweep.py [--type q4_K] [--tune ks,stages,warps,rg,grid ...]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import qformats as qf  # noqa: E402

pkg = ge.load_package()
from chatllm_cpp_b200 import kernels as K  # noqa: E402

TYPES = {"q4_K": qf.Q4_K, "q4_0": qf.Q4_0, "q8_0": qf.Q8_0}


def gemv_bytes(t, k, m):
    return m * qf.row_size(t, k) + 4 * k + 4 * m


def bench(t, k, m, n=1, iters=50, min_bytes=600e6):
    wb = m * qf.row_size(t, k)
    copies = max(2, int(np.ceil(min_bytes / wb)))
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    # random bytes are fine for timing, but fp16 scales must be finite: build one valid matrix and replicate
    w0 = K.upload_weights(t, qf.random_blocks(t, min(m, 1024), k, seed=1).repeat((m + min(m, 1024) - 1) // min(m, 1024), axis=0)[:m], k, m)
    ws = [w0] + [w0.clone() for _ in range(copies - 1)]
    x = torch.randn((n, k), device="cuda")
    q = K.quantize_act(t, x)
    y = torch.empty((n, m), device="cuda")
    for i in range(3):
        K.mul_mat_q(t, ws[i % copies], k, m, q, n, out=y)
    torch.cuda.synchronize()
    # capture the launches in a CUDA graph so Python/ctypes launch overhead is not what we time
    stream = torch.cuda.Stream()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(graph, stream=stream):
            for i in range(iters):
                K.mul_mat_q(t, ws[i % copies], k, m, q, n, out=y)
    graph.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, gemv_bytes(t, k, m) / ms / 1e6


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--type", default="q4_K")
    ap.add_argument("--tune", nargs="*", default=["0,0,0,0,0"])
    ap.add_argument("--shapes", default="4096x4096,4096x1024,4096x14336,14336x4096,4096x128256")
    ap.add_argument("--n", type=int, default=1)
    a = ap.parse_args()
    t = TYPES[a.type]
    shapes = [tuple(int(v) for v in s.split("x")) for s in a.shapes.split(",")]
    for tune in a.tune:
        ks, st, wp, rg, grid = (int(v) for v in tune.split(","))
        pkg.lib().b200_gemv_set_tuning(ks, st, wp, rg, grid)
        for (k, m) in shapes:
            ms, gbs = bench(t, k, m, n=a.n)
            print(json.dumps({"type": a.type, "k": k, "m": m, "n": a.n, "tune": tune, "us": round(ms * 1e3, 2), "GBps": round(gbs, 1)}), flush=True)


if __name__ == "__main__":
    main()
