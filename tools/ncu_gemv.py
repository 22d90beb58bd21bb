"""Tiny driver for ncu captures: runs a few launches of one GEMV shape. Usage: ncu ... python tools/ncu_gemv.py q4_K 4096 14336 [tune]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge, qformats as qf
pkg = ge.load_package()
from chatllm_cpp_b200 import kernels as K
t = {"q4_K": qf.Q4_K, "q4_0": qf.Q4_0, "q8_0": qf.Q8_0}[sys.argv[1]]
k, m = int(sys.argv[2]), int(sys.argv[3])
if len(sys.argv) > 4:
    pkg.lib().b200_gemv_set_tuning(*[int(v) for v in sys.argv[4].split(",")])
mm = min(m, 1024)
w0 = K.upload_weights(t, qf.random_blocks(t, mm, k, seed=1).repeat((m + mm - 1) // mm, axis=0)[:m], k, m)
ws = [w0] + [w0.clone() for _ in range(5)]
x = torch.randn((1, k), device="cuda"); q = K.quantize_act(t, x); y = torch.empty((1, m), device="cuda")
for i in range(6):
    K.mul_mat_q(t, ws[i], k, m, q, 1, out=y)
torch.cuda.synchronize()
