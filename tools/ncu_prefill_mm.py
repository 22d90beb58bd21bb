"""Tiny driver for ncu captures of the prompt matmul: python tools/ncu_prefill_mm.py q4_K 4096 14336 2048"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge, qformats as qf
pkg = ge.load_package(); L = pkg.lib()
from chatllm_cpp_b200 import kernels as K
t = {"q4_K": qf.Q4_K, "q4_0": qf.Q4_0, "q8_0": qf.Q8_0}[sys.argv[1]]
k, m, n = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mm = min(m, 1024)
w = K.upload_weights(t, qf.random_blocks(t, mm, k, seed=1).repeat((m + mm - 1) // mm, axis=0)[:m], k, m)
x = torch.randn((n, k), device="cuda")
pq = torch.empty(L.b200_pact_col_bytes(t, k) * n, dtype=torch.uint8, device="cuda")
y = torch.empty((n, m), device="cuda")
st = torch.cuda.current_stream().cuda_stream
L.b200_quantize_plain(t, x.data_ptr(), k, k, n, pq.data_ptr(), st)
for _ in range(3):
    L.b200_mul_mat_q_batched_tc(t, w.data_ptr(), k, m, pq.data_ptr(), n, y.data_ptr(), m, 0, st)
torch.cuda.synchronize()
