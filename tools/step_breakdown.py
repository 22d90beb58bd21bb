"""Warm per-launch cost of every kernel type of the fused decode step (Llama-3-8B Q4_K shapes, n_kv = 4097): each kernel
is captured 64x back-to-back in a CUDA graph (same stream => dependent launches, like in the real step) and replayed."""
import ctypes as C, math, os, sys, json
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge
pkg = ge.load_package(); L = pkg.lib()
from chatllm_cpp_b200 import session as S
if len(sys.argv) > 1: L.b200_gemv_set_tuning(*[int(v) for v in sys.argv[1].split(",")])
cfg = S.make_config("llama3-8b", pkg.Q4_K, layers=16, max_len=4352)   # 16 layers: q/k/v 400 MB, o 150 MB, gate+up 1.06 GB, down 528 MB rotate — every class exceeds the 126 MB L2 (r01 used 4 layers: o and q/k/v sat in L2)
s = S.DecodeSession(cfg, seed=0); s.fill_kv_random(4096); s._ptr_arrays(); s.pos.fill_(4096)
c = cfg; W = s.layers; q = s.qact.data_ptr(); hd = c.head_dim
def st(): return torch.cuda.current_stream().cuda_stream
NL = len(W)
kern = {
 "add_rmsnorm_quant": (2, lambda i: L.b200_add_rmsnorm_quant(c.wtype, s.x.data_ptr(), s.o.data_ptr(), W[i % NL].attn_norm.data_ptr(), s.x.data_ptr(), 0, q, c.hidden, 1, c.eps, st())),
 "gemv_qkv(multi)": (1, lambda i: L.b200_mul_mat_q_multi(c.wtype, 0, 3, W[i % NL].qkv["W"], W[i % NL].qkv["m"], W[i % NL].qkv["y"], W[i % NL].qkv["ld"], W[i % NL].qkv["b"], c.hidden, q, 1, st())),
 "rope_kv_store": (1, lambda i: L.b200_rope_kv_store(s.q.data_ptr(), s.k.data_ptr(), s.v.data_ptr(), s.pos.data_ptr(), 0, W[i % NL].kc.data_ptr(), W[i % NL].vc.data_ptr(), c.heads, c.kv_heads, hd, c.rope_mode, c.rope_theta, c.kv_hidden, c.max_len, st())),
 "attn_decode_quant(3 launches)": (1, lambda i: L.b200_attn_decode_quant(s.q.data_ptr(), W[i % NL].kc.data_ptr(), W[i % NL].vc.data_ptr(), s.att.data_ptr(), s.scratch.data_ptr(), c.heads, c.kv_heads, hd, 4097, c.kv_hidden, c.max_len, 1.0 / math.sqrt(hd), c.wtype, q, st())),
 "gemv_o": (1, lambda i: L.b200_mul_mat_q(c.wtype, W[i % NL].wo.data_ptr(), c.hidden, c.hidden, q, 1, s.o.data_ptr(), c.hidden, 0, st())),
 "gemv_gate_up(paired)": (1, lambda i: L.b200_mul_mat_q_multi(c.wtype, 1, 2, W[i % NL].gu["W"], W[i % NL].gu["m"], W[i % NL].gu["y"], W[i % NL].gu["ld"], W[i % NL].gu["b"], c.hidden, q, 1, st())),
 "quantize(ffn)": (1, lambda i: L.b200_quantize_act(c.wtype, s.gate.data_ptr(), c.ffn, c.ffn, 1, q, st())),
 "gemv_down": (1, lambda i: L.b200_mul_mat_q(c.wtype, W[i % NL].wdown.data_ptr(), c.ffn, c.hidden, q, 1, s.o.data_ptr(), c.hidden, 0, st())),
}
tot = 0
for name, (per_layer, fn) in kern.items():
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        with torch.cuda.graph(g, stream=stream):
            for i in range(64): fn(i)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [g.replay() for _ in range(5)]; e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 / 64 * 1000
    tot += us * per_layer
    print(f"{name:28s} {us:7.2f} us/call  x{per_layer}/layer")
print(f"sum per layer {tot:.1f} us -> x32 = {tot*32/1000:.2f} ms (+ lm_head ~48 us)")
