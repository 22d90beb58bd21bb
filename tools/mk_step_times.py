"""Per-step time breakdown of the persistent decode kernel (B200_MK_TIMES=1): SM-clock stamps of CTA 0 after every step's grid barrier.
    B200_MK_TIMES=1 python tools/mk_step_times.py [layers]"""
import ctypes as C
import os
import sys

os.environ["B200_MK_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import __graft_entry__ as ge
pkg = ge.load_package()
from chatllm_cpp_b200 import session as S

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = S.make_config("llama3-8b", pkg.Q4_K, layers=layers, max_len=4096 + 256)
sess = S.DecodeSession(cfg, seed=0, fused=3)
sess.fill_kv_random(4096, seed=1)
sess.mk_advance = True
sess.tok.fill_(5); sess.pos.fill_(4096)
for _ in range(5):
    sess.enqueue(0)
torch.cuda.synchronize()
n = 1 + 7 * layers + 3
NX = 80
acc = np.zeros(n - 1)
prof = np.zeros(NX)
reps = 10
mhz = 1.0
for _ in range(reps):
    sess.enqueue(0)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (n + NX))()
    got = pkg.lib().b200_decode_plan_times(sess._plan, buf, n + NX, torch.cuda.current_stream().cuda_stream)
    t = np.array(buf[:n], dtype=np.float64)
    acc[:n - 1] += np.diff(t)
    prof = np.array(buf[n:n + NX], dtype=np.float64)   # accumulated since plan creation
acc /= reps
clk_ghz = float(os.environ.get("SM_GHZ", "1.965"))
us = acc / (clk_ghz * 1e3)
names = ["qkv", "scores", "pv", "tail", "o", "gate_up", "down"]
print(f"embed {us[0]:.2f} us")
per = us[1:1 + 7 * layers].reshape(layers, 7)
for j, nme in enumerate(names):
    print(f"{nme:8s} mean {per[:, j].mean():7.2f} us   min {per[:, j].min():7.2f}   max {per[:, j].max():7.2f}")
print(f"layer    mean {per.sum(axis=1).mean():7.2f} us")
print(f"head {us[1 + 7 * layers]:.2f} us  finalize {us[2 + 7 * layers]:.2f} us   total {us.sum():.1f} us (assuming {clk_ghz} GHz SM clock)")

# per-phase-kind split of one warp's time (warp 3 of the first and of the last CTA), accumulated over every launch since plan creation
kn = ["qkv", "o", "gate_up", "down", "head"]
for c, cta in enumerate(("CTA 0", "last CTA")):
    for k in range(5):
        a = prof[(c * 5 + k) * 8:(c * 5 + k) * 8 + 8]
        if a[4] > 0:
            f = 1.0 / a[4] / (clk_ghz * 1e3)
            print(f"{cta:9s} {kn[k]:8s} per phase: prologue {a[5] * f:6.2f} us  wait-for-data {a[0] * f:6.2f}  compute {a[1] * f:6.2f}  issue+cursor {a[2] * f:6.2f}   items {a[3] / a[4]:.2f}")
