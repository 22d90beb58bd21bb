"""Find the first graph node whose output differs between the CPU backend and the plugin: runs the host harness with
--trace on both and compares per-node checksums.  usage: trace_diff.py ARCH QUANT [prefill] [decode]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
arch, quant = sys.argv[1], sys.argv[2]
prefill = sys.argv[3] if len(sys.argv) > 3 else "5"
decode = sys.argv[4] if len(sys.argv) > 4 else "1"
model = f"/tmp/td-{arch}-{quant}.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", arch, "--quant", quant, "--out", model, "--max_length", "512"], check=True, capture_output=True)
import numpy as np
def run(ngl, out):
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "16", "--prefill", prefill, "--decode", decode, "--max_length", "512", "--trace", out, "--trace_full", out + ".bin"], capture_output=True, text=True)
    if p.returncode: print("harness failed", p.stderr[-500:])
    return [l.split() for l in open(out) if not l.startswith("#")]
a, b = run("0", "/tmp/td_cpu.txt"), run("all", "/tmp/td_gpu.txt")
print(len(a), len(b), "nodes traced")
# element-wise comparison of every contiguous node output
fa, fb = np.fromfile("/tmp/td_cpu.txt.bin", dtype=np.uint8), np.fromfile("/tmp/td_gpu.txt.bin", dtype=np.uint8)
off = 0; shown = 0
SZ = {"f32": 4, "f16": 2, "i32": 4}
for la in a:
    if la[4] != "c" or la[2] not in SZ: continue
    ne = [int(v) for v in la[3].strip("[]").split(",")]
    nb = ne[0] * ne[1] * ne[2] * ne[3] * SZ[la[2]]
    dt = {"f32": np.float32, "f16": np.float16, "i32": np.int32}[la[2]]
    xa = fa[off:off + nb].view(dt).astype(np.float64); xb = fb[off:off + nb].view(dt).astype(np.float64)
    off += nb
    fin = np.isfinite(xa) & np.isfinite(xb)
    if not fin.any(): continue
    err = np.abs(xa[fin] - xb[fin]).max() / (np.abs(xa[fin]).max() + 1e-30)
    if err > 1e-4 and la[1] not in ("RESHAPE", "VIEW", "PERMUTE", "TRANSPOSE"):
        print("ELEMENTWISE", la[0], la[1], la[2], la[3], f"max rel err {err:.3e}", "at", int(np.argmax(np.abs(xa - xb)))); shown += 1
        if shown >= 10: break
shown = 99
for la, lb in zip(a, b):
    fa = float(la[-1].split("=")[1]); fb = float(lb[-1].split("=")[1])
    sa = float(la[-2].split("=")[1]); sb = float(lb[-2].split("=")[1])
    bad = abs(fa - fb) > 1e-3 * max(abs(fa), 1e-6) or abs(sa - sb) > 1e-3 * max(abs(fa), 1e-6)
    if la[1] != lb[1] or la[3] != lb[3]:
        print("STRUCTURE DIFF", la, lb); break
    if bad and la[4] == "c":
        print("DIFF", " ".join(la), "| gpu", lb[-2], lb[-1]); shown += 1
        if shown > 12: break
print("done")
