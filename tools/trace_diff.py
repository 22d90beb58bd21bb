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
def run(ngl, out):
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "16", "--prefill", prefill, "--decode", decode, "--max_length", "512", "--trace", out], capture_output=True, text=True)
    if p.returncode: print("harness failed", p.stderr[-500:])
    return [l.split() for l in open(out)]
a, b = run("0", "/tmp/td_cpu.txt"), run("all", "/tmp/td_gpu.txt")
print(len(a), len(b), "nodes traced")
shown = 0
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
