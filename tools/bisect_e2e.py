"""Bisect an end-to-end logits mismatch: run the host harness on CPU and through the plugin with individual op kinds
forced back to the CPU (B200_DISABLE_OPS), print the max relative logit error for each."""
import json, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
arch, quant, vocab = sys.argv[1], sys.argv[2], int(sys.argv[3])
layers = sys.argv[4] if len(sys.argv) > 4 else "0"
prefill = sys.argv[5] if len(sys.argv) > 5 else "37"
model = f"/tmp/bis-{arch}-{quant}.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", arch, "--quant", quant, "--out", model, "--max_length", "512", "--layers", layers], check=True, capture_output=True)
def run(ngl, dump, env=None):
    e = dict(os.environ); e.update(env or {})
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "16", "--prefill", prefill, "--decode", "3", "--max_length", "512", "--dump", dump], capture_output=True, text=True, env=e)
    if p.returncode:
        print("  harness failed:", p.stderr.strip().splitlines()[-1][:200] if p.stderr.strip() else p.returncode)
        return None
    return np.fromfile(dump, dtype=np.float32).reshape(-1, vocab)
a = run("0", "/tmp/bis_cpu.bin")
for ops in ["", "MUL_MAT_Q", "MUL_MAT_F", "GET_ROWS", "ROPE", "RMS_NORM", "SOFT_MAX", "SET_ROWS", "CPY", "ADD", "MUL", "UNARY", "SCALE", "DIAG_MASK_INF", "CONT"]:
    b = run("all", "/tmp/bis_gpu.bin", {"B200_DISABLE_OPS": ops})
    if b is None:
        continue
    rel = np.abs(a - b).max(axis=1) / np.abs(a).max(axis=1)
    print(f"{arch} {quant} disabled=[{ops}] rel={np.array2string(rel, precision=2)}", flush=True)
