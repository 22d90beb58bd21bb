"""Dump one graph node (dst/src0/src1) from the CPU run and the plugin run and check both against the oracle port.
usage: node_check.py ARCH QUANT NODE [prefill] [decode]"""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qformats as qf
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
arch, quant, node = sys.argv[1], sys.argv[2], sys.argv[3]
prefill = sys.argv[4] if len(sys.argv) > 4 else "37"
decode = sys.argv[5] if len(sys.argv) > 5 else "1"
T = {"q4_K": qf.Q4_K, "q4_0": qf.Q4_0, "q8_0": qf.Q8_0}[quant]
model = f"/tmp/nc-{arch}-{quant}.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", arch, "--quant", quant, "--out", model, "--max_length", "512"], check=True, capture_output=True)
res = {}
for ngl, tag in (("0", "cpu"), ("all", "gpu")):
    tr = f"/tmp/nc_{tag}.txt"
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "16", "--prefill", prefill, "--decode", decode, "--max_length", "512", "--trace", tr, "--trace_dump", node], capture_output=True, text=True)
    meta = [l for l in open(tr) if l.startswith("# dump")]
    print(tag, [m.strip() for m in meta])
    res[tag] = {k: np.fromfile(f"{tr}.node{node}.{k}", dtype=np.uint8) for k in ("dst", "src0", "src1")}
c, g = res["cpu"], res["gpu"]
print("src0 identical:", np.array_equal(c["src0"], g["src0"]), " src1 identical:", np.array_equal(c["src1"], g["src1"]))
x_c = c["src1"].view(np.float32); x_g = g["src1"].view(np.float32)
print("src1 max abs diff:", np.abs(x_c - x_g).max(), "max |x|:", np.abs(x_c).max())
k = x_c.size; m = c["dst"].size // 4
w = c["src0"].reshape(m, -1)
for tag, x, y in (("cpu", x_c, c["dst"].view(np.float32)), ("gpu", x_g, g["dst"].view(np.float32))):
    ref = qf.port_mul_mat(T, w, k, m, x.reshape(1, k), variant=1)[0]
    print(tag, "dst vs oracle port on ITS OWN src1: max rel", np.abs(ref - y).max() / np.abs(ref).max())
print("dst cpu vs gpu max rel", np.abs(c["dst"].view(np.float32) - g["dst"].view(np.float32)).max() / np.abs(c["dst"].view(np.float32)).max())
d = np.abs(x_c - x_g); idx = np.argsort(-d)[:8]; print("largest src1 diffs at", idx, x_c[idx], x_g[idx])
