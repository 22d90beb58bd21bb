import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
model = "/tmp/dbg.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", "tiny-test", "--quant", "q4_K", "--out", model, "--max_length", "512"], check=True, capture_output=True)
def run(env):
    e = dict(os.environ); e.update(env); e["B200_TRACE"] = "1"
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", "all", "--threads", "8", "--prefill", "9", "--decode", "0", "--max_length", "512"], capture_output=True, text=True, env=e)
    return [l for l in p.stderr.splitlines() if l.startswith("B200TRACE")]
a = run({}); b = run({"B200_FUSE_OFF": "1"})
from collections import Counter
def keyed(lines):
    c = Counter(); out = {}
    for l in lines:
        k0 = (l.split()[1], l.split()[2]); c[k0] += 1
        out[k0 + (c[k0],)] = l
    return out
da, db = keyed(a), keyed(b)
n = 0
for k, l in db.items():
    if k in da:
        sa = da[k].split("sum=")[1]; sb = l.split("sum=")[1]
        flag = "" if sa == sb else "   <<<< DIFF"
        if flag or n == 0: print(l[:120], "| fused:", sa, flag)
        if flag: n += 1
        if n > 12: break
