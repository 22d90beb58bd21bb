import os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
model = "/tmp/dbg.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", "tiny-test", "--quant", "q4_K", "--out", model, "--max_length", "512"], check=True, capture_output=True)
def run(ngl, env, pf):
    out = "/tmp/dbg_logits.bin"
    e = dict(os.environ); e.update(env)
    p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "8", "--prefill", str(pf), "--decode", "2", "--max_length", "512", "--dump", out], capture_output=True, text=True, env=e)
    assert p.returncode == 0, p.stderr[-500:]
    return np.fromfile(out, dtype=np.float32).reshape(-1, 512)
for pf in (5, 8, 9, 12, 33):
    a = run("0", {}, pf)
    for name, env in [("default", {}), ("NO_MMQ", {"B200_NO_MMQ": "1"}), ("NORM_NOQ+NO_MMQ", {"B200_NO_MMQ": "1", "B200_NORM_NOQ": "1"}), ("NORM_NOQ", {"B200_NORM_NOQ": "1"}),
                      ("FUSE_OFF=1", {"B200_FUSE_OFF": "1"})]:
        b = run("all", env, pf)
        print(pf, name, (np.abs(a - b).max(axis=1) / np.abs(a).max(axis=1)), flush=True)
