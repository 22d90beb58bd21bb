import os, subprocess, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
for quant in ("q4_K", "q4_0"):
    model = f"/tmp/dbg_{quant}.bin"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", "tiny-test", "--quant", quant, "--out", model, "--max_length", "512"], check=True, capture_output=True)
    def run(ngl, env, pf, seed):
        out = "/tmp/dbg_logits.bin"
        e = dict(os.environ); e.update(env)
        p = subprocess.run([H, "--model", model, "--ggml_dir", RUN, "--ngl", ngl, "--threads", "8", "--prefill", str(pf), "--decode", "2", "--max_length", "512", "--dump", out,
                            "--seed", str(seed)], capture_output=True, text=True, env=e)
        assert p.returncode == 0, p.stderr[-500:]
        return np.fromfile(out, dtype=np.float32).reshape(-1, 512)
    for seed in range(1, 9):
        for pf in (9, 20):
            a = run("0", {}, pf, seed)
            r = [float((np.abs(a - run("all", env, pf, seed)).max(axis=1) / np.abs(a).max(axis=1)).max()) for env in ({}, {"B200_NO_MMQ": "1"}, {"B200_FUSE_OFF": "1"})]
            print(quant, "seed", seed, "pf", pf, "default %.2e  no_mmq %.2e  fuse_off1 %.2e" % tuple(r), flush=True)
