"""CPU (harness --trace) vs plugin (B200_TRACE, fusion ON) per-node checksums, aligned greedily by (op, shape)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H = os.path.join(ROOT, "oracle/_ref/bin/host_harness"); RUN = os.path.join(ROOT, "oracle/_ref/run")
arch, quant, pf = (sys.argv + ["tiny-test", "q4_K", "9"])[1:4]
model = "/tmp/dbg.bin"
subprocess.run([sys.executable, os.path.join(ROOT, "tools/make_model.py"), "--arch", arch, "--quant", quant, "--out", model, "--max_length", "512"], check=True, capture_output=True)
base = [H, "--model", model, "--ggml_dir", RUN, "--threads", "8", "--prefill", pf, "--decode", "0", "--max_length", "512"]
subprocess.run(base + ["--ngl", "0", "--trace", "/tmp/cpu_trace.txt"], capture_output=True, text=True)
cpu = []
for l in open("/tmp/cpu_trace.txt"):
    if l.startswith("#"): continue
    f = l.split()
    cpu.append((f[1], f[3], float(f[-2].split("=")[1]), float(f[-1].split("=")[1])))
e = dict(os.environ); e["B200_TRACE"] = "1"
for k in sys.argv[4:]:
    e[k.split("=")[0]] = k.split("=")[1]
p = subprocess.run(base + ["--ngl", "all"], capture_output=True, text=True, env=e)
gpu = []
for l in p.stderr.splitlines():
    if not l.startswith("B200TRACE"): continue
    f = l.split()
    ne = [x for x in f if x.startswith("[")][0]
    gpu.append((f[2], ne, float(f[-2].split("=")[1]), float(f[-1].split("=")[1]), f[1]))
print(len(cpu), "cpu nodes", len(gpu), "gpu nodes")
ci = 0; bad = 0
for g in gpu:
    j = ci
    while j < len(cpu) and not (cpu[j][0] == g[0] and cpu[j][1] == g[1]): j += 1
    if j == len(cpu): print("no match for", g); continue
    ci = j + 1
    c = cpu[j]
    err = max(abs(c[2] - g[2]), abs(c[3] - g[3])) / max(c[3], 1e-20)
    if err > 3e-6:
        print(f"node {g[4]} {g[0]} {g[1]} cpu sum={c[2]:.9g} abs={c[3]:.9g} | gpu sum={g[2]:.9g} abs={g[3]:.9g} rel={err:.2e}")
        bad += 1
        if bad > 14: break
