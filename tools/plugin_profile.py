"""decode through the plugin with B200_PROFILE=1: host enqueue time, GPU time and launches per whole-model graph"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
m = bench.ensure_model_file()
e = dict(os.environ); e["B200_PROFILE"] = "1"
for kv in sys.argv[1:]:
    e[kv.split("=")[0]] = kv.split("=")[1]
p = subprocess.run([bench.HARNESS, "--model", m, "--ggml_dir", bench.RUNDIR, "--ngl", "all", "--threads", "16", "--prefill", "0", "--fake_prefill", "4096", "--decode", "70",
                    "--skip", "6", "--max_length", "4352"], capture_output=True, text=True, env=e)
print(p.stdout[-330:]); print("\n".join([l for l in p.stderr.splitlines() if "B200PROF" in l][-3:]))
