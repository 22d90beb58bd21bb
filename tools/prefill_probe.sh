set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import bench, json, subprocess, time
m = bench.ensure_model_file()
for pf, batch in [(512, 512), (2048, 2048), (4096, 4096), (4096, 512)]:
    cmd = [bench.HARNESS, "--model", m, "--ggml_dir", bench.RUNDIR, "--ngl", "all", "--threads", "16", "--prefill", str(pf), "--batch", str(batch),
           "--decode", "24", "--skip", "8", "--max_length", "4352"]
    t = time.time()
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(pf, batch, p.returncode, time.time() - t, p.stdout.strip().splitlines()[-1][:600] if p.stdout.strip() else p.stderr[-800:], flush=True)
PY
