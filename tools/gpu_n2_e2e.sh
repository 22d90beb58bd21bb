#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_n2_e2e.sh' : the driver's N=2 command, e2e leg included (host process drives both devices)
OUT=gpurun_out/r02_n2e2e
mkdir -p $OUT
( timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus ${NG:-2} --steps 64 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err ); echo "bench N=2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_n2e2e/bench_n2.json").read().strip().splitlines()[-1]); print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", d.get("e2e"))
except Exception as e: print("failed", e); print(open("gpurun_out/r02_n2e2e/bench_n2.err").read()[-1500:])
PY
