#!/bin/bash
# whole GPU suite + the headline bench on one B200:  gpurun --timeout 1200 -- 'bash tools/gpu_full_check.sh [tag]'
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 700 python -m pytest tests -m gpu -q -x ) > $OUT/pytest_gpu.log 2>&1; echo "pytest -m gpu: rc=$? $(tail -1 $OUT/pytest_gpu.log)"
grep -E "^(FAILED|ERROR)|Error" $OUT/pytest_gpu.log | head -10
( timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err ); echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$OUT/bench.json")); e = d.get("e2e", {})
    print("value", d["value"], "tok/s", d["ms_per_step"], "ms | e2e", e.get("value"), "| roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "| cpu", d.get("cpu_baseline", {}).get("value"))
except Exception as ex: print("bench failed", ex); print(open("$OUT/bench.err").read()[-1500:])
PY
