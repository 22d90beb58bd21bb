#!/bin/bash
# GEMV change check: kernel parity tests, per-shape sweep, warm per-launch step breakdown, decode bench.  gpurun --timeout 600 -- 'bash tools/gpu_gemv_check.sh tag'
TAG=${1:-r02_gemv}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_session.py -m gpu -q -x ) > $OUT/pytest.log 2>&1; echo "pytest kernels: rc=$? $(tail -1 $OUT/pytest.log)"
grep -E "^(FAILED|ERROR)" $OUT/pytest.log | head
( timeout 200 python tools/gemv_sweep.py ) > $OUT/gemv_sweep.txt 2>&1; tail -6 $OUT/gemv_sweep.txt
( timeout 200 python tools/step_breakdown.py ) > $OUT/step_breakdown.txt 2>&1; tail -10 $OUT/step_breakdown.txt
( timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu > $OUT/bench.json 2> $OUT/bench.err ); python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "tok/s", d["ms_per_step"], "ms roofline", d["roofline"]["achieved"], d["roofline"]["frac"])
except Exception as ex: print("bench failed", ex); print(open("$OUT/bench.err").read()[-1500:])
PY
