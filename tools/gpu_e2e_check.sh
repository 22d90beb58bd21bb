#!/bin/bash
# plugin change check: the host-driven e2e tests, then the bench with its e2e leg.  gpurun --timeout 900 -- 'bash tools/gpu_e2e_check.sh tag'
TAG=${1:-r02_e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_e2e_host.py tests/test_plugin_ops.py -m gpu -q -x ) > $OUT/pytest.log 2>&1; echo "pytest e2e: rc=$? $(tail -1 $OUT/pytest.log)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest.log | head -20
( timeout 400 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err ); python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1]); e = d.get("e2e", {})
    print("value", d["value"], "tok/s", d["ms_per_step"], "ms | e2e", e, "| roofline", d["roofline"]["frac"], "| parity", d.get("parity"))
except Exception as ex: print("bench failed", ex); print(open("$OUT/bench.err").read()[-1500:])
PY
B200_GRAPH=0 timeout 400 python bench.py --no-cpu --steps 8 > $OUT/bench_nograph.json 2> $OUT/bench_nograph.err; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_nograph.json").read().strip().splitlines()[-1]); print("B200_GRAPH=0 e2e", d.get("e2e"))
except Exception as ex: print("bench failed", ex)
PY
