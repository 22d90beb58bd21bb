#!/bin/bash
# the driver's round-end commands on one B200: smoke(), the bench line, the reference arm.  gpurun --timeout 900 -- 'bash tools/gpu_bench_line.sh tag'
TAG=${1:-r02i}
OUT=gpurun_out/${TAG}_line
mkdir -p $OUT
( timeout 200 python -m pytest tests/test_session.py -m gpu -q -x ) > $OUT/pytest_session.log 2>&1; echo "test_session: rc=$? $(tail -1 $OUT/pytest_session.log)"
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke.log 2>&1; echo "smoke rc=$? $(tail -2 $OUT/smoke.log | tr '\n' ' ')"
( timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err ); echo "bench rc=$?"
( timeout 400 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err ); echo "reference arm rc=$?"; tail -c 600 $OUT/bench_reference.json
python - <<PY
import json
d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1]); e = d.get("e2e", {})
print("value", d["value"], d["ms_per_step"], "launches", d["gpu_launches"], "| e2e", e.get("value"), e.get("ms_per_step"), "| roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "| cpu", d.get("cpu_baseline", {}).get("value"), "| clocks", d.get("clocks"))
PY
