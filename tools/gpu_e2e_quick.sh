#!/bin/bash
OUT=gpurun_out/r02i_quick; mkdir -p $OUT
( timeout 400 python -m pytest tests/test_e2e_host.py -m gpu -q -x -k "replay or persistent or short_prompt or no_graph_node" ) > $OUT/pytest.log 2>&1; echo "pytest e2e subset: rc=$? $(tail -1 $OUT/pytest.log)"
grep -E "^(FAILED|ERROR)|^E  " $OUT/pytest.log | head -10
( timeout 300 python bench.py --no-cpu --steps 32 > $OUT/bench.json 2> $OUT/bench.err ); python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1]); e = d.get("e2e", {}); print("value", d["value"], "e2e", e.get("value"), e.get("ms_per_step"))
except Exception as ex: print("bench failed", ex)
PY
