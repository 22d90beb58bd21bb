#!/bin/bash
OUT=gpurun_out/r02_mk
mkdir -p $OUT
( timeout 240 python -m pytest tests/test_session.py -m gpu -x -q -k "3" ) > $OUT/pytest_session_mk.log 2>&1; echo "session[mk]: rc=$? $(tail -1 $OUT/pytest_session_mk.log)"
( timeout 300 python -m pytest tests/test_decode_mk.py -m gpu -q -s ) > $OUT/pytest_decode_mk.log 2>&1; echo "decode_mk: rc=$? $(tail -1 $OUT/pytest_decode_mk.log)"
grep -E "vs oracle|FAILED|Error|error" $OUT/pytest_session_mk.log $OUT/pytest_decode_mk.log | head -20
( timeout 200 python tools/mk_step_times.py 32 ) > $OUT/step_times_32b.txt 2>&1; tail -12 $OUT/step_times_32b.txt
( timeout 200 python bench.py --no-e2e --no-cpu --steps 32 > $OUT/bench_mk2.json 2> $OUT/bench_mk2.err ); python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_mk/bench_mk2.json")); print("mk", d["value"], "tok/s", d["ms_per_step"], "ms  frac", d["roofline"]["frac"], d["roofline"].get("mk"))
except Exception as e: print("bench failed", e); print(open("gpurun_out/r02_mk/bench_mk2.err").read()[-1500:])
PY
