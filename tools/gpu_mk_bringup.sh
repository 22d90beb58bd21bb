#!/bin/bash
# first GPU run of the persistent decode kernel: small tests first, under a short timeout (a hang must not become a strike)
OUT=gpurun_out/r02_mk
mkdir -p $OUT
( timeout 240 python -m pytest tests/test_session.py -m gpu -x -q -k "3" ) > $OUT/pytest_session_mk.log 2>&1; echo "session[mk]: rc=$? $(tail -1 $OUT/pytest_session_mk.log)"
( timeout 300 python -m pytest tests/test_decode_mk.py -m gpu -x -q -s ) > $OUT/pytest_decode_mk.log 2>&1; echo "decode_mk: rc=$? $(tail -1 $OUT/pytest_decode_mk.log)"
grep -E "rel errors|FAILED|Error|error" $OUT/pytest_session_mk.log $OUT/pytest_decode_mk.log | head -20
( timeout 200 python bench.py --no-e2e --no-cpu --steps 32 > $OUT/bench_mk.json 2> $OUT/bench_mk.err ); echo "bench mk rc=$?"; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r02_mk/bench_mk.json")); print("mk", d["value"], "tok/s", d["ms_per_step"], "ms  frac", d["roofline"]["frac"], d["roofline"].get("mk"))
except Exception as e: print("bench failed", e); print(open("gpurun_out/r02_mk/bench_mk.err").read()[-1500:])
PY
for t in "16,3" "12,3" "8,4" "16,2"; do
  ks=${t%,*}; st=${t#*,}
  B200_MK_KS=$ks B200_MK_STAGES=$st timeout 120 python bench.py --no-e2e --no-cpu --steps 32 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ks,stages $t', d['value'], 'tok/s', d['roofline']['frac'], d['roofline']['mk'])" 2>/dev/null || echo "tune $t failed"
done | tee $OUT/mk_tune.txt
