#!/bin/bash
OUT=gpurun_out/r02_mk
mkdir -p $OUT
( timeout 240 python -m pytest tests/test_session.py -m gpu -x -q -k "3" ) > $OUT/pytest_session_mk.log 2>&1; echo "session[mk]: rc=$? $(tail -1 $OUT/pytest_session_mk.log)"
( timeout 300 python -m pytest tests/test_decode_mk.py -m gpu -q -s ) > $OUT/pytest_decode_mk.log 2>&1; echo "decode_mk: rc=$? $(tail -1 $OUT/pytest_decode_mk.log)"
grep -E "vs oracle|FAILED|Error|error" $OUT/pytest_session_mk.log $OUT/pytest_decode_mk.log | head -20
for cfg in "0 1" "0 0" "4 1" "8 1"; do
  set -- $cfg
  echo "=== L2AHEAD=$1 KVPREFETCH=$2"; ( B200_MK_L2AHEAD=$1 B200_MK_KVPREFETCH=$2 timeout 200 python tools/mk_step_times.py 32 ) 2>&1 | grep -v "last CTA" | tail -15 | tee $OUT/step_times_v5_$1_$2.txt
done
