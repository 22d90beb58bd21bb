#!/bin/bash
OUT=gpurun_out/r02_mk
mkdir -p $OUT
( timeout 200 python tools/mk_step_times.py 32 ) > $OUT/step_times_32.txt 2>&1; cat $OUT/step_times_32.txt | tail -12
( timeout 300 python -m pytest tests/test_decode_mk.py -m gpu -x -q -s ) > $OUT/pytest_decode_mk.log 2>&1; echo "decode_mk: rc=$? $(tail -1 $OUT/pytest_decode_mk.log)"
grep -E "rel errors|FAILED|Error|error" $OUT/pytest_decode_mk.log | head -10
( B200_MK_COOP=0 timeout 100 python tools/mk_step_times.py 8 ) 2>&1 | tail -3
