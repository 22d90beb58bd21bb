#!/bin/bash
# Short end-of-round confirmation (one gpurun call): full GPU parity suite, the bench line, the PRELOAD=0 control of the
# timing-sensitive tests, the warm per-kernel breakdown and the ncu launch list of the timed steps.
TAG=${1:-r01f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 330 python -m pytest tests -m gpu -q --durations=5 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"; grep -E "^(FAILED|ERROR)" $OUT/pytest_gpu.log | head -20
( time timeout 300 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; tail -c 2600 $OUT/bench.json
( B200_ATTN_PRELOAD=0 timeout 150 python -m pytest tests/test_session.py tests/test_e2e_host.py -m gpu -q -k "graph_replay or short_prompt or tinyllama" ) > $OUT/pytest_preload0.log 2>&1
echo "preload=0 control: $(grep -E 'passed|failed|error' $OUT/pytest_preload0.log | tail -1)"
timeout 100 python tools/step_breakdown.py > $OUT/step_breakdown.txt 2>&1; tail -11 $OUT/step_breakdown.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
echo "ncu list rc=$? rows=$(wc -l < $OUT/launches.csv)"
