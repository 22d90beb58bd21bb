#!/bin/bash
# One gpurun call that refreshes everything the round is judged on: GPU parity tests, the bench line, the warm per-kernel
# breakdown, the ncu launch list of the bench command and one `ncu --set full` capture of the attention kernels.
#   gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt

( time timeout ${PYTEST_TIMEOUT:-540} python -m pytest tests -m gpu -x -q --durations=15 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"

( time timeout 420 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; tail -c 2500 $OUT/bench.json

timeout 120 python tools/step_breakdown.py > $OUT/step_breakdown.txt 2>&1
cat $OUT/step_breakdown.txt | tail -14
( echo "B200_ATTN_OLD_PV=1:"; B200_ATTN_OLD_PV=1 timeout 120 python tools/step_breakdown.py 2>&1 | grep attn
  echo "B200_ATTN_PRELOAD=0:"; B200_ATTN_PRELOAD=0 timeout 120 python tools/step_breakdown.py 2>&1 | grep attn ) > $OUT/step_breakdown_ab.txt; cat $OUT/step_breakdown_ab.txt

# host-side breakdown of the e2e arm (B200PROF lines: enqueue time / GPU time / launches per whole-model graph), staged vs synchronous inputs
( timeout 200 python tools/plugin_profile.py; echo "--- B200_SYNC_INPUTS=1"; timeout 200 python tools/plugin_profile.py B200_SYNC_INPUTS=1 ) > $OUT/plugin_profile.txt 2>&1
cat $OUT/plugin_profile.txt

timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
echo "ncu list rc=$? rows=$(wc -l < $OUT/launches.csv)"

timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn -c 4 -f -o $OUT/attn_full \
    python tools/ncu_attn.py > $OUT/ncu_attn.log 2>&1
echo "ncu attn rc=$?"; ls -la $OUT
