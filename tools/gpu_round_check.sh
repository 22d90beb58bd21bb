#!/bin/bash
# One gpurun call that refreshes everything the round is judged on: GPU parity tests, the bench line, the warm per-kernel
# breakdown, the ncu launch list of the bench command and one `ncu --set full` capture of the attention kernels.
#   gpurun --timeout 900 -- 'bash tools/gpu_round_check.sh [tag]'
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
nproc > $OUT/nproc.txt

( time timeout ${PYTEST_TIMEOUT:-540} python -m pytest tests -m gpu -q --durations=8 ${PYTEST_ARGS:-} ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"; grep -E "^(FAILED|ERROR)" $OUT/pytest_gpu.log | head -20

( time timeout 420 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; tail -c 2500 $OUT/bench.json

timeout 120 python tools/step_breakdown.py > $OUT/step_breakdown.txt 2>&1
cat $OUT/step_breakdown.txt | tail -14

# host-side breakdown of the e2e arm (B200PROF lines: enqueue time / GPU time / launches per whole-model graph), staged vs synchronous inputs
( timeout 200 python tools/plugin_profile.py ) > $OUT/plugin_profile.txt 2>&1
cat $OUT/plugin_profile.txt

timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
echo "ncu list rc=$? rows=$(wc -l < $OUT/launches.csv)"

# prompt path: launch list of a 2048-token prompt (batch 512) through the plugin, 4-layer Llama-3-8B-shaped model
python tools/make_model.py --arch llama3-8b --quant q4_K --layers 4 --out /tmp/l3_4.bin --max_length 4352 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/prefill_launches.csv \
    oracle/_ref/bin/host_harness --model /tmp/l3_4.bin --ggml_dir oracle/_ref/run --ngl all --threads 16 --prefill 2048 --batch 512 --decode 1 \
    --max_length 4352 > $OUT/ncu_prefill.log 2>&1
echo "ncu prefill rc=$?"; python tools/ncu_launch_summary.py $OUT/prefill_launches.csv | head -16
oracle/_ref/bin/host_harness --model /tmp/l3_4.bin --ggml_dir oracle/_ref/run --ngl all --threads 16 --prefill 2048 --batch 512 --decode 4 --max_length 4352 2>/dev/null | tail -1 | cut -c1-400

timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn -c 4 -f -o $OUT/attn_full \
    python tools/ncu_attn.py > $OUT/ncu_attn.log 2>&1
echo "ncu attn rc=$?"; ls -la $OUT
