#!/bin/bash
# One gpurun call that refreshes what round 2 is judged on (1 x B200):  gpurun --timeout 1500 -- 'bash tools/gpu_round2_final.sh [tag]'
TAG=${1:-r02f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1; nproc > $OUT/nproc.txt
( time timeout 700 python -m pytest tests -m gpu -q --durations=8 ) > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? : $(grep -E 'passed|failed|error' $OUT/pytest_gpu.log | tail -1)"; grep -E "^(FAILED|ERROR)" $OUT/pytest_gpu.log | head -20
( time timeout 600 python bench.py ) > $OUT/bench.json 2> $OUT/bench.err
echo "bench rc=$?"; tail -c 3000 $OUT/bench.json
timeout 200 python tools/step_breakdown.py > $OUT/step_breakdown.txt 2>&1; tail -11 $OUT/step_breakdown.txt
timeout 200 python tools/gemv_sweep.py > $OUT/gemv_sweep.txt 2>&1; tail -8 $OUT/gemv_sweep.txt
( timeout 200 python tools/plugin_profile.py ) > $OUT/plugin_profile.txt 2>&1; grep -E "B200PROF|decode_ms_mean" $OUT/plugin_profile.txt | cut -c1-220 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_bench.log 2>&1
echo "ncu list rc=$? rows=$(wc -l < $OUT/launches.csv)"; python tools/ncu_launch_summary.py $OUT/launches.csv | head -14
# prompt path: launch list of a 2048-token prompt (batch 512) through the plugin, 4-layer Llama-3-8B-shaped model (tcgen05 prompt matmul)
python tools/make_model.py --arch llama3-8b --quant q4_K --layers 4 --out /tmp/l3_4.bin --max_length 4352 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/prefill_launches.csv \
    oracle/_ref/bin/host_harness --model /tmp/l3_4.bin --ggml_dir oracle/_ref/run --ngl all --threads 16 --prefill 2048 --batch 512 --decode 1 \
    --max_length 4352 > $OUT/ncu_prefill.log 2>&1
echo "ncu prefill rc=$?"; python tools/ncu_launch_summary.py $OUT/prefill_launches.csv | head -10
# full captures: the decode GEMV (DRAM traffic for roofline.traffic) and the tcgen05 prompt matmul (tensor-pipe activity)
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemv_q_kernel -s 2 -c 3 -f -o $OUT/gemv_full python tools/ncu_gemv.py q4_K 4096 14336 > $OUT/ncu_gemv.log 2>&1; echo "ncu gemv rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mmq_tc -s 1 -c 2 -f -o $OUT/mmq_tc_full python tools/ncu_prefill_mm.py q4_K 4096 14336 2048 > $OUT/ncu_mmq.log 2>&1; echo "ncu mmq_tc rc=$?"
ls -la $OUT | head -30
