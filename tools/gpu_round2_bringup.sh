#!/bin/bash
# First gpurun call of round 2: bring up what round 1 wrote but could not run (its B200 minutes were spent), in order of value.
#   gpurun --timeout 900 -- 'bash tools/gpu_round2_bringup.sh'
OUT=gpurun_out/r02_bringup
mkdir -p $OUT
# 1. the default suite must still be green on this box
( timeout 300 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_default.log 2>&1; echo "default: $(tail -1 $OUT/pytest_default.log)"
# 2. opt-in fusion: kernels + session, then the same fusion inside the plugin over the whole plugin / e2e suites
( B200_TEST_FUSED2=1 timeout 200 python -m pytest tests/test_fused2_optin.py -m gpu -q -k "not tcgen05" ) > $OUT/pytest_fused2.log 2>&1; echo "fused2 kernels/session: $(tail -1 $OUT/pytest_fused2.log)"
# 2b. tcgen05 prompt matmul: its own process and a short timeout (a wrong descriptor would hang on the MMA barrier), smallest case first
( B200_TEST_FUSED2=1 timeout 60 python -m pytest tests/test_fused2_optin.py -m gpu -q -x -k tcgen05 ) > $OUT/pytest_tcgen05.log 2>&1; echo "tcgen05 prefill: rc=$? $(tail -1 $OUT/pytest_tcgen05.log)"
( B200_FUSE2=1 timeout 300 python -m pytest tests/test_plugin_ops.py tests/test_e2e_host.py -m gpu -q ) > $OUT/pytest_plugin_fuse2.log 2>&1; echo "plugin B200_FUSE2=1: $(tail -1 $OUT/pytest_plugin_fuse2.log)"
# 2c. cluster V.P (DSMEM reduction + in-cluster quantization, attention = 2 launches)
( B200_ATTN_CLUSTER=1 B200_TEST_FUSED2=1 timeout 200 python -m pytest tests/test_fused2_optin.py tests/test_gpu_kernels.py tests/test_session.py -m gpu -q -k "cluster_pv or attn or graph_replay or decode_steps" ) > $OUT/pytest_cluster.log 2>&1; echo "cluster V.P: $(tail -1 $OUT/pytest_cluster.log)"
B200_ATTN_CLUSTER=1 timeout 100 python tools/step_breakdown.py 2>&1 | grep attn
grep -E "^(FAILED|ERROR)" $OUT/pytest_fused2.log $OUT/pytest_cluster.log $OUT/pytest_tcgen05.log $OUT/pytest_plugin_fuse2.log | head -20
# 3. what it buys: device-resident step and e2e, default vs opt-in
python bench.py --no-e2e --no-cpu > $OUT/bench_default.json 2>> $OUT/bench.err; python bench.py --no-e2e --no-cpu --fused2 > $OUT/bench_fused2.json 2>> $OUT/bench.err
python - <<'PY'
import json
for n in ("default", "fused2"):
    try:
        d = json.load(open(f"gpurun_out/r02_bringup/bench_{n}.json")); print(n, d["value"], "tok/s", d["ms_per_step"], "ms", d["gpu_launches"] // d["steps"], "launches/step")
    except Exception as e: print(n, "failed", e)
PY
( timeout 200 python tools/plugin_profile.py; echo "--- B200_FUSE2=1"; timeout 200 python tools/plugin_profile.py B200_FUSE2=1 ) > $OUT/plugin_profile.txt 2>&1; grep -E "decode_ms_mean|B200PROF|---" $OUT/plugin_profile.txt | cut -c1-200 | tail -8
# 3b. GEMV residency experiment: smaller rings so two CTAs share an SM and consecutive GEMVs overlap
for t in "16,2,8,4,0" "8,2,8,4,296" "8,3,8,4,296" "12,2,8,4,296" "8,2,4,4,592" "8,4,4,4,592"; do
  python bench.py --no-e2e --no-cpu --steps 32 --tune $t 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tune $t', d['value'], 'tok/s  gemv', d['roofline']['achieved'], 'GB/s')" 2>/dev/null || echo "tune $t failed"
done | tee $OUT/gemv_tune_sweep.txt
# 4. 2-device layer split through the plugin (only on a --gpus 2 call)
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  ( B200_TEST_MULTI_GPU=1 timeout 200 python -m pytest tests/test_e2e_host.py -m gpu -q -k layer_split ) > $OUT/pytest_2gpu.log 2>&1; echo "2-GPU split: $(tail -1 $OUT/pytest_2gpu.log)"
fi
