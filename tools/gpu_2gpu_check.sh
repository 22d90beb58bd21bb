#!/bin/bash
# gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_2gpu_check.sh'
OUT=gpurun_out/r02_2gpu
mkdir -p $OUT
nvidia-smi -L | head -3
( timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x ) > $OUT/pytest_multi_gpu.log 2>&1; echo "multi_gpu: rc=$? $(tail -1 $OUT/pytest_multi_gpu.log)"
( timeout 300 python -m pytest tests/test_e2e_host.py -m gpu -q -k "layer_split or persistent or replay" ) > $OUT/pytest_e2e_split.log 2>&1; echo "e2e split/persistent: rc=$? $(tail -1 $OUT/pytest_e2e_split.log)"
grep -E "^(FAILED|ERROR)|Error|whole-token" $OUT/pytest_multi_gpu.log $OUT/pytest_e2e_split.log | head -12
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 64 --warmup 5 > $OUT/bench_n2.json 2> $OUT/bench_n2.err ); echo "bench N=2 rc=$?"
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 64 --warmup 5 --no-peer --steps 32 > $OUT/bench_n2_nccl.json 2> $OUT/bench_n2_nccl.err ); echo "bench N=2 (NCCL send/recv) rc=$?"
python - <<'PY'
import json
for n in ("bench_n2", "bench_n2_nccl"):
    try:
        d = json.loads(open(f"gpurun_out/r02_2gpu/{n}.json").read().strip().splitlines()[-1]); print(n, d["value"], "tok/s", d["ms_per_step"], "ms")
    except Exception as e: print(n, "failed", e); print(open(f"gpurun_out/r02_2gpu/{n}.err").read()[-1200:])
PY
