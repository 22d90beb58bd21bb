#!/bin/bash
# The reference's OWN CUDA backend (ggml/src/ggml-cuda compiled unmodified for sm_100 by `make -C oracle refcuda`) next to this repo's plugin,
# same unmodified host, same synthetic Llama-3-8B Q4_K file, same token stream: decode at n_past = 4096 after a real 4096-token prompt.
# ("the existing GPU kernel" comparison.)   gpurun --timeout 900 -- 'bash tools/refcuda_compare.sh'
OUT=gpurun_out/r02_refcuda
mkdir -p $OUT
M=/tmp/b200_llama3-8b_q4_K.bin
[ -f $M ] || python tools/make_model.py --arch llama3-8b --quant q4_K --out $M --max_length 4352 > /dev/null
H=oracle/_ref/bin/host_harness
export LD_LIBRARY_PATH=/usr/local/cuda/lib64:$LD_LIBRARY_PATH
for tag in ours refcuda; do
  rd=oracle/_ref/run; [ $tag = refcuda ] && rd=oracle/_ref/run_refcuda
  timeout 400 $H --model $M --ggml_dir $rd --ngl all --threads 16 --prefill 4096 --batch 512 --decode 69 --skip 5 --max_length 4352 > $OUT/$tag.json 2> $OUT/$tag.err || echo "$tag failed: $(tail -3 $OUT/$tag.err)"
  python - <<PY
import json
try:
    r = json.loads(open("$OUT/$tag.json").read().strip().splitlines()[-1])
    print("$tag", "device0", r["device0"], "| decode", round(1000.0 / r["decode_ms_mean_after_skip"], 1), "tok/s", round(r["decode_ms_mean_after_skip"], 3), "ms | prefill 4096 tokens", round(r["prefill_ms"], 1), "ms", "| load", r["load_ms"], "ms")
except Exception as e: print("$tag", "no result", e)
PY
done | tee $OUT/summary.txt
