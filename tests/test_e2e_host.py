"""End-to-end drop-in test (-m gpu): the UNMODIFIED reference host (oracle/_ref/bin/host_harness = chatllm objects +
a token-id driver) runs the same synthetic GGMM model once on its own CPU backend (-ngl 0, the oracle) and once with
every layer offloaded through the boundary to libggml-cuda.so (-ngl all).  Logits must agree within 1e-3 relative
(BASELINE.json north star) at every step, prefill and decode."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "host_harness")
RUNDIR = os.path.join(ROOT, "oracle", "_ref", "run")

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(HARNESS), reason="oracle/_ref/bin/host_harness not built")]


def run_host(model, ngl, dump, prefill, decode, max_length=512, threads=16, extra_env=None, batch=4096):
    env = dict(os.environ)
    env.update(extra_env or {})
    cmd = [HARNESS, "--model", model, "--ggml_dir", RUNDIR, "--ngl", ngl, "--threads", str(threads), "--prefill", str(prefill),
           "--decode", str(decode), "--max_length", str(max_length), "--dump", dump, "--batch", str(batch)]
    p = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.loads(p.stdout.strip().splitlines()[-1]), p.stderr


def make_model(tmp, arch, quant, layers=0, max_length=512):
    out = os.path.join(tmp, f"{arch}-{quant}.bin")
    cmd = [sys.executable, os.path.join(ROOT, "tools", "make_model.py"), "--arch", arch, "--quant", quant, "--out", out,
           "--max_length", str(max_length)]
    if layers:
        cmd += ["--layers", str(layers)]
    subprocess.run(cmd, check=True, capture_output=True)
    return out


def compare(tmp, model, vocab, prefill, decode, max_length=512):
    cpu_dump, gpu_dump = os.path.join(tmp, "cpu.bin"), os.path.join(tmp, "gpu.bin")
    rc, _ = run_host(model, "0", cpu_dump, prefill, decode, max_length)
    rg, err = run_host(model, "all", gpu_dump, prefill, decode, max_length, extra_env={"GGML_SCHED_DEBUG": "1"})
    assert rg["device0"] == "CUDA0" and rg["devices"] >= 2
    a = np.fromfile(cpu_dump, dtype=np.float32).reshape(-1, vocab)
    b = np.fromfile(gpu_dump, dtype=np.float32).reshape(-1, vocab)
    assert a.shape == b.shape and a.shape[0] == decode + 1
    assert np.isfinite(b).all()
    rel = np.abs(a - b).max(axis=1) / np.abs(a).max(axis=1)
    assert rel.max() <= 1e-3, rel
    assert (a.argmax(axis=1) == b.argmax(axis=1)).all()
    return rel.max(), rc, rg, err


@pytest.mark.parametrize("arch,quant", [("tiny-test", "q4_K"), ("tiny-test", "q4_0"), ("tiny-test", "q8_0"), ("qwen2-test", "q4_0"),
                                        ("qwen2-test", "q4_K")])
def test_tiny_models_logits_match_cpu(tmp_path, arch, quant):
    model = make_model(str(tmp_path), arch, quant)
    rel, rc, rg, err = compare(str(tmp_path), model, 512, prefill=37, decode=6)
    print(arch, quant, "max rel logit err", rel)


def test_no_graph_node_runs_on_cpu(tmp_path):
    """With -ngl all every compute split of the decode graph must be assigned to CUDA0 (the CPU only feeds inputs)."""
    model = make_model(str(tmp_path), "tiny-test", "q4_K")
    _, err = run_host(model, "all", os.path.join(str(tmp_path), "g.bin"), 8, 2, extra_env={"GGML_SCHED_DEBUG": "2"})
    splits = [l for l in err.splitlines() if l.startswith("## SPLIT")]
    assert splits, err[-1500:]
    assert all("CUDA0" in l for l in splits), [l for l in splits if "CUDA0" not in l][:5]


def test_tinyllama_q8_0_real_shape_few_layers(tmp_path):
    """TinyLlama-1.1B shapes (BASELINE.json configs[0]) with 4 of 22 layers, 300-token prefill + decode"""
    model = make_model(str(tmp_path), "tinyllama-1.1b", "q8_0", layers=4, max_length=512)
    rel, rc, rg, _ = compare(str(tmp_path), model, 32000, prefill=300, decode=4)
    print("tinyllama q8_0 max rel", rel, "cpu ms/tok", rc["decode_ms_median"], "gpu ms/tok", rg["decode_ms_median"])


def test_llama3_8b_q4k_real_shape_few_layers(tmp_path):
    """Llama-3-8B shapes (configs[1]) with 2 of 32 layers"""
    model = make_model(str(tmp_path), "llama3-8b", "q4_K", layers=2, max_length=512)
    rel, rc, rg, _ = compare(str(tmp_path), model, 128256, prefill=100, decode=3)
    print("llama3-8b q4_K max rel", rel, "cpu ms/tok", rc["decode_ms_median"], "gpu ms/tok", rg["decode_ms_median"])
