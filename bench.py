#!/usr/bin/env python
"""bench.py — decode tokens/s @4096 ctx, Q4_K Llama-3-8B (BASELINE.json metric), on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" = one decode token of the whole model at n_past = 4096:
  value    whole-job tokens/s with everything resident in HBM: the decode step (chatllm.cpp_b200/session.py, all kernels
           through the C ABI) captured in one CUDA graph and replayed K times; timed with CUDA events on the replay
           stream.  Inputs >> L2 (4.2 GB of weights + 0.5 GB of KV per token), so no L2 flush is needed.
  e2e      the same metric through the reference-facing boundary with HOST buffers: the unmodified chatllm host
           (oracle/_ref/bin/host_harness) drives libggml-cuda.so; every step uploads the token id / positions and reads
           the 513 KB logits row back (h2d/d2h bytes counted from those tensors); wall clock of the host loop.
  roofline the dominant kernel gemv_q_kernel<FmtQ4K,...>: algorithmic bytes of all quantized matmuls of a token
           (m*(k/256)*144 + 4k + 4m each) / CUDA-event time of exactly those launches replayed alone,
           vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's own ggml CPU backend (oracle/_ref, `kind: reference`) on the same file, bounded sample.
--impl reference: the reference CPU path is the arm being timed (rank 0 only).
Multi-GPU (N>1, torchrun): the path is one sequential stream, sharded BY LAYER (KV cache sharded with it); the hidden
state row is handed over with NCCL send/recv.  Total work is fixed as N grows -> scaling "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODEL, QUANT, N_PAST = "llama3-8b", "q4_K", 4096
HARNESS = os.path.join(ROOT, "oracle", "_ref", "bin", "host_harness")
RUNDIR = os.path.join(ROOT, "oracle", "_ref", "run")


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    def __init__(self, idx=0):
        super().__init__(daemon=True)
        self.idx, self.rows, self._stop_evt = idx, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([v.strip() for v in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def ensure_model_file(layers=0):
    path = f"/tmp/b200_{MODEL}_{QUANT}{'_L%d' % layers if layers else ''}.bin"
    if not os.path.exists(path):
        cmd = [sys.executable, os.path.join(ROOT, "tools", "make_model.py"), "--arch", MODEL, "--quant", QUANT, "--out", path + ".tmp",
               "--max_length", str(N_PAST + 256)]
        if layers:
            cmd += ["--layers", str(layers)]
        subprocess.run(cmd, check=True, capture_output=True)
        os.replace(path + ".tmp", path)
    return path


def run_harness(model, ngl, decode, threads, extra=(), real_prefill=False, timeout=3000):
    # real_prefill: the 4096-token prompt is actually evaluated (plugin arm: prompt GEMM path, ~6 s); otherwise the host is
    # told 4096 positions are cached (KV buffers zero-filled) — decode cost does not depend on the cached VALUES
    pf = ["--prefill", str(N_PAST), "--batch", "512"] if real_prefill else ["--prefill", "0", "--fake_prefill", str(N_PAST)]
    cmd = [HARNESS, "--model", model, "--ggml_dir", RUNDIR, "--ngl", ngl, "--threads", str(threads)] + pf + [
           "--decode", str(decode), "--max_length", str(N_PAST + 256)] + list(extra)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError("host_harness failed: " + p.stderr[-1500:])
    return json.loads(p.stdout.strip().splitlines()[-1])


def cpu_reference_arm(steps, warmup, probe=(8, 16, 32, 64)):
    """the reference's own CPU implementation of the path (oracle/_ref = unmodified reference build) on the host cores.
    ggml's thread pool spin-waits and does not scale monotonically on a many-core host (measured r01: 64 threads on the 128-vCPU GPU box
    are 3x SLOWER than 8 threads on an 8-vCPU container), so the thread count is chosen by a short probe — the reference gets its
    best configuration, which is what 'all the host threads it can use' has to mean for a fair ratio."""
    cores = os.cpu_count() or 1
    model = ensure_model_file()
    cands = sorted({t for t in probe if t <= cores} | ({cores} if cores <= 16 else set()))
    tried = {}
    for t in cands:
        try:
            r = run_harness(model, "0", 3, t, ["--skip", "1"])
            tried[t] = r["decode_ms_mean_after_skip"]
        except Exception:  # noqa: BLE001
            continue
    threads = min(tried, key=tried.get) if tried else min(cores, 64)
    r = run_harness(model, "0", warmup + steps, threads, ["--skip", str(warmup)])
    ms = r["decode_ms_mean_after_skip"]
    return {"value": 1000.0 / ms, "ms_per_step": ms, "cores": threads, "host_cores": cores,
            "threads_probed_ms_per_token": {str(k): round(v, 1) for k, v in tried.items()},
            "sample": f"{steps} decode tokens of the full {MODEL} {QUANT} model at n_past={N_PAST} (KV cache zero-filled, no prefill), {threads} threads "
                      f"(best of a probe over {list(tried)} threads on {cores} host cores)"}


def clocks_during(fn, idx=0, min_s=1.0):
    """run fn() repeatedly for >= min_s while sampling nvidia-smi clocks; returns the clocks dict"""
    import torch
    s = ClockSampler(idx); s.start()
    t_end = time.time() + min_s
    while time.time() < t_end:
        fn()
        torch.cuda.synchronize()
    return s.stop()


def run_config3(a):
    """BASELINE.json configs[2]: Qwen2.5-7B Q4_0, prefill 2048 + decode 512, 1 x B200.
    value  = decode tokens/s over 512 greedy steps after position 2048 (persistent kernel, device-resident);
    prefill = the prompt's quantized matmuls (28 layers x 7 matrices at n = 2048 columns: 2*n*sum(m*k) = 28.96 TFLOP) through
              b200_quantize_plain + b200_mul_mat_q_batched, timed with CUDA events -> TFLOP/s vs the measured bf16 tensor peak (roofline);
    e2e    = the unmodified chatllm host through libggml-cuda.so: real 2048-token prompt, then 512 decode steps."""
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package(); L = pkg.lib()
    from chatllm_cpp_b200 import session as S
    torch.cuda.set_device(0)
    model, wtype, P, D = "qwen2.5-7b", pkg.Q4_0, 2048, 512
    if os.environ.get("B200_CFG3_MODEL"):   # e.g. "llama3-8b:q4_K": the same measurement on another architecture / format
        model, qn = os.environ["B200_CFG3_MODEL"].split(":")
        wtype = {"q4_K": pkg.Q4_K, "q4_0": pkg.Q4_0, "q8_0": pkg.Q8_0}[qn]
    cfg = S.make_config(model, wtype, layers=a.layers or None, max_len=P + D + 64)
    sess = S.DecodeSession(cfg, seed=0, fused=(3 if a.mk else True))
    sess.fill_kv_random(P + D, seed=1)
    sess.mk_advance = True
    sess.tok.fill_(12345 % cfg.vocab)
    graph = sess.capture(P + D // 2)      # the per-op step is captured at the middle position of the 512-token decode (its grids depend on n_kv)
    sess.pos.fill_(P if a.mk else P + D // 2)
    for _ in range(max(a.warmup, 3)):
        graph.replay()
    torch.cuda.synchronize()
    sess.pos.fill_(P if a.mk else P + D // 2)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(D):
        graph.replay()
    e1.record(); torch.cuda.synchronize()
    dec_ms = e0.elapsed_time(e1) / D
    pk, pk_kind = peaks()
    tok_bytes = cfg.weight_bytes_per_token() + cfg.kv_bytes_per_token(P + D / 2)
    # ---- prefill matmuls (the tensor-core path), every layer's own weights: working set 4 GB >> L2
    st = torch.cuda.current_stream().cuda_stream
    n = P
    xs = {k: torch.randn((n, k), device="cuda") for k in (cfg.hidden, cfg.ffn)}
    pact = {k: torch.empty(L.b200_pact_col_bytes(wtype, k) * n, dtype=torch.uint8, device="cuda") for k in xs}
    ybuf = torch.empty(n * max(cfg.ffn, cfg.hidden), dtype=torch.float32, device="cuda")
    shapes = []
    for W in sess.layers:
        shapes += [(W.wq, cfg.hidden, cfg.hidden), (W.wk, cfg.hidden, cfg.kv_hidden), (W.wv, cfg.hidden, cfg.kv_hidden), (W.wo, cfg.hidden, cfg.hidden),
                   (W.wgate, cfg.hidden, cfg.ffn), (W.wup, cfg.hidden, cfg.ffn), (W.wdown, cfg.ffn, cfg.hidden)]
    tc = os.environ.get("B200_MMQ_TCGEN05", "1") != "0"   # default: the tcgen05 kernel (the library default)
    mm = L.b200_mul_mat_q_batched_tc if tc else L.b200_mul_mat_q_batched

    def prefill_mm():
        rc = 0
        for k in xs:
            rc |= L.b200_quantize_plain(wtype, xs[k].data_ptr(), k, k, n, pact[k].data_ptr(), st)
        for (w, k, m) in shapes:
            rc |= mm(wtype, w.data_ptr(), k, m, pact[k].data_ptr(), n, ybuf.data_ptr(), m, 0, st)
        if rc:
            raise RuntimeError(f"prefill matmul failed rc={rc}")
    for _ in range(2):
        prefill_mm()
    torch.cuda.synchronize()
    reps = 3
    e0.record()
    for _ in range(reps):
        prefill_mm()
    e1.record(); torch.cuda.synchronize()
    pf_ms = e0.elapsed_time(e1) / reps
    flop = 2.0 * n * sum(m * k for (_, k, m) in shapes)
    tfl = flop / pf_ms / 1e9
    clocks = clocks_during(lambda: prefill_mm(), 0, 1.0)
    out = {"metric": "decode tokens/s after a 2048-token prompt (Q4_0 Qwen2.5-7B); prefill tensor-pipe %", "value": round(1000.0 / dec_ms, 2), "unit": "tokens/s",
           "n_gpus": 1, "steps": D, "warmup": max(a.warmup, 3), "ms_per_step": round(dec_ms, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "int8 x int4 -> int32 dot, fp32 scales/accumulate", "data": "synthetic (seeded random valid Q4_0 blocks, random F16 KV cache)",
           "config": {"workload": "Qwen2.5-7B Q4_0 prefill 2048 + decode 512, 1 x B200 (BASELINE.json configs[2])", "n_past": P, "decode_steps": D,
                      "l2": "inputs (4.2 GB per token / per prompt pass) exceed L2; no flush needed"},
           "gpu_launches": D * sess.launches_per_step, "clocks": clocks,
           "frac_of_hbm_roofline_whole_token": round(1000.0 / dec_ms * tok_bytes / 1e9 / pk["hbm_gbs"], 4),
           "roofline": {"bound": "tensor", "kernel": ("mmq_tc_kernel (tcgen05.mma kind::i8)" if tc else "mmq_kernel (mma.sync m16n8k32.s8)"), "achieved": round(tfl, 1),
                        "peak": pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), "unit": "TFLOP/s",
                        "frac": round(tfl / pk.get("bf16_tflops_sustained", pk["bf16_tflops"]), 4), "peak_kind": pk_kind + " bf16 dense sustained (int8 peak is higher)",
                        "algorithmic_TFLOP_per_pass": round(flop / 1e12, 2), "ms_per_pass": round(pf_ms, 2), "launches": len(shapes), "traffic": None,
                        "prompt_tokens_per_s_matmuls_only": round(n / pf_ms * 1e3, 1)}}
    if not a.no_e2e and os.path.exists(HARNESS):
        try:
            del sess, graph
            torch.cuda.empty_cache()
            path = f"/tmp/b200_{model}_q4_0{'_L%d' % a.layers if a.layers else ''}.bin"
            if not os.path.exists(path):
                cmd = [sys.executable, os.path.join(ROOT, "tools", "make_model.py"), "--arch", model, "--quant", "q4_0", "--out", path + ".tmp", "--max_length", str(P + D + 64)]
                if a.layers:
                    cmd += ["--layers", str(a.layers)]
                subprocess.run(cmd, check=True, capture_output=True)
                os.replace(path + ".tmp", path)
            cmdl = [HARNESS, "--model", path, "--ggml_dir", RUNDIR, "--ngl", "all", "--threads", "16", "--prefill", str(P), "--batch", str(P), "--decode", str(D),
                    "--max_length", str(P + D + 64), "--skip", "3"]
            p = subprocess.run(cmdl, capture_output=True, text=True, timeout=3000)
            r = json.loads(p.stdout.strip().splitlines()[-1])
            out["e2e"] = {"value": round(1000.0 / r["decode_ms_mean_after_skip"], 2), "unit": "tokens/s", "ms_per_step": round(r["decode_ms_mean_after_skip"], 4),
                          "h2d_bytes_per_step": 4 + 4 * cfg.layers, "d2h_bytes_per_step": 4 * cfg.vocab,
                          "prefill": {"tokens": P, "ms": r["prefill_ms"], "tokens_per_s": round(P / r["prefill_ms"] * 1e3, 1)},
                          "path": "unmodified chatllm host -> libggml-cuda.so"}
        except Exception as ex:  # noqa: BLE001
            out["e2e"] = {"value": None, "error": str(ex)[-300:]}
    print(json.dumps(out))


def run_config5(a):
    """BASELINE.json configs[4]: Mixtral-8x7B Q4_K MoE decode, 1 x B200 (26 GB resident, 7.17 GB streamed per token).
    The model runs through the drop-in boundary (router -> top-2 -> expert-indexed GEMV, MODE 2 / 3 of gemv.cu): `value` is the unmodified
    host's decode loop (the only whole-model MoE driver in this repo), and the roofline is the expert-indexed GEMV kernel on the real
    shapes, replayed alone with CUDA events."""
    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package(); L = pkg.lib()
    torch.cuda.set_device(0)
    wtype, hidden, ffn, n_exp, top, layers = pkg.Q4_K, 4096, 14336, 8, 2, (a.layers or 32)
    st = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda"); gen.manual_seed(0)
    from chatllm_cpp_b200 import session as S
    # expert stacks of `rot` layers (each 3 x 8 x 33 MB = 793 MB) so that the replayed stream never sits in L2
    rot = 4
    stacks = [(S.synth_weights_device(wtype, n_exp * ffn, hidden, gen), S.synth_weights_device(wtype, n_exp * ffn, hidden, gen),
               S.synth_weights_device(wtype, n_exp * hidden, ffn, gen)) for _ in range(rot)]
    ids = torch.tensor([3, 6], dtype=torch.int32, device="cuda")
    xq = torch.empty(L.b200_qact_col_bytes(wtype, hidden), dtype=torch.uint8, device="cuda")
    gq = torch.empty(L.b200_qact_col_bytes(wtype, ffn) * top, dtype=torch.uint8, device="cuda")
    x = torch.randn((1, hidden), device="cuda"); gx = torch.randn((top, ffn), device="cuda")
    L.b200_quantize_act(wtype, x.data_ptr(), hidden, hidden, 1, xq.data_ptr(), st)
    L.b200_quantize_act(wtype, gx.data_ptr(), ffn, ffn, top, gq.data_ptr(), st)
    yg = torch.empty((top, ffn), dtype=torch.float32, device="cuda"); yd = torch.empty((top, hidden), dtype=torch.float32, device="cuda")

    def moe_layers():
        rc = 0
        st = torch.cuda.current_stream().cuda_stream   # (the capture stream while the graph is being recorded)
        for i in range(layers):
            wg, wu, wd = stacks[i % rot]
            rc |= L.b200_mul_mat_q_id(wtype, 1, wg.data_ptr(), wu.data_ptr(), hidden, ffn, n_exp, ids.data_ptr(), top, xq.data_ptr(), 1, yg.data_ptr(), ffn, st)
            rc |= L.b200_mul_mat_q_id(wtype, 0, wd.data_ptr(), 0, ffn, hidden, n_exp, ids.data_ptr(), top, gq.data_ptr(), top, yd.data_ptr(), hidden, st)
        if rc:
            raise RuntimeError(f"mul_mat_q_id failed rc={rc}")
    moe_layers(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); gs = torch.cuda.Stream()
    with torch.cuda.stream(gs):
        with torch.cuda.graph(g, stream=gs):
            moe_layers()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    exp_bytes = layers * top * (2 * ffn * (hidden // 256) * 144 + hidden * (ffn // 256) * 144)
    pk, pk_kind = peaks()
    ach = exp_bytes / ms / 1e6
    clocks = clocks_during(lambda: g.replay(), 0, 1.0)
    out = {"metric": "decode tokens/s (Q4_K Mixtral-8x7B, top-2 of 8 experts)", "value": None, "unit": "tokens/s", "n_gpus": 1, "steps": a.steps, "warmup": max(a.warmup, 3),
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8 x int4 -> int32 dot, fp32 scales/accumulate",
           "data": "synthetic (seeded random valid Q4_K blocks)",
           "config": {"workload": "Mixtral-8x7B Q4_K MoE decode, 1 x B200 (BASELINE.json configs[4])", "l2": "expert stacks of 4 layers rotate (3.2 GB) in the kernel-only leg; whole model 26 GB in the e2e leg"},
           "clocks": clocks, "gpu_launches": 2 * layers * reps,
           "roofline": {"bound": "hbm", "kernel": "gemv_q_kernel<FmtQ4K, MODE 2/3> (expert-indexed, ids on device)", "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": round(ach / pk["hbm_gbs"], 4), "peak_kind": pk_kind, "launches": 2 * layers, "avg_launch_us": round(ms * 1e3 / (2 * layers), 2),
                        "algorithmic_MB_per_token_experts": round(exp_bytes / 1e6, 1), "traffic": None}}
    if not a.no_e2e and os.path.exists(HARNESS):
        try:
            del stacks
            torch.cuda.empty_cache()
            path = f"/tmp/b200_mixtral-8x7b_q4_K{'_L%d' % a.layers if a.layers else ''}.bin"
            if not os.path.exists(path):
                cmd = [sys.executable, os.path.join(ROOT, "tools", "make_model.py"), "--arch", "mixtral-8x7b", "--quant", "q4_K", "--out", path + ".tmp", "--max_length", "1024"]
                if a.layers:
                    cmd += ["--layers", str(a.layers)]
                subprocess.run(cmd, check=True, capture_output=True)
                os.replace(path + ".tmp", path)
            cmdl = [HARNESS, "--model", path, "--ggml_dir", RUNDIR, "--ngl", "all", "--threads", "16", "--prefill", "64", "--batch", "64", "--decode", str(a.warmup + a.steps),
                    "--max_length", "1024", "--skip", str(a.warmup)]
            env = dict(os.environ); env["B200_STATS"] = "1"
            p = subprocess.run(cmdl, capture_output=True, text=True, timeout=3000, env=env)
            r = json.loads(p.stdout.strip().splitlines()[-1])
            e_ms = r["decode_ms_mean_after_skip"]
            tok_bytes = 7171.1e6 + 2 * 1024 * 2 * 64 * 32
            out["value"] = round(1000.0 / e_ms, 2); out["ms_per_step"] = round(e_ms, 4)
            out["frac_of_hbm_roofline_whole_token"] = round(1000.0 / e_ms * tok_bytes / 1e9 / pk["hbm_gbs"], 4)
            out["e2e"] = {"value": round(1000.0 / e_ms, 2), "unit": "tokens/s", "ms_per_step": round(e_ms, 4), "h2d_bytes_per_step": 4 + 4 * 32, "d2h_bytes_per_step": 4 * 32000,
                          "load_ms": r.get("load_ms"), "path": "unmodified chatllm host -> libggml-cuda.so (node-by-node path with fusion: MoE is not a persistent-kernel shape yet)",
                          "n_past": "64..", "stats": [l for l in p.stderr.splitlines() if l.startswith("B200STATS")][:1]}
        except Exception as ex:  # noqa: BLE001
            out["e2e"] = {"value": None, "error": str(ex)[-300:]}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="llama3-8b-q4_K-decode4096", choices=["llama3-8b-q4_K-decode4096", "qwen2.5-7b-q4_0-prefill2048", "mixtral-8x7b-q4_K-decode"],
                    help="default = the headline (BASELINE.json configs[1]); the other two print the secondary lines for configs[2] / configs[4]")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--tune", default="", help="ks,stages,warps,rg,grid for the GEMV pipeline (0 = default)")
    ap.add_argument("--unfused", action="store_true")
    ap.add_argument("--mk", action="store_true", help="the persistent whole-token kernel (csrc/decode_mk.cu) instead of the per-op step replayed as a CUDA graph; "
                    "measured slower on B200 (DESIGN.md §7.1), kept as an opt-in")
    ap.add_argument("--no-peer", action="store_true", help="multi-GPU A/B aid: hand the hidden row over with torch.distributed send/recv instead of NVLink mailboxes")
    ap.add_argument("--layers", type=int, default=0, help="debug: fewer layers (result is then NOT the BASELINE config)")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 3)
    if a.config != "llama3-8b-q4_K-decode4096" and a.impl == "ours":
        if int(os.environ.get("RANK", "0")) == 0:
            (run_config3 if a.config.startswith("qwen") else run_config5)(a)
        return
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    base = {"metric": "decode tokens/s @4096ctx (Q4_K Llama-3-8B)", "unit": "tokens/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int8 x int4 -> int32 dot, fp32 scales/accumulate",
            "data": "synthetic (seeded random valid Q4_K blocks, random F16 KV cache; no checkpoints offline)",
            "config": {"workload": "Llama-3-8B Q4_K single-token decode at n_past=4096 (BASELINE.json configs[1])", "n_past": N_PAST,
                       "parallelism": "1 GPU" if a.gpus == 1 else f"layer-sharded x{a.gpus} (KV sharded by layer; one CUDA graph per rank: peer_wait kernel -> layers -> peer_send kernel storing the hidden row / next token into the next rank's CUDA-IPC mailbox over NVLink)",
                       "l2": "inputs (4.76 GB per token) exceed L2; no flush needed"}}

    if a.impl == "reference":
        if rank != 0:
            return
        steps = min(a.steps, 8)  # bounded sample: the CPU path runs ~0.1-0.3 s per token
        r = cpu_reference_arm(steps, min(a.warmup, 3))
        out = dict(base)
        out.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"], "steps": steps,
                    "cpu_baseline": {"value": r["value"], "unit": "tokens/s", "cores": r["cores"], "host_cores": r["host_cores"], "kind": "reference",
                                     "sample": r["sample"], "threads_probed_ms_per_token": r["threads_probed_ms_per_token"]},
                    "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        print(json.dumps(out))
        return

    import torch
    import __graft_entry__ as ge
    pkg = ge.load_package()
    pkg.lib()  # raises if the CUDA extension is missing: no fallback
    from chatllm_cpp_b200 import session as S
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (no CPU fallback)")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    if a.tune:
        pkg.lib().b200_gemv_set_tuning(*[int(v) for v in a.tune.split(",")])
    use_mk = a.mk and not a.unfused
    CTX_ROOM = 1024   # the device-side greedy loop walks one position per step: warm-up + timed steps must fit behind n_past = 4096
    if use_mk and a.warmup + a.steps + 2 > CTX_ROOM:
        raise SystemExit(f"--warmup + --steps must stay below {CTX_ROOM - 2}")
    cfg = S.make_config(MODEL, pkg.Q4_K, layers=a.layers or None, max_len=N_PAST + (CTX_ROOM if use_mk else 256))
    from chatllm_cpp_b200 import sharding
    lo, hi = sharding.plan_layers(cfg.layers, world)[rank]
    sess = S.DecodeSession(cfg, seed=0, layer_lo=lo, layer_hi=hi, first=(rank == 0), last=(rank == world - 1), fused=(3 if use_mk else not a.unfused))
    sess.device_argmax = True   # every N: a step is "token in -> next token out" (greedy sampling on the device, b200_argmax)
    sess.mk_advance = True   # persistent kernel only: tok <- argmax(logits), pos <- pos + 1 on the device (greedy decoding, growing KV cache)
    sess.fill_kv_random(N_PAST, seed=rank)
    sess.tok.fill_(12345 % cfg.vocab)

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local) if rank == 0 else None
    if world == 1:
        graph = sess.capture(N_PAST)
        sess.pos.fill_(N_PAST)
        for _ in range(a.warmup):
            graph.replay()
        sync_all()
        if sampler: sampler.start()
        torch.cuda.cudart().cudaProfilerStart()   # `ncu --profile-from-start off` then sees exactly the timed steps (profiles/)
        e0.record()
        for _ in range(a.steps):
            graph.replay()
        e1.record()
        sync_all()
        torch.cuda.cudart().cudaProfilerStop()
    elif not a.no_peer:
        # layer-sharded: every rank replays ONE CUDA graph per token (peer_wait -> its layers -> peer_send); the hidden row / next token move
        # between the ranks through NVLink-mapped CUDA-IPC mailboxes (sharding.PeerRing) — no host, no NCCL on the data path
        ring = sharding.PeerRing(rank, world, cfg.hidden, pkg.lib())
        sess.attach_ring(ring)
        if rank == 0:
            ring.view("tok", "<i4").fill_(12345 % cfg.vocab)
        sess.pos.fill_(N_PAST)
        sync_all()
        graph = sess.capture(N_PAST)           # one eager step (a real token: every rank does it) + capture
        for _ in range(a.warmup):
            graph.replay()
        sync_all()
        if sampler: sampler.start()
        e0.record()
        for _ in range(a.steps):
            graph.replay()
        e1.record()
        sync_all()
        if int(ring.view("status", "<i8").item()) != 0 or (use_mk and sess.mk_status() != 0):
            raise SystemExit(f"rank {rank}: a peer flag / grid barrier timed out")
    else:
        pipe = sharding.Pipeline(rank, world, sess.x, lambda _x: sess.enqueue(N_PAST), tok_buf=sess.tok)
        one_step = pipe.step
        sess.pos.fill_(N_PAST)
        for _ in range(a.warmup):
            one_step()
        sync_all()
        if sampler: sampler.start()
        e0.record()
        for _ in range(a.steps):
            one_step()
        pipe.drain()
        e1.record()
        sync_all()
    ms_total = e0.elapsed_time(e1)
    # the timed region is only ~0.1 s (64 steps): keep the same step running, untimed, until the 0.2 s-period nvidia-smi sampler
    # has seen the clocks / throttle reasons UNDER THIS LOAD a few times (world == 1 only: the ranks of a sharded run stay in lock step)
    if sampler and world == 1:
        t_tail = time.time() + 1.2
        while time.time() < t_tail:
            sess.pos.fill_(N_PAST)   # keep the device-side position inside the allocated context
            for _ in range(32):
                graph.replay()
            torch.cuda.synchronize()
    clocks = sampler.stop() if sampler else None
    if clocks is not None:
        clocks["sampled_over"] = "timed steps + 1.2 s of the same step replayed untimed" if world == 1 else "timed steps"
    if dist:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        n_launch = torch.tensor([sess.launches_per_step], device="cuda")
        dist.all_reduce(n_launch)
        launches = int(n_launch.item()) * a.steps
    else:
        launches = sess.launches_per_step * a.steps
    ms = ms_total / a.steps
    value = 1000.0 / ms

    # ---- roofline of the dominant kernel: exactly the token's quantized-matmul launches, replayed alone (rank 0's shard)
    import ctypes
    L = pkg.lib()
    shapes = []
    for W in sess.layers:
        shapes += [(W.wq, cfg.hidden, cfg.hidden), (W.wk, cfg.hidden, cfg.kv_hidden), (W.wv, cfg.hidden, cfg.kv_hidden), (W.wo, cfg.hidden, cfg.hidden),
                   (W.wgate, cfg.hidden, cfg.ffn), (W.wup, cfg.hidden, cfg.ffn), (W.wdown, cfg.ffn, cfg.hidden)]
    if sess.last:
        shapes.append((sess.lm_head, cfg.hidden, cfg.vocab))
    xq = {k: None for k in (cfg.hidden, cfg.ffn)}
    for k in xq:
        x = torch.randn((1, k), device="cuda")
        q = torch.empty(L.b200_qact_col_bytes(cfg.wtype, k), dtype=torch.uint8, device="cuda")
        L.b200_quantize_act(cfg.wtype, x.data_ptr(), k, k, 1, q.data_ptr(), torch.cuda.current_stream().cuda_stream)
        xq[k] = q
    ybuf = torch.empty(max(cfg.vocab, cfg.ffn), dtype=torch.float32, device="cuda")

    def gemv_all():
        st = torch.cuda.current_stream().cuda_stream
        for (w, k, m) in shapes:
            L.b200_mul_mat_q(cfg.wtype, w.data_ptr(), k, m, xq[k].data_ptr(), 1, ybuf.data_ptr(), m, 0, st)
    gemv_all(); torch.cuda.synchronize()
    gg = torch.cuda.CUDAGraph(); gs = torch.cuda.Stream()
    with torch.cuda.stream(gs):
        with torch.cuda.graph(gg, stream=gs):
            gemv_all()
    for _ in range(3):
        gg.replay()
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(5, min(a.steps, 20))
    g0.record()
    for _ in range(reps):
        gg.replay()
    g1.record(); torch.cuda.synchronize()
    gemv_ms = g0.elapsed_time(g1) / reps
    b, s = S.BLK[cfg.wtype]
    gemv_bytes = sum(m * (k // b) * s + 4 * k + 4 * m for (_, k, m) in shapes)
    pk, pk_kind = peaks()
    achieved = gemv_bytes / gemv_ms / 1e6
    roofline = {"bound": "hbm", "kernel": "gemv_q_kernel<FmtQ4K>", "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / pk["hbm_gbs"], 4), "peak_kind": pk_kind + " (MEASURED_PEAKS.json hbm_gbs)" if pk_kind == "measured" else "fallback",
                "frac_of_nominal_8TBps": round(achieved / 8000.0, 4), "launches": len(shapes), "avg_launch_us": round(gemv_ms * 1e3 / len(shapes), 2),
                "algorithmic_MB_per_token_shard": round(gemv_bytes / 1e6, 1),
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the `ncu --set full` capture of this kernel on the gate shape
                # (profiles/r02i_gemv_q4k_gate_ncu_full_summary.json: 33.074 MB read, 0 written, for 33.03 MB algorithmic), scaled to the mean launch
                "traffic": round(gemv_bytes / len(shapes) * 33.074176 / 33.030144),
                "share_of_step": round(gemv_ms / ms, 3)}
    if use_mk and world == 1:
        # the dominant kernel IS the step: one launch of decode_mk_kernel streams every weight byte of the token plus the KV cache.
        # algorithmic bytes per launch = sum over the quantized matmuls of m*(k/256)*144 + 4k + 4m + the F16 K/V rows read
        n_kv_mean = N_PAST + a.warmup + (a.steps + 1) / 2.0
        mk_bytes = cfg.weight_bytes_per_token() + cfg.kv_bytes_per_token(n_kv_mean)
        mk_ach = mk_bytes / ms / 1e6
        roofline = {"bound": "hbm", "kernel": "decode_mk_kernel<FmtQ4K,128> (persistent, 1 launch per token)", "achieved": round(mk_ach, 1), "peak": pk["hbm_gbs"],
                    "unit": "GB/s", "frac": round(mk_ach / pk["hbm_gbs"], 4),
                    "peak_kind": pk_kind + " (MEASURED_PEAKS.json hbm_gbs)" if pk_kind == "measured" else "fallback",
                    "frac_of_nominal_8TBps": round(mk_ach / 8000.0, 4), "launches": 1, "avg_launch_us": round(ms * 1e3, 2),
                    "algorithmic_MB_per_launch": round(mk_bytes / 1e6, 1), "traffic": None, "share_of_step": 1.0, "mk": getattr(sess, "mk_info", None),
                    "per_op_gemv_kernel_alone": {"achieved": round(achieved, 1), "frac": round(achieved / pk["hbm_gbs"], 4), "launches": len(shapes),
                                                 "note": "round-1 gemv_q_kernel, one launch per matmul, replayed alone (the fallback path for shapes the persistent kernel declines)"}}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return
    out = dict(base)
    tok_bytes = cfg.weight_bytes_per_token() + cfg.kv_bytes_per_token(N_PAST + (a.warmup + (a.steps + 1) / 2.0 if use_mk else 0))
    out.update({"value": round(value, 2), "ms_per_step": round(ms, 4), "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "frac_of_hbm_roofline_whole_token": round(value * tok_bytes / 1e9 / pk["hbm_gbs"], 4),
                "bytes_per_token_MB": round(tok_bytes / 1e6, 1)})
    if a.layers:
        out["config"]["workload"] += f" [DEBUG: {cfg.layers} layers only]"

    # ---- e2e through the drop-in boundary.  The reference host is ONE process that drives every device itself (-ngl "0:16,prolog;1:16,epilog",
    # docs/gpu.md): at N > 1 rank 0 runs it over all N GPUs after the other ranks have left (their memory is released with them).
    if not a.no_e2e and os.path.exists(HARNESS):
        try:
            if use_mk and sess.mk_status() != 0:
                raise SystemExit("persistent kernel reported a grid-barrier timeout")
            del sess, graph, gg
            torch.cuda.empty_cache()
            model = ensure_model_file(a.layers)
            ngl = "all"
            if world > 1:
                from chatllm_cpp_b200 import sharding as _sh
                ngl = _sh.ngl_spec(cfg.layers, world)
                time.sleep(3.0)   # the other ranks are exiting
            r = run_harness(model, ngl, a.warmup + a.steps, 16, ["--skip", str(a.warmup)], real_prefill=True, timeout=600)
            e_ms = r["decode_ms_mean_after_skip"]
            out["e2e"] = {"value": round(1000.0 / e_ms, 2), "unit": "tokens/s", "ms_per_step": round(e_ms, 4),
                          "h2d_bytes_per_step": 4 + 4 * cfg.layers, "d2h_bytes_per_step": 4 * cfg.vocab,
                          "path": "unmodified chatllm host (graph rebuild + ggml sched per token) -> libggml-cuda.so; pageable host buffers of the host app"
                                  + (f"; layers split over {world} devices with -ngl {ngl}" if world > 1 else ""),
                          "plugin_execution": "every one-token graph -> decode plan -> ONE replayed CUDA graph of the per-op kernels, re-parameterised for the next "
                                              "token while the current one runs (csrc/decode_graph.cu; B200_GRAPH=0 = node-by-node launches)",
                          "prefill": {"tokens": r.get("prefill_tokens"), "ms": r.get("prefill_ms"), "note": "real 4096-token prompt through the plugin (batch 512) before the timed decode"}}
            # parity of the benchmarked FILE: the same short real prompt through the plugin and on the reference's CPU backend, logits of the
            # prompt's last token and of 3 decode steps compared (the 4096-token prompt itself would take the CPU minutes)
            try:
                if world > 1:
                    raise RuntimeError("reported at N = 1 only")
                import numpy as np
                dumps = {}
                for tag, ngl, rd in (("gpu", "all", RUNDIR), ("cpu", "0", RUNDIR + "_avx512"), ("cpu_avx2", "0", RUNDIR + "_avx2")):
                    dp = f"/tmp/b200_parity_{tag}.bin"
                    cmd = [HARNESS, "--model", model, "--ggml_dir", rd, "--ngl", ngl, "--threads", "32", "--prefill", "24", "--decode", "3", "--max_length",
                           str(N_PAST + 256), "--dump", dp, "--seed", "3"]
                    try:
                        subprocess.run(cmd, capture_output=True, text=True, timeout=600, check=True)
                        dumps[tag] = np.fromfile(dp, dtype=np.float32).reshape(-1, cfg.vocab)
                    except Exception:  # noqa: BLE001  (no AVX-512 on this host: the cpu arm falls back to the default module directory)
                        if tag == "cpu":
                            subprocess.run(cmd[:4] + [RUNDIR] + cmd[5:], capture_output=True, text=True, timeout=600, check=True)
                            dumps[tag] = np.fromfile(dp, dtype=np.float32).reshape(-1, cfg.vocab)
                relf = lambda a, b: np.abs(a - b).max(axis=1) / np.abs(b).max(axis=1)
                rel = relf(dumps["gpu"], dumps["cpu"])
                out["parity"] = {"max_rel": float(rel.max()), "per_eval_rel": [round(float(v), 5) for v in rel],
                                 "argmax_equal": bool((dumps["gpu"].argmax(1) == dumps["cpu"].argmax(1)).all()),
                                 "reference_self_spread_avx2_vs_avx512": (round(float(relf(dumps["cpu_avx2"], dumps["cpu"]).max()), 5) if "cpu_avx2" in dumps else None),
                                 "what": "logits of a 24-token prompt + 3 decode steps of the FULL 32-layer synthetic file: plugin vs the reference CPU backend, next to the "
                                         "reference's own AVX2-vs-AVX-512 spread on the same inputs.  A random-weight 32-layer network amplifies a 1e-7 difference to ~1e-1 "
                                         "(the reference disagrees with itself that much), so the 1e-3 bar is enforced per layer at these shapes (tests/test_decode_mk.py, "
                                         "teacher-forced vs the oracle: 4e-4 / 5e-7 of the layer's update) and on shallow models (tests/test_e2e_host.py)"}
            except Exception as ex:  # noqa: BLE001
                out["parity"] = {"max_rel": None, "error": str(ex)[-200:]}
        except Exception as ex:  # noqa: BLE001
            out["e2e"] = {"value": None, "error": str(ex)[-300:]}
    if not a.no_cpu and world == 1 and os.path.exists(HARNESS):
        try:
            r = cpu_reference_arm(4, 1, probe=(8, 16, 32, 64))
            out["cpu_baseline"] = {"value": round(r["value"], 3), "unit": "tokens/s", "cores": r["cores"], "host_cores": r["host_cores"], "kind": "reference",
                                   "sample": r["sample"], "threads_probed_ms_per_token": r["threads_probed_ms_per_token"]}
        except Exception as ex:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "error": str(ex)[-300:]}
    print(json.dumps(out))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
