"""OPT-IN tests (-m gpu AND B200_TEST_FUSED2=1) for the round-2 fusion written at the end of round 1 and not yet run on a GPU:
RMSNorm + activation quantization in the consumer GEMV's prologue, residual add in the producer GEMV's epilogue
(chatllm.cpp_b200/csrc/normquant.cuh, b200_gemv_fused, DecodeSession(fused=2): 9 launches per layer instead of 11).
They stay out of the default suite until they have passed once on a B200:
    B200_TEST_FUSED2=1 python -m pytest tests/test_fused2_optin.py -m gpu            # kernels + session
    B200_FUSE2=1 python -m pytest tests/test_plugin_ops.py tests/test_e2e_host.py -m gpu   # the same fusion inside the plugin (graph_compute)
then flip the defaults (DecodeSession(fused=2) in bench.py, B200_FUSE2 default on) and move these into the regular files."""
import ctypes as C
import os

import numpy as np
import pytest

import qformats as qf

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("B200_TEST_FUSED2"), reason="opt-in: set B200_TEST_FUSED2=1")]


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _setup():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from chatllm_cpp_b200 import kernels as K
    return pkg, K


arr = lambda ct, v: (ct * len(v))(*v)


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("k,ms", [(4096, [4096, 1024, 1024]), (1024, [512, 128, 128]), (14336, [4096]), (256, [64])])
def test_norm_prologue_concat_vs_separate_kernels(wtype, k, ms):
    """q/k/v-style launch: W_i . Q(rms_norm(x) * w) computed with the in-kernel prologue vs rms_norm -> quantize -> multi GEMV"""
    import torch
    pkg, K = _setup()
    rng = np.random.default_rng(k + wtype)
    ws = [qf.random_blocks(wtype, m, k, rng=rng) for m in ms]
    wd = [K.upload_weights(wtype, w, k, m) for w, m in zip(ws, ms)]
    x = rng.standard_normal((1, k)).astype(np.float32) * 3; x[0, rng.integers(0, k, 4)] *= 20
    nw = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    xd, nwd = _t(x), _t(nw)
    ys = [torch.zeros((1, m), device="cuda") for m in ms]
    L = pkg.lib()
    rc = L.b200_gemv_fused(wtype, 0, len(ms), arr(C.c_void_p, [w.data_ptr() for w in wd]), arr(C.c_int64, ms), arr(C.c_void_p, [y.data_ptr() for y in ys]),
                           arr(C.c_int64, ms), 0, 0, k, 0, xd.data_ptr(), nwd.data_ptr(), 1e-5, 0)
    assert rc == 0
    xn = K.rms_norm(xd, nwd, 1e-5)
    for w, m, y in zip(wd, ms, ys):
        ref = K.mul_mat(wtype, w, k, m, xn).cpu().numpy()
        assert np.abs(y.cpu().numpy() - ref).max() <= 3e-5 * np.abs(ref).max()
    # and against the oracle end to end
    xn_ref = np.zeros_like(x); qf.port().oq_rms_norm(x.ctypes.data, nw.ctypes.data, xn_ref.ctypes.data, k, 1, 1e-5)
    ref0 = qf.port_mul_mat(wtype, ws[0], k, ms[0], xn_ref)
    assert np.abs(ys[0].cpu().numpy() - ref0).max() <= 1e-3 * np.abs(ref0).max()   # a flipped code is possible: north-star bound


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0])
def test_norm_prologue_paired_swiglu(wtype):
    import torch
    pkg, K = _setup()
    rng = np.random.default_rng(77 + wtype)
    k, m = 4096, 14336
    pool = qf.random_blocks(wtype, 1024, k, rng=rng)
    wg, wu = pool[rng.integers(0, 1024, m)], pool[rng.integers(0, 1024, m)]
    dg, du = K.upload_weights(wtype, wg, k, m), K.upload_weights(wtype, wu, k, m)
    x = rng.standard_normal((1, k)).astype(np.float32); nw = (1 + 0.1 * rng.standard_normal(k)).astype(np.float32)
    xd, nwd = _t(x), _t(nw)
    y = torch.zeros((1, m), device="cuda")
    L = pkg.lib()
    rc = L.b200_gemv_fused(wtype, 1, 2, arr(C.c_void_p, [dg.data_ptr(), du.data_ptr()]), arr(C.c_int64, [m, m]), arr(C.c_void_p, [y.data_ptr(), 0]),
                           arr(C.c_int64, [m, m]), 0, 0, k, 0, xd.data_ptr(), nwd.data_ptr(), 1e-5, 0)
    assert rc == 0
    xn = K.rms_norm(xd, nwd, 1e-5)
    ref = K.silu_mul(K.mul_mat(wtype, dg, k, m, xn), K.mul_mat(wtype, du, k, m, xn)).cpu().numpy()
    assert np.abs(y.cpu().numpy() - ref).max() <= 3e-5 * np.abs(ref).max()


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q8_0])
@pytest.mark.parametrize("k,m", [(4096, 4096), (14336, 4096), (512, 6)])
def test_residual_epilogue_in_place(wtype, k, m):
    """x += W . act  is bit-identical to the plain GEMV followed by an fp32 add (same operation order)"""
    import torch
    pkg, K = _setup()
    rng = np.random.default_rng(k + m + wtype)
    w = qf.random_blocks(wtype, m, k, rng=rng); wd = K.upload_weights(wtype, w, k, m)
    act = _t(rng.standard_normal((1, k)).astype(np.float32)); q = K.quantize_act(wtype, act)
    res = _t(rng.standard_normal((1, m)).astype(np.float32))
    plain = K.mul_mat_q(wtype, wd, k, m, q, 1)
    x = res.clone()
    xp = arr(C.c_void_p, [x.data_ptr()])
    rc = pkg.lib().b200_gemv_fused(wtype, 0, 1, arr(C.c_void_p, [wd.data_ptr()]), arr(C.c_int64, [m]), xp, arr(C.c_int64, [m]), 0, xp, k, q.data_ptr(), 0, 0, 0.0, 0)
    assert rc == 0
    assert torch.equal(x, plain + res)


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0])
def test_session_fused2_matches_fused1_and_graph_replay(wtype):
    import torch
    import __graft_entry__ as ge
    ge.load_package()
    from chatllm_cpp_b200 import session as S
    cfg = S.Config(wtype, 1024, 512, 8, 2, 3, 1024, max_len=128)
    a, b = S.DecodeSession(cfg, seed=3, fused=True), S.DecodeSession(cfg, seed=3, fused=2)
    for s in (a, b):
        s.fill_kv_random(100, seed=1)
    la, lb = a.step(7, 100).clone(), b.step(7, 100).clone()
    rel = float((la - lb).abs().max() / la.abs().max())
    assert rel <= 2e-2 and np.isfinite(lb.cpu().numpy()).all(), rel   # 1e-7 when no activation code flips (sum-of-squares order differs)
    print("fused2 vs fused1 rel", rel)
    assert b.launches_per_step == 1 + 3 * 9 + 1
    g = b.capture(100)
    b.logits.zero_()
    g.replay(); torch.cuda.synchronize()
    assert torch.equal(lb, b.logits)


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0, qf.Q8_0])
@pytest.mark.parametrize("k,m,n", [(256, 128, 128), (1024, 130, 64), (4096, 300, 200), (4096, 4096, 512), (14336, 512, 130)])
def test_tcgen05_prefill_equals_mma_sync(wtype, k, m, n):
    """csrc/prefill_tc.cu (tcgen05.mma kind::i8, three exact int8 planes) must reproduce mmq_kernel (mma.sync): the int32 super-block
    sums are the same integers and the fp32 rescale is the same expression in the same order, so the two agree to fp32 rounding of
    that expression (nvcc may contract it into FMAs differently in the two kernels: a few ulp, 2e-6 allowed; bit-equality is reported).
    Run under a timeout: a wrong descriptor can hang the MMA barrier."""
    import torch
    pkg, K = _setup()
    rng = np.random.default_rng(k + m + n)
    pool = qf.random_blocks(wtype, min(m, 512), k, rng=rng)
    w = pool[rng.integers(0, pool.shape[0], m)]
    wd = K.upload_weights(wtype, w, k, m)
    x = rng.standard_normal((n, k)).astype(np.float32); x[rng.random((n, k)) < 1e-3] *= 20
    xd = _t(x)
    L = pkg.lib()
    pq = torch.empty(L.b200_pact_col_bytes(wtype, k) * n, dtype=torch.uint8, device="cuda")
    assert L.b200_quantize_plain(wtype, xd.data_ptr(), k, k, n, pq.data_ptr(), 0) == 0
    bias = _t(rng.standard_normal(m).astype(np.float32))
    y0 = torch.zeros((n, m), device="cuda"); y1 = torch.full((n, m), float("nan"), device="cuda")
    assert L.b200_mul_mat_q_batched(wtype, wd.data_ptr(), k, m, pq.data_ptr(), n, y0.data_ptr(), m, bias.data_ptr(), 0) == 0
    assert L.b200_mul_mat_q_batched_tc(wtype, wd.data_ptr(), k, m, pq.data_ptr(), n, y1.data_ptr(), m, bias.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    print("tcgen05 vs mma.sync bit-identical:", bool(torch.equal(y0, y1)), "max abs diff", float((y0 - y1).abs().max()))
    assert torch.isfinite(y1).all()
    assert float((y0 - y1).abs().max()) <= 2e-6 * float(y0.abs().max())


@pytest.mark.parametrize("wtype", [qf.Q4_K, qf.Q4_0])
@pytest.mark.parametrize("heads,kvh,hd,n_kv", [(32, 8, 128, 4097), (32, 8, 128, 300), (32, 8, 128, 8192), (32, 4, 64, 777), (8, 2, 64, 1500)])
def test_cluster_pv_equals_split_pv(wtype, heads, kvh, hd, n_kv):
    """B200_ATTN_CLUSTER=1 (thread-block-cluster V.P with DSMEM reduction and in-cluster quantization, no tail launch) against the default
    three-launch attention: same P, same products; only the order in which the position slabs are summed differs (512- instead of
    256-position slabs) -> outputs within 1e-6 of the output scale, quantized codes equal except where a value sits on a rounding tie.
    The flag is read once per process: run as  B200_ATTN_CLUSTER=1 B200_TEST_FUSED2=1 pytest -k cluster_pv  (the reference result is
    produced by a child process without the flag)."""
    import subprocess, sys, json, tempfile, torch
    if not os.environ.get("B200_ATTN_CLUSTER"):
        pytest.skip("set B200_ATTN_CLUSTER=1 for this process")
    pkg, K = _setup()
    code = f"""
import os, sys, numpy as np, torch
os.environ.pop('B200_ATTN_CLUSTER', None)
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import test_fused2_optin as T
np.save(sys.argv[1], T.run_attn_quant({wtype}, {heads}, {kvh}, {hd}, {n_kv}))
"""
    with tempfile.TemporaryDirectory() as d:
        ref_path = os.path.join(d, "ref.npy")
        env = {k: v for k, v in os.environ.items() if k != "B200_ATTN_CLUSTER"}
        subprocess.run([sys.executable, "-c", code, ref_path], check=True, env=env, timeout=120)
        ref = np.load(ref_path)
    got = run_attn_quant(wtype, heads, kvh, hd, n_kv)
    k = heads * hd
    assert np.abs(got[:k] - ref[:k]).max() <= 1e-6 * np.abs(ref[:k]).max()
    assert (got[k:] != ref[k:]).mean() <= 2e-3   # quantized codes / scales / sums (as floats): ties only


def run_attn_quant(wtype, heads, kvh, hd, n_kv):
    """attention output (k floats) followed by the decoded qact (codes, scales, block sums) as one float array"""
    import torch
    import test_gpu_kernels as TG
    pkg, K = _setup()
    rng = np.random.default_rng(n_kv + heads + wtype)
    max_len = ((n_kv + 255) // 256) * 256 + 256
    kv_hidden = kvh * hd
    kd = _t(rng.standard_normal((max_len, kv_hidden)).astype(np.float16))
    vd = _t(rng.standard_normal((kv_hidden, max_len)).astype(np.float16))
    qd = _t(rng.standard_normal((heads, hd)).astype(np.float32) * 2)
    L = pkg.lib()
    scratch = torch.empty(L.b200_attn_decode_scratch_bytes(heads, max_len) // 4 + 16, dtype=torch.float32, device="cuda")
    out = torch.zeros((1, heads * hd), device="cuda")
    q = torch.zeros(L.b200_qact_col_bytes(wtype, heads * hd), dtype=torch.uint8, device="cuda")
    assert L.b200_attn_decode_quant(qd.data_ptr(), kd.data_ptr(), vd.data_ptr(), out.data_ptr(), scratch.data_ptr(), heads, kvh, hd, n_kv, kv_hidden,
                                    max_len, 1.0 / np.sqrt(hd), wtype, q.data_ptr(), 0) == 0
    torch.cuda.synchronize()
    qs, d, bs = TG._decode_qact(q.cpu().numpy()[None, :], wtype, heads * hd)
    return np.concatenate([out.cpu().numpy().reshape(-1), qs.astype(np.float32).reshape(-1), d.reshape(-1), bs.astype(np.float32).reshape(-1)])
