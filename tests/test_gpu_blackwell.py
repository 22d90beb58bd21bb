"""Blackwell-specific kernels (-m gpu), both measured in round 2 and on by default since:
  * the tcgen05 prompt matmul (csrc/prefill_tc.cu: tcgen05.mma kind::i8, accumulators in TMEM) against the mma.sync kernel it replaces;
  * the thread-block-cluster V.P of the decode attention (DSMEM reduction + in-cluster quantization, attention = 2 launches) against the
    split V.P + tail launch (B200_ATTN_CLUSTER=0)."""
import ctypes as C
import os

import numpy as np
import pytest

import qformats as qf

pytestmark = pytest.mark.gpu


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _setup():
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from chatllm_cpp_b200 import kernels as K
    return pkg, K


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
    """the default thread-block-cluster V.P (DSMEM reduction and in-cluster quantization, no tail launch) against the three-launch attention
    (B200_ATTN_CLUSTER=0, run in a child process: the flag is read once per process): same P, same products; only the order in which the
    position slabs are summed differs (512- instead of 256-position slabs) -> outputs within 1e-6 of the output scale, quantized codes
    equal except where a value sits on a rounding tie (measured r02: <= 0.26 % of the codes at 8192 positions)."""
    import subprocess, sys, json, tempfile, torch
    pkg, K = _setup()
    code = f"""
import os, sys, numpy as np, torch
os.environ['B200_ATTN_CLUSTER'] = '0'
sys.path.insert(0, {os.path.dirname(os.path.abspath(__file__))!r}); sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})
import test_gpu_blackwell as T
np.save(sys.argv[1], T.run_attn_quant({wtype}, {heads}, {kvh}, {hd}, {n_kv}))
"""
    with tempfile.TemporaryDirectory() as d:
        ref_path = os.path.join(d, "ref.npy")
        env = dict(os.environ); env["B200_ATTN_CLUSTER"] = "0"
        subprocess.run([sys.executable, "-c", code, ref_path], check=True, env=env, timeout=120)
        ref = np.load(ref_path)
    got = run_attn_quant(wtype, heads, kvh, hd, n_kv)
    k = heads * hd
    assert np.abs(got[:k] - ref[:k]).max() <= 1e-6 * np.abs(ref[:k]).max()
    assert (got[k:] != ref[k:]).mean() <= 4e-3   # quantized codes / scales / sums (as floats): ties only


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
