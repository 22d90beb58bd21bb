"""Layer-sharded decode over NVLink mailboxes (-m gpu, needs >= 2 GPUs): two processes, one per GPU, each running its layer range with the
persistent kernel; the hidden row and the next token travel through CUDA-IPC mapped peer memory (sharding.PeerRing, b200_decode_io
wait_flag / send_x / send_tok).  Tokens and logits must equal the single-GPU run of the same model bit for bit."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _n_cuda():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:  # noqa: BLE001
        return 0


pytestmark = [pytest.mark.gpu, pytest.mark.skipif(_n_cuda() < 2, reason="needs two GPUs")]

N_LAYERS, STEPS, N_PAST = 4, 6, 200


def _weights_fn(torch, S, cfg):
    def weights(i, name, m, k):
        gen = torch.Generator(device="cuda")
        gen.manual_seed(1000 * (i + 2) + sum(map(ord, name)))
        if k == 0:
            return 1 + 0.1 * torch.randn(m, device="cuda", generator=gen)
        return S.synth_weights_device(cfg.wtype, m, k, gen)
    return weights


def _fill_kv(torch, sess, lo):
    for j, W in enumerate(sess.layers):
        gen = torch.Generator(device="cuda"); gen.manual_seed(77 + lo + j)
        W.kc[:N_PAST] = torch.randn((N_PAST, sess.cfg.kv_hidden), device="cuda", generator=gen).to(torch.float16)
        W.vc[:, :N_PAST] = torch.randn((sess.cfg.kv_hidden, N_PAST), device="cuda", generator=gen).to(torch.float16)


def _worker(rank, world, port, out, fused):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    from chatllm_cpp_b200 import session as S, sharding
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = S.Config(pkg.Q4_K, 2048, 1024, 8, 2, N_LAYERS, 2816, max_len=N_PAST + 64)
    ref_toks, ref_logits = None, None
    if rank == 0:   # the single-GPU answer
        full = S.DecodeSession(cfg, weights=_weights_fn(torch, S, cfg), fused=fused)
        _fill_kv(torch, full, 0)
        ref_toks, ref_logits = [], []
        if fused == 3:
            full.mk_advance = True
            full.tok.fill_(7); full.pos.fill_(N_PAST)
            for _ in range(STEPS):
                full.enqueue(0); torch.cuda.synchronize()
                ref_toks.append(int(full.next_tok.item())); ref_logits.append(full.logits.cpu().clone())
        else:
            tok = 7
            for i in range(STEPS):
                lg = full.step(tok, N_PAST + i)
                tok = int(lg.argmax().item())
                ref_toks.append(tok); ref_logits.append(lg.cpu().clone())
        del full
    lo, hi = sharding.plan_layers(N_LAYERS, world)[rank]
    sess = S.DecodeSession(cfg, weights=_weights_fn(torch, S, cfg), layer_lo=lo, layer_hi=hi, first=(rank == 0), last=(rank == world - 1), fused=fused)
    _fill_kv(torch, sess, lo)
    ring = sharding.PeerRing(rank, world, cfg.hidden, pkg.lib())
    sess.attach_ring(ring)
    if rank == 0:
        ring.view("tok", "<i4").fill_(7)
    sess.pos.fill_(N_PAST)
    torch.cuda.synchronize(); dist.barrier()
    toks, logits = [], []
    for i in range(STEPS):
        if fused != 3:
            sess.pos.fill_(N_PAST + i)      # the per-op step takes the position from the host (it sizes the attention grids)
        sess.enqueue(N_PAST + i)
        if rank == world - 1:
            torch.cuda.synchronize()
            toks.append(int(sess.next_tok.item())); logits.append(sess.logits.cpu().clone())
    torch.cuda.synchronize()
    assert sess.mk_status() == 0 and int(ring.view("status", "<i8").item()) == 0
    dist.barrier()
    if rank == 0:
        out.put(("ref", ref_toks, [l.numpy() for l in ref_logits]))
        assert int(ring.view("tok", "<i4").item()) == ref_toks[-1]   # the last rank's final token arrived in rank 0's mailbox
    if rank == world - 1:
        out.put(("ring", toks, [l.numpy() for l in logits]))
    dist.barrier()
    ring.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("fused", [True, 3])
def test_two_gpu_peer_ring_equals_single_gpu(fused):
    import numpy as np
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, out, fused)) for r in range(world)]
    [p.start() for p in procs]
    got = dict()
    for _ in range(2):
        tag, toks, logits = out.get(timeout=300)
        got[tag] = (toks, logits)
    [p.join(timeout=120) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert got["ring"][0] == got["ref"][0], (got["ring"][0], got["ref"][0])
    for a, b in zip(got["ring"][1], got["ref"][1]):
        assert np.array_equal(a, b)      # the same kernels on both sides, the hidden row crosses NVLink unchanged: bit-identical
