"""N>1 host logic on CPU: world_size-2 gloo run of the layer-sharded pipeline (plan + hidden-row hand-off)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, n_layers, steps, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import __graft_entry__ as ge
    ge.load_package()
    from chatllm_cpp_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = sharding.plan_layers(n_layers, world)[rank]
    x = torch.zeros(8)
    tok = torch.zeros(1, dtype=torch.int32)

    def run(buf):
        if rank == 0:
            buf.fill_(1.0)                      # "embedding"
        for layer in range(lo, hi):
            buf.mul_(1.5).add_(float(layer))    # stand-in for a decoder layer: order-sensitive
    order = []

    def run_logged(buf):
        if rank == 0:
            order.append(int(tok.item()))     # the token fed back by the last rank (0 on the first step)
        run(buf)
        if rank == world - 1:
            tok.fill_(pipe.steps + 1)          # "sampled" token of this step
    pipe = sharding.Pipeline(rank, world, x, run_logged, tok_buf=tok)
    for _ in range(steps):
        pipe.step()
    pipe.drain()
    if rank == 0:
        assert order == list(range(steps)), order   # step i+1 started only after step i's token arrived
    if rank == world - 1:
        out.put(x.clone())
    dist.barrier()
    dist.destroy_process_group()


def test_plan_layers_balanced_and_contiguous():
    import __graft_entry__ as ge
    ge.load_package()
    from chatllm_cpp_b200 import sharding
    for n, w in ((32, 1), (32, 2), (32, 8), (22, 4), (28, 8), (3, 2)):
        p = sharding.plan_layers(n, w)
        assert p[0][0] == 0 and p[-1][1] == n and all(a[1] == b[0] for a, b in zip(p, p[1:]))
        sizes = [hi - lo for lo, hi in p]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_pipeline_matches_sequential():
    n_layers, steps, world = 7, 3, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_layers, steps, out)) for r in range(world)]
    [p.start() for p in procs]
    got = out.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ref = torch.ones(8)
    for layer in range(n_layers):
        ref.mul_(1.5).add_(float(layer))
    assert torch.equal(got, ref)


def test_ring_io_flags_are_consistent():
    """PeerRing address arithmetic (no GPU): whatever flag / buffer rank r sends to is exactly what rank (r + 1) % world waits on / reads."""
    import __graft_entry__ as ge
    ge.load_package()
    from chatllm_cpp_b200 import sharding
    hidden = 4096
    L = sharding.mailbox_layout(hidden)
    assert L["x_flag"] % 8 == 0 and L["tok_flag"] % 8 == 0 and L["tok"] % 4 == 0 and L["bytes"] >= L["tok"] + 4
    assert sharding.ring_io(0, 1, hidden, 1 << 20, 0) == {}
    for world in (2, 4, 8):
        bases = [(r + 1) << 24 for r in range(world)]
        ios = [sharding.ring_io(r, world, hidden, bases[r], bases[(r + 1) % world]) for r in range(world)]
        for r in range(world):
            nxt = (r + 1) % world
            assert ios[r]["send_flag"] == ios[nxt]["wait_flag"], (world, r)
            if r < world - 1:
                assert ios[r]["send_x"] == bases[nxt] + L["x"] and ios[r]["send_tok"] == 0
            else:
                assert ios[r]["send_tok"] == bases[0] + L["tok"] and ios[r]["send_x"] == 0
        # rank 0 starts step t when the token of step t-1 has arrived (offset 0); rank r > 0 needs the hidden row of step t itself (offset 1)
        assert [io["wait_offset"] for io in ios] == [0] + [1] * (world - 1)


def test_ngl_spec_places_layers_like_the_rank_plan():
    """bench.py's e2e leg at N > 1 hands the ONE host process the same layer placement the ranks of the device-resident arm use"""
    import re
    import __graft_entry__ as ge
    ge.load_package()
    from chatllm_cpp_b200 import sharding
    assert sharding.ngl_spec(32, 1) == "all"
    assert sharding.ngl_spec(32, 2) == "0:16,prolog;1:16,epilog"
    assert sharding.ngl_spec(32, 4) == "0:8,prolog;1:8;2:8;3:8,epilog"
    for n, w in [(32, 8), (28, 8), (22, 4), (2, 2), (3, 8)]:
        spec = sharding.ngl_spec(n, w)
        plan = sharding.plan_layers(n, w)
        seen = 0
        for part in spec.split(";"):
            m = re.fullmatch(r"(\d+):(.*)", part)
            assert m, spec
            d, items = int(m.group(1)), m.group(2).split(",")
            cnt = sum(int(i) for i in items if i.isdigit())
            assert cnt == plan[d][1] - plan[d][0], (spec, plan)
            assert ("prolog" in items) == (d == 0) and ("epilog" in items) == (d == w - 1), spec
            seen += cnt
        assert seen == n, spec
