"""Device-resident decode step of the Llama family, built only from the C-ABI kernels (include/chatllm_b200.h).

This is the host-side mirror of what the reference's graph does for one token
(HeterogeneousModel::forward src/models.cpp:1399-1424 -> LMBlock1Forward::forward src/layers.cpp:2719-2761 ->
LMFinalSteps::forward src/models.cpp:1736-1785), with the same tensor layouts (F16 K cache row-per-position, V cache
transposed) so that results can be checked against the oracle step for step.  It exists for two reasons:
  * `value` in bench.py: the whole decode step with every input resident in HBM, replayed as ONE CUDA graph
    (the reference rebuilds and re-schedules a ~1000-node ggml graph per token on the host);
  * layer-sharded multi-GPU decode (BASELINE.json configs[3]): one process per GPU, each owning a contiguous layer range
    and its KV-cache shard, the single hidden-state row handed over with NCCL send/recv.
torch is used for device memory, streams, CUDA-graph capture and torch.distributed only.
"""
import math

import numpy as np
import torch

from . import F16, F32, Q4_0, Q4_K, Q8_0, DecodeIO, DecodeLayer, DecodeModel, lib
from . import kernels as K

BLK = {Q4_0: (32, 18), Q8_0: (32, 34), Q4_K: (256, 144)}


def synth_weights_device(wtype, m, k, gen, scale=None):
    """Random *valid* quant blocks generated directly on the GPU, already in the device layout (native for Q4_K,
    per-row SoA for Q4_0/Q8_0).  Statistics as tests/qformats.random_blocks (zero-mean, std ~0.02)."""
    dev = "cuda"
    nb = k // BLK[wtype][0]
    if wtype == Q4_K:
        w = torch.randint(0, 256, (m, nb, 144), dtype=torch.uint8, device=dev, generator=gen)
        d = (torch.rand((m, nb), device=dev, generator=gen) + 0.5) * 1.2e-4
        w[:, :, 0:2] = d.to(torch.float16).view(torch.uint8).reshape(m, nb, 2)
        w[:, :, 2:4] = (d * 7.5).to(torch.float16).view(torch.uint8).reshape(m, nb, 2)
        # mins = scales (zero-mean sub-blocks): bytes 4..7 := bytes 0..3 of the 12 scale bytes, high nibbles mirrored
        sc = w[:, :, 4:8].clone()
        w[:, :, 8:12] = sc
        lo = w[:, :, 12:16] & 0x0F
        w[:, :, 12:16] = lo | (lo << 4)
        return w.reshape(m, nb * 144)
    if wtype == Q4_0:
        qs = torch.randint(0, 256, (m, nb * 16), dtype=torch.uint8, device=dev, generator=gen)
        d = ((torch.rand((m, nb), device=dev, generator=gen) + 0.5) * 4.3e-3).to(torch.float16).view(torch.uint8).reshape(m, nb * 2)
        return torch.cat([qs, d], dim=1).contiguous()
    if wtype == Q8_0:
        qs = torch.randint(0, 256, (m, nb * 32), dtype=torch.uint8, device=dev, generator=gen)
        d = ((torch.rand((m, nb), device=dev, generator=gen) + 0.5) * 2.7e-4).to(torch.float16).view(torch.uint8).reshape(m, nb * 2)
        return torch.cat([qs, d], dim=1).contiguous()
    raise ValueError(wtype)


class Config:
    def __init__(self, wtype, vocab, hidden, heads, kv_heads, layers, ffn, rope_theta=10000.0, rope_mode=0, eps=1e-5, max_len=4352,
                 bias=False, name="custom"):
        self.wtype, self.vocab, self.hidden, self.heads, self.kv_heads = wtype, vocab, hidden, heads, kv_heads
        self.layers, self.ffn, self.rope_theta, self.rope_mode, self.eps = layers, ffn, rope_theta, rope_mode, eps
        self.max_len, self.bias, self.name = max_len, bias, name
        self.head_dim = hidden // heads
        self.kv_hidden = kv_heads * self.head_dim

    def gemv_shapes(self):
        """[(k, m)] of every quantized matmul of one token, in execution order (per layer, then lm_head)."""
        per_layer = [(self.hidden, self.hidden), (self.hidden, self.kv_hidden), (self.hidden, self.kv_hidden), (self.hidden, self.hidden),
                     (self.hidden, self.ffn), (self.hidden, self.ffn), (self.ffn, self.hidden)]
        return per_layer * self.layers + [(self.hidden, self.vocab)]

    def weight_bytes_per_token(self):
        b, s = BLK[self.wtype]
        return sum(m * (k // b) * s + 4 * k + 4 * m for (k, m) in self.gemv_shapes())

    def kv_bytes_per_token(self, n_kv):
        return 2 * self.kv_hidden * 2 * n_kv * self.layers


CONFIGS = {
    "llama3-8b": dict(vocab=128256, hidden=4096, heads=32, kv_heads=8, layers=32, ffn=14336, rope_theta=500000.0, rope_mode=0),
    "tinyllama-1.1b": dict(vocab=32000, hidden=2048, heads=32, kv_heads=4, layers=22, ffn=5632, rope_theta=10000.0, rope_mode=0),
    "qwen2.5-7b": dict(vocab=152064, hidden=3584, heads=28, kv_heads=4, layers=28, ffn=18944, rope_theta=1000000.0, rope_mode=2, bias=True),
}


class LayerWeights:
    pass


class DecodeSession:
    """Owns layers [layer_lo, layer_hi) of the model (all of them on one GPU), their KV-cache shard and the scratch."""

    def __init__(self, cfg, weights=None, seed=0, layer_lo=0, layer_hi=None, first=True, last=True, fused=True):
        self.cfg = cfg
        self.fused = fused
        self.lo, self.hi = layer_lo, cfg.layers if layer_hi is None else layer_hi
        self.first, self.last = first, last
        c = cfg
        gen = torch.Generator(device="cuda"); gen.manual_seed(seed + 1000 * layer_lo)
        f32 = lambda *s: torch.empty(s, dtype=torch.float32, device="cuda")
        self.layers = []
        for i in range(self.lo, self.hi):
            L = LayerWeights()
            get = (lambda name, m, k: weights(i, name, m, k)) if weights else (lambda name, m, k: synth_weights_device(c.wtype, m, k, gen))
            L.wq, L.wk, L.wv = get("q", c.hidden, c.hidden), get("k", c.kv_hidden, c.hidden), get("v", c.kv_hidden, c.hidden)
            L.wo = get("o", c.hidden, c.hidden)
            L.wgate, L.wup, L.wdown = get("gate", c.ffn, c.hidden), get("up", c.ffn, c.hidden), get("down", c.hidden, c.ffn)
            nrm = (lambda name, n: weights(i, name, n, 0)) if weights else (lambda name, n: 1 + 0.1 * torch.randn(n, device="cuda", generator=gen))
            L.attn_norm, L.ffn_norm = nrm("attn_norm", c.hidden).float(), nrm("ffn_norm", c.hidden).float()
            if c.bias:
                bs = (lambda name, n: weights(i, name, n, 0)) if weights else (lambda name, n: 0.02 * torch.randn(n, device="cuda", generator=gen))
                L.bq, L.bk, L.bv = bs("bq", c.hidden).float(), bs("bk", c.kv_hidden).float(), bs("bv", c.kv_hidden).float()
            else:
                L.bq = L.bk = L.bv = None
            L.kc = torch.zeros((c.max_len, c.kv_hidden), dtype=torch.float16, device="cuda")
            L.vc = torch.zeros((c.kv_hidden, c.max_len), dtype=torch.float16, device="cuda")
            self.layers.append(L)
        if first:
            self.embed = weights(-1, "embed", c.vocab, c.hidden) if weights else synth_weights_device(c.wtype, c.vocab, c.hidden, gen)
        if last:
            self.final_norm = (weights(-1, "final_norm", c.hidden, 0) if weights else 1 + 0.1 * torch.randn(c.hidden, device="cuda", generator=gen)).float()
            self.lm_head = weights(-1, "lm_head", c.vocab, c.hidden) if weights else synth_weights_device(c.wtype, c.vocab, c.hidden, gen)
            self.logits = f32(1, c.vocab)
        # activations / scratch (fixed addresses -> graph-capturable)
        self.tok = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.pos = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.x = f32(1, c.hidden); self.xn = f32(1, c.hidden)
        self.q = f32(1, c.hidden); self.k = f32(1, c.kv_hidden); self.v = f32(1, c.kv_hidden)
        self.att = f32(1, c.hidden); self.o = f32(1, c.hidden)
        self.gate = f32(1, c.ffn); self.up = f32(1, c.ffn)
        self.scratch = torch.empty(lib().b200_attn_decode_scratch_bytes(c.heads, c.max_len) // 4, dtype=torch.float32, device="cuda")
        cbmax = max(lib().b200_qact_col_bytes(c.wtype, c.hidden), lib().b200_qact_col_bytes(c.wtype, c.ffn))
        self.qact = torch.empty(cbmax, dtype=torch.uint8, device="cuda")
        self.next_tok = torch.zeros(1, dtype=torch.int32, device="cuda")
        self.launches_per_step = 0
        self._graph = None
        self._plan = None

    # ---- one token ------------------------------------------------------------------------------------------------
    def _mm(self, w, k, m, x, out, bias=None):
        c = self.cfg
        q = self.qact
        L = lib()
        st = torch.cuda.current_stream().cuda_stream
        rc = L.b200_quantize_act(c.wtype, x.data_ptr(), k, k, 1, q.data_ptr(), st)
        rc |= L.b200_mul_mat_q(c.wtype, w.data_ptr(), k, m, q.data_ptr(), 1, out.data_ptr(), m, 0 if bias is None else bias.data_ptr(), st)
        if rc:
            raise RuntimeError(f"mul_mat failed rc={rc}")
        self._n += 2

    def enqueue_step(self, n_past):
        """Enqueue the kernels of one decode step for the token in self.tok at position n_past (host int: it sizes the
        attention grids).  self.x holds the incoming hidden state when this shard is not the first."""
        c = self.cfg
        L = lib()
        st = torch.cuda.current_stream().cuda_stream
        self._n = 0
        hd, n_kv = c.head_dim, n_past + 1
        if self.first:
            L.b200_get_rows(c.wtype, self.embed.data_ptr(), c.hidden, self.tok.data_ptr(), 1, self.x.data_ptr(), st); self._n += 1
        for W in self.layers:
            L.b200_rms_norm(self.x.data_ptr(), W.attn_norm.data_ptr(), self.xn.data_ptr(), c.hidden, 1, c.eps, st); self._n += 1
            self._mm(W.wq, c.hidden, c.hidden, self.xn, self.q, W.bq)
            self._mm(W.wk, c.hidden, c.kv_hidden, self.xn, self.k, W.bk)
            self._mm(W.wv, c.hidden, c.kv_hidden, self.xn, self.v, W.bv)
            for t, nh in ((self.k, c.kv_heads), (self.q, c.heads)):
                L.b200_rope(t.data_ptr(), t.data_ptr(), self.pos.data_ptr(), 0, hd, nh, 1, hd, nh * hd, hd, nh * hd, hd, c.rope_mode, 0,
                            c.rope_theta, 1.0, 0.0, 1.0, 32.0, 1.0, st); self._n += 1
            L.b200_kv_store(self.k.data_ptr(), self.v.data_ptr(), W.kc.data_ptr(), W.vc.data_ptr(), c.kv_hidden, c.kv_hidden, c.max_len, n_past, st)
            L.b200_attn_decode(self.q.data_ptr(), W.kc.data_ptr(), W.vc.data_ptr(), self.att.data_ptr(), self.scratch.data_ptr(), c.heads, c.kv_heads,
                               hd, n_kv, c.kv_hidden, c.max_len, 1.0 / math.sqrt(hd), st); self._n += 4
            self._mm(W.wo, c.hidden, c.hidden, self.att, self.o)
            L.b200_add(self.x.data_ptr(), self.o.data_ptr(), self.x.data_ptr(), c.hidden, st); self._n += 1
            L.b200_rms_norm(self.x.data_ptr(), W.ffn_norm.data_ptr(), self.xn.data_ptr(), c.hidden, 1, c.eps, st); self._n += 1
            self._mm(W.wgate, c.hidden, c.ffn, self.xn, self.gate)
            self._mm(W.wup, c.hidden, c.ffn, self.xn, self.up)
            L.b200_silu_mul(self.gate.data_ptr(), self.up.data_ptr(), self.gate.data_ptr(), c.ffn, st); self._n += 1
            self._mm(W.wdown, c.ffn, c.hidden, self.gate, self.o)
            L.b200_add(self.x.data_ptr(), self.o.data_ptr(), self.x.data_ptr(), c.hidden, st); self._n += 1
        if self.last:
            L.b200_rms_norm(self.x.data_ptr(), self.final_norm.data_ptr(), self.xn.data_ptr(), c.hidden, 1, c.eps, st); self._n += 1
            self._mm(self.lm_head, c.hidden, c.vocab, self.xn, self.logits)
        self.launches_per_step = self._n

    # ---- fused step: 10 launches per layer ------------------------------------------------------------------------
    def _ptr_arrays(self):
        """host-side pointer tables for b200_mul_mat_q_multi (kept alive for the lifetime of the session / graph)"""
        import ctypes as C
        c = self.cfg
        self._tables = []
        for W in self.layers:
            def arr(ctype, vals):
                a = (ctype * len(vals))(*vals)
                return a
            W.qkv = dict(W=arr(C.c_void_p, [W.wq.data_ptr(), W.wk.data_ptr(), W.wv.data_ptr()]), m=arr(C.c_int64, [c.hidden, c.kv_hidden, c.kv_hidden]),
                         y=arr(C.c_void_p, [self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr()]), ld=arr(C.c_int64, [c.hidden, c.kv_hidden, c.kv_hidden]),
                         b=arr(C.c_void_p, [0 if W.bq is None else W.bq.data_ptr(), 0 if W.bk is None else W.bk.data_ptr(), 0 if W.bv is None else W.bv.data_ptr()]))
            W.gu = dict(W=arr(C.c_void_p, [W.wgate.data_ptr(), W.wup.data_ptr()]), m=arr(C.c_int64, [c.ffn, c.ffn]),
                        y=arr(C.c_void_p, [self.gate.data_ptr(), 0]), ld=arr(C.c_int64, [c.ffn, c.ffn]), b=arr(C.c_void_p, [0, 0]))

    def attach_ring(self, ring):
        """sharding.PeerRing: later shards read / update the hidden row in their NVLink-visible mailbox; pointer tables are rebuilt"""
        self.ring = ring
        if ring is not None and ring.world > 1 and not self.first:
            self.x = ring.view("x").view(1, self.cfg.hidden)
        if hasattr(self, "_tables"):
            del self._tables

    def attn_launches(self, n_kv):
        """kernels b200_attn_decode_quant launches for this shape (the predicate of csrc/fused.cu attn_decode_mma_t): tensor-core scores + the
        thread-block-cluster V.P (2), or scores + split V.P + the tail that sums the partials (3)"""
        import os
        c = self.cfg
        cluster = os.environ.get("B200_ATTN_CLUSTER", "1") != "0" and os.environ.get("B200_ATTN_NO_MMA", "0") == "0"
        return 2 if cluster and (c.heads // c.kv_heads * c.head_dim) % 256 == 0 and (n_kv + 511) // 512 <= 16 else 3

    def enqueue_step_fused(self, n_past):
        c = self.cfg
        L = lib()
        st = torch.cuda.current_stream().cuda_stream
        if not hasattr(self, "_tables"):
            self._ptr_arrays()
        n = 0
        hd, n_kv = c.head_dim, n_past + 1
        q = self.qact.data_ptr()
        rc = 0
        ring = getattr(self, "ring", None)
        pio = ring.io if ring is not None and ring.world > 1 else None
        if pio:  # layer-sharded: wait (on the device) until the previous shard's hidden row / the last shard's token is in our mailbox
            rc |= L.b200_peer_wait(pio["wait_flag"], ring.seq_ptr, pio["wait_offset"], ring.status_ptr, st); n += 1
        if self.first:
            rc |= L.b200_get_rows(c.wtype, self.embed.data_ptr(), c.hidden, ring.tok_ptr if pio else self.tok.data_ptr(), 1, self.x.data_ptr(), st); n += 1
        pending = 0  # residual branch output not yet added into x
        for W in self.layers:
            rc |= L.b200_add_rmsnorm_quant(c.wtype, self.x.data_ptr(), pending, W.attn_norm.data_ptr(), self.x.data_ptr() if pending else 0, 0, q,
                                           c.hidden, 1, c.eps, st); n += 1
            t = W.qkv
            rc |= L.b200_mul_mat_q_multi(c.wtype, 0, 3, t["W"], t["m"], t["y"], t["ld"], t["b"], c.hidden, q, 1, st); n += 1
            rc |= L.b200_rope_kv_store(self.q.data_ptr(), self.k.data_ptr(), self.v.data_ptr(), self.pos.data_ptr(), 0, W.kc.data_ptr(), W.vc.data_ptr(),
                                       c.heads, c.kv_heads, hd, c.rope_mode, c.rope_theta, c.kv_hidden, c.max_len, st); n += 1
            # scores, split V.P, and the tail that sums the partials AND quantizes the result for the o-projection
            rc |= L.b200_attn_decode_quant(self.q.data_ptr(), W.kc.data_ptr(), W.vc.data_ptr(), self.att.data_ptr(), self.scratch.data_ptr(), c.heads,
                                           c.kv_heads, hd, n_kv, c.kv_hidden, c.max_len, 1.0 / math.sqrt(hd), c.wtype, q, st); n += self.attn_launches(n_kv)
            rc |= L.b200_mul_mat_q(c.wtype, W.wo.data_ptr(), c.hidden, c.hidden, q, 1, self.o.data_ptr(), c.hidden, 0, st); n += 1
            rc |= L.b200_add_rmsnorm_quant(c.wtype, self.x.data_ptr(), self.o.data_ptr(), W.ffn_norm.data_ptr(), self.x.data_ptr(), 0, q, c.hidden, 1, c.eps, st); n += 1
            t = W.gu
            rc |= L.b200_mul_mat_q_multi(c.wtype, 1, 2, t["W"], t["m"], t["y"], t["ld"], t["b"], c.hidden, q, 1, st); n += 1
            rc |= L.b200_quantize_act(c.wtype, self.gate.data_ptr(), c.ffn, c.ffn, 1, q, st); n += 1
            rc |= L.b200_mul_mat_q(c.wtype, W.wdown.data_ptr(), c.ffn, c.hidden, q, 1, self.o.data_ptr(), c.hidden, 0, st); n += 1
            pending = self.o.data_ptr()
        if self.last:
            rc |= L.b200_add_rmsnorm_quant(c.wtype, self.x.data_ptr(), pending, self.final_norm.data_ptr(), self.x.data_ptr() if pending else 0, 0, q,
                                           c.hidden, 1, c.eps, st); n += 1
            rc |= L.b200_mul_mat_q(c.wtype, self.lm_head.data_ptr(), c.hidden, c.vocab, q, 1, self.logits.data_ptr(), c.vocab, 0, st); n += 1
            if pio or getattr(self, "device_argmax", False):   # greedy sampling on the device (src/models.cpp:1026-1031 does it on the host)
                rc |= L.b200_argmax(self.logits.data_ptr(), c.vocab, self.next_tok.data_ptr(), st); n += 1
        elif pending:
            rc |= L.b200_add(self.x.data_ptr(), pending, self.x.data_ptr(), c.hidden, st); n += 1
        if pio:  # hand the result to the next shard through its NVLink-mapped mailbox and raise its flag
            if self.last:
                rc |= L.b200_peer_send(0, 0, 0, self.next_tok.data_ptr(), pio["send_tok"], pio["send_flag"], ring.seq_ptr, 0, st)
            else:
                rc |= L.b200_peer_send(self.x.data_ptr(), pio["send_x"], c.hidden, 0, 0, pio["send_flag"], ring.seq_ptr, 0, st)
            n += 1
        if rc:
            raise RuntimeError(f"fused step failed rc={rc}")
        self.launches_per_step = n

    # ---- fused=3: the whole step as ONE persistent kernel (csrc/decode_mk.cu, b200_decode_step) ----------------------------------
    def _mk_plan(self):
        import ctypes as C
        c = self.cfg
        n = len(self.layers)
        arr = (DecodeLayer * max(n, 1))()
        p = lambda t: 0 if t is None else t.data_ptr()
        for i, W in enumerate(self.layers):
            arr[i] = DecodeLayer(p(W.wq), p(W.wk), p(W.wv), p(W.wo), p(W.wgate), p(W.wup), p(W.wdown), p(W.bq), p(W.bk), p(W.bv),
                                 p(W.attn_norm), p(W.ffn_norm), p(W.kc), p(W.vc))
        m = DecodeModel(c.wtype, n, c.hidden, c.heads, c.kv_heads, c.head_dim, c.ffn, c.vocab, c.rope_mode, 0, c.rope_theta, c.eps, 1.0 / math.sqrt(c.head_dim),
                        c.kv_hidden, c.max_len, arr, p(self.embed) if self.first else 0, p(self.final_norm) if self.last else 0,
                        p(self.lm_head) if self.last else 0, 0)
        err = C.c_int(0)
        plan = lib().b200_decode_plan_create(C.byref(m), c.max_len, C.byref(err))
        if not plan:
            raise RuntimeError(f"b200_decode_plan_create failed err={err.value}")
        self._plan = plan
        self._plan_keep = (arr, m)
        g, sm, ns, stg, ks = (C.c_int(0) for _ in range(5))
        lib().b200_decode_plan_info(plan, C.byref(g), C.byref(sm), C.byref(ns), C.byref(stg), C.byref(ks))
        self.mk_info = dict(grid=g.value, smem_bytes=sm.value, n_steps=ns.value, stages=stg.value, ks=ks.value)

    def enqueue_step_mk(self, n_past=None, advance=False, step_begin=0, step_end=0):
        """n_past is NOT needed: the kernel reads the position from self.pos on the device (one captured graph serves every position)."""
        import ctypes as C
        if self._plan is None:
            self._mk_plan()
        ring = getattr(self, "ring", None)   # sharding.PeerRing: the hidden row / token arrive in (and leave through) NVLink-mapped mailboxes
        x_ptr = ring.x_ptr if ring is not None and not self.first else self.x.data_ptr()
        tok_ptr = (ring.tok_ptr if ring is not None and ring.world > 1 else self.tok.data_ptr()) if self.first else 0
        pio = ring.io if ring is not None else {}
        flags = (1 if advance else 0) if not pio else 2   # sharded: every rank advances its own position; the token travels through the ring
        io = DecodeIO(tok_ptr, self.pos.data_ptr(), -1, -1, x_ptr, self.logits.data_ptr() if self.last else 0,
                      self.next_tok.data_ptr() if self.last else 0, flags, step_begin, step_end,
                      pio.get("wait_flag", 0), pio.get("send_x", 0), pio.get("send_flag", 0), pio.get("send_tok", 0), pio.get("wait_offset", 0), 0)
        rc = lib().b200_decode_step(self._plan, C.byref(io), torch.cuda.current_stream().cuda_stream)
        if rc:
            raise RuntimeError(f"b200_decode_step failed rc={rc}")
        self.launches_per_step = 1

    def mk_status(self):
        return lib().b200_decode_plan_status(self._plan, torch.cuda.current_stream().cuda_stream) if self._plan else 0

    def __del__(self):
        try:
            if self._plan:
                lib().b200_decode_plan_destroy(self._plan)
                self._plan = None
        except Exception:  # noqa: BLE001  (interpreter shutdown)
            pass

    def enqueue(self, n_past):
        if self.fused == 3:
            self.enqueue_step_mk(n_past, advance=getattr(self, "mk_advance", False))
        elif self.fused:
            self.enqueue_step_fused(n_past)
        else:
            self.enqueue_step(n_past)

    def step(self, token, n_past):
        """eager step with host token / position (tests)"""
        self.tok.fill_(int(token)); self.pos.fill_(int(n_past))
        self.enqueue(n_past)
        return self.logits if self.last else self.x

    def capture(self, n_past):
        """capture the whole step at a fixed n_past into one CUDA graph (bench: `value`)"""
        self.pos.fill_(int(n_past))
        self.enqueue(n_past)  # warm-up outside capture (first-use cudaFuncSetAttribute etc.)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                self.enqueue(n_past)
        self._graph = g
        return g

    def fill_kv_random(self, n, seed=0):
        gen = torch.Generator(device="cuda"); gen.manual_seed(seed)
        for W in self.layers:
            W.kc[:n] = torch.randn((n, self.cfg.kv_hidden), device="cuda", generator=gen).to(torch.float16)
            W.vc[:, :n] = torch.randn((self.cfg.kv_hidden, n), device="cuda", generator=gen).to(torch.float16)


def make_config(name, wtype, layers=None, max_len=4352):
    d = dict(CONFIGS[name])
    if layers:
        d["layers"] = layers
    return Config(wtype, max_len=max_len, name=name, **d)
