"""Fabricates synthetic chatllm GGMM model files with the real architectures' shapes (no checkpoints can be downloaded).

Layout follows the reference writer convert.py:1329-1358 (header), :570-602 (tensor records), :1566-1582 + :1640-1650
(Llama-3 config), reader src/models.cpp:1995-2047 / src/chat.cpp:1425-1459.  Weights are random
*valid* quant blocks (tests/qformats.random_blocks: zero-mean, std ~0.02) drawn from a seeded pool; 1-D tensors are F32.
Both backends read the same bytes, which is all parity needs.

    python tools/make_model.py --arch llama3-8b --quant q4_K --out /tmp/l3.bin [--layers N] [--max_length 4352]
"""
import argparse
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import qformats as qf  # noqa: E402

MODEL_TYPE = {"llama3": 0x1700, "qwen2": 0x710, "mixtral": 0x601}
QUANT = {"q4_K": qf.Q4_K, "q4_0": qf.Q4_0, "q8_0": qf.Q8_0}

ARCHS = {
    # name: (family, vocab, hidden, heads, kv_heads, layers, ffn, rope_theta)
    "tiny-test": ("llama3", 512, 256, 4, 2, 2, 512, 10000.0),
    "tinyllama-1.1b": ("llama3", 32000, 2048, 32, 4, 22, 5632, 10000.0),
    "llama3-8b": ("llama3", 128256, 4096, 32, 8, 32, 14336, 500000.0),
    "qwen2.5-7b": ("qwen2", 152064, 3584, 28, 4, 28, 18944, 1000000.0),
    "qwen2-test": ("qwen2", 512, 256, 4, 2, 2, 512, 1000000.0),
    # Mixtral: 8 experts, top-2 (the reference's model class hard-codes both, models/mistral.h:143-146), sliding window 4096
    "mixtral-test": ("mixtral", 512, 256, 4, 2, 2, 512, 1000000.0),
    "mixtral-8x7b": ("mixtral", 32000, 4096, 32, 8, 32, 14336, 1000000.0),
}
N_EXPERTS, EXPERTS_PER_TOK = 8, 2


def bytes_to_unicode():
    """GPT-2 byte <-> unicode table (the byte-level BPE alphabet BPEProcessor2 expects, src/tokenizer.cpp)."""
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b); cs.append(256 + n); n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def write_vocab_sp(f, vocab_size):
    """i32 len | bytes | f32 score ... -1   (sentencepiece-style list read by BPEProcessor1::Load -> load_vocab_list(has_score=true,
    has_type=false), src/tokenizer.cpp:310-440; llama v2 / Mistral / Mixtral tokenizers, models/llama.cpp:30-33)"""
    for i in range(vocab_size):
        t = f"<t{i}>".encode()
        f.write(struct.pack("i", len(t))); f.write(t); f.write(struct.pack("<f", -float(i)))
    f.write(struct.pack("i", -1))


def write_vocab_bpe2(f, vocab_size):
    """i32 len | bytes | u8 type ... -1 | merges ... -1   (convert.py:1107-1123; reader src/tokenizer.cpp:310-370, :475-489)"""
    b2u = bytes_to_unicode()
    toks = [b2u[b].encode("utf-8") for b in range(256)]
    toks += [f"<t{i}>".encode() for i in range(256, vocab_size)]
    for t in toks[:vocab_size]:
        f.write(struct.pack("i", len(t))); f.write(t); f.write(struct.pack("B", 1))  # token_type::NORMAL
    f.write(struct.pack("i", -1))
    f.write(struct.pack("i", -1))  # no merges: every character is its own token


class Pool:
    """Seeded pool of random valid blocks; rows are cut from it at random block-aligned offsets."""
    def __init__(self, t, seed, pool_rows=256, k=4096):
        self.t, self.rng = t, np.random.default_rng(seed)
        self.k = k
        self.blocks = qf.random_blocks(t, pool_rows, k, rng=self.rng).reshape(-1)
        self.bb = qf.BLK[t][1]

    def rows(self, m, k):
        rb = qf.row_size(self.t, k)
        nblk_total = self.blocks.size // self.bb
        nb = rb // self.bb
        starts = self.rng.integers(0, nblk_total - nb, size=m)
        idx = (starts[:, None] + np.arange(nb)[None, :]).reshape(-1)
        return self.blocks.reshape(nblk_total, self.bb)[idx].reshape(m, rb)


def dump_tensor(f, name, shape, gtype, data_bytes):
    nb = name.encode()
    f.write(struct.pack("i", len(nb))); f.write(nb)
    f.write(struct.pack("i" * (2 + len(shape)), len(shape), *shape, gtype))
    pos = (f.tell() + 15) // 16 * 16
    f.seek(pos)
    f.write(data_bytes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="tiny-test", choices=sorted(ARCHS))
    ap.add_argument("--quant", default="q4_K", choices=sorted(QUANT))
    ap.add_argument("--out", required=True)
    ap.add_argument("--layers", type=int, default=0, help="override the number of layers (0 = architecture default)")
    ap.add_argument("--max_length", type=int, default=4352)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()

    family, vocab, hidden, heads, kv_heads, layers, ffn, theta = ARCHS[a.arch]
    if a.layers:
        layers = a.layers
    t = QUANT[a.quant]
    hd = hidden // heads
    pool = Pool(t, a.seed)
    rng = np.random.default_rng(a.seed + 1)

    def qmat(m, k):
        out = np.empty((m, qf.row_size(t, k)), dtype=np.uint8)
        for r0 in range(0, m, 16384):
            out[r0:r0 + 16384] = pool.rows(min(16384, m - r0), k)
        return out.tobytes()

    def vec(n, kind):
        v = (1.0 + 0.1 * rng.standard_normal(n)) if kind == "norm" else 0.02 * rng.standard_normal(n)
        return v.astype(np.float32).tobytes()

    with open(a.out, "wb") as f:
        f.write(b"ggmm")
        f.write(struct.pack("i" * 4, 1, 0, 0, 0))
        meta = json.dumps({"model_name": f"synthetic-{a.arch}", "model_native_name": a.arch}).encode()
        meta += b"\0" * ((-len(meta)) % 4)
        f.write(meta)

        def mark(off):
            size = f.tell(); f.seek(off); f.write(struct.pack("i", size)); f.seek(0, 2)

        mark(8)
        f.write(struct.pack("ii", MODEL_TYPE[family], 1))
        f.write(struct.pack("i" * 11, t, vocab, hidden, heads, layers, ffn, a.max_length, 1, 2, -1, -1))
        if family == "llama3":
            f.write(struct.pack("i", kv_heads)); f.write(struct.pack("<f", theta))
        elif family == "qwen2":
            f.write(struct.pack("ii", kv_heads, a.max_length)); f.write(struct.pack("<f", theta))
        elif family == "mixtral":  # mistral::Config + mixtral::Config (models/mistral.h:5-10, :45-49)
            f.write(struct.pack("ii", kv_heads, 4096)); f.write(struct.pack("<f", theta)); f.write(struct.pack("ii", EXPERTS_PER_TOK, N_EXPERTS))
        mark(12)
        if family == "mixtral":
            write_vocab_sp(f, vocab)
        else:
            write_vocab_bpe2(f, vocab)
        mark(16)

        dump_tensor(f, "model.embed_tokens.weight", (vocab, hidden), t, qmat(vocab, hidden))
        for i in range(layers):
            p = f"model.layers.{i}."
            dump_tensor(f, p + "input_layernorm.weight", (hidden,), 0, vec(hidden, "norm"))
            if family == "mixtral":
                # merged expert stacks [n_expert, rows, k] (loader: models/mistral.h:112-114 -> src/chat.cpp:1606-1613) + the router
                dump_tensor(f, p + "mlp.experts_down.weight", (N_EXPERTS, hidden, ffn), t, qmat(N_EXPERTS * hidden, ffn))
                dump_tensor(f, p + "mlp.experts_gate.weight", (N_EXPERTS, ffn, hidden), t, qmat(N_EXPERTS * ffn, hidden))
                dump_tensor(f, p + "mlp.experts_up.weight", (N_EXPERTS, ffn, hidden), t, qmat(N_EXPERTS * ffn, hidden))
                dump_tensor(f, p + "block_sparse_moe.gate.weight", (N_EXPERTS, hidden), t, qmat(N_EXPERTS, hidden))
            else:
                dump_tensor(f, p + "mlp.down_proj.weight", (hidden, ffn), t, qmat(hidden, ffn))
                dump_tensor(f, p + "mlp.gate_proj.weight", (ffn, hidden), t, qmat(ffn, hidden))
                dump_tensor(f, p + "mlp.up_proj.weight", (ffn, hidden), t, qmat(ffn, hidden))
            dump_tensor(f, p + "post_attention_layernorm.weight", (hidden,), 0, vec(hidden, "norm"))
            dump_tensor(f, p + "self_attn.k_proj.weight", (kv_heads * hd, hidden), t, qmat(kv_heads * hd, hidden))
            dump_tensor(f, p + "self_attn.o_proj.weight", (hidden, hidden), t, qmat(hidden, hidden))
            dump_tensor(f, p + "self_attn.q_proj.weight", (hidden, hidden), t, qmat(hidden, hidden))
            dump_tensor(f, p + "self_attn.v_proj.weight", (kv_heads * hd, hidden), t, qmat(kv_heads * hd, hidden))
            if family == "qwen2":
                dump_tensor(f, p + "self_attn.q_proj.bias", (hidden,), 0, vec(hidden, "bias"))
                dump_tensor(f, p + "self_attn.k_proj.bias", (kv_heads * hd,), 0, vec(kv_heads * hd, "bias"))
                dump_tensor(f, p + "self_attn.v_proj.bias", (kv_heads * hd,), 0, vec(kv_heads * hd, "bias"))
        dump_tensor(f, "model.norm.weight", (hidden,), 0, vec(hidden, "norm"))
        dump_tensor(f, "lm_head.weight", (vocab, hidden), t, qmat(vocab, hidden))
    print(json.dumps({"out": a.out, "bytes": os.path.getsize(a.out), "arch": a.arch, "quant": a.quant, "layers": layers}))


if __name__ == "__main__":
    main()
