"""Layer sharding for multi-GPU decode (BASELINE.json configs[3]): the path is one sequential stream, so it is split BY
LAYER exactly like the reference's `-ngl "0:16;1:16"` (docs/gpu.md:36-52, src/backend.cpp:578-652): rank r owns a contiguous
layer range with its weights and its KV-cache shard; the single [hidden] F32 row is handed to the next rank with
torch.distributed send/recv (NCCL on GPUs, gloo in the CPU tests).  No collective is needed anywhere else."""
import torch.distributed as dist


def plan_layers(n_layers, world):
    """contiguous, balanced ranges: the first (n_layers % world) ranks get one extra layer"""
    base, extra = divmod(n_layers, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


class Pipeline:
    """run_shard(x) -> x for this rank's layers; step() moves ONE token through all ranks.

    Single-stream decode is sequential: token i+1 is sampled from the logits of token i.  The last rank therefore returns
    the next token id to rank 0 (`tok`, 1 x int32) and rank 0 waits for it before it starts the next step — without this
    feedback edge the ranks would pipeline independent tokens and report a throughput no single stream can have."""

    def __init__(self, rank, world, hidden_buf, run_shard, tok_buf=None):
        self.rank, self.world, self.x, self.run, self.tok = rank, world, hidden_buf, run_shard, tok_buf
        self.steps = 0

    def step(self):
        if self.world > 1 and self.tok is not None and self.rank == 0 and self.steps > 0:
            dist.recv(self.tok, src=self.world - 1)
        if self.rank > 0:
            dist.recv(self.x, src=self.rank - 1)
        self.run(self.x)
        if self.rank < self.world - 1:
            dist.send(self.x, dst=self.rank + 1)
        elif self.world > 1 and self.tok is not None:
            dist.send(self.tok, dst=0)
        self.steps += 1

    def drain(self):
        """consume the token of the final step so no message is left in flight"""
        if self.world > 1 and self.tok is not None and self.rank == 0 and self.steps > 0:
            dist.recv(self.tok, src=self.world - 1)
