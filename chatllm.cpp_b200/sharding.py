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
    """run_shard(x) -> x for this rank's layers; step() moves one token through all ranks"""

    def __init__(self, rank, world, hidden_buf, run_shard):
        self.rank, self.world, self.x, self.run = rank, world, hidden_buf, run_shard

    def step(self):
        if self.rank > 0:
            dist.recv(self.x, src=self.rank - 1)
        self.run(self.x)
        if self.rank < self.world - 1:
            dist.send(self.x, dst=self.rank + 1)
