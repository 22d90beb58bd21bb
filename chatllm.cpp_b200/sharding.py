"""Layer sharding for multi-GPU decode (BASELINE.json configs[3]): the path is one sequential stream, so it is split BY
LAYER exactly like the reference's `-ngl "0:16;1:16"` (docs/gpu.md:36-52, src/backend.cpp:578-652): rank r owns a contiguous
layer range with its weights and its KV-cache shard; the single [hidden] F32 row is handed to the next rank with
torch.distributed send/recv (gloo in the CPU tests, and the fallback on GPUs).  No collective is needed anywhere else.

On GPUs the hand-off does not go through the host at all (PeerRing): every rank exports a small device mailbox with CUDA IPC, maps its
successor's, and the persistent decode kernel of rank r stores the hidden row straight into rank r+1's mailbox over NVLink and raises
its flag; rank r+1's kernel — already launched, its weight rings already streaming — spins on that LOCAL flag.  The last rank returns
the next token to rank 0 the same way.  Each rank just replays one CUDA graph per token (include/chatllm_b200.h, b200_decode_io)."""
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


def ngl_spec(n_layers, world):
    """the reference host's `-ngl` argument (docs/gpu.md:36-52: `[id:]layer_specs[;id:layer_specs]..`) that places the layers on `world` devices
    exactly as plan_layers() shards them: the embedding ("prolog") with the first range, final norm + lm_head ("epilog") with the last"""
    if world <= 1:
        return "all"
    parts = []
    for d, (lo, hi) in enumerate(plan_layers(n_layers, world)):
        spec = [str(hi - lo)] if hi > lo else []
        if d == 0:
            spec.append("prolog")
        if d == world - 1:
            spec.append("epilog")
        if spec:
            parts.append(f"{d}:" + ",".join(spec))
    return ";".join(parts)


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


def mailbox_layout(hidden):
    """byte offsets inside a rank's device mailbox: the incoming hidden row, the flag raised when it has arrived, and (rank 0) the
    incoming next token with its flag.  Flags are 64-bit launch counters, never reset."""
    x_bytes = hidden * 4
    return dict(x=0, x_flag=x_bytes, tok_flag=x_bytes + 8, tok=x_bytes + 16, seq=x_bytes + 24, status=x_bytes + 32, bytes=x_bytes + 64)


def ring_io(rank, world, hidden, base_self, base_next):
    """b200_decode_io hand-off fields of `rank` (device addresses): what it waits on in its own mailbox and where it stores its result
    in the mailbox of rank (rank + 1) % world.  Rank 0 waits for the token of the PREVIOUS step (wait_offset 0), later ranks for the
    hidden row of THIS step (wait_offset 1); the last rank sends the token to rank 0 instead of a hidden row."""
    if world == 1:
        return {}
    L = mailbox_layout(hidden)
    io = dict(wait_flag=base_self + (L["tok_flag"] if rank == 0 else L["x_flag"]), wait_offset=0 if rank == 0 else 1, send_x=0, send_tok=0)
    if rank < world - 1:
        io.update(send_x=base_next + L["x"], send_flag=base_next + L["x_flag"])
    else:
        io.update(send_tok=base_next + L["tok"], send_flag=base_next + L["tok_flag"])
    return io


class PeerRing:
    """CUDA-IPC mailboxes of a layer-sharded decode (one process per GPU, torch.distributed only for the one-time handle exchange)."""

    def __init__(self, rank, world, hidden, lib):
        import ctypes as C
        self.rank, self.world, self.hidden, self.lib = rank, world, hidden, lib
        L = mailbox_layout(hidden)
        self.base, self.next_base = C.c_void_p(0), C.c_void_p(0)
        handle = C.create_string_buffer(64)
        rc = lib.b200_ipc_alloc(L["bytes"], C.byref(self.base), handle)
        if rc:
            raise RuntimeError(f"b200_ipc_alloc failed rc={rc}")
        self.layout = L
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw))
            nxt = C.create_string_buffer(handles[(rank + 1) % world], 64)
            rc = lib.b200_ipc_open(nxt, C.byref(self.next_base))
            if rc:
                raise RuntimeError(f"b200_ipc_open failed rc={rc} (peer access between the GPUs of this node is required)")
        self.io = ring_io(rank, world, hidden, self.base.value, self.next_base.value or 0)

    def view(self, name, dtype="<f4", count=None):
        """zero-copy torch view of a mailbox field (host-side initialisation / inspection)"""
        import torch

        class _CAI:
            pass
        o = _CAI()
        n = count if count is not None else (self.hidden if name == "x" else 1)
        o.__cuda_array_interface__ = dict(shape=(n,), typestr=dtype, data=(self.base.value + self.layout[name], False), version=3)
        return torch.as_tensor(o, device="cuda")

    @property
    def seq_ptr(self):
        """tokens this rank has completed (device counter used by b200_peer_wait / b200_peer_send)"""
        return self.base.value + self.layout["seq"]

    @property
    def status_ptr(self):
        return self.base.value + self.layout["status"]

    @property
    def x_ptr(self):
        return self.base.value + self.layout["x"]

    @property
    def tok_ptr(self):
        return self.base.value + self.layout["tok"]

    def close(self):
        if self.next_base.value:
            self.lib.b200_ipc_close(self.next_base)
            self.next_base.value = None
        if self.base.value:
            self.lib.b200_ipc_free(self.base)
            self.base.value = None
