"""Vocab-parallel lm_head across the GPUs of one node (one process per GPU, NCCL).

north_star / SURVEY.md 8(e): each rank owns a tile-granular slice of lm_head.weight; every
rank sees all token rows; only the per-token (max, sum-exp, target-logit) triples cross
NVLink in the forward (12 bytes per token per rank, `_gather_stats` in logprob.py), dW stays
local to the slice owner and the partial dHidden sums are all-reduced once per step.
"""
from typing import List, Tuple

TILE = 256  # vocabulary columns per MMA tile (GemmShape::BN)


def vocab_slices(V: int, world: int) -> List[Tuple[int, int]]:
    """Whole 256-column tiles per rank (not V/world columns: 152064/8 = 19008 = 74.25 tiles);
    the first `rem` ranks take one extra tile; the ragged last tile goes to the last rank."""
    tiles = -(-V // TILE)
    base, rem = divmod(tiles, world)
    out, t0 = [], 0
    for r in range(world):
        t1 = t0 + base + (1 if r < rem else 0)
        out.append((min(t0 * TILE, V), min(t1 * TILE, V)))
        t0 = t1
    return out


def shard_weight(weight, rank: int, world: int):
    """(local rows of lm_head.weight as a contiguous tensor, v_offset)."""
    v0, v1 = vocab_slices(weight.shape[0], world)[rank]
    return weight[v0:v1].contiguous(), v0


class PeerExchange:
    """Forward exchange of the vocab-parallel path fused into the merge kernel.

    Each rank's K1 writes its per-token (max, sum-exp, target-logit) triple straight into a
    symmetric (peer-mapped, NVLink P2P) buffer; after ONE stream-ordered cross-rank barrier the
    merge kernel of every rank loads all ranks' triples directly over NVLink
    (`o3v_lmhead_merge_stats_peers`).  No NCCL all-gather, no gathered copy.  Two slots alternate
    so that a rank can write call k+2 while a slow peer still reads call k (a rank only gets past
    barrier k+1 after every peer has finished merge k on its stream).

    Pass an instance as the `group=` argument of `logprob.fused_logprob` / `fused_logprob_gspo`;
    `.group` is the torch.distributed group used for the remaining collective (dHidden all-reduce).
    """

    def __init__(self, group, capacity_tokens: int, hidden_size: int = 0, device=None, allreduce_ctas: int = 0):
        """`hidden_size` > 0 also allocates a peer-mapped [capacity, hidden] bf16 dHidden buffer: the
        partial dHidden of every token chunk is then all-reduced by `o3v_allreduce_bf16_peers` on a side
        stream WHILE the dW GEMM of the chunk runs (the kernel has no smem and ~43 registers, so its
        CTAs share SMs with the persistent GEMM CTAs), instead of one exposed NCCL all-reduce."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group = group
        self.world = dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("PeerExchange supports up to 16 ranks (one NVLink domain)")
        self.cap = int(capacity_tokens)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((2, 3, self.cap), dtype=torch.float32, device=device)
        self.handle = symm.rendezvous(self.buf, group)
        self.rank = self.handle.rank
        self._ptrs = [int(p) for p in self.handle.buffer_ptrs]
        self._calls = 0
        self.hidden_size = int(hidden_size)
        self.dh = None
        if self.hidden_size > 0:
            self.dh = symm.empty((self.cap, self.hidden_size), dtype=torch.bfloat16, device=device)
            self.dh_handle = symm.rendezvous(self.dh, group)
            self._dh_ptrs = [int(p) for p in self.dh_handle.buffer_ptrs]
            self.side = torch.cuda.Stream(device=device)
            self.ar_ctas = int(allreduce_ctas) if allreduce_ctas else torch.cuda.get_device_properties(device).multi_processor_count

    def dh_view(self, tokens: int, hidden: int):
        """[tokens, hidden] bf16 view of the peer-mapped dHidden buffer (valid until the next step)."""
        if self.dh is None or hidden != self.hidden_size or tokens > self.cap:
            return None
        return self.dh[:tokens]

    def allreduce_dh_async(self, row0: int, rows: int):
        """Sum rows [row0, row0 + rows) of every rank's dHidden buffer, on the side stream, ordered
        after everything enqueued so far on the current stream."""
        import ctypes
        import torch
        from . import _lib
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        off = row0 * self.hidden_size * 2
        arr = (ctypes.c_void_p * self.world)(*[p + off for p in self._dh_ptrs])
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.dh_handle.barrier(channel=2)          # every rank has written its partial sums of these rows
            with torch.cuda.device(self.dh.device):
                _lib.call("o3v_allreduce_bf16_peers", 1, _lib.load().o3v_allreduce_bf16_peers, arr, self.world,
                          self.rank, rows * self.hidden_size, self.ar_ctas,
                          ctypes.c_void_p(self.side.cuda_stream))
            self.dh_handle.barrier(channel=3)          # every rank's results have landed in every buffer

    def wait_allreduce(self):
        import torch
        torch.cuda.current_stream().wait_stream(self.side)

    def next_slot(self) -> int:
        slot = self._calls & 1
        self._calls += 1
        return slot

    def local_stats(self, slot: int):
        """This rank's [3, cap] triple buffer of `slot` (K1 writes rows [:, :T])."""
        return self.buf[slot]

    def merge(self, slot: int, T: int):
        import ctypes
        import torch
        from . import _lib
        from .gspo import _p, _stream
        if T > self.cap:
            raise ValueError("PeerExchange capacity %d < %d tokens" % (self.cap, T))
        self.handle.barrier(channel=slot)               # every rank's K1 for this call has written its triple
        logp = torch.empty(T, dtype=torch.float32, device=self.buf.device)
        lse = torch.empty(T, dtype=torch.float32, device=self.buf.device)
        off = slot * 3 * self.cap * 4
        arr = (ctypes.c_void_p * self.world)(*[p + off for p in self._ptrs])
        with torch.cuda.device(self.buf.device):
            _lib.call("o3v_lmhead_merge_stats_peers", 1, _lib.load().o3v_lmhead_merge_stats_peers, arr, self.world,
                      self.cap, T, _p(logp), _p(lse), _stream())
        return logp, lse
