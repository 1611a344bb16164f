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

    def __init__(self, group, capacity_tokens: int, device=None):
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
