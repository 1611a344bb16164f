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


def token_owner_rows(tokens: int, world: int, rank: int) -> Tuple[int, int, int]:
    """(rows_per_owner, lo, hi): the data-parallel layout the reduce-scatter of dHidden assumes (include/o3v.h,
    o3v_lmhead_bwd_dhidden_scatter): rank r owns the contiguous token rows [r * ceil(T / P), (r + 1) * ceil(T / P))
    clipped to T."""
    rpo = -(-tokens // world)
    return rpo, min(rank * rpo, tokens), min((rank + 1) * rpo, tokens)


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

    def __init__(self, group, capacity_tokens: int, hidden_size: int = 0, device=None, allreduce_ctas: int = 0,
                 dh_mode: str = "all_reduce"):
        """`hidden_size` > 0 also allocates peer-mapped dHidden storage [capacity, hidden] bf16, used in one of three ways:

        dh_mode="all_reduce": every rank ends with the full dHidden.  The partial dHidden of every token chunk is
        all-reduced by `o3v_allreduce_bf16_peers` on a side stream WHILE the dW GEMM of the chunk runs (the kernel
        has no smem and ~43 registers, so its CTAs share SMs with the persistent GEMM CTAs).

        dh_mode="reduce_scatter" (the data-parallel layout of the reference's launch: rank r owns token
        rows [r * ceil(T / P), ...) and only needs ITS rows of dHidden): same partial buffers, but every owner PULLS
        its rows from all peers (`o3v_reduce_scatter_bf16_peers`: NVLink loads, fp32 sum in rank order, local
        store) on the side stream beside the dW GEMM: half the NVLink bytes of the all-reduce and no write fan-out.

        dh_mode="reduce_scatter_fused": the K2a epilogue itself stores every output tile into the owner's slot buffer
        over NVLink (`o3v_lmhead_bwd_dhidden_scatter`: GEMM and transfer are ONE kernel); after one barrier each
        owner sums its P slots locally (`o3v_sum_slots_bf16`).  Bit-identical to "reduce_scatter"; measured on
        8 x B200 (profiles/r2_ab_dh_collective_n8.jsonl) the remote stores stall the epilogue of the 256x512 tiles
        (one accumulator stage: no MMA runs under it): 48.7 ms per c2 step against 39.6 ms, so it is an option, not
        the default."""
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
        if dh_mode not in ("all_reduce", "reduce_scatter", "reduce_scatter_fused"):
            raise ValueError("dh_mode must be 'all_reduce', 'reduce_scatter' or 'reduce_scatter_fused'")
        self.dh_mode = dh_mode
        self.slot_rows = -(-self.cap // self.world)
        self._owned = None
        if self.hidden_size > 0 and dh_mode == "reduce_scatter_fused":
            self.dh_slots = symm.empty((self.world, self.slot_rows, self.hidden_size), dtype=torch.bfloat16, device=device)
            self.dh_handle = symm.rendezvous(self.dh_slots, group)
            self._slot_ptrs = [int(p) for p in self.dh_handle.buffer_ptrs]
            self.side = torch.cuda.Stream(device=device)
            self.ar_ctas = int(allreduce_ctas) if allreduce_ctas else torch.cuda.get_device_properties(device).multi_processor_count
        elif self.hidden_size > 0:
            self.dh = symm.empty((self.cap, self.hidden_size), dtype=torch.bfloat16, device=device)
            self.dh_handle = symm.rendezvous(self.dh, group)
            self._dh_ptrs = [int(p) for p in self.dh_handle.buffer_ptrs]
            self.side = torch.cuda.Stream(device=device)
            self.ar_ctas = int(allreduce_ctas) if allreduce_ctas else torch.cuda.get_device_properties(device).multi_processor_count

    # ------------------------------------------------------------------ reduce-scatter of dHidden (fused into K2a)
    def owner_rows(self, tokens: int, rank=None):
        """(rows_per_owner, lo, hi): token rows owned by `rank` (default: this rank) of a step of `tokens` rows."""
        return token_owner_rows(tokens, self.world, self.rank if rank is None else rank)

    def dhidden_scatter(self, dlogits, weight, row0: int, tokens_total: int):
        """K2a of one chunk (rows row0 .. row0 + dlogits.shape[0] of the step): dH tiles go straight to their owners."""
        import ctypes
        import torch
        from . import _lib
        from .gspo import _p, _stream
        if self.hidden_size != weight.shape[1] or tokens_total > self.cap:
            raise ValueError("PeerExchange was sized for %d tokens x %d" % (self.cap, self.hidden_size))
        T, V = dlogits.shape
        rpo = self.owner_rows(tokens_total)[0]
        arr = (ctypes.c_void_p * self.world)(*self._slot_ptrs)
        with torch.cuda.device(dlogits.device):
            _lib.call("o3v_lmhead_bwd_dhidden_scatter", 1, _lib.load().o3v_lmhead_bwd_dhidden_scatter, _p(dlogits),
                      dlogits.stride(0), _p(weight), T, V, self.hidden_size, arr, self.world, self.rank, rpo,
                      self.slot_rows, int(row0), _stream())

    def dhidden_reduce_async(self, tokens_total: int):
        """After the LAST K2a of the step: barrier (every rank's tiles have landed in my slots), then the local sum
        of the P slots on the side stream (overlaps the dW GEMM that follows on the current stream).  Returns the
        [rows_owned, hidden] bf16 result; call `wait_allreduce()` before using it."""
        import ctypes
        import torch
        from . import _lib
        from .gspo import _p
        rpo, lo, hi = self.owner_rows(tokens_total)
        out = torch.empty(hi - lo, self.hidden_size, dtype=torch.bfloat16, device=self.dh_slots.device)
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.dh_handle.barrier(channel=2)
            with torch.cuda.device(self.dh_slots.device):
                _lib.call("o3v_sum_slots_bf16", 1, _lib.load().o3v_sum_slots_bf16, _p(self.dh_slots), self.world,
                          self.slot_rows, hi - lo, self.hidden_size, _p(out), self.ar_ctas,
                          ctypes.c_void_p(self.side.cuda_stream))
            self.dh_handle.barrier(channel=3)          # nobody overwrites my slots (next step's K2a) before I have read them
        out.record_stream(self.side)
        return out

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

    def reduce_scatter_dh_async(self, row0: int, rows: int, tokens_total: int):
        """dh_mode="reduce_scatter": after K2a of the chunk [row0, row0 + rows) has written this rank's partial sums
        into the peer-mapped buffer, the owners of those rows pull and sum them (side stream, beside the dW GEMM).
        Returns the [rows_owned, hidden] bf16 tensor of the step (complete after the last chunk + wait_allreduce())."""
        import ctypes
        import torch
        from . import _lib
        from .gspo import _p
        _, lo, hi = self.owner_rows(tokens_total)
        if row0 == 0 or self._owned is None or self._owned.shape[0] != hi - lo:
            self._owned = torch.empty(hi - lo, self.hidden_size, dtype=torch.bfloat16, device=self.dh.device)
        out = self._owned
        a, b = max(lo, row0), min(hi, row0 + rows)
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        arr = (ctypes.c_void_p * self.world)(*self._dh_ptrs)
        with torch.cuda.stream(self.side):
            self.side.wait_event(ev)
            self.dh_handle.barrier(channel=2)          # every rank has written its partial sums of these rows
            if b > a:
                with torch.cuda.device(self.dh.device):
                    _lib.call("o3v_reduce_scatter_bf16_peers", 1, _lib.load().o3v_reduce_scatter_bf16_peers, arr,
                              self.world, a * self.hidden_size, (b - a) * self.hidden_size, _p(out[a - lo:b - lo]),
                              self.ar_ctas, ctypes.c_void_p(self.side.cuda_stream))
            self.dh_handle.barrier(channel=3)          # nobody overwrites its partial sums (next step) before all have read
        out.record_stream(self.side)
        return out

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


class PeerGather:
    """All-gather of per-rank ROW SLICES of several tensors by copy-engine pushes into peer-mapped buffers
    (in the reference's data-parallel launch every rank holds the hidden states / ids / reference
    log-probs of ITS sequences, the vocab-parallel head needs all rows on every rank).

    Each rank copies its slice into its own buffer and PUSHES it into every peer's buffer with plain device-to-device
    copies (cudaMemcpyAsync over NVLink, executed by the copy engines): no SM is taken from the persistent GEMM
    kernels of a step that is running, unlike NCCL all-gather kernels (round 1: e2e at N=8 rose from 83 % to 94 % of
    the device-resident rate).  `slots` buffers alternate so that step i+1 can be gathered while step i computes.
    """

    def __init__(self, group, shapes_dtypes, device=None, slots: int = 2):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.group, self.world = group, dist.get_world_size(group)
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.bufs, self.views, self.handles = [], [], []
        for _ in range(slots):
            bs, vs, hs = [], [], []
            for shape, dtype in shapes_dtypes:
                b = symm.empty(tuple(shape), dtype=dtype, device=device)
                h = symm.rendezvous(b, group)
                bs.append(b)
                hs.append(h)
                vs.append([h.get_buffer(r, tuple(shape), dtype) for r in range(self.world)])
            self.bufs.append(bs)
            self.views.append(vs)
            self.handles.append(hs)
        self.rank = self.handles[0][0].rank

    def gather(self, slot: int, locals_, row_lo: int):
        """Enqueue on the CURRENT stream: wait until every rank is done with `slot`, copy `locals_[i]` (rows
        row_lo .. row_lo + len) into all ranks' buffers, barrier.  Returns the list of full tensors of `slot`
        (complete on every rank once the stream reaches this point)."""
        h0 = self.handles[slot][0]
        h0.barrier(channel=4 + 2 * (slot & 1))
        for i, loc in enumerate(locals_):
            hi = row_lo + loc.shape[0]
            src = self.bufs[slot][i][row_lo:hi]
            if loc.data_ptr() != src.data_ptr():             # (the caller may have filled its own rows in place)
                src.copy_(loc, non_blocking=True)
            for r in range(self.world):
                if r != self.rank:
                    self.views[slot][i][r][row_lo:hi].copy_(src, non_blocking=True)
        h0.barrier(channel=5 + 2 * (slot & 1))
        return self.bufs[slot]
