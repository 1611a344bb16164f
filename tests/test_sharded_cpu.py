"""Host-side logic of the vocab-parallel path on the CPU: slice planning, and the
(max, sum-exp, target-logit) exchange protocol over a world_size-2 gloo group."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import logps as ologps
from oracle import synth


@pytest.mark.parametrize("V", [152064, 151936, 256, 1000, 19008])
@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_vocab_slices_partition(V, world):
    from open_o3_video_b200.sharded import TILE, vocab_slices
    s = vocab_slices(V, world)
    assert len(s) == world and s[0][0] == 0 and s[-1][1] == V
    assert all(a[1] == b[0] for a, b in zip(s, s[1:]))
    assert all(a % TILE == 0 for a, b in s if b > a)        # every non-empty slice starts on an MMA tile boundary
    sizes = [b - a for a, b in s]
    assert max(sizes) - min(sizes) <= 2 * TILE              # balanced to within a tile (+ the ragged tail)
    if V >= world * TILE:
        assert min(sizes) > 0                               # (fewer tiles than ranks leaves ranks empty: unsupported)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, T, H, V, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from open_o3_video_b200 import logprob, sharded
        hidden, weight, targets = synth.head_inputs(T, H, V, seed=5)
        w_local, v0 = sharded.shard_weight(weight, rank, world)
        # per-slice statistics exactly as K1 defines them (include/o3v.h), computed with torch
        z = hidden @ w_local.T
        m = z.max(dim=1).values
        s = torch.exp(z - m[:, None]).sum(dim=1)
        tc = targets - v0
        own = (tc >= 0) & (tc < w_local.shape[0])
        zt = torch.where(own, z.gather(1, tc.clamp(0, w_local.shape[0] - 1)[:, None])[:, 0], torch.zeros(T))
        parts = logprob._gather_stats(torch.stack([m, s, zt]), dist.group.WORLD)      # the product's exchange
        assert parts.shape == (world, 3, T)
        # merge formula of o3v_lmhead_merge_stats
        M = parts[:, 0].max(dim=0).values
        S = (parts[:, 1] * torch.exp(parts[:, 0] - M)).sum(dim=0)
        logp = parts[:, 2].sum(dim=0) - (M + torch.log(S))
        ref, _ = ologps.token_logps(hidden, weight, targets)
        err = (logp - ref).abs().max().item()
        if rank == 0:
            out.put(err)
    finally:
        dist.destroy_process_group()


def test_stats_exchange_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 96, 64, 1000, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) < 1e-5


@pytest.mark.parametrize("T,world", [(768, 8), (1000, 8), (7, 4), (131072, 8), (5, 8)])
def test_token_owner_rows_partition(T, world):
    from open_o3_video_b200.sharded import token_owner_rows
    rows = [token_owner_rows(T, world, r) for r in range(world)]
    assert rows[0][1] == 0 and rows[-1][2] == T
    assert all(a[2] == b[1] for a, b in zip(rows, rows[1:]))           # contiguous, in rank order
    assert all(hi - lo <= rpo for rpo, lo, hi in rows)
    # the owner formula of the K2a epilogue (include/o3v.h): owner = min(row / rows_per_owner, P - 1)
    rpo = rows[0][0]
    for t in (0, T // 2, T - 1):
        owner = min(t // rpo, world - 1)
        assert rows[owner][1] <= t < rows[owner][2]


def _rs_worker(rank, world, port, T, H, out):
    """The reduce-scatter protocol of the vocab-parallel backward with the device parts replaced by torch: every rank
    holds a partial dHidden for ALL rows, 'stores' row r into slot `rank` of the owner's buffer (here: all_to_all of
    the owners' row ranges), the owner sums its slots in rank order."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from open_o3_video_b200.sharded import token_owner_rows
        g = torch.Generator().manual_seed(100 + rank)
        partial = torch.randn(T, H, generator=g).bfloat16()
        rpo = token_owner_rows(T, world, 0)[0]
        slot_rows = rpo
        send = []
        for owner in range(world):
            _, lo, hi = token_owner_rows(T, world, owner)
            blk = torch.zeros(slot_rows, H, dtype=torch.bfloat16)
            blk[: hi - lo] = partial[lo:hi]
            send.append(blk)
        slots = [torch.empty(slot_rows, H, dtype=torch.bfloat16) for _ in range(world)]
        dist.all_to_all(slots, send) if dist.get_backend() != "gloo" else [dist.gather(send[o], slots if rank == o else None, dst=o) for o in range(world)]
        _, lo, hi = token_owner_rows(T, world, rank)
        acc = torch.zeros(hi - lo, H)
        for p in range(world):                                           # fp32 in rank order, as o3v_sum_slots_bf16
            acc += slots[p][: hi - lo].float()
        mine = acc.bfloat16()
        full = partial.float().clone()
        dist.all_reduce(full)                                            # reference: sum over ranks of all rows
        err = (mine.float() - full[lo:hi].bfloat16().float()).abs().max().item()
        out.put((rank, err, hi - lo))
    finally:
        dist.destroy_process_group()


def test_reduce_scatter_protocol_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rs_worker, args=(r, 2, port, 101, 16, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert [g[2] for g in got] == [51, 50]
    assert all(g[1] <= 2 ** -7 for g in got)                             # summation order differs by one bf16 rounding at most
