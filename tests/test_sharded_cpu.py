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
