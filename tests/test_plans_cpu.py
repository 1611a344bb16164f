"""Host-side scheduling arithmetic of the fused step (no GPU): how the K2b tail rows are split over the idle CTA pairs
(logprob._dw_tail_plan) and how a step is cut into token chunks (logprob.plan_chunks)."""
from open_o3_video_b200 import logprob, sharded


def test_dw_tail_plan_on_the_benchmarked_shapes():
    plan = lambda V, H, T: logprob._dw_tail_plan(V, H, T, "cuda")
    # whole waves of the 74 CTA pairs already: one launch
    assert plan(18944, 3584, 131072) is None and plan(37888, 4096, 131072) is None
    assert plan(9496, 1792, 2900) is None
    # 75 x 7 = 7 x 74 + 7 (one rank of c2 at 8 GPUs), 149 x 8 = 16 x 74 + 8 (c3 at 4 GPUs): last block K-split
    assert plan(19200, 3584, 131072) == (74 * 256, 10)
    assert plan(38144, 4096, 131072) == (148 * 256, 9)
    # the full heads on one GPU: 594 blocks = 592 + 2 (the ragged last block of the 8B head included)
    assert plan(152064, 3584, 32768) == (592 * 256, 5)
    assert plan(151936, 4096, 32768) == (592 * 256, 4)
    # short K ranges are not split (at least 4096 tokens per split)
    assert plan(19200, 3584, 4096) is None and plan(19200, 3584, 8192) == (74 * 256, 2)
    # every real vocabulary slice gets a plan or fills its waves
    for V in (152064, 151936):
        for world in (2, 4, 8):
            for v0, v1 in sharded.vocab_slices(V, world):
                p = plan(v1 - v0, 4096, 65536)
                assert p is None or (0 < p[0] < v1 - v0 and p[0] % 256 == 0 and 2 <= p[1] <= 16)


def test_plan_chunks_whole_sequences_within_the_budget(monkeypatch):
    for N, Tc, H, V in ((64, 2048, 3584, 152064), (128, 4096, 4096, 151936), (16, 16384, 4096, 151936), (4, 512, 3584, 152064),
                        (7, 3000, 1024, 50000), (1, 100000, 3584, 152064)):
        budget = logprob.auto_chunk_tokens(V)
        for tune in (False, True):
            monkeypatch.setattr(logprob, "PLAN_CHUNKS", tune)
            plan = logprob.plan_chunks(N, Tc, H, V, budget)
            assert sum(plan) == N and all(p >= 1 for p in plan)
            assert all(p * Tc <= max(budget, Tc) for p in plan)          # (a single over-long sequence is its own chunk)
    monkeypatch.setattr(logprob, "PLAN_CHUNKS", False)
    assert logprob.plan_chunks(64, 2048, 3584, 152064, logprob.auto_chunk_tokens(152064)) == [16] * 4
    assert logprob.plan_chunks(128, 4096, 4096, 19200, logprob.auto_chunk_tokens(19200)) == [64, 64]
    monkeypatch.setattr(logprob, "PLAN_CHUNKS", True)
    # the cut with the least wave quantisation of K1 / K2a (measured: no gain under the power cap, hence off by default)
    assert logprob.plan_chunks(64, 2048, 3584, 152064, logprob.auto_chunk_tokens(152064)) == [17, 17, 17, 13]
    # vocab-sharded ranks never tune (every rank must cut the step the same way)
    assert logprob.plan_chunks(64, 2048, 3584, 152064, logprob.auto_chunk_tokens(152064), tune=False) == [16] * 4


def test_reduce_scatter_rows_cover_every_token_once():
    """dh_mode="reduce_scatter": after K2a of the chunk [row0, row0 + rows) every owner pulls the intersection of ITS rows
    with the chunk (sharded.PeerExchange.reduce_scatter_dh_async); over the chunks of a step and the ranks of the group
    every token row must be pulled exactly once, into the right place of the owner's [hi - lo, H] result."""
    for T, world, chunk in ((131072, 8, 131072), (524288, 8, 262144), (524288, 4, 131072), (101, 2, 40), (7, 8, 3), (1000, 3, 999)):
        seen = [0] * T
        for rank in range(world):
            rpo, lo, hi = sharded.token_owner_rows(T, world, rank)
            assert rpo == -(-T // world) and 0 <= lo <= hi <= T
            filled = [0] * (hi - lo)
            for row0 in range(0, T, chunk):
                rows = min(chunk, T - row0)
                a, b = max(lo, row0), min(hi, row0 + rows)          # as in reduce_scatter_dh_async
                for t in range(a, b):
                    seen[t] += 1
                    filled[t - lo] += 1
            assert all(f == 1 for f in filled)
        assert all(s == 1 for s in seen)
