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
