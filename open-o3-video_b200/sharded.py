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
