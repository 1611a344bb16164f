"""Fused lm_head -> log-softmax -> gather (K1) with its chunked fused backward (K2), and the
whole-step `fused_logprob_gspo`.

Replaces the reference's `_get_per_token_logps`
(src/r1-v/src/open_r1/trainer/grpo_trainer.py:371-384, whose `model(...).logits` ends in
transformers' lm_head = nn.Linear(H, V, bias=False)) and its autograd backward.  The
[tokens x vocab] logits are never materialised in the forward; the backward materialises
bf16 dlogits for one token chunk at a time ("backward").
"""
import ctypes
from typing import Optional

import torch

from . import _lib
from .gspo import _need_cuda, _p, _stream, gspo_raw

# token chunk of the fused fwd+bwd: bounds the bf16 dlogits buffer (chunk x V_local x 2 bytes)
DEFAULT_CHUNK_TOKENS = 32768
# chunk_tokens=0 sizes the chunk from this buffer budget instead (10 GB = 32768 tokens of a 152k
# vocabulary on one GPU; a vocab-sharded rank with V/8 columns takes the whole batch in one chunk)
CHUNK_BUFFER_BYTES = 10 << 30


def auto_chunk_tokens(v_local: int) -> int:
    return max(128, int(CHUNK_BUFFER_BYTES // (2 * v_local)))
# autograd path (`fused_logprob` + a separate backward): the forward may keep its bf16 logits for the backward
# (executed flops = algorithmic 6*T*H*V) or drop them and recompute per chunk (8*T*H*V, nothing of size [T, V]
# alive between forward and backward).  They are kept only when they fit BOTH this cap and a quarter of the
# memory that is free at forward time, so a 7B backbone's activations are never squeezed by the head.
SAVE_LOGITS_BYTES = 24 << 30
SAVE_LOGITS_FREE_FRACTION = 0.25


PLAN_CHUNKS = False     # measured on c2 (profiles/r2_ab_chunk_plan_c2.jsonl): no gain under the power cap, see plan_chunks


_plan_vocab_cache = {}


def _plan_vocab(V: int, group, dev) -> int:
    """Rows of the LARGEST vocabulary slice of the group (slices differ by one 256-row tile): the chunk size derived
    from it is the same on every rank.  One all-reduce per (group, V), then cached."""
    if group is None:
        return V
    key = (id(group), V)
    if key not in _plan_vocab_cache:
        import torch.distributed as dist
        t = torch.tensor([V], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=_pg(group))
        _plan_vocab_cache[key] = int(t.item())
    return _plan_vocab_cache[key]


def plan_chunks(N: int, Tc: int, H: int, V: int, chunk_tokens: int, dev=None, tune: bool = True):
    """Sequences per token chunk of the fused step.  Chunks hold whole sequences and at most `chunk_tokens` tokens; among
    the few ways to cut N sequences the one with the least WAVE QUANTISATION of the persistent grids is taken: K2a runs
    ceil(T_c / 256) x ceil(H / 512) tiles on #SM / 2 CTA pairs and K1 ceil(T_c / 128) x g items of ceil(tiles / g) tiles
    on #SM CTAs (g as chosen by the library, csrc/lmhead.cu:fwd_groups).  c2 on one GPU: 4 x 16 sequences = 128 x 7 =
    12.1 -> 13 waves of K2a (7 % idle); 17 + 17 + 17 + 13 sequences = 12.9 -> 13 and 9.8 -> 10 waves.
    OFF by default (PLAN_CHUNKS): measured on c2, K2a does get 3-5 % faster with the better cuts, but the step does not
    (300-306 ms for every cut): under the 1 kW cap the step is energy-bound, SMs that idle in a tail wave cost no energy
    and their power goes to the busy ones, so removing idle time only moves clock between kernels."""
    s_max = max(1, min(N, chunk_tokens // Tc))
    n_min = -(-N // s_max)
    even = -(-N // n_min)
    base = [even] * (N // even) + ([N % even] if N % even else [])
    if not (PLAN_CHUNKS and tune) or N == 1:
        return base
    sms = 148 if dev is None or not torch.cuda.is_available() else torch.cuda.get_device_properties(dev).multi_processor_count
    pairs, tiles = sms // 2, -(-V // 256)

    def cost(seqs):                                           # wave-quantised time of K1 + K2a, in units of one
        t = seqs * Tc                                         # 128 x 256 x H tile of MMA work
        mb, nt = -(-t // 256), -(-H // 512)
        k2a = -(-(mb * nt) // pairs) * tiles * 4 / 2          # waves x (256x512 pair tile over V) per SM
        m = -(-t // 128)
        k1 = min(-(-(m * g) // sms) * -(-tiles // g) for g in range(4, 17))
        g4 = -(-(m * 4) // sms) * -(-tiles // 4)
        if k1 * 100 > g4 * 96:
            k1 = g4                                           # the library keeps 4 groups unless another count saves 4 %
        return k1 + k2a

    best, best_cost = base, sum(cost(c) for c in base)
    for n_chunks in (n_min, n_min + 1):
        for first in range(max(1, s_max - 3), s_max + 1):     # k full-size chunks + one remainder
            k, rest = divmod(N, first)
            plan = [first] * k + ([rest] if rest else [])
            if len(plan) != n_chunks:
                continue
            c = sum(cost(x) for x in plan)
            if c * 100 < best_cost * 98:                      # only for a clear (>= 2 %) gain
                best, best_cost = plan, c
    return best


def save_logits_budget(device) -> int:
    free, _ = torch.cuda.mem_get_info(device)
    return int(min(SAVE_LOGITS_BYTES, free * SAVE_LOGITS_FREE_FRACTION))


_side_streams = {}


def _side_stream(dev):
    key = torch.device(dev).index
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=dev)
    return _side_streams[key]


def _check_head(hidden, weight, targets):
    _need_cuda(hidden, weight, targets)
    if hidden.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise TypeError("hidden and weight must be bf16 (got %s, %s)" % (hidden.dtype, weight.dtype))
    if hidden.dim() != 2 or weight.dim() != 2 or hidden.shape[1] != weight.shape[1]:
        raise ValueError("hidden [T, H], weight [V, H] expected")
    if targets.shape != (hidden.shape[0],):
        raise ValueError("targets must be [T]")


def lmhead_stats(hidden, weight, targets, v_offset=0, logits_out=None, stats_out=None, row_ref=None, row_keep=None):
    """K1 on one vocab slice -> stats [3, T] fp32 (row max, sum exp(z - max), target logit).
    `logits_out` ([T, ld] bf16, optional) also receives the bf16 logits or, with `row_ref` ([T] fp32), the
    exponentials exp(z - row_ref) (zeros for rows with `row_keep` == 0) that the exp-store backward consumes;
    `stats_out` (a contiguous [3, T] fp32 buffer, e.g. in peer-mapped memory) receives the triple instead of a
    new tensor."""
    T, H = hidden.shape
    V = weight.shape[0]
    lib = _lib.load()
    dev = hidden.device
    stats = torch.empty(3, T, dtype=torch.float32, device=dev) if stats_out is None else stats_out
    ws = torch.empty(int(lib.o3v_lmhead_fwd_workspace_bytes(T, V, H)), dtype=torch.uint8, device=dev)
    ld = 0 if logits_out is None else logits_out.stride(0)
    with torch.cuda.device(dev):
        if row_ref is None:
            _lib.call("o3v_lmhead_fwd" if logits_out is None else "o3v_lmhead_fwd+store", 2, lib.o3v_lmhead_fwd,
                      _p(hidden), _p(weight), _p(targets), T, V, H, int(v_offset), _p(stats),
                      _p(logits_out), ld, _p(ws), ws.numel(), _stream())
        else:
            _lib.call("o3v_lmhead_fwd+store", 2, lib.o3v_lmhead_fwd_exp,
                      _p(hidden), _p(weight), _p(targets), T, V, H, int(v_offset), _p(stats),
                      _p(logits_out), ld, _p(row_ref), _p(row_keep), _p(ws), ws.numel(), _stream())
    return stats


# ---- exp-store backward (include/o3v.h): K1 stores E = exp(z - row_ref), the softmax backward is folded into the K2a
# epilogue and the B operand of K2b by linearity: no elementwise pass over the [T, V] chunk.
ROW_REF_SAMPLE = 256       # vocabulary rows sampled (strided) for the per-row reference
ROW_REF_MARGIN = 16.0      # reference = sampled maximum + margin: overflow only if a logit exceeds the sample by > 100


def sample_weight(weight: torch.Tensor) -> torch.Tensor:
    """ROW_REF_SAMPLE rows of this rank's lm_head slice, strided over the slice (gathered once per step, 1.8 MB)."""
    V = weight.shape[0]
    stride = max(1, V // ROW_REF_SAMPLE)
    return weight[::stride][:ROW_REF_SAMPLE].contiguous()


def row_reference(hidden, w_sample, targets) -> torch.Tensor:
    """[T] fp32: max over the sampled vocabulary rows of hidden . w^T (K1 itself on a [256, H] weight) + margin."""
    return lmhead_stats(hidden, w_sample, targets, v_offset=1 << 40)[0] + ROW_REF_MARGIN


def softmax_rows(lse, grad_logp, targets, row_ref, v_offset, V):
    """-> (records [T, 4] int32 opaque, order [T] int64): per-token (g, a, target column) of the exp-store backward and
    the stable order of the tokens by target column that makes the one-hot scatter into dW deterministic."""
    T = lse.shape[0]
    rows = torch.empty(T, 4, dtype=torch.int32, device=lse.device)
    key = torch.empty(T, dtype=torch.int64, device=lse.device)
    with torch.cuda.device(lse.device):
        _lib.call("o3v_lmhead_softmax_rows", 1, _lib.load().o3v_lmhead_softmax_rows, _p(lse), _p(grad_logp), _p(targets),
                  _p(row_ref), int(v_offset), int(V), T, _p(rows), _p(key), _stream())
    return rows, torch.sort(key, stable=True).indices


def bwd_dhidden_exp(expz, rows, weight, out=None, fp32=False, scatter=None):
    """dH = a * (E . W) + g * W[tcol] (K2a with the softmax backward in its epilogue).  `scatter` =
    (PeerExchange, row0, tokens_total): tiles are stored at their token owners instead of `out`."""
    T, V = expz.shape
    H = weight.shape[1]
    lib = _lib.load()
    with torch.cuda.device(expz.device):
        if scatter is None:
            if out is None:
                out = torch.empty(T, H, dtype=torch.float32 if fp32 else torch.bfloat16, device=expz.device)
            _lib.call("o3v_lmhead_bwd_dhidden_exp", 1, lib.o3v_lmhead_bwd_dhidden_exp, _p(expz), expz.stride(0), _p(rows),
                      _p(weight), T, V, H, _p(out), 1 if out.dtype == torch.float32 else 0, None, 0, 0, 0, 0, 0, _stream())
            return out
        ex, row0, total = scatter
        arr = (ctypes.c_void_p * ex.world)(*ex._slot_ptrs)
        _lib.call("o3v_lmhead_bwd_dhidden_exp", 1, lib.o3v_lmhead_bwd_dhidden_exp, _p(expz), expz.stride(0), _p(rows),
                  _p(weight), T, V, H, None, 0, arr, ex.world, ex.rank, ex.owner_rows(total)[0], ex.slot_rows, int(row0),
                  _stream())
    return None


# K2b runs 256-row x 512-column tiles on a persistent grid of #SM / 2 CTA pairs.  When the vocabulary slice has one or
# two 256-row blocks more than a whole number of waves (19200 rows x 3584: 75 x 7 = 7 x 74 + 7 tiles), the last wave keeps
# 7 of 74 pairs busy for a full tile time: 12 % of the kernel at 8 GPUs.  The host layer then runs the main launch on
# the rows that fill whole waves and splits the K (token) range of the remaining rows over the idle pairs: S plain
# launches on S streams into S fp32 slabs, summed into dW in slab order (o3v_add_slabs_f32: deterministic).
DW_TAIL_SPLIT = True
_tail_streams = {}


def _dw_tail_plan(V: int, H: int, T: int, dev):
    """-> (rows of the main launch, K splits) or None when the tile count already fills its waves."""
    if not DW_TAIL_SPLIT:
        return None
    sms = torch.cuda.get_device_properties(dev).multi_processor_count if torch.cuda.is_available() else 148
    workers = sms // 2
    mb, nt = -(-V // 256), -(-H // 512)
    for tail in (1, 2):
        m_main = mb - tail
        if m_main >= 1 and (m_main * nt) % workers == 0 and 2 * tail * nt <= workers:
            splits = min(workers // (tail * nt), T // 4096)       # at least 64 k-blocks per split
            return (m_main * 256, splits) if splits >= 2 else None
    return None


def bwd_dweight_exp(expz, rows, order, hidden, d_weight, accumulate, scratch=None):
    """dW (+)= E^T . (a * hidden) + one-hot scatter (pre-scale, K2b, scatter: three launches; plus the K-split
    launches of the last vocabulary rows when they would otherwise run as a nearly empty wave, `_dw_tail_plan`)."""
    T, V = expz.shape
    H = hidden.shape[1]
    dev = expz.device
    if scratch is None or scratch.numel() < T * H:
        scratch = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
    lib = _lib.load()
    plan = _dw_tail_plan(V, H, T, dev)
    with torch.cuda.device(dev):
        if plan is None:
            _lib.call("o3v_lmhead_bwd_dweight_exp", 3, lib.o3v_lmhead_bwd_dweight_exp, _p(expz), expz.stride(0),
                      _p(rows), _p(order), _p(hidden), T, V, H, _p(d_weight), 1 if accumulate else 0, _p(scratch), _stream())
            return d_weight
        v_main, splits = plan
        if not accumulate:
            d_weight[v_main:].zero_()                      # the scatter and the slab sum below accumulate into these rows
        # rows [0, v_main): pre-scale of ALL token rows into `scratch`, main GEMM, one-hot scatter (every row of dW)
        _lib.call("o3v_lmhead_bwd_dweight_exp", 3, lib.o3v_lmhead_bwd_dweight_exp, _p(expz), expz.stride(0),
                  _p(rows), _p(order), _p(hidden), T, v_main, H, _p(d_weight), 1 if accumulate else 0, _p(scratch), _stream())
        main = torch.cuda.current_stream(dev)
        key = torch.device(dev).index
        pool = _tail_streams.setdefault(key, [])
        while len(pool) < splits:
            pool.append(torch.cuda.Stream(device=dev))
        ready = torch.cuda.Event()
        ready.record(main)
        slabs = torch.empty(splits, V - v_main, H, dtype=torch.float32, device=dev)
        scaled = scratch.view(-1)[: T * H].view(T, H)
        per = -(-T // splits)
        step = -(-per // 64) * 64                          # token ranges in whole 64-row k-blocks
        used = 0
        for i in range(splits):
            t0, t1 = i * step, min(T, (i + 1) * step)
            if t1 <= t0:
                break
            st = pool[i]
            with torch.cuda.stream(st):
                st.wait_event(ready)
                _lib.call("o3v_lmhead_bwd_dweight_tail", 1, lib.o3v_lmhead_bwd_dweight, _p(expz[t0:t1, v_main:]),
                          expz.stride(0), _p(scaled[t0:t1]), t1 - t0, V - v_main, H, _p(slabs[i]), 0,
                          ctypes.c_void_p(st.cuda_stream))
            main.wait_stream(st)
            used += 1
        _lib.call("o3v_add_slabs_f32", 1, lib.o3v_add_slabs_f32, _p(slabs), used, (V - v_main) * H, _p(d_weight[v_main:]),
                  _stream())
    return d_weight


def merge_stats(parts):
    """parts [P, 3, T] -> (logp [T], lse [T])."""
    P, _, T = parts.shape
    logp = torch.empty(T, dtype=torch.float32, device=parts.device)
    lse = torch.empty(T, dtype=torch.float32, device=parts.device)
    with torch.cuda.device(parts.device):
        _lib.call("o3v_lmhead_merge_stats", 1, _lib.load().o3v_lmhead_merge_stats, _p(parts), P, T, _p(logp),
                  _p(lse), _stream())
    return logp, lse


def _is_peer(group):
    return group is not None and hasattr(group, "merge") and hasattr(group, "local_stats")


def _pg(group):
    """The torch.distributed group behind `group` (a ProcessGroup or a sharded.PeerExchange)."""
    return group.group if _is_peer(group) else group


def _stats_to_logp(hidden, weight, targets, v_offset, logits_out, group, row_ref=None, row_keep=None):
    """K1 on this rank's vocab slice + the cross-rank merge -> (logp, lse) over the full vocabulary."""
    kw = dict(row_ref=row_ref, row_keep=row_keep)
    if _is_peer(group):
        T = hidden.shape[0]
        if T > group.cap:
            raise ValueError("PeerExchange capacity %d < %d tokens" % (group.cap, T))
        slot = group.next_slot()
        # the kernel writes rows of length T; the peer buffer has row stride cap: stage through a [3, T]
        # view only when T == cap, else write compactly and let merge use row_stride = cap
        buf = group.local_stats(slot)
        if T == group.cap:
            lmhead_stats(hidden, weight, targets, v_offset, logits_out, stats_out=buf, **kw)
        else:
            st = lmhead_stats(hidden, weight, targets, v_offset, logits_out, **kw)
            buf[:, :T].copy_(st)
        return group.merge(slot, T)
    stats = lmhead_stats(hidden, weight, targets, v_offset, logits_out, **kw)
    return merge_stats(_gather_stats(stats, group))


def _gather_stats(stats, group):
    """Vocab-parallel exchange: only the per-token (max, sum-exp, target-logit) triples cross
    NVLink (12 bytes per token per rank)."""
    if group is None:
        return stats.unsqueeze(0)
    import torch.distributed as dist
    world = dist.get_world_size(group)
    out = torch.empty((world * stats.shape[0],) + tuple(stats.shape[1:]), dtype=stats.dtype, device=stats.device)
    dist.all_gather_into_tensor(out, stats.contiguous(), group=group)     # rank-major concatenation along dim 0
    return out.view((world,) + tuple(stats.shape))


def dlogits_(logits, lse, grad_logp, targets, v_offset=0, V=None):
    """In place: logits[t, v] <- g[t] * (onehot - softmax)."""
    T = logits.shape[0]
    V = logits.shape[1] if V is None else V
    with torch.cuda.device(logits.device):
        _lib.call("o3v_lmhead_dlogits", 1, _lib.load().o3v_lmhead_dlogits, _p(logits), T, V, logits.stride(0),
                  _p(lse), _p(grad_logp), _p(targets), int(v_offset), _stream())
    return logits


# Softmax backward fused into the operand pipeline of the two backward GEMMs (o3v_lmhead_bwd_*_fused): the bf16
# logits chunk is consumed as stored by K1 and never rewritten by an elementwise pass.  Measured on c2 (round 2,
# profiles/r2_notes.md): correct but SLOWER under the 1 kW cap (every P element is re-exponentiated by each of the
# 7 + 7 CTAs columns that sweep it: 350 ms per step against 316 ms), so the default stays the separate in-place pass
# dlogits_ -> plain GEMMs.
FUSE_DLOGITS = False
# How the chunked backward forms P = g * (onehot - softmax):
#   "exp"      K1 stores exp(z - row_ref); rescale + one-hot folded into the K2a epilogue and the K2b operands
#              (o3v_lmhead_*_exp, default: no pass over the [T, V] chunk between K1 and the two GEMMs)
#   "dlogits"  K1 stores z, an elementwise in-place pass turns it into P (round 1), or FUSE_DLOGITS rewrites the
#              A tiles in shared memory inside the GEMMs
BACKWARD = "exp"


def softmax_bwd_rows(lse, grad_logp, targets, v_offset, V):
    """Per-token records of the fused softmax backward (o3v_lmhead_softmax_bwd_rows) -> opaque [T, 4] int32 tensor."""
    T = lse.shape[0]
    rows = torch.empty(T, 4, dtype=torch.int32, device=lse.device)
    with torch.cuda.device(lse.device):
        _lib.call("o3v_lmhead_softmax_bwd_rows", 1, _lib.load().o3v_lmhead_softmax_bwd_rows, _p(lse), _p(grad_logp),
                  _p(targets), int(v_offset), int(V), T, _p(rows), _stream())
    return rows


def _rows_of(softmax_bwd, V):
    """`softmax_bwd` is either the records themselves or (lse, grad_logp, targets, v_offset)."""
    if isinstance(softmax_bwd, torch.Tensor):
        return softmax_bwd
    lse, g, tgt, v_off = softmax_bwd
    return softmax_bwd_rows(lse.contiguous(), g.contiguous(), tgt.contiguous(), v_off, V)


def bwd_dhidden(dlogits, weight, out=None, fp32=False, softmax_bwd=None):
    """dH = P . W.  `softmax_bwd` = (lse, grad_logp, targets, v_offset) or the records `softmax_bwd_rows` made of
    them: `dlogits` then holds the raw bf16 LOGITS and P is formed inside the GEMM's operand pipeline."""
    T, V = dlogits.shape
    H = weight.shape[1]
    if out is None:
        out = torch.empty(T, H, dtype=torch.float32 if fp32 else torch.bfloat16, device=dlogits.device)
    lib = _lib.load()
    with torch.cuda.device(dlogits.device):
        if softmax_bwd is None:
            _lib.call("o3v_lmhead_bwd_dhidden", 1, lib.o3v_lmhead_bwd_dhidden, _p(dlogits), dlogits.stride(0),
                      _p(weight), T, V, H, _p(out), 1 if out.dtype == torch.float32 else 0, _stream())
        else:
            rows = _rows_of(softmax_bwd, V)
            _lib.call("o3v_lmhead_bwd_dhidden_fused", 1, lib.o3v_lmhead_bwd_dhidden_fused, _p(dlogits),
                      dlogits.stride(0), _p(rows), _p(weight), T, V, H, _p(out),
                      1 if out.dtype == torch.float32 else 0, _stream())
    return out


def bwd_dweight(dlogits, hidden, d_weight, accumulate, softmax_bwd=None):
    """dW (+)= P^T . hidden; `softmax_bwd` as in bwd_dhidden."""
    T, V = dlogits.shape
    H = hidden.shape[1]
    lib = _lib.load()
    with torch.cuda.device(dlogits.device):
        if softmax_bwd is None:
            _lib.call("o3v_lmhead_bwd_dweight", 1, lib.o3v_lmhead_bwd_dweight, _p(dlogits), dlogits.stride(0),
                      _p(hidden), T, V, H, _p(d_weight), 1 if accumulate else 0, _stream())
        else:
            rows = _rows_of(softmax_bwd, V)
            _lib.call("o3v_lmhead_bwd_dweight_fused", 1, lib.o3v_lmhead_bwd_dweight_fused, _p(dlogits),
                      dlogits.stride(0), _p(rows), _p(hidden), T, V, H, _p(d_weight),
                      1 if accumulate else 0, _stream())
    return d_weight


class _FusedLogprobFn(torch.autograd.Function):
    """logp[t] = log_softmax(hidden[t] . W^T)[targets[t]] with a chunked fused backward."""

    @staticmethod
    def forward(ctx, hidden, weight, targets, v_offset, group, chunk_tokens):
        T, H = hidden.shape
        V = weight.shape[0]
        need_grad = hidden.requires_grad or weight.requires_grad
        keep = need_grad and (T * V * 2 <= save_logits_budget(hidden.device))
        logits = torch.empty(T, V, dtype=torch.bfloat16, device=hidden.device) if keep else None
        ctx.mode = BACKWARD
        ctx.row_ref = None
        if keep and ctx.mode == "exp":
            ctx.row_ref = row_reference(hidden, sample_weight(weight), targets)
        logp, lse = _stats_to_logp(hidden, weight, targets, v_offset, logits, group, row_ref=ctx.row_ref)
        ctx.save_for_backward(hidden, weight, targets, lse)
        ctx.logits, ctx.v_offset, ctx.group, ctx.chunk_tokens = logits, v_offset, group, chunk_tokens
        ctx.mark_non_differentiable(lse)
        return logp, lse

    @staticmethod
    def backward(ctx, g_logp, _g_lse):
        hidden, weight, targets, lse = ctx.saved_tensors
        T, H = hidden.shape
        V = weight.shape[0]
        g = g_logp.to(torch.float32).contiguous()
        d_hidden = torch.empty(T, H, dtype=torch.bfloat16, device=hidden.device)
        d_weight = torch.empty(V, H, dtype=torch.float32, device=hidden.device)
        chunk = min(T, ctx.chunk_tokens)
        zbuf = None if ctx.logits is not None else torch.empty(chunk, V, dtype=torch.bfloat16, device=hidden.device)
        first = True
        exp_mode = ctx.mode == "exp"
        w_sample = sample_weight(weight) if (exp_mode and ctx.logits is None) else None
        scratch = torch.empty(chunk, H, dtype=torch.bfloat16, device=hidden.device) if exp_mode else None
        for s in range(0, T, chunk):
            e = min(T, s + chunk)
            r = None if ctx.row_ref is None else ctx.row_ref[s:e]
            if ctx.logits is not None:
                z = ctx.logits[s:e]
            else:                                   # recompute this chunk's logits (forward kept nothing)
                z = zbuf[: e - s]
                if exp_mode:
                    r = row_reference(hidden[s:e], w_sample, targets[s:e])
                lmhead_stats(hidden[s:e], weight, targets[s:e], ctx.v_offset, z, row_ref=r)
            if exp_mode:
                rows, order = softmax_rows(lse[s:e], g[s:e], targets[s:e], r, ctx.v_offset, V)
                bwd_dhidden_exp(z, rows, weight, out=d_hidden[s:e])
                bwd_dweight_exp(z, rows, order, hidden[s:e], d_weight, accumulate=not first, scratch=scratch)
                first = False
                continue
            sb = None
            if FUSE_DLOGITS:
                sb = softmax_bwd_rows(lse[s:e], g[s:e], targets[s:e], ctx.v_offset, V)
            else:
                dlogits_(z, lse[s:e], g[s:e], targets[s:e], ctx.v_offset)
            bwd_dhidden(z, weight, out=d_hidden[s:e], softmax_bwd=sb)
            bwd_dweight(z, hidden[s:e], d_weight, accumulate=not first, softmax_bwd=sb)
            first = False
        ctx.logits = None
        if ctx.group is not None:                   # partial sums over vocab slices
            import torch.distributed as dist
            dist.all_reduce(d_hidden, group=_pg(ctx.group))
        return d_hidden, d_weight.to(weight.dtype), None, None, None, None


def fused_logprob(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor, *, v_offset: int = 0,
                  group=None, chunk_tokens: int = DEFAULT_CHUNK_TOKENS, return_lse: bool = False):
    """hidden [T, H] bf16, weight [V, H] bf16 (rows v_offset.. of lm_head.weight when vocab
    sharded over `group`), targets [T] int64 global ids -> logp [T] fp32 (autograd-enabled)."""
    hidden = hidden.contiguous()
    weight = weight.contiguous()
    targets = targets.to(torch.int64).contiguous()
    _check_head(hidden, weight, targets)
    logp, lse = _FusedLogprobFn.apply(hidden, weight, targets, int(v_offset), group, int(chunk_tokens))
    return (logp, lse) if return_lse else logp


def per_token_logps(hidden: torch.Tensor, weight: torch.Tensor, input_ids: torch.Tensor, **kw) -> torch.Tensor:
    """The reference's contract (grpo_trainer.py:371-384): hidden [B, L, H] (final hidden
    states of `input_ids` [B, L]) -> [B, L-1]: log p(input_ids[:, t+1] | ..t)."""
    B, L, H = hidden.shape
    h = hidden[:, :-1, :].reshape(B * (L - 1), H)          # :376 drop the last position
    tgt = input_ids[:, 1:].reshape(B * (L - 1))            # :377 drop the first id
    return fused_logprob(h, weight, tgt, **kw).view(B, L - 1)


def sft_cross_entropy(hidden: torch.Tensor, weight: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100,
                      num_items_in_batch=None, **kw) -> torch.Tensor:
    """Causal-LM cross-entropy of the SFT stage on the same fused head (the loss the reference's
    `MySFTTrainer.compute_loss`, sft_multi_task.py:402-409, gets from the HF model for the labels
    built at :387-398): shift by one, ignore `ignore_index`, mean over the remaining targets (or
    sum / num_items_in_batch).  hidden [B, L, H] bf16, labels [B, L].  Ignored positions are
    compacted away before the GEMM, so pad / visual tokens cost nothing; autograd reaches hidden
    and weight through `fused_logprob`."""
    B, L, H = hidden.shape
    tgt = labels[:, 1:].reshape(-1)
    keep = (tgt != ignore_index).nonzero(as_tuple=True)[0]
    h = hidden[:, :-1, :].reshape(B * (L - 1), H).index_select(0, keep)
    if keep.numel() == 0:
        return (h.sum() * 0.0).to(torch.float32)
    logp = fused_logprob(h, weight, tgt.index_select(0, keep), **kw)
    denom = keep.numel() if num_items_in_batch is None else num_items_in_batch
    return -logp.sum() / denom


def fused_logprob_gspo(hidden: torch.Tensor, weight: torch.Tensor, completion_ids: torch.Tensor,
                       ref_per_token_logps: torch.Tensor, completion_mask: torch.Tensor,
                       rewards_per_func: torch.Tensor, num_generations: int, beta: float,
                       epsilon_low: float = 0.2, epsilon_high: float = 0.2, gspo: bool = True,
                       old_per_token_logps: Optional[torch.Tensor] = None, *, v_offset: int = 0, group=None,
                       chunk_tokens: int = DEFAULT_CHUNK_TOKENS, need_grad: bool = True,
                       d_weight_out: Optional[torch.Tensor] = None, overlap_dlogits: bool = False,
                       fuse_dlogits: Optional[bool] = None, backward: Optional[str] = None):
    """The whole policy-objective step in one call: per-token log-probs, KL, group advantages,
    GSPO loss AND the gradients w.r.t. hidden and lm_head.weight (grpo_trainer.py:612-613,
    635-636, 658-706 + their backward), chunked over whole sequences.

    hidden [N, Tc, H] bf16: final hidden states at the positions that PREDICT each completion
    token; completion_ids / ref / mask / old: [N, Tc]; rewards_per_func [N, F].
    Returns dict(loss, per_token_logps, advantages, mean_kl, completion_length, reward_std,
    d_hidden [N, Tc, H] bf16, d_weight [V, H] fp32).  With `group=` a `sharded.PeerExchange(dh_mode="reduce_scatter" / "reduce_scatter_fused")`
    d_hidden is [hi - lo, H]: the token rows `d_hidden_rows = (lo, hi)` this rank owns (data-parallel layout).

    Per chunk (`backward="exp"`, the default `BACKWARD`): row reference (K1 on 256 sampled vocabulary rows) -> K1 storing
    exp(z - ref) -> merge -> K3 (loss, d loss / d logp) -> per-token records + sort -> K2a with the softmax backward in
    its epilogue -> pre-scaled hidden, K2b, one-hot scatter.  `backward="dlogits"`: K1 stores z, then either the in-place
    elementwise pass (round 1) or, with `fuse_dlogits=True`, the shared-memory transform inside the GEMMs.
    """
    fuse = FUSE_DLOGITS if fuse_dlogits is None else bool(fuse_dlogits)
    mode = BACKWARD if (backward is None and fuse_dlogits is None) else (backward or "dlogits")
    if mode not in ("exp", "dlogits"):
        raise ValueError("backward must be 'exp' or 'dlogits'")
    exp_mode = need_grad and mode == "exp"
    if exp_mode:
        fuse = False
    N, Tc, H = hidden.shape
    V = weight.shape[0]
    dev = hidden.device
    hidden2 = hidden.reshape(N * Tc, H).contiguous()
    weight = weight.contiguous()
    targets = completion_ids.to(torch.int64).reshape(N * Tc).contiguous()
    _check_head(hidden2, weight, targets)
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    ref, old, rpf = f32(ref_per_token_logps), f32(old_per_token_logps), f32(rewards_per_func)
    mask = completion_mask.to(torch.int32).contiguous()

    if not chunk_tokens:
        chunk_tokens = auto_chunk_tokens(_plan_vocab(V, group, dev))
    # sequences per chunk (whole sequences only); vocab-sharded ranks must all cut the step the same way
    seq_plan = plan_chunks(N, Tc, H, V, chunk_tokens, dev, tune=group is None)
    seqs = max(seq_plan)
    n_chunks = len(seq_plan)
    logp = torch.empty(N, Tc, dtype=torch.float32, device=dev)
    zbuf = torch.empty(seqs * Tc, V, dtype=torch.bfloat16, device=dev) if need_grad else None
    d_hidden = None
    peer_dh = False
    dh_mode = getattr(group, "dh_mode", "") if (need_grad and _is_peer(group) and group.hidden_size == H) else ""
    scatter_dh = dh_mode == "reduce_scatter_fused"          # K2a epilogue stores tiles at their owners
    pull_dh = dh_mode == "reduce_scatter"                   # owners pull their rows beside the dW GEMM
    if scatter_dh and fuse_dlogits:
        raise ValueError("fuse_dlogits and the reduce-scatter K2a cannot be combined")
    if need_grad and not scatter_dh:
        d_hidden = group.dh_view(N * Tc, H) if _is_peer(group) else None      # peer-mapped: overlapped all-reduce
        peer_dh = d_hidden is not None
        if d_hidden is None:
            d_hidden = torch.empty(N * Tc, H, dtype=torch.bfloat16, device=dev)
    d_weight = None
    if need_grad:
        d_weight = d_weight_out if d_weight_out is not None else torch.empty(V, H, dtype=torch.float32, device=dev)
    state = {}
    # Software pipeline over chunks (overlap_dlogits): the HBM-bound dlogits pass of chunk c runs on a side
    # stream beside the tensor-bound K1 of chunk c+1 (its CTAs use no shared memory and co-reside with the
    # persistent GEMM CTAs), then the two backward GEMMs of chunk c follow on the main stream.  Needs a second
    # logits buffer.  Order of work on the device:  F0 | F1 + D0 | B0 | F2 + D1 | B1 | ... | D(n-1) | B(n-1).
    pipelined = bool(need_grad and overlap_dlogits and n_chunks > 1 and not fuse)
    zbuf2 = torch.empty_like(zbuf) if pipelined else None
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if pipelined else None

    dh_owned = [None]
    w_sample = sample_weight(weight) if exp_mode else None
    scratch = torch.empty(seqs * Tc, H, dtype=torch.bfloat16, device=dev) if exp_mode else None
    mask_flat = mask.view(-1)

    def backward_gemms_exp(ci, s, e, z, lse, g, r):
        rows, order = softmax_rows(lse, g, targets[s:e], r, v_offset, V)
        if scatter_dh:
            bwd_dhidden_exp(z, rows, weight, scatter=(group, s, N * Tc))     # epilogue stores each tile at its owner
            if e == N * Tc:
                dh_owned[0] = group.dhidden_reduce_async(N * Tc)
        else:
            bwd_dhidden_exp(z, rows, weight, out=d_hidden[s:e])
        if pull_dh:
            dh_owned[0] = group.reduce_scatter_dh_async(s, e - s, N * Tc)
        elif peer_dh:
            group.allreduce_dh_async(s, e - s)
        bwd_dweight_exp(z, rows, order, hidden2[s:e], d_weight, (ci > 0) or (d_weight_out is not None), scratch)

    def backward_gemms(ci, s, e, z, sb=None):
        if scatter_dh:
            group.dhidden_scatter(z, weight, s, N * Tc)       # K2a whose epilogue stores each tile at its owner (NVLink)
            if e == N * Tc:
                dh_owned[0] = group.dhidden_reduce_async(N * Tc)   # barrier + local slot sum beside the dW GEMM below
        else:
            bwd_dhidden(z, weight, out=d_hidden[s:e], softmax_bwd=sb)
        if pull_dh:
            dh_owned[0] = group.reduce_scatter_dh_async(s, e - s, N * Tc)   # runs beside the dW GEMM below
        elif peer_dh:
            group.allreduce_dh_async(s, e - s)                # runs beside the dW GEMM below
        bwd_dweight(z, hidden2[s:e], d_weight, accumulate=(ci > 0) or (d_weight_out is not None), softmax_bwd=sb)

    pending = None                                            # chunk whose backward GEMMs are still to be enqueued
    starts = [sum(seq_plan[:i]) for i in range(n_chunks)]
    for ci, n0 in enumerate(starts):
        n1 = n0 + seq_plan[ci]
        s, e = n0 * Tc, n1 * Tc
        z = None
        if need_grad:
            z = (zbuf2 if (pipelined and (ci & 1)) else zbuf)[: e - s]
        r = row_reference(hidden2[s:e], w_sample, targets[s:e]) if exp_mode else None
        lp, lse = _stats_to_logp(hidden2[s:e], weight, targets[s:e], v_offset, z, group, row_ref=r,
                                 row_keep=mask_flat[s:e] if exp_mode else None)
        logp[n0:n1] = lp.view(n1 - n0, Tc)
        state, g, _ = gspo_raw(logp[n0:n1], ref[n0:n1], mask[n0:n1], rpf, num_generations, beta, epsilon_low,
                               epsilon_high, gspo, None if old is None else old[n0:n1], want_grad=need_grad,
                               want_kl=False, N_total=N, seq_offset=n0, state=state)
        if not need_grad:
            continue
        if exp_mode:
            backward_gemms_exp(ci, s, e, z, lse, g.view(-1), r)
            continue
        if fuse:
            backward_gemms(ci, s, e, z, softmax_bwd_rows(lse, g.view(-1), targets[s:e], v_offset, V))
            continue
        if not pipelined:
            dlogits_(z, lse, g.view(-1), targets[s:e], v_offset)
            backward_gemms(ci, s, e, z)
            continue
        if pending is not None:                               # B(c-1): its dlogits ran beside this chunk's K1
            main.wait_event(pending[-1])
            backward_gemms(*pending[:4])
        last = ci == n_chunks - 1
        done = torch.cuda.Event()
        if last:                                              # nothing left to hide it behind
            dlogits_(z, lse, g.view(-1), targets[s:e], v_offset)
            done.record(main)
        else:
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                dlogits_(z, lse, g.view(-1), targets[s:e], v_offset)
                done.record(side)
        pending = (ci, s, e, z, lse, g, done)                 # lse / g stay referenced until the side stream is done
    if pending is not None:
        main.wait_event(pending[-1])
        backward_gemms(*pending[:4])
    rows = None
    if scatter_dh or pull_dh:
        group.wait_allreduce()
        d_hidden = dh_owned[0]
        rows = group.owner_rows(N * Tc)[1:]
    elif need_grad and peer_dh:
        group.wait_allreduce()
    elif need_grad and group is not None:
        import torch.distributed as dist
        dist.all_reduce(d_hidden, group=_pg(group))
    return dict(loss=state["loss"].reshape(()), per_token_logps=logp, advantages=state["adv"],
                mean_kl=state["mean_kl"].reshape(()), completion_length=state["clen"], reward_std=state["rstd"],
                d_hidden=None if d_hidden is None else (d_hidden if (scatter_dh or pull_dh) else d_hidden.view(N, Tc, H)),
                d_hidden_rows=rows, d_weight=d_weight)


class _FusedPolicyStepFn(torch.autograd.Function):
    """`fused_logprob_gspo` attached to autograd at (hidden, weight): the forward runs the whole chunked step and
    already holds dHidden / dW; the backward only scales them by the incoming gradient of the loss."""

    @staticmethod
    def forward(ctx, hidden, weight, completion_ids, ref, mask, rpf, G, beta, eps_lo, eps_hi, gspo, old, opts):
        need = hidden.requires_grad or weight.requires_grad
        out = fused_logprob_gspo(hidden.detach(), weight.detach(), completion_ids, ref, mask, rpf, G, beta, eps_lo,
                                 eps_hi, gspo, old, need_grad=need, **opts)
        ctx.need = need
        if need:
            ctx.save_for_backward(out["d_hidden"], out["d_weight"])
        ctx.w_dtype = weight.dtype
        extras = (out["per_token_logps"], out["advantages"], out["mean_kl"], out["completion_length"],
                  out["reward_std"])
        ctx.mark_non_differentiable(*extras)
        return (out["loss"].clone(),) + extras

    @staticmethod
    def backward(ctx, g_loss, *_):
        if not ctx.need:
            return (None,) * 13
        d_hidden, d_weight = ctx.saved_tensors
        g = g_loss.to(torch.float32)
        dh = d_hidden if _is_one(g) else (d_hidden.float() * g).to(d_hidden.dtype)
        dw = (d_weight if _is_one(g) else d_weight * g).to(ctx.w_dtype)
        return (dh, dw) + (None,) * 11


def _is_one(g) -> bool:
    # gradient accumulation scales the loss by 1/steps: only skip the multiply when it is known to be a no-op
    # WITHOUT reading the device (no sync)
    return False


def fused_policy_step(hidden, weight, completion_ids, ref_per_token_logps, completion_mask, rewards_per_func,
                      num_generations, beta, epsilon_low=0.2, epsilon_high=0.2, gspo=True,
                      old_per_token_logps=None, **opts):
    """Autograd-enabled `fused_logprob_gspo`: returns the same dict (without d_hidden / d_weight); `loss.backward()`
    delivers dHidden to whatever produced `hidden` (the backbone) and dW to `weight` (lm_head.weight)."""
    res = _FusedPolicyStepFn.apply(hidden, weight, completion_ids, ref_per_token_logps, completion_mask,
                                   rewards_per_func, int(num_generations), float(beta), float(epsilon_low),
                                   float(epsilon_high), bool(gspo), old_per_token_logps, opts)
    keys = ("loss", "per_token_logps", "advantages", "mean_kl", "completion_length", "reward_std")
    return dict(zip(keys, res))
