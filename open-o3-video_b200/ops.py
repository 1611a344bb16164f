"""`torch.ops.o3v.*`: the C-ABI entry points as registered PyTorch operators (SURVEY.md 8b,
"torch-level ops to expose").

The functions in logprob.py / gspo.py / rewards.py are what the trainer drop-in calls; this module
registers the same launches with `torch.library` so that they have a schema, a fake (meta)
implementation and an autograd formula, i.e. they can be traced by `torch.compile` / `torch.export`
around the backbone instead of breaking the graph.  Nothing here computes anything itself: every
operator body is a call into libo3v.so through the wrappers of this package.

  o3v::lmhead_logprob(hidden[T,H] bf16, weight[V,H] bf16, targets[T] i64, v_offset, keep_logits)
        -> (logp[T] f32, lse[T] f32, logits[T,V] bf16 or [0,V])                      K1 + merge
  o3v::lmhead_logprob_backward(grad_logp, hidden, weight, targets, lse, logits, v_offset, chunk_tokens)
        -> (d_hidden[T,H] bf16, d_weight[V,H] f32)                                   dlogits + K2a + K2b
        (the operators keep the bf16 logits and run the in-place softmax-backward pass: the "dlogits" mode of
        logprob.BACKWARD; the wrappers default to the exp-store mode, same parity bar)
  o3v::eos_mask(completion_ids[N,Tc] i64, eos_id) -> (eos_idx[N] i64, mask[N,Tc] i32)   K3a
  o3v::gspo_objective(logp, ref, mask, rewards_per_func, old?, G, beta, eps_lo, eps_hi, gspo)
        -> (loss[], grad_logp[N,Tc], advantages[N], mean_kl[], completion_len[N] i32, reward_std[N])   K3
  o3v::policy_step(hidden[N,Tc,H], weight, completion_ids, ref, mask, rewards_per_func, old?, G, beta,
                   eps_lo, eps_hi, gspo, chunk_tokens)
        -> (loss[], logp[N,Tc], advantages[N], mean_kl[], d_hidden[N,Tc,H] bf16, d_weight[V,H] f32)   whole step
`lmhead_logprob` has an autograd formula (its backward is `lmhead_logprob_backward`); `gspo_objective`
and `policy_step` return their gradients as outputs.
"""
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import gspo as _gspo
from . import logprob as _logprob

__all__ = ["lmhead_logprob", "lmhead_logprob_backward", "eos_mask", "gspo_objective", "policy_step", "fused_logprob"]


@torch.library.custom_op("o3v::lmhead_logprob", mutates_args=())
def lmhead_logprob(hidden: Tensor, weight: Tensor, targets: Tensor, v_offset: int, keep_logits: bool
                   ) -> Tuple[Tensor, Tensor, Tensor]:
    hidden, weight, targets = hidden.contiguous(), weight.contiguous(), targets.to(torch.int64).contiguous()
    _logprob._check_head(hidden, weight, targets)
    T, V = hidden.shape[0], weight.shape[0]
    logits = torch.empty(T if keep_logits else 0, V, dtype=torch.bfloat16, device=hidden.device)
    logp, lse = _logprob._stats_to_logp(hidden, weight, targets, v_offset, logits if keep_logits else None, None)
    return logp, lse, logits


@lmhead_logprob.register_fake
def _(hidden, weight, targets, v_offset, keep_logits):
    T, V = hidden.shape[0], weight.shape[0]
    f = lambda *s, dt=torch.float32: hidden.new_empty(s, dtype=dt)
    return f(T), f(T), f(T if keep_logits else 0, V, dt=torch.bfloat16)


@torch.library.custom_op("o3v::lmhead_logprob_backward", mutates_args=())
def lmhead_logprob_backward(grad_logp: Tensor, hidden: Tensor, weight: Tensor, targets: Tensor, lse: Tensor,
                            logits: Tensor, v_offset: int, chunk_tokens: int) -> Tuple[Tensor, Tensor]:
    hidden, weight, targets = hidden.contiguous(), weight.contiguous(), targets.to(torch.int64).contiguous()
    T, H = hidden.shape
    V = weight.shape[0]
    g = grad_logp.to(torch.float32).contiguous()
    d_hidden = torch.empty(T, H, dtype=torch.bfloat16, device=hidden.device)
    d_weight = torch.empty(V, H, dtype=torch.float32, device=hidden.device)
    have = logits.shape[0] == T
    chunk = max(1, min(T, chunk_tokens))
    zbuf = None if have else torch.empty(chunk, V, dtype=torch.bfloat16, device=hidden.device)
    for i, s in enumerate(range(0, T, chunk)):
        e = min(T, s + chunk)
        if have:
            z = logits[s:e].clone()                          # operators must not mutate their inputs
        else:                                                # the forward kept nothing: recompute this chunk
            z = zbuf[: e - s]
            _logprob.lmhead_stats(hidden[s:e], weight, targets[s:e], v_offset, z)
        _logprob.dlogits_(z, lse[s:e], g[s:e], targets[s:e], v_offset)
        _logprob.bwd_dhidden(z, weight, out=d_hidden[s:e])
        _logprob.bwd_dweight(z, hidden[s:e], d_weight, accumulate=i > 0)
    return d_hidden, d_weight


@lmhead_logprob_backward.register_fake
def _(grad_logp, hidden, weight, targets, lse, logits, v_offset, chunk_tokens):
    return (hidden.new_empty(hidden.shape, dtype=torch.bfloat16),
            hidden.new_empty(weight.shape, dtype=torch.float32))


def _lmhead_setup(ctx, inputs, output):
    hidden, weight, targets, v_offset, keep_logits = inputs
    _, lse, logits = output
    ctx.save_for_backward(hidden, weight, targets, lse, logits)
    ctx.v_offset = v_offset


def _lmhead_backward(ctx, g_logp, _g_lse, _g_logits):
    hidden, weight, targets, lse, logits = ctx.saved_tensors
    d_hidden, d_weight = lmhead_logprob_backward(g_logp, hidden, weight, targets, lse, logits, ctx.v_offset,
                                                 _logprob.DEFAULT_CHUNK_TOKENS)
    return d_hidden, d_weight.to(weight.dtype), None, None, None


lmhead_logprob.register_autograd(_lmhead_backward, setup_context=_lmhead_setup)


def fused_logprob(hidden: Tensor, weight: Tensor, targets: Tensor, v_offset: int = 0) -> Tensor:
    """Same contract as logprob.fused_logprob (single GPU), built from the registered operators."""
    T, V = hidden.shape[0], weight.shape[0]
    keep = hidden.requires_grad or weight.requires_grad
    keep = keep and T * V * 2 <= _logprob.SAVE_LOGITS_BYTES
    return lmhead_logprob(hidden, weight, targets, v_offset, keep)[0]


@torch.library.custom_op("o3v::eos_mask", mutates_args=())
def eos_mask(completion_ids: Tensor, eos_id: int) -> Tuple[Tensor, Tensor]:
    return _gspo.eos_mask(completion_ids, eos_id)


@eos_mask.register_fake
def _(completion_ids, eos_id):
    N, Tc = completion_ids.shape
    return completion_ids.new_empty((N,), dtype=torch.int64), completion_ids.new_empty((N, Tc), dtype=torch.int32)


@torch.library.custom_op("o3v::gspo_objective", mutates_args=())
def gspo_objective(logp: Tensor, ref_logp: Tensor, mask: Tensor, rewards_per_func: Tensor, old_logp: Optional[Tensor],
                   num_generations: int, beta: float, epsilon_low: float, epsilon_high: float, gspo: bool
                   ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    state, grad, _ = _gspo.gspo_raw(f32(logp), f32(ref_logp), mask.to(torch.int32).contiguous(), f32(rewards_per_func),
                                    num_generations, beta, epsilon_low, epsilon_high, gspo, f32(old_logp),
                                    want_grad=True, want_kl=False)
    return (state["loss"].reshape(()), grad, state["adv"], state["mean_kl"].reshape(()), state["clen"], state["rstd"])


@gspo_objective.register_fake
def _(logp, ref_logp, mask, rewards_per_func, old_logp, num_generations, beta, epsilon_low, epsilon_high, gspo):
    N, Tc = logp.shape
    f = lambda *s, dt=torch.float32: logp.new_empty(s, dtype=dt)
    return f(), f(N, Tc), f(N), f(), f(N, dt=torch.int32), f(N)


@torch.library.custom_op("o3v::policy_step", mutates_args=())
def policy_step(hidden: Tensor, weight: Tensor, completion_ids: Tensor, ref_logp: Tensor, mask: Tensor,
                rewards_per_func: Tensor, old_logp: Optional[Tensor], num_generations: int, beta: float,
                epsilon_low: float, epsilon_high: float, gspo: bool, chunk_tokens: int
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor, Tensor]:
    out = _logprob.fused_logprob_gspo(hidden, weight, completion_ids, ref_logp, mask, rewards_per_func, num_generations,
                                      beta, epsilon_low, epsilon_high, gspo, old_logp, chunk_tokens=chunk_tokens)
    return (out["loss"], out["per_token_logps"], out["advantages"], out["mean_kl"], out["d_hidden"], out["d_weight"])


@policy_step.register_fake
def _(hidden, weight, completion_ids, ref_logp, mask, rewards_per_func, old_logp, num_generations, beta,
      epsilon_low, epsilon_high, gspo, chunk_tokens):
    N, Tc, H = hidden.shape
    f = lambda *s, dt=torch.float32: hidden.new_empty(s, dtype=dt)
    return f(), f(N, Tc), f(N), f(), f(N, Tc, H, dt=torch.bfloat16), f(*weight.shape)
