"""EOS mask and the GSPO objective (K3a / K3), host side.

Mirrors the inline block of the reference's compute_loss
(src/r1-v/src/open_r1/trainer/grpo_trainer.py:590-596, 635-636, 658, 675-681, 691-706,
711, 737): same argument meaning, same results, one CUDA launch instead of ~25.
"""
import ctypes
from typing import NamedTuple, Optional

import torch

from . import _lib


def _p(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("open-o3-video_b200 ops take CUDA tensors only (no CPU fallback)")


def eos_mask(completion_ids: torch.Tensor, eos_token_id: int):
    """grpo_trainer.py:590-596 -> (eos_idx int64 [N], completion_mask int32 [N, Tc])."""
    _need_cuda(completion_ids)
    ids = completion_ids.to(torch.int64).contiguous()
    N, Tc = ids.shape
    eos_idx = torch.empty(N, dtype=torch.int64, device=ids.device)
    mask = torch.empty(N, Tc, dtype=torch.int32, device=ids.device)
    with torch.cuda.device(ids.device):
        _lib.call("o3v_eos_mask", 1, _lib.load().o3v_eos_mask, _p(ids), N, Tc, int(eos_token_id), _p(eos_idx),
                  _p(mask), _stream())
    return eos_idx, mask


class GspoOutput(NamedTuple):
    loss: torch.Tensor               # scalar, autograd-connected to per_token_logps
    advantages: torch.Tensor         # [N]
    mean_kl: torch.Tensor            # scalar (metric, grpo_trainer.py:737)
    completion_length: torch.Tensor  # [N] int32 (grpo_trainer.py:711)
    reward_std: torch.Tensor         # [N] std of the sequence's group, repeat_interleaved (:679)
    per_token_kl: torch.Tensor       # [N, Tc]


def gspo_raw(logp, ref, mask, rewards_per_func, num_generations, beta, epsilon_low, epsilon_high, gspo,
             old=None, want_grad=True, want_kl=True, *, N_total=None, seq_offset=0, state=None):
    """One K3 launch on rows [seq_offset, seq_offset + n_seq) of a step of N_total sequences.

    `state` (dict) carries the outputs shared by the range calls of one step."""
    _need_cuda(logp, ref, mask, rewards_per_func, old)
    n_seq, Tc = logp.shape
    N = n_seq if N_total is None else N_total
    dev = logp.device
    lib = _lib.load()
    if state is None:
        state = {}
    if not state:
        state.update(
            loss=torch.zeros(1, dtype=torch.float32, device=dev),
            mean_kl=torch.zeros(1, dtype=torch.float32, device=dev),
            adv=torch.empty(N, dtype=torch.float32, device=dev),
            rstd=torch.empty(N, dtype=torch.float32, device=dev),
            clen=torch.empty(N, dtype=torch.int32, device=dev),
            ws=torch.empty(int(lib.o3v_gspo_workspace_bytes(N)), dtype=torch.uint8, device=dev))
    grad = torch.empty(n_seq, Tc, dtype=torch.float32, device=dev) if want_grad else None
    kl = torch.empty(n_seq, Tc, dtype=torch.float32, device=dev) if want_kl else None
    rpf = rewards_per_func
    F = rpf.shape[1]
    with torch.cuda.device(dev):
        _lib.call("o3v_gspo_fwd_bwd", 2, lib.o3v_gspo_fwd_bwd,
                  _p(logp), _p(old), _p(ref), _p(mask), _p(rpf), N, Tc, F, int(num_generations),
                  int(seq_offset), n_seq, float(beta), float(epsilon_low), float(epsilon_high), 1 if gspo else 0,
                  _p(state["loss"]), _p(state["mean_kl"]), _p(state["adv"]), _p(state["rstd"]), _p(state["clen"]),
                  _p(grad), _p(kl), _p(state["ws"]), state["ws"].numel(), _stream())
    return state, grad, kl


class _AttachGrad(torch.autograd.Function):
    """loss value and d loss / d logp were both produced by the same K3 launch; this only
    wires them into autograd."""

    @staticmethod
    def forward(ctx, logp, loss, grad):
        ctx.save_for_backward(grad)
        return loss.clone()

    @staticmethod
    def backward(ctx, g_loss):
        (grad,) = ctx.saved_tensors
        return grad * g_loss, None, None


def gspo_loss(per_token_logps: torch.Tensor, ref_per_token_logps: torch.Tensor,
              completion_mask: torch.Tensor, rewards_per_func: torch.Tensor, num_generations: int,
              beta: float, epsilon_low: float = 0.2, epsilon_high: float = 0.2, gspo: bool = True,
              old_per_token_logps: Optional[torch.Tensor] = None) -> GspoOutput:
    """KL + group advantages + GSPO / token-level clipped objective (grpo_trainer.py:635-706).

    per_token_logps / ref / old: [N, Tc]; completion_mask [N, Tc] int; rewards_per_func
    [N, F] (one column per reward function, summed as at :658).  `old_per_token_logps=None`
    is the reference's behaviour (`per_token_logps.detach()`, :691).  Computed in fp32.
    """
    f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
    logp = per_token_logps.to(torch.float32).contiguous()
    want_grad = torch.is_grad_enabled() and logp.requires_grad
    state, grad, kl = gspo_raw(logp.detach(), f32(ref_per_token_logps), completion_mask.to(torch.int32).contiguous(),
                               f32(rewards_per_func), int(num_generations), float(beta), float(epsilon_low),
                               float(epsilon_high), bool(gspo), f32(old_per_token_logps), want_grad=want_grad)
    loss = state["loss"].reshape(())
    if want_grad:
        loss = _AttachGrad.apply(logp, loss, grad)
    return GspoOutput(loss, state["adv"], state["mean_kl"].reshape(()), state["clen"], state["rstd"], kl)
