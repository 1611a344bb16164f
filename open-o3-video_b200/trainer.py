"""Drop-in for the hot section of the reference trainer
(src/r1-v/src/open_r1/trainer/grpo_trainer.py, class Qwen2VLGRPOTrainer).

Usage (the only change a user of the reference makes):

    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    class Trainer(O3VB200TrainerMixin, Qwen2VLGRPOTrainer): pass

`_get_per_token_logps(self, model, input_ids, **kwargs)` keeps the reference signature and
return contract ([B, L-1] log-probs of input_ids[:, 1:], grpo_trainer.py:371-384) but calls the
model with `lm_head` tapped (so the final hidden states come back instead of `[B, L, V]` logits) and
runs the fused lm_head/log-softmax/gather kernel on them.  `compute_policy_loss` is the numeric block of
compute_loss (:635-636, 658, 675-681, 691-706) plus its metrics (:711-738) in one launch.
Everything else in compute_loss (vision prep, generate, decode, reward callables) is the
reference's own code and is untouched.
"""
from collections import defaultdict
from typing import Optional

import torch

from . import gspo as _gspo
from . import logprob as _logprob


def _unwrap(model):
    """Strip DDP / DeepSpeed / accelerate wrappers (`.module`)."""
    while hasattr(model, "module") and isinstance(getattr(model, "module"), torch.nn.Module):
        model = model.module
    return model


class _HeadTap(torch.nn.Module):
    """Stands in for `lm_head` during one forward: remembers the hidden states it is handed and returns a
    zero-width logits tensor, so that nothing of size [B, L, V] is ever computed."""

    def __init__(self):
        super().__init__()
        self.hidden = None

    def forward(self, hidden_states):
        self.hidden = hidden_states
        return hidden_states[..., :0]


def final_hidden_states(model, input_ids, **kwargs) -> torch.Tensor:
    """The hidden states `lm_head` would be applied to ([B, L, H]), without the head.

    The reference calls `model(input_ids, **kwargs).logits` with the vision kwargs (`pixel_values_videos`,
    `video_grid_thw`, grpo_trainer.py:375, :603-611).  Where the vision tower is merged differs between
    transformers versions (inside `ForConditionalGeneration.forward` at the commit the reference pins, inside
    `.model` in later releases), so the full forward is run unchanged with `lm_head` swapped for a tap that records
    its input; this works for every causal-LM class that ends in `self.lm_head(hidden_states)`."""
    m = _unwrap(model)
    head = getattr(m, "lm_head", None)
    if head is None:
        raise RuntimeError("O3VB200TrainerMixin needs a model with an `lm_head` (got %s); there is no "
                           "logits-materialising fallback" % type(m).__name__)
    tap = _HeadTap()
    m.lm_head = tap
    try:
        model(input_ids, **kwargs)
    finally:
        m.lm_head = head
    if tap.hidden is None:
        raise RuntimeError("%s.forward never called lm_head" % type(m).__name__)
    return tap.hidden


def lm_head_weight(model) -> torch.Tensor:
    m = _unwrap(model)
    head = getattr(m, "lm_head", None)
    if head is None and hasattr(m, "get_output_embeddings"):
        head = m.get_output_embeddings()
    if head is None or getattr(head, "bias", None) is not None:
        raise RuntimeError("expected lm_head = nn.Linear(H, V, bias=False)")
    return head.weight


class O3VB200TrainerMixin:
    # set by compute_loss before the call to skip the prompt positions the caller throws away at
    # grpo_trainer.py:613 / :626 (`[:, prompt_length - 1:]`); None = project every position
    o3v_prompt_length: Optional[int] = None

    def _get_per_token_logps(self, model, input_ids, **kwargs):
        hidden = final_hidden_states(model, input_ids, **kwargs)
        weight = lm_head_weight(model)
        B, L, H = hidden.shape
        skip = self.o3v_prompt_length
        start = 0 if not skip else max(int(skip) - 1, 0)
        h = hidden[:, start:L - 1, :].to(torch.bfloat16)
        tgt = input_ids[:, start + 1:]
        lp = _logprob.fused_logprob(h.reshape(-1, H), weight.to(torch.bfloat16), tgt.reshape(-1)).view(B, L - 1 - start)
        if start == 0:
            return lp
        # the discarded prompt part is never read by the caller; keep the [B, L-1] contract
        return torch.cat([lp.new_zeros(B, start), lp], dim=1)

    def compute_policy_loss(self, per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func,
                            old_per_token_logps=None):
        """grpo_trainer.py:635-738 without the text work: returns the loss and appends the
        reference's metric keys to `self._metrics`."""
        G = getattr(self, "num_generations", None) or self.args.num_generations
        out = _gspo.gspo_loss(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func, G,
                              getattr(self, "beta", 0.04), getattr(self, "epsilon_low", 0.2),
                              getattr(self, "epsilon_high", 0.2), getattr(self, "gspo", True), old_per_token_logps)
        metrics = getattr(self, "_metrics", None)
        if metrics is None:
            metrics = self._metrics = defaultdict(list)
        gather = self.accelerator.gather_for_metrics if hasattr(self, "accelerator") else (lambda t: t)
        rewards = rewards_per_func.sum(dim=1)
        metrics["completion_length"].append(gather(out.completion_length).float().mean().item())       # :711
        per_func = gather(rewards_per_func).mean(0)                                                     # :714
        for i, f in enumerate(getattr(self, "reward_funcs", [])):
            metrics["rewards/%s" % getattr(f, "__name__", str(i))].append(per_func[i].item())          # :720
        gathered = gather(rewards)                                                                     # :722
        per_dev = gathered.view(-1, G)
        metrics["all_wrong"].append((per_dev <= 1).all(dim=1).float().mean().item())                    # :726-731
        metrics["all_correct"].append((per_dev >= 2).all(dim=1).float().mean().item())
        metrics["reward"].append(gathered.mean().item())                                               # :734
        metrics["reward_std"].append(gather(out.reward_std).mean().item())                             # :735
        metrics["kl"].append(gather(out.mean_kl).mean().item())                                        # :738
        return out.loss
