"""Oracle: EOS mask, KL, group advantages, GSPO objective and metrics.

Test infrastructure (see oracle/__init__.py).  Line-for-line torch restatement
of the block that is INLINE in the reference's compute_loss and therefore cannot
be imported (it needs a live model/processor/generate):
  src/r1-v/src/open_r1/trainer/grpo_trainer.py
    :590-596  first-EOS mask
    :635-636  per-token KL
    :658      reward sum over functions
    :675-681  group mean / std / advantages
    :691-706  GSPO (or token-level) ratio, clip, loss
    :711,:737 metrics (completion length, mean KL)
The only generalisation: `per_token_logps.detach()` at :691 becomes the
`old_per_token_logps` argument, defaulting to exactly that.
"""
import torch


def eos_mask(completion_ids: torch.Tensor, eos_token_id: int):
    """grpo_trainer.py:590-596 -> (eos_idx int64 [N], completion_mask int32 [N, Tc])."""
    is_eos = completion_ids == eos_token_id                                           # :591
    eos_idx = torch.full((is_eos.size(0),), is_eos.size(1), dtype=torch.long)         # :593
    eos_idx[is_eos.any(dim=1)] = is_eos.int().argmax(dim=1)[is_eos.any(dim=1)]        # :594
    sequence_indices = torch.arange(is_eos.size(1)).expand(is_eos.size(0), -1)        # :595
    completion_mask = (sequence_indices <= eos_idx.unsqueeze(1)).int()                # :596
    return eos_idx, completion_mask


def per_token_kl(per_token_logps, ref_per_token_logps):
    """grpo_trainer.py:635-636."""
    x_clamped = torch.clamp(ref_per_token_logps - per_token_logps, min=-10, max=10)
    return torch.exp(x_clamped) - x_clamped - 1


def group_advantages(rewards_per_func: torch.Tensor, num_generations: int):
    """grpo_trainer.py:658, 675-681 -> (rewards, advantages, std_grouped_rewards)."""
    rewards = rewards_per_func.sum(dim=1)                                             # :658
    mean_grouped_rewards = rewards.view(-1, num_generations).mean(dim=1)              # :675
    std_grouped_rewards = rewards.view(-1, num_generations).std(dim=1)                # :676
    mean_grouped_rewards = mean_grouped_rewards.repeat_interleave(num_generations, dim=0)
    std_grouped_rewards = std_grouped_rewards.repeat_interleave(num_generations, dim=0)
    advantages = (rewards - mean_grouped_rewards) / (std_grouped_rewards + 1e-4)      # :681
    return rewards, advantages, std_grouped_rewards


def gspo_objective(per_token_logps, ref_per_token_logps, completion_mask, advantages,
                   beta: float, epsilon_low: float = 0.2, epsilon_high: float = 0.2,
                   gspo: bool = True, old_per_token_logps=None):
    """grpo_trainer.py:635-636, 691-706, 711, 737.

    Returns dict(loss, per_token_kl, mean_kl, completion_length).  `loss` carries
    autograd history when `per_token_logps.requires_grad`.
    """
    kl = per_token_kl(per_token_logps, ref_per_token_logps)                           # :635-636
    old = per_token_logps.detach() if old_per_token_logps is None else old_per_token_logps
    log_ratio = per_token_logps - old                                                 # :691
    if gspo:                                                                          # :692
        log_importance_weights = (log_ratio * completion_mask).sum(-1) / completion_mask.sum(-1).clamp(min=1.0)
        log_importance_weights = log_importance_weights.unsqueeze(-1)                 # :694
    else:
        log_importance_weights = log_ratio                                            # :696
    coef_1 = torch.exp(log_importance_weights)                                        # :698
    coef_2 = torch.clamp(coef_1, 1 - epsilon_low, 1 + epsilon_high)                   # :699
    per_token_loss1 = coef_1 * advantages.unsqueeze(1)                                # :701
    per_token_loss2 = coef_2 * advantages.unsqueeze(1)                                # :702
    per_token_loss = -torch.min(per_token_loss1, per_token_loss2)                     # :703
    per_token_loss = per_token_loss + beta * kl                                       # :704
    loss = ((per_token_loss * completion_mask).sum(-1) / completion_mask.sum(-1).clamp(min=1.0)).mean()  # :706
    completion_length = completion_mask.sum(1)                                        # :711
    mean_kl = ((kl * completion_mask).sum(dim=1) / completion_mask.sum(dim=1)).mean() # :737
    return dict(loss=loss, per_token_kl=kl, mean_kl=mean_kl, completion_length=completion_length)


def gspo_step(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func,
              num_generations: int, beta: float, epsilon_low: float = 0.2,
              epsilon_high: float = 0.2, gspo: bool = True, old_per_token_logps=None):
    """Advantages + objective in the reference's order (:658 -> :706)."""
    rewards, advantages, std = group_advantages(rewards_per_func, num_generations)
    out = gspo_objective(per_token_logps, ref_per_token_logps, completion_mask, advantages,
                         beta, epsilon_low, epsilon_high, gspo, old_per_token_logps)
    out.update(rewards=rewards, advantages=advantages, reward_std=std)
    return out


def analytic_grad(per_token_logps, ref_per_token_logps, completion_mask, advantages,
                  beta: float, epsilon_low: float = 0.2, epsilon_high: float = 0.2,
                  old_per_token_logps=None):
    """Closed form of d loss / d per_token_logps for the GSPO branch.

    Not reference code: a second derivation used to cross-check autograd through
    gspo_objective and to document what the CUDA backward must produce.
    """
    lp, ref, m = per_token_logps.detach().double(), ref_per_token_logps.double(), completion_mask.double()
    old = lp if old_per_token_logps is None else old_per_token_logps.double()
    A = advantages.double()
    N = lp.shape[0]
    n = m.sum(-1).clamp(min=1.0)
    nz = (m.sum(-1) > 0).double()
    s = ((lp - old) * m).sum(-1) / n
    c1 = torch.exp(s)
    gate = torch.where(A > 0, (c1 <= 1 + epsilon_high).double(),
                       torch.where(A < 0, (c1 >= 1 - epsilon_low).double(),
                                   # A == 0: torch.minimum splits the tie, the clamp branch may be dead; A=0 kills it anyway
                                   torch.ones_like(c1)))
    x = ref - lp
    inr = ((x >= -10) & (x <= 10)).double()
    dkl = (1.0 - torch.exp(torch.clamp(x, -10, 10))) * inr
    g = m / (n[:, None] * N) * ((-A * c1 * gate * nz)[:, None] + beta * dkl)
    return g
