"""Oracle: per-token log-probs = lm_head -> log_softmax -> gather.

Test infrastructure (see oracle/__init__.py).  Restates, in plain torch on the
CPU, what the reference computes at
  src/r1-v/src/open_r1/trainer/grpo_trainer.py:371-384  (_get_per_token_logps)
with the third-party `model(...).logits` replaced by its last layer
`lm_head = nn.Linear(H, V, bias=False)` (transformers @ 336dc69d, not vendored in
the reference; `F.linear(hidden, W)`).

"The reference result" is this code executed in fp32 on bf16-representable
inputs.
"""
import torch
import torch.nn.functional as F


def per_token_logps(hidden: torch.Tensor, weight: torch.Tensor,
                    input_ids: torch.Tensor) -> torch.Tensor:
    """[B, L, H], [V, H], [B, L] int64 -> [B, L-1].

    grpo_trainer.py:375   logits = model(input_ids, **kwargs).logits
    grpo_trainer.py:376   logits = logits[:, :-1, :]
    grpo_trainer.py:377   input_ids = input_ids[:, 1:]
    grpo_trainer.py:380-383  per row: log_softmax(-1), gather(dim=1, ids)
    grpo_trainer.py:384   torch.stack
    """
    logits = F.linear(hidden, weight)
    logits = logits[:, :-1, :]
    ids = input_ids[:, 1:]
    out = []
    for logits_row, ids_row in zip(logits, ids):
        log_probs = logits_row.log_softmax(dim=-1)
        out.append(torch.gather(log_probs, dim=1, index=ids_row.unsqueeze(1)).squeeze(1))
    return torch.stack(out)


def token_logps(hidden: torch.Tensor, weight: torch.Tensor, targets: torch.Tensor,
                chunk: int = 1024):
    """Flat form used by the kernels' tests: hidden [T, H], targets [T] -> (logp [T], lse [T]).

    Same arithmetic as per_token_logps after the caller's shift (each row of
    `hidden` already paired with its next-token id); chunked over T so that the
    [T, V] logits of the large configs fit in host memory.
    """
    logp = torch.empty(hidden.shape[0], dtype=hidden.dtype)
    lse = torch.empty(hidden.shape[0], dtype=hidden.dtype)
    for s in range(0, hidden.shape[0], chunk):
        z = F.linear(hidden[s:s + chunk], weight)
        l = torch.logsumexp(z, dim=-1)
        lp = z.log_softmax(dim=-1)
        logp[s:s + chunk] = torch.gather(lp, 1, targets[s:s + chunk, None])[:, 0]
        lse[s:s + chunk] = l
    return logp, lse
