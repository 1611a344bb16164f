"""Oracle: the SFT stage's causal-LM cross-entropy.

Test infrastructure (see oracle/__init__.py).  The reference's SFT trainer
(src/r1-v/src/open_r1/sft_multi_task.py:402-409, `MySFTTrainer.compute_loss`) delegates to
trl's SFTTrainer -> the HF model's own loss with `labels` built at :387-398 (input_ids with pad and
visual tokens set to -100).  That loss lives in the third-party `transformers` package (pinned
336dc69d, not vendored): `ForCausalLMLoss` = shift by one, `cross_entropy(ignore_index=-100)`,
mean over the non-ignored targets (or sum / num_items_in_batch).  Restated here in fp32 torch.
"""
import torch
import torch.nn.functional as F


def causal_lm_loss(hidden, weight, labels, ignore_index=-100, num_items_in_batch=None):
    logits = F.linear(hidden, weight).float()                       # lm_head
    shift_logits = logits[:, :-1, :].reshape(-1, logits.shape[-1])
    shift_labels = labels[:, 1:].reshape(-1)
    reduction = "sum" if num_items_in_batch is not None else "mean"
    loss = F.cross_entropy(shift_logits, shift_labels, ignore_index=ignore_index, reduction=reduction)
    if num_items_in_batch is not None:
        loss = loss / num_items_in_batch
    return loss
