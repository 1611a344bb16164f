"""Drop-in for the hot section of the reference trainer
(src/r1-v/src/open_r1/trainer/grpo_trainer.py, class Qwen2VLGRPOTrainer).

Usage (the only change a user of the reference makes, INTEGRATION.md level 1):

    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    class Trainer(O3VB200TrainerMixin, Qwen2VLGRPOTrainer): pass

What the mixin overrides, with the reference's own signatures:

* `compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None)` (:402-405, same
  `ValueError` on `return_outputs`).  It records the prompt length the reference computes as a local (:583) by
  watching the `generate` call of the unwrapped model (:581-582), so that both `_get_per_token_logps` calls
  (:612, :622-631) project ONLY the completion rows the caller keeps (`[:, prompt_length - 1:]`, :613 / :626) —
  at the reference's shapes (16 384 prompt + 768 completion tokens, run_grpo_video.sh:21-23) 95.5 % of the head
  work is prompt rows that are thrown away — and then delegates to the reference's own `compute_loss`.
* `_get_per_token_logps(self, model, input_ids, **kwargs) -> [B, L-1]` (:371-384).  The model's own forward runs
  unchanged, but for the duration of the call `lm_head.forward` is replaced by the fused
  lm_head / log-softmax / gather kernel (K1, `logprob.fused_logprob`): the `[B, L, V]` logits never exist.  The
  head MODULE stays in place, so whatever hooks wrap it still fire (DeepSpeed ZeRO-3 gathers `lm_head.weight`
  before the module's forward and again before its backward, PEFT / accelerate wrappers resolve to the real head).

Level 2 (`integration/patch_reference.py`) rewrites the three inline blocks of the reference's
`compute_loss` into calls of the methods below, because an inline block cannot be overridden by inheritance:

* `o3v_completion_mask(completion_ids)`                         :590-596  (K3a)
* `compute_policy_loss(per_token_logps, ref, mask, rewards_per_func)`   :635-636, 658, 675-681, 691-738  (K3)
* `o3v_fused_policy_loss(model, prompt_completion_ids, prompt_length, ref, mask, rewards_per_func, **vision)`
  the policy pass, the loss AND their backward in one chunked fused step (K1 + K3 + K2a + K2b), attached to
  autograd at the hidden states (Liger-style: the gradients are produced in the forward and handed to autograd).
"""
import contextlib
from collections import defaultdict
from typing import Optional

import torch

from . import gspo as _gspo
from . import logprob as _logprob


def _unwrap_chain(model):
    """model, model.module, model.module.module ... (DDP / DeepSpeed / accelerate wrappers)."""
    out = [model]
    while True:
        inner = getattr(model, "module", None)
        if not isinstance(inner, torch.nn.Module):
            inner = model.__dict__.get("_orig_mod") if hasattr(model, "__dict__") else None     # torch.compile wrapper
        if not isinstance(inner, torch.nn.Module):
            inner = getattr(model, "_modules", {}).get("_orig_mod")
        if not isinstance(inner, torch.nn.Module):
            return out
        model = inner
        out.append(model)


def _unwrap(model):
    return _unwrap_chain(model)[-1]


def find_lm_head(model) -> torch.nn.Module:
    """The module that IS the head: the value of the `_modules['lm_head']` entry of whichever submodule owns it.
    A `getattr(model, 'lm_head')` is not enough: wrappers such as PeftModel forward unknown attributes to the
    wrapped model through `__getattr__`, so reading works but assigning would register a second module on the
    wrapper (round-1 advisor finding)."""
    m = _unwrap(model)
    for mod in m.modules():                     # modules() yields m itself first, then depth-first
        head = mod._modules.get("lm_head")
        if head is not None:
            return head
    get = getattr(m, "get_output_embeddings", None)
    head = get() if callable(get) else None
    if head is None:
        raise RuntimeError("O3VB200TrainerMixin needs a model with an `lm_head` (got %s); there is no "
                           "logits-materialising fallback" % type(m).__name__)
    return head


def lm_head_weight(model) -> torch.Tensor:
    head = find_lm_head(model)
    # PEFT `modules_to_save` wraps the head; the trainable copy is the active adapter's module
    inner = getattr(head, "modules_to_save", None)
    if inner is not None and getattr(head, "active_adapter", None) in inner:
        head = inner[head.active_adapter]
    if getattr(head, "bias", None) is not None or not hasattr(head, "weight"):
        raise RuntimeError("expected lm_head = nn.Linear(H, V, bias=False)")
    return head.weight


def _check_gathered(weight):
    """DeepSpeed ZeRO-3 (the reference's launch: run_grpo_video.sh:20, local_scripts/zero3.json) keeps a 0-element
    placeholder in `param.data` outside the owning module's forward / backward.  Inside the patched
    `lm_head.forward` the module's ZeRO-3 pre-forward hook has gathered it; if it is still a placeholder the hooks
    are not installed on this module and there is nothing sensible to compute with."""
    if hasattr(weight, "ds_id") and weight.numel() == 0:
        raise RuntimeError(
            "lm_head.weight is a DeepSpeed ZeRO-3 partitioned placeholder (ds_shape=%s) inside lm_head.forward: "
            "the ZeRO-3 module hooks did not gather it.  Call the model through the DeepSpeed engine, or wrap the "
            "call in deepspeed.zero.GatheredParameters([lm_head.weight]) for a forward-only pass."
            % (tuple(getattr(weight, "ds_shape", ())),))
    if weight.dim() != 2:
        raise RuntimeError("lm_head.weight must be [V, H] (got shape %s)" % (tuple(weight.shape),))


class _HeadPatch:
    """Replaces `lm_head.forward` (instance attribute, the module object itself is untouched) for one model call.

    `fn(head, hidden_states) -> tensor [B, L', 1]` computes on the final hidden states; what it returns becomes
    the model's `.logits`, so gradients reach it THROUGH the module's output (which is where DeepSpeed's
    pre-backward hooks sit)."""

    def __init__(self, model, fn):
        self.head = find_lm_head(model)
        inner = getattr(self.head, "modules_to_save", None)
        if inner is not None and getattr(self.head, "active_adapter", None) in inner:
            self.head = inner[self.head.active_adapter]
        self.fn = fn
        self.called = False
        self.result = None

    def __enter__(self):
        self._had = "forward" in self.head.__dict__
        self._old = self.head.__dict__.get("forward")

        def forward(hidden_states, *a, **kw):
            self.called = True
            self.result = self.fn(self.head, hidden_states)
            return self.result

        self.head.forward = forward
        return self

    def __exit__(self, *exc):
        if self._had:
            self.head.forward = self._old
        else:
            del self.head.forward
        return False


def _logits_of(out, tap):
    if not tap.called:
        raise RuntimeError("the model's forward never called lm_head")
    logits = getattr(out, "logits", None)
    if logits is None and isinstance(out, (tuple, list)) and len(out):
        logits = out[0]
    if not isinstance(logits, torch.Tensor) or logits.shape != tap.result.shape:
        logits = tap.result                      # a wrapper re-packed the output: use what the head returned
    return logits


def final_hidden_states(model, input_ids, **kwargs) -> torch.Tensor:
    """The hidden states `lm_head` is applied to ([B, L, H]), without the head (diagnostics / tests).

    The reference calls `model(input_ids, **kwargs).logits` with the vision kwargs (`pixel_values_videos`,
    `video_grid_thw`, grpo_trainer.py:375, :603-611).  Where the vision tower is merged differs between
    transformers versions, so the full forward is run unchanged with only `lm_head.forward` replaced."""
    seen = {}

    def grab(head, hidden_states):
        seen["h"] = hidden_states
        return hidden_states[..., :1]

    with _HeadPatch(model, grab) as tap:
        model(input_ids, **kwargs)
    if not tap.called:
        raise RuntimeError("%s.forward never called lm_head" % type(_unwrap(model)).__name__)
    return seen["h"]


@contextlib.contextmanager
def _generate_probe(trainer, model):
    """For the duration of the reference's compute_loss: note the width of the `input_ids` handed to `generate`
    (grpo_trainer.py:582; `prompt_length = prompt_ids.size(1)` at :583 is the same number because :570-571 truncate
    `prompt_inputs['input_ids']` itself) on every wrapper level `unwrap_model_for_generation` may yield."""
    patched = []
    for m in _unwrap_chain(model):
        orig = getattr(m, "generate", None)
        if not callable(orig):
            continue
        had = "generate" in m.__dict__

        def wrapped(*a, _orig=orig, **kw):
            ids = kw.get("input_ids", kw.get("inputs", a[0] if a else None))
            out = _orig(*a, **kw)
            if isinstance(ids, torch.Tensor) and ids.dim() == 2:
                trainer.o3v_prompt_length = int(ids.shape[1])
            return out

        patched.append((m, had, m.__dict__.get("generate")))
        object.__setattr__(m, "generate", wrapped)
    try:
        yield
    finally:
        for m, had, old in patched:
            if had:
                object.__setattr__(m, "generate", old)
            else:
                object.__delattr__(m, "generate")
        trainer.o3v_prompt_length = None


class O3VB200TrainerMixin:
    # completion rows only: set by compute_loss (from the generate call) for the duration of one step to skip the
    # prompt positions the caller throws away at grpo_trainer.py:613 / :626; None = project every position
    o3v_prompt_length: Optional[int] = None
    # token chunk of the fused backward (bounds the bf16 logits buffer), 0 = sized from a 10 GB budget
    o3v_chunk_tokens: int = 0

    # ------------------------------------------------------------------ level 1: pure inheritance
    def compute_loss(self, model, inputs, return_outputs=False, num_items_in_batch=None):
        if return_outputs:
            raise ValueError("The GRPOTrainer does not support returning outputs")       # :404-405
        with _generate_probe(self, model):
            return super().compute_loss(model, inputs, return_outputs=return_outputs,
                                        num_items_in_batch=num_items_in_batch)

    def _get_per_token_logps(self, model, input_ids, **kwargs):
        skip = self.o3v_prompt_length
        L = input_ids.shape[1]
        start = 0 if not skip else min(max(int(skip) - 1, 0), L - 1)

        def head_fn(head, hidden_states):
            w = head.weight
            _check_gathered(w)
            B, Lh, H = hidden_states.shape
            if Lh != L:
                raise RuntimeError("lm_head saw %d positions for %d input ids (logits_to_keep?)" % (Lh, L))
            h = hidden_states[:, start:L - 1, :].to(torch.bfloat16)                      # :376 + the caller's slice
            tgt = input_ids[:, start + 1:]                                               # :377
            lp = _logprob.fused_logprob(h.reshape(-1, H), w.to(torch.bfloat16), tgt.reshape(-1))
            return lp.view(B, L - 1 - start, 1)

        with _HeadPatch(model, head_fn) as tap:
            out = model(input_ids, **kwargs)
        lp = _logits_of(out, tap)[..., 0]
        if start == 0:
            return lp
        # the discarded prompt part is never read by the caller; keep the [B, L-1] contract
        return torch.cat([lp.new_zeros(lp.shape[0], start), lp], dim=1)

    # ------------------------------------------------------------------ level 2: the patched inline blocks
    def o3v_completion_mask(self, completion_ids):
        """grpo_trainer.py:590-596 in one launch (K3a) -> completion_mask int32 [N, Tc]."""
        return _gspo.eos_mask(completion_ids, self.processing_class.eos_token_id)[1]

    def _o3v_scalars(self):
        G = getattr(self, "num_generations", None) or self.args.num_generations
        return (G, getattr(self, "beta", 0.04), getattr(self, "epsilon_low", 0.2), getattr(self, "epsilon_high", 0.2),
                getattr(self, "gspo", True))

    def _o3v_log_metrics(self, rewards_per_func, completion_length, reward_std, mean_kl):
        """grpo_trainer.py:711-738: same keys, same reductions, same `.item()` host reads."""
        G = self._o3v_scalars()[0]
        metrics = getattr(self, "_metrics", None)
        if metrics is None:
            metrics = self._metrics = defaultdict(list)
        gather = self.accelerator.gather_for_metrics if hasattr(self, "accelerator") else (lambda t: t)
        rewards = rewards_per_func.sum(dim=1)                                                           # :658
        metrics["completion_length"].append(gather(completion_length).float().mean().item())           # :711
        per_func = gather(rewards_per_func).mean(0)                                                     # :714
        for i, f in enumerate(getattr(self, "reward_funcs", [])):
            if isinstance(f, torch.nn.Module):                                                         # :716-717
                name = f.config._name_or_path.split("/")[-1]
            else:
                name = getattr(f, "__name__", str(i))
            metrics["rewards/%s" % name].append(per_func[i].item())                                    # :720
        gathered = gather(rewards)                                                                     # :722
        num_devices = gathered.size(0) // G                                                            # :724
        per_dev = gathered.view(num_devices, G)
        metrics["all_wrong"].append((per_dev <= 1).all(dim=1).sum().item() / num_devices)               # :726-731
        metrics["all_correct"].append((per_dev >= 2).all(dim=1).sum().item() / num_devices)
        metrics["reward"].append(gathered.mean().item())                                               # :734
        metrics["reward_std"].append(gather(reward_std).mean().item())                                 # :735
        metrics["kl"].append(gather(mean_kl).mean().item())                                            # :738

    def compute_policy_loss(self, per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func,
                            old_per_token_logps=None):
        """grpo_trainer.py:635-738 without the text work: returns the loss (autograd-connected to
        `per_token_logps`) and appends the reference's metric keys to `self._metrics`."""
        G, beta, el, eh, gs = self._o3v_scalars()
        out = _gspo.gspo_loss(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func, G, beta, el, eh,
                              gs, old_per_token_logps)
        self._o3v_log_metrics(rewards_per_func, out.completion_length, out.reward_std, out.mean_kl)
        return out.loss

    def o3v_fused_policy_loss(self, model, prompt_completion_ids, prompt_length, ref_per_token_logps, completion_mask,
                              rewards_per_func, **kwargs):
        """The policy pass (:612-613) and the loss block (:635-738) as ONE fused chunked step inside the model's own
        forward: K1 + K3 + dlogits + K2a + K2b run where `lm_head` would, dHidden / dW are handed to autograd, the
        returned loss back-propagates into the backbone and into `lm_head.weight` through the head module (so ZeRO-3
        / DDP gradient hooks see it)."""
        G, beta, el, eh, gs = self._o3v_scalars()
        L = prompt_completion_ids.shape[1]
        start = min(max(int(prompt_length) - 1, 0), L - 1)
        info = {}

        def head_fn(head, hidden_states):
            w = head.weight
            _check_gathered(w)
            B, Lh, H = hidden_states.shape
            h = hidden_states[:, start:L - 1, :].to(torch.bfloat16)
            ids = prompt_completion_ids[:, start + 1:]
            res = _logprob.fused_policy_step(h, w.to(torch.bfloat16), ids, ref_per_token_logps, completion_mask,
                                             rewards_per_func, G, beta, el, eh, gs,
                                             chunk_tokens=self.o3v_chunk_tokens or _logprob.DEFAULT_CHUNK_TOKENS)
            info.update(res)
            return res["loss"].reshape(1, 1, 1)

        with _HeadPatch(model, head_fn) as tap:
            out = model(prompt_completion_ids, **kwargs)
        loss = _logits_of(out, tap).reshape(())
        self._o3v_log_metrics(rewards_per_func, info["completion_length"], info["reward_std"], info["mean_kl"])
        self.o3v_last_step = info
        return loss
