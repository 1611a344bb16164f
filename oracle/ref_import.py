"""Import the UNMODIFIED reference modules from /root/reference through stubs.

Test infrastructure (see oracle/__init__.py).  Only usable in the build
container: /root/reference does not exist on the GPU box, so nothing that runs
there may call this.  It is used by tests/golden/gen_golden.py to produce the
committed fixtures and by tests (skipped when the reference is absent) to
re-check the oracle against the live reference.

The reference cannot be imported as-is: reward_func.py:6
needs `rouge_score` (used only by ans_acc_reward, out of scope) and
grpo_trainer.py:59-62 needs `trl`.  Both are replaced by empty stub modules;
no reference source is copied or modified.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("O3V_REFERENCE_ROOT", "/root/reference")
_R1V = os.path.join(REFERENCE_ROOT, "src", "r1-v")
_OPEN_R1 = os.path.join(_R1V, "src", "open_r1")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(_OPEN_R1, "reward_func.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reward_func():
    """The reference's reward_func module (reward_func.py), rouge stubbed."""
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _stub("rouge_score", rouge_scorer=None)
    if _OPEN_R1 not in sys.path:
        sys.path.insert(0, _OPEN_R1)
    return importlib.import_module("reward_func")


def load_trainer_class():
    """The reference's Qwen2VLGRPOTrainer class (grpo_trainer.py), trl stubbed.

    Only `_get_per_token_logps` (grpo_trainer.py:371-384) is callable without a
    live model: it never touches `self`, so it is used unbound.
    """
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    _stub("trl")
    _stub("trl.data_utils", apply_chat_template=None, is_conversational=None,
          maybe_apply_chat_template=None)
    _stub("trl.models", create_reference_model=None, prepare_deepspeed=None,
          unwrap_model_for_generation=None)
    _stub("trl.trainer")
    _stub("trl.trainer.grpo_config", GRPOConfig=object)
    _stub("trl.trainer.utils", generate_model_card=None, get_comet_experiment_url=None)
    if _R1V not in sys.path:
        sys.path.insert(0, _R1V)
    mod = importlib.import_module("src.open_r1.trainer.grpo_trainer")
    return mod.Qwen2VLGRPOTrainer


class FakeLMHeadModel:
    """Stands in for the HF model: `model(input_ids).logits = hidden @ W^T`.

    The backbone is out of scope; the reference's `model(...)` ends in
    `lm_head = nn.Linear(H, V, bias=False)` (transformers
    modeling_qwen2_5_vl.py), which is all the hot path sees.
    """

    def __init__(self, hidden, weight):
        self.hidden, self.weight = hidden, weight

    def __call__(self, input_ids, **kwargs):
        import torch.nn.functional as F
        return types.SimpleNamespace(logits=F.linear(self.hidden, self.weight))


def load_vstar_functions():
    """Numeric functions of the reference's V-STAR scorer (eval/test/eval_vstar.py:90-198).

    The module itself cannot be imported (it parses argv and loads a 72B judge model at import
    time, :10-25), so the five pure functions are compiled from the reference file's own AST and
    executed in a namespace holding what they use (np, ast).  Nothing is copied into this repo."""
    import ast as _ast
    import numpy as _np
    path = os.path.join(REFERENCE_ROOT, "eval", "test", "eval_vstar.py")
    if not os.path.isfile(path):
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    tree = _ast.parse(open(path).read(), filename=path)
    want = {"calculate_temporal_iou", "compute_iou", "calculate_bbox_iou", "calculate_spatial_metrics",
            "calculate_spatial_random"}
    body = [n for n in tree.body if isinstance(n, _ast.FunctionDef) and n.name in want]
    assert {n.name for n in body} == want
    ns = {"np": _np, "ast": _ast}
    exec(compile(_ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in want})


# ---------------------------------------------------------------------------------------------------
# The INLINE numeric block of compute_loss, executed from the reference file itself
# ---------------------------------------------------------------------------------------------------
_LOSS_BLOCK_RANGES = {
    # name: (first line, last line, identifiers that must appear: guards against a shifted file)
    "mask": (590, 596, ("is_eos", "eos_idx", "completion_mask")),
    "kl": (635, 636, ("x_clamped", "per_token_kl")),
    "sum": (658, 658, ("rewards_per_func.sum",)),
    "adv": (675, 681, ("mean_grouped_rewards", "std_grouped_rewards", "advantages")),
    "objective": (691, 706, ("log_ratio", "coef_1", "coef_2", "per_token_loss", "loss")),
    "mean_kl": (737, 737, ("mean_kl",)),
}


def _reference_lines(first, last, must):
    import textwrap
    path = os.path.join(_OPEN_R1, "trainer", "grpo_trainer.py")
    with open(path) as f:
        lines = f.readlines()
    text = "".join(lines[first - 1:last])
    for ident in must:
        if ident not in text:
            raise RuntimeError("grpo_trainer.py:%d-%d does not contain %r: the reference file moved" % (first, last, ident))
    return textwrap.dedent(text)


class _OldPolicyTensor:
    pass


def load_loss_block():
    """The reference's own source lines grpo_trainer.py:590-596, 635-636, 658, 675-681, 691-706, 737, compiled
    from the file where it lies and executed in a namespace that supplies what the lines read (`self`, `torch`,
    the tensors).  Nothing is copied into this repository.  Returns (mask_block, loss_block):

      mask_block(completion_ids, eos_token_id) -> dict(eos_idx, completion_mask)
      loss_block(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func, num_generations, beta,
                 epsilon_low, epsilon_high, gspo, old_per_token_logps=None) -> dict(...)

    The one substitution: the reference's old policy IS the current one (`per_token_logps.detach()`, :691).  To
    drive the off-policy (clipping) side of the same lines, `old_per_token_logps` is delivered through that very
    `.detach()` call: `per_token_logps` is handed in as a tensor subclass whose `detach()` returns it."""
    import torch
    if not reference_available():
        raise RuntimeError("reference not present at %s" % REFERENCE_ROOT)
    code = {k: compile(_reference_lines(a, b, must), "grpo_trainer.py:%d-%d" % (a, b), "exec")
            for k, (a, b, must) in _LOSS_BLOCK_RANGES.items()}

    class _Logps(torch.Tensor):
        def detach(self):
            old = getattr(self, "_o3v_old", None)
            return old if old is not None else super().detach()

    def mask_block(completion_ids, eos_token_id):
        self = types.SimpleNamespace(processing_class=types.SimpleNamespace(eos_token_id=eos_token_id),
                                     accelerator=types.SimpleNamespace(device=completion_ids.device))
        ns = dict(torch=torch, self=self, completion_ids=completion_ids)
        exec(code["mask"], ns)
        return dict(eos_idx=ns["eos_idx"], completion_mask=ns["completion_mask"])

    def loss_block(per_token_logps, ref_per_token_logps, completion_mask, rewards_per_func, num_generations, beta,
                   epsilon_low=0.2, epsilon_high=0.2, gspo=True, old_per_token_logps=None):
        self = types.SimpleNamespace(num_generations=num_generations, beta=beta, epsilon_low=epsilon_low,
                                     epsilon_high=epsilon_high, gspo=gspo)
        lp = per_token_logps
        if old_per_token_logps is not None:
            lp = per_token_logps.as_subclass(_Logps)
            lp._o3v_old = old_per_token_logps
        ns = dict(torch=torch, self=self, per_token_logps=lp, ref_per_token_logps=ref_per_token_logps,
                  completion_mask=completion_mask, rewards_per_func=rewards_per_func)
        for k in ("kl", "sum", "adv", "objective", "mean_kl"):
            exec(code[k], ns)
        plain = lambda t: t.as_subclass(torch.Tensor) if isinstance(t, torch.Tensor) else t
        return {k: plain(ns[k]) for k in ("per_token_kl", "rewards", "advantages", "std_grouped_rewards", "loss",
                                          "mean_kl")}

    return mask_block, loss_block
