"""The drop-in trainer mixin on the CPU box.

* the REAL reference class (stub-imported from /root/reference, skipped when absent) runs its own compute_loss on
  the fakes of oracle/host_trainer.py; its loss / metrics pin (a) `oracle.host_trainer.HostTrainer`, the
  restatement the GPU tests use as host class, and (b) the committed golden `tests/golden/compute_loss_small.json`;
* `class T(O3VB200TrainerMixin, <reference class>)`: compute_loss records the prompt length from the generate call
  and only B*G*Tc completion rows reach the head, loss unchanged;
* `integration/patch_reference.py` applied to a temp copy of the reference file, stub-imported, both modes;
* lm_head lookup through PEFT-like wrappers, a ZeRO-3-like partitioned parameter, head restored on errors.
The CUDA launches are replaced by torch doubles here (no GPU): what is tested is the host logic; the same scenarios
run with the real kernels in tests/test_gpu_trainer.py.
"""
import importlib.util
import json
import os
import sys
import types

import pytest
import torch

from oracle import gspo as ogspo, host_trainer as ht, ref_import

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "compute_loss_small.json")
needs_ref = pytest.mark.skipif(not ref_import.reference_available(), reason="reference not present")


# ------------------------------------------------------------------------------------------ torch doubles
class Doubles:
    """CPU stand-ins for the C-ABI launches, same signatures as the package functions they replace."""

    def __init__(self, monkeypatch):
        from open_o3_video_b200 import gspo, logprob
        self.rows = []
        self.fused_steps = 0
        monkeypatch.setattr(logprob, "fused_logprob", self.fused_logprob)
        monkeypatch.setattr(logprob, "fused_policy_step", self.fused_policy_step)
        monkeypatch.setattr(gspo, "eos_mask", lambda ids, eos: ogspo.eos_mask(ids, eos))
        monkeypatch.setattr(gspo, "gspo_loss", self.gspo_loss)

    def fused_logprob(self, hidden, weight, targets, **kw):
        assert hidden.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
        self.rows.append(hidden.shape[0])
        z = hidden.float() @ weight.float().T
        return z.log_softmax(-1).gather(1, targets[:, None])[:, 0]

    def gspo_loss(self, lp, ref, mask, rpf, G, beta, el, eh, gs, old=None):
        from open_o3_video_b200.gspo import GspoOutput
        o = ogspo.gspo_step(lp.float(), ref.float(), mask, rpf, G, beta, el, eh, gs, old)
        return GspoOutput(o["loss"], o["advantages"], o["mean_kl"], o["completion_length"], o["reward_std"],
                          o["per_token_kl"])

    def fused_policy_step(self, hidden, weight, ids, ref, mask, rpf, G, beta, el, eh, gs, old=None, **opts):
        self.fused_steps += 1
        N, Tc, H = hidden.shape
        lp = self.fused_logprob(hidden.reshape(-1, H), weight, ids.reshape(-1)).view(N, Tc)
        o = ogspo.gspo_step(lp, ref.float(), mask, rpf, G, beta, el, eh, gs, old)
        return dict(loss=o["loss"], per_token_logps=lp.detach(), advantages=o["advantages"], mean_kl=o["mean_kl"],
                    completion_length=o["completion_length"], reward_std=o["reward_std"])


def _models():
    return ht.FakeVLModel(seed=11), ht.FakeVLModel(seed=12).eval()


def _run(cls, gspo=True, return_model=False):
    model, ref = _models()
    t = ht.configure(cls.__new__(cls), model, ref, gspo=gspo)
    loss = t.compute_loss(model, [ht.make_example()])
    loss.backward()
    out = dict(loss=loss.item(), metrics={k: v[0] for k, v in t._metrics.items()},
               g_head=model.lm_head.weight.grad.norm().item(), g_embed=model.embed.weight.grad.norm().item())
    return (out, model, t) if return_model else out


def _ref_class():
    cls = ref_import.load_trainer_class()
    ht.install_reference_fakes(sys.modules[cls.__module__])
    return cls


# ------------------------------------------------------------------------------------------ pins
@needs_ref
@pytest.mark.parametrize("gspo", [True, False])
def test_host_trainer_restatement_equals_live_reference(gspo):
    a = _run(_ref_class(), gspo)
    b = _run(ht.HostTrainer, gspo)
    assert a == b                      # loss, every metric and the gradient norms: bit-equal


def test_host_trainer_matches_committed_golden():
    with open(GOLDEN) as f:
        gold = json.load(f)
    for key, gspo in (("gspo", True), ("grpo", False)):
        got = _run(ht.HostTrainer, gspo)
        assert got["loss"] == float(gold[key]["loss"])
        assert got["metrics"] == {k: float(v) for k, v in gold[key]["metrics"].items()}


@needs_ref
def test_golden_is_what_the_live_reference_gives():
    with open(GOLDEN) as f:
        gold = json.load(f)
    got = _run(_ref_class(), True)
    assert got["loss"] == float(gold["gspo"]["loss"])


# ------------------------------------------------------------------------------------------ level 1
def _mixin_on(base):
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    return type("O3V" + base.__name__, (O3VB200TrainerMixin, base), {})


@pytest.mark.parametrize("which", ["reference", "host"])
def test_mixin_compute_loss_projects_completion_rows_only(which, monkeypatch):
    if which == "reference" and not ref_import.reference_available():
        pytest.skip("reference not present")
    base = _ref_class() if which == "reference" else ht.HostTrainer
    want = _run(base)
    d = Doubles(monkeypatch)
    got, model, t = _run(_mixin_on(base), return_model=True)
    B, G, Tc = 1, 4, model.completion_len
    assert d.rows == [B * G * Tc, B * G * Tc]                   # policy pass + reference pass, no prompt rows
    assert t.o3v_prompt_length is None                          # reset after the step
    assert "generate" not in model.__dict__ and "forward" not in model.lm_head.__dict__
    assert got["metrics"].keys() == want["metrics"].keys()
    assert abs(got["loss"] - want["loss"]) <= 1e-5 * abs(want["loss"])
    for k, v in want["metrics"].items():
        assert abs(got["metrics"][k] - v) <= 1e-5 * max(abs(v), 1e-3), k
    assert abs(got["g_head"] - want["g_head"]) <= 1e-4 * want["g_head"]
    assert abs(got["g_embed"] - want["g_embed"]) <= 1e-4 * want["g_embed"]


def test_mixin_without_compute_loss_keeps_the_full_contract(monkeypatch):
    """`_get_per_token_logps` alone (no prompt length known): [B, L-1], every position, == the reference method."""
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    d = Doubles(monkeypatch)
    model, _ = _models()
    ids = torch.randint(8, 1000, (3, 17))
    pv = torch.randn(16, 12)
    want = ht.HostTrainer()._get_per_token_logps(model, ids, pixel_values=pv)
    got = O3VB200TrainerMixin()._get_per_token_logps(model, ids, pixel_values=pv)
    assert got.shape == (3, 16) and d.rows == [3 * 16]
    assert torch.allclose(got, want, atol=1e-5)
    m = O3VB200TrainerMixin()
    m.o3v_prompt_length = 9
    part = m._get_per_token_logps(model, ids, pixel_values=pv)
    assert part.shape == (3, 16) and d.rows[-1] == 3 * (17 - 9)        # the 8 completion tokens of each row
    assert torch.equal(part[:, :8], torch.zeros(3, 8)) and torch.allclose(part[:, 8:], want[:, 8:], atol=1e-5)


def test_compute_loss_signature_and_return_outputs_error():
    import inspect
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    sig = inspect.signature(O3VB200TrainerMixin.compute_loss)
    assert list(sig.parameters) == ["self", "model", "inputs", "return_outputs", "num_items_in_batch"]
    assert sig.parameters["return_outputs"].default is False and sig.parameters["num_items_in_batch"].default is None
    with pytest.raises(ValueError, match="does not support returning outputs"):
        _mixin_on(ht.HostTrainer)().compute_loss(None, [], return_outputs=True)


# ------------------------------------------------------------------------------------------ level 2
def _load_patched(mode, tmp_path):
    spec = importlib.util.spec_from_file_location("patch_reference", os.path.join(ROOT, "integration", "patch_reference.py"))
    pr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pr)
    src_path = os.path.join(ref_import.REFERENCE_ROOT, "src", "r1-v", "src", "open_r1", "trainer", "grpo_trainer.py")
    with open(src_path) as f:
        src = f.read()
    out = pr.patch_source(src, mode)
    with pytest.raises(pr.PatchError):
        pr.patch_source(out, mode)                              # refuses to patch twice
    ref_import.load_trainer_class()                             # installs the trl stubs + sys.path
    dst = tmp_path / ("grpo_trainer_o3v_%s.py" % mode)
    dst.write_text(out)
    spec = importlib.util.spec_from_file_location("grpo_trainer_o3v_%s" % mode, str(dst))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ht.install_reference_fakes(mod)
    return mod.Qwen2VLGRPOTrainer


@needs_ref
@pytest.mark.parametrize("mode", ["calls", "fused"])
def test_patched_reference_file_runs_and_matches(mode, tmp_path, monkeypatch):
    want = _run(_ref_class())
    d = Doubles(monkeypatch)
    cls = _load_patched(mode, tmp_path)
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    assert issubclass(cls, O3VB200TrainerMixin)
    got, model, _ = _run(cls, return_model=True)
    n = 4 * model.completion_len
    assert d.rows == [n, n] and d.fused_steps == (1 if mode == "fused" else 0)
    assert abs(got["loss"] - want["loss"]) <= 1e-5 * abs(want["loss"])
    for k, v in want["metrics"].items():
        assert abs(got["metrics"][k] - v) <= 1e-5 * max(abs(v), 1e-3), k
    assert abs(got["g_head"] - want["g_head"]) <= 1e-4 * want["g_head"]
    assert abs(got["g_embed"] - want["g_embed"]) <= 1e-4 * want["g_embed"]


@pytest.mark.parametrize("cls", [ht.HostTrainerPatched, ht.HostTrainerFused])
def test_patched_host_sequences_match_the_unpatched_one(cls, monkeypatch):
    want = _run(ht.HostTrainer)
    Doubles(monkeypatch)
    got = _run(_mixin_on(cls))
    assert abs(got["loss"] - want["loss"]) <= 1e-5 * abs(want["loss"])
    assert abs(got["g_embed"] - want["g_embed"]) <= 1e-4 * want["g_embed"]


# ------------------------------------------------------------------------------------------ wrappers
class PeftLike(torch.nn.Module):
    """Forwards unknown attributes to the wrapped model the way peft.PeftModel does."""

    def __init__(self, model):
        super().__init__()
        self.base_model = torch.nn.ModuleDict({"model": model})

    def __getattr__(self, name):
        try:
            return super().__getattr__(name)
        except AttributeError:
            return getattr(self.base_model["model"], name)

    def forward(self, *a, **kw):
        return self.base_model["model"](*a, **kw)


def test_head_is_found_and_restored_through_a_peft_like_wrapper(monkeypatch):
    from open_o3_video_b200.trainer import O3VB200TrainerMixin, find_lm_head, lm_head_weight
    d = Doubles(monkeypatch)
    inner, _ = _models()
    model = PeftLike(inner)
    keys = set(model.state_dict().keys())
    assert find_lm_head(model) is inner.lm_head and lm_head_weight(model) is inner.lm_head.weight
    ids = torch.randint(8, 1000, (2, 9))
    got = O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    want = ht.HostTrainer()._get_per_token_logps(inner, ids)
    assert d.rows == [2 * 8] and torch.allclose(got, want, atol=1e-5)
    assert set(model.state_dict().keys()) == keys               # no second lm_head registered on the wrapper
    assert "lm_head" not in model._modules and "forward" not in inner.lm_head.__dict__


def test_head_forward_is_restored_on_error():
    from open_o3_video_b200.trainer import final_hidden_states

    class Broken(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.lm_head = torch.nn.Linear(4, 8, bias=False)

        def forward(self, input_ids, **kw):
            raise ValueError("boom")

    m = Broken()
    with pytest.raises(ValueError):
        final_hidden_states(m, torch.zeros(1, 2, dtype=torch.long))
    assert "forward" not in m.lm_head.__dict__ and isinstance(m.lm_head, torch.nn.Linear)


class _SavedWeightLogprob(torch.autograd.Function):
    """Double with the autograd structure of logprob._FusedLogprobFn: saves the weight PARAMETER and reads it again
    in backward."""

    @staticmethod
    def forward(ctx, hidden, weight, targets):
        ctx.save_for_backward(hidden, weight, targets)
        z = hidden.float() @ weight.float().T
        return z.log_softmax(-1).gather(1, targets[:, None])[:, 0]

    @staticmethod
    def backward(ctx, g):
        hidden, weight, targets = ctx.saved_tensors
        assert weight.numel() > 0, "backward ran while the weight was partitioned"
        z = hidden.float() @ weight.float().T
        p = -z.softmax(-1) * g[:, None]
        p[torch.arange(len(targets)), targets] += g
        return (p @ weight.float()).to(hidden.dtype), (p.T @ hidden.float()).to(weight.dtype), None


def test_zero3_like_partitioned_head_weight(monkeypatch):
    """DeepSpeed ZeRO-3 keeps a 0-element placeholder in lm_head.weight outside the module's forward / backward and
    gathers it in module hooks.  Because the drop-in computes INSIDE lm_head.forward and returns through the module's
    output, those hooks bracket both the fused forward and its backward."""
    from open_o3_video_b200 import logprob
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    monkeypatch.setattr(logprob, "fused_logprob", lambda h, w, t, **kw: _SavedWeightLogprob.apply(h, w, t))
    model = ht.FakeVLModel(seed=5).bfloat16()
    ids = torch.randint(8, 1000, (2, 9))
    want = ht.HostTrainer()._get_per_token_logps(model.float(), ids).detach()
    model = model.bfloat16()
    w = model.lm_head.weight
    full = w.data.clone()
    w.ds_id, w.ds_shape = 0, full.shape
    empty = torch.empty(0, dtype=full.dtype)
    events = []

    def gather(*_):
        events.append("gather")
        w.data = full

    def release(*_):
        events.append("release")
        w.data = empty

    w.data = empty
    # without hooks: a specific error, not a shape crash deep inside a kernel wrapper
    with pytest.raises(RuntimeError, match="ZeRO-3 partitioned placeholder"):
        O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    model.lm_head.register_forward_pre_hook(gather)
    model.lm_head.register_forward_hook(release)
    model.lm_head.register_full_backward_pre_hook(gather)
    model.lm_head.register_full_backward_hook(release)
    got = O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    assert w.numel() == 0 and events == ["gather", "release"]
    got.sum().backward()
    assert events[2] == "gather" and w.grad is not None and tuple(w.grad.shape) == tuple(full.shape)
    assert torch.allclose(got.float(), want, atol=0.05)


# ------------------------------------------------------------------------------------------ real HF class
def test_head_patch_on_real_qwen2_5_vl_with_video_inputs(monkeypatch):
    tf = pytest.importorskip("transformers")
    if not hasattr(tf, "Qwen2_5_VLForConditionalGeneration"):
        pytest.skip("transformers without Qwen2.5-VL")
    from open_o3_video_b200.trainer import O3VB200TrainerMixin, final_hidden_states, lm_head_weight
    cfg = tf.Qwen2_5_VLConfig(
        text_config=dict(vocab_size=600, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                         num_key_value_heads=2, max_position_embeddings=256,
                         rope_scaling={"type": "mrope", "mrope_section": [2, 3, 3]}),
        vision_config=dict(depth=1, hidden_size=32, intermediate_size=64, num_heads=2, out_hidden_size=64,
                           fullatt_block_indexes=[0], window_size=56),
        video_token_id=599, image_token_id=598, vision_start_token_id=597, vision_end_token_id=596)
    torch.manual_seed(0)
    model = tf.Qwen2_5_VLForConditionalGeneration(cfg).eval()
    vc = cfg.vision_config
    ids = torch.randint(0, 500, (2, 12))
    ids[:, 2], ids[:, 3:7], ids[:, 7] = 597, 599, 596                   # <vision_start> 4 video tokens <vision_end>
    kwargs = dict(pixel_values_videos=torch.randn(2 * 16, 3 * vc.temporal_patch_size * vc.patch_size ** 2),
                  video_grid_thw=torch.tensor([[1, 4, 4], [1, 4, 4]]), attention_mask=torch.ones_like(ids))
    with torch.no_grad():
        logits = model(ids, **kwargs).logits                            # what the reference materialises
        hidden = final_hidden_states(model, ids, **kwargs)
        text_only = final_hidden_states(model, ids)
    assert "forward" not in model.lm_head.__dict__                      # the head is back to normal
    assert hidden.shape == (2, 12, 64)
    assert torch.allclose(hidden.float() @ lm_head_weight(model).float().T, logits.float(), atol=1e-4)
    assert not torch.allclose(hidden, text_only)                        # the vision tower really ran
    # the whole drop-in method against the reference method on the real class
    d = Doubles(monkeypatch)
    with torch.no_grad():
        want = ht.HostTrainer()._get_per_token_logps(model, ids, **kwargs)
        got = O3VB200TrainerMixin()._get_per_token_logps(model, ids, **kwargs)
    assert d.rows == [2 * 11] and torch.allclose(got, want, atol=2e-2)   # bf16 cast of hidden / W in the drop-in
