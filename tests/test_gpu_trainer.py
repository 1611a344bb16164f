"""The drop-in trainer mixin with the REAL kernels (B200).

The reference does not exist on the GPU box, so the host class is `oracle.host_trainer.HostTrainer`, the
restatement of the reference's compute_loss call sequence that tests/test_trainer_cpu.py pins bit-for-bit against
the live reference class (and against tests/golden/compute_loss_small.json, produced by the live reference)."""
import json
import os

import pytest
import torch

from oracle import host_trainer as ht

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gold():
    with open(os.path.join(ROOT, "tests", "golden", "compute_loss_small.json")) as f:
        return json.load(f)


def _mixin_on(base):
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    return type("O3V" + base.__name__, (O3VB200TrainerMixin, base), {})


def _run(cls, device, gspo=True):
    from open_o3_video_b200 import _lib
    model, ref = ht.FakeVLModel(seed=11).to(device), ht.FakeVLModel(seed=12).eval().to(device)
    t = ht.configure(cls.__new__(cls), model, ref, device=device, gspo=gspo)
    trace = _lib.Trace()
    _lib.trace = trace
    try:
        loss = t.compute_loss(model, [ht.make_example()])
        loss.backward()
        torch.cuda.synchronize()
    finally:
        _lib.trace = None
    return dict(loss=loss.item(), metrics={k: v[0] for k, v in t._metrics.items()}, model=model, trace=trace, trainer=t)


def _cpu_reference(gspo=True):
    model, ref = ht.FakeVLModel(seed=11), ht.FakeVLModel(seed=12).eval()
    t = ht.configure(ht.HostTrainer(), model, ref, gspo=gspo)
    t.compute_loss(model, [ht.make_example()]).backward()
    return model


@pytest.mark.parametrize("host", ["HostTrainer", "HostTrainerPatched", "HostTrainerFused"])
@pytest.mark.parametrize("gspo", [True, False])
def test_dropin_compute_loss_matches_reference_golden(host, gspo):
    """Level 1 (pure inheritance) and level 2 (patched call sites, both modes): loss and metrics == what the LIVE
    reference produced (golden), 1e-3 relative (bf16 head, north_star); gradients into the backbone and the head
    == the fp32 CPU run of the reference sequence; only B*G*Tc completion rows go through K1."""
    gold = _gold()["gspo" if gspo else "grpo"]
    got = _run(_mixin_on(getattr(ht, host)), "cuda", gspo)
    want_loss = float(gold["loss"])
    assert abs(got["loss"] - want_loss) <= 1e-3 * abs(want_loss), (got["loss"], want_loss)
    for k, v in gold["metrics"].items():
        assert abs(got["metrics"][k] - float(v)) <= 1e-3 * max(abs(float(v)), 1e-3), k
    model = got["model"]
    n = 4 * model.completion_len                                        # B=1 prompt x G=4 x Tc
    k1 = [ints[0] for name, ints in got["trace"].calls             # K1 over the whole vocabulary (not the 256-row
          if name.startswith("o3v_lmhead_fwd") and ints[1] == model.vocab]   # sample of the row reference)
    assert k1 and all(rows == n for rows in k1), k1                     # never the 4 x (Lp + Tc - 1) rows of the reference
    assert len(k1) == 2 or (host != "HostTrainerFused" and len(k1) == 3)   # policy + ref (+ recompute in backward)
    names = {name for name, _ in got["trace"].calls}
    if host != "HostTrainer":
        assert "o3v_eos_mask" in names and "o3v_gspo_fwd_bwd" in names
    assert {"o3v_lmhead_bwd_dhidden_exp", "o3v_lmhead_bwd_dweight_exp"} <= names     # the default (exp-store) backward
    assert got["trainer"].o3v_prompt_length is None
    cpu = _cpu_reference(gspo)
    for name in ("embed", "mix", "lm_head"):
        a = getattr(model, name).weight.grad.float().cpu()
        b = getattr(cpu, name).weight.grad
        assert (a - b).norm() <= 2e-2 * b.norm(), (name, ((a - b).norm() / b.norm()).item())


def test_zero3_like_partitioned_head_weight_with_the_real_kernels():
    """The autograd Function behind K1/K2 saves the lm_head weight PARAMETER; with the weight partitioned away
    between forward and backward (ZeRO-3 style module hooks), the backward GEMMs must see the re-gathered data."""
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    model = ht.FakeVLModel(seed=5).cuda()
    model.lm_head.to(torch.bfloat16)
    ids = torch.randint(8, 1000, (2, 33), device="cuda")
    plain = O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    plain.sum().backward()
    g_ref = model.lm_head.weight.grad.clone()
    e_ref = model.embed.weight.grad.clone()
    model.zero_grad()
    w = model.lm_head.weight
    full = w.data.clone()
    empty = torch.empty(0, dtype=full.dtype, device="cuda")
    w.ds_id, w.ds_shape = 0, full.shape

    def gather(*_):
        w.data = full

    def release(*_):
        w.data = empty

    w.data = empty
    with pytest.raises(RuntimeError, match="ZeRO-3 partitioned placeholder"):
        O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    model.lm_head.register_forward_pre_hook(gather)
    model.lm_head.register_forward_hook(release)
    model.lm_head.register_full_backward_pre_hook(gather)
    model.lm_head.register_full_backward_hook(release)
    got = O3VB200TrainerMixin()._get_per_token_logps(model, ids)
    assert w.numel() == 0
    got.sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(got, plain)
    assert torch.equal(w.grad, g_ref) and torch.equal(model.embed.weight.grad, e_ref)
