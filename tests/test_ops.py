"""`torch.ops.o3v.*` (open-o3-video_b200/ops.py): schemas and fake implementations on the CPU box, equality
with the plain wrappers and `torch.library.opcheck` on the GPU."""
import numpy as np
import pytest
import torch

from oracle import gspo as ogspo
from oracle import synth


def test_ops_registered_with_schemas_and_fake_kernels():
    from open_o3_video_b200 import ops  # noqa: F401
    from torch._subclasses.fake_tensor import FakeTensorMode
    for name in ("lmhead_logprob", "lmhead_logprob_backward", "eos_mask", "gspo_objective", "policy_step"):
        assert hasattr(torch.ops.o3v, name), name
    assert "Tensor? old_logp" in str(torch.ops.o3v.policy_step.default._schema)
    with FakeTensorMode():
        dev = "cuda"
        h = torch.empty(64, 128, dtype=torch.bfloat16, device=dev)
        w = torch.empty(1000, 128, dtype=torch.bfloat16, device=dev)
        t = torch.empty(64, dtype=torch.int64, device=dev)
        lp, lse, lg = torch.ops.o3v.lmhead_logprob(h, w, t, 0, True)
        assert lp.shape == (64,) and lse.dtype == torch.float32 and lg.shape == (64, 1000) and lg.dtype == torch.bfloat16
        assert torch.ops.o3v.lmhead_logprob(h, w, t, 0, False)[2].shape == (0, 1000)
        dh, dw = torch.ops.o3v.lmhead_logprob_backward(lp, h, w, t, lse, lg, 0, 32)
        assert dh.shape == h.shape and dh.dtype == torch.bfloat16 and dw.shape == w.shape and dw.dtype == torch.float32
        ids = torch.empty(4, 16, dtype=torch.int64, device=dev)
        idx, mask = torch.ops.o3v.eos_mask(ids, 3)
        assert idx.shape == (4,) and mask.shape == (4, 16) and mask.dtype == torch.int32
        f = torch.empty(4, 16, device=dev)
        out = torch.ops.o3v.gspo_objective(f, f, mask, torch.empty(4, 3, device=dev), None, 2, 0.04, 0.2, 0.2, True)
        assert [tuple(x.shape) for x in out] == [(), (4, 16), (4,), (), (4,), (4,)]
        out = torch.ops.o3v.policy_step(torch.empty(4, 16, 128, dtype=torch.bfloat16, device=dev), w, ids, f, mask,
                                        torch.empty(4, 3, device=dev), None, 2, 0.04, 0.2, 0.2, True, 32768)
        assert [tuple(x.shape) for x in out] == [(), (4, 16), (4,), (), (4, 16, 128), (1000, 128)]


def test_ops_refuse_cpu_tensors():
    from open_o3_video_b200 import ops  # noqa: F401
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        torch.ops.o3v.eos_mask(torch.zeros(2, 4, dtype=torch.int64), 1)


@pytest.mark.gpu
def test_ops_match_the_wrappers_and_pass_opcheck(monkeypatch):
    from open_o3_video_b200 import gspo, logprob, ops
    T, H, V = 200, 256, 5000
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=3)
    h = hidden.cuda().bfloat16()
    w = weight.cuda().bfloat16()
    t = targets.cuda()
    ha, wa = h.clone().requires_grad_(True), w.clone().requires_grad_(True)
    hb, wb = h.clone().requires_grad_(True), w.clone().requires_grad_(True)
    # the registered operators keep the bf16 LOGITS and run the in-place softmax-backward pass ("dlogits" mode of the
    # wrappers: bit-identical); the wrappers' default "exp" mode keeps exp(z - ref) instead (same bf16 bar)
    monkeypatch.setattr(logprob, "BACKWARD", "dlogits")
    la = logprob.fused_logprob(ha, wa, t)
    lb = ops.fused_logprob(hb, wb, t)
    assert torch.equal(la, lb)
    g = torch.randn(T, device="cuda") * 1e-3
    la.backward(g)
    lb.backward(g)
    assert torch.equal(ha.grad, hb.grad) and torch.equal(wa.grad, wb.grad)
    monkeypatch.setattr(logprob, "BACKWARD", "exp")
    hc, wc = h.clone().requires_grad_(True), w.clone().requires_grad_(True)
    lc = logprob.fused_logprob(hc, wc, t)
    assert torch.equal(lc, lb)
    lc.backward(g)
    for a, b in ((hc.grad, hb.grad), (wc.grad, wb.grad)):
        assert ((a.float() - b.float()).norm() / b.float().norm()).item() < 1e-2
    # recompute path of the backward operator (forward kept no logits) == saved-logits path
    lp, lse, _ = torch.ops.o3v.lmhead_logprob(h, w, t, 0, False)
    dh, dw = torch.ops.o3v.lmhead_logprob_backward(g, h, w, t, lse, torch.empty(0, V, dtype=torch.bfloat16, device="cuda"),
                                                   0, 64)
    np.testing.assert_allclose(dh.float().cpu().numpy(), ha.grad.float().cpu().numpy(), rtol=0, atol=2e-2 * ha.grad.abs().max().item())
    torch.library.opcheck(torch.ops.o3v.lmhead_logprob, (h, w, t, 0, True), test_utils=("test_schema", "test_faketensor"))
    torch.library.opcheck(torch.ops.o3v.lmhead_logprob_backward, (g, h, w, t, lse, torch.empty(0, V, dtype=torch.bfloat16, device="cuda"), 0, 64),
                          test_utils=("test_schema", "test_faketensor"))
    # objective and whole step
    N, Tc, G = 8, 25, 4
    d = synth.gspo_inputs(N, Tc, G, off_policy=True, seed=9)
    _, mask = ogspo.eos_mask(d["ids"], d["eos_id"])
    cu = lambda x: x.cuda()
    idx, m2 = torch.ops.o3v.eos_mask(cu(d["ids"]), d["eos_id"])
    assert torch.equal(m2.cpu(), mask)
    loss, grad, adv, kl, clen, rstd = torch.ops.o3v.gspo_objective(cu(d["logp"]), cu(d["ref"]), m2, cu(d["rewards_per_func"]),
                                                                   cu(d["old"]), G, 0.04, 0.2, 0.2, True)
    lp_g = cu(d["logp"]).requires_grad_(True)
    ref = gspo.gspo_loss(lp_g, cu(d["ref"]), m2, cu(d["rewards_per_func"]), G, 0.04, 0.2, 0.2, True, cu(d["old"]))
    ref.loss.backward()
    assert torch.equal(loss, ref.loss.detach()) and torch.equal(grad, lp_g.grad) and torch.equal(adv, ref.advantages)
    torch.library.opcheck(torch.ops.o3v.gspo_objective, (cu(d["logp"]), cu(d["ref"]), m2, cu(d["rewards_per_func"]), None, G,
                                                         0.04, 0.2, 0.2, True), test_utils=("test_schema", "test_faketensor"))
    hid = h[:N * Tc].view(N, Tc, H)
    ids = (cu(d["ids"]) % V)
    a = logprob.fused_logprob_gspo(hid, w, ids, cu(d["ref"]), m2, cu(d["rewards_per_func"]), G, 0.04, chunk_tokens=100)
    b = torch.ops.o3v.policy_step(hid, w, ids, cu(d["ref"]), m2, cu(d["rewards_per_func"]), None, G, 0.04, 0.2, 0.2, True, 100)
    for x, k in zip(b, ("loss", "per_token_logps", "advantages", "mean_kl", "d_hidden", "d_weight")):
        assert torch.equal(x, a[k]), k


@pytest.mark.gpu
def test_ops_trace_under_torch_compile_without_graph_breaks():
    """fullgraph=True: dynamo + AOT autograd must trace through the registered operators (fake kernels, autograd
    formula) with no graph break; results equal eager."""
    from open_o3_video_b200 import ops
    T, H, V = 128, 128, 2000
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=11)
    h0, w, t = hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda()

    def head(h, w, t):
        lp, lse, _ = torch.ops.o3v.lmhead_logprob(h * 1.0, w, t, 0, True)
        return (lp * 2.0).sum()

    ha = h0.clone().requires_grad_(True)
    eager = head(ha, w, t)
    eager.backward()
    hb = h0.clone().requires_grad_(True)
    compiled = torch.compile(head, fullgraph=True, backend="aot_eager")(hb, w, t)
    compiled.backward()
    assert torch.equal(eager, compiled) and torch.equal(ha.grad, hb.grad)
