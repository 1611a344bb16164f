"""GPU parity of the whole fused step (`fused_logprob_gspo`) against the oracle: loss and
log-probs within the north_star tolerances, gradients against autograd through the oracle."""
import numpy as np
import pytest
import torch

from oracle import gspo as ogspo
from oracle import logps as ologps
from oracle import synth

pytestmark = pytest.mark.gpu


def _oracle_step(hidden, weight, ids, ref, mask, rpf, G, beta, old, gspo=True):
    N, Tc, H = hidden.shape
    h = hidden.clone().requires_grad_(True)
    w = weight.clone().requires_grad_(True)
    z = h.view(N * Tc, H) @ w.T
    lp = z.log_softmax(-1).gather(1, ids.view(-1, 1))[:, 0].view(N, Tc)
    out = ogspo.gspo_step(lp, ref, mask, rpf, G, beta, 0.2, 0.2, gspo, old)
    out["loss"].backward()
    return out, lp.detach(), h.grad, w.grad


def _inputs(N, Tc, G, H, V, off_policy, seed=0):
    hidden, weight, targets = synth.head_inputs(N * Tc, H, V, seed=seed + 17)
    d = synth.gspo_inputs(N, Tc, G, off_policy=off_policy, vocab=V + 1000, eos_id=V - 1, seed=seed)
    ids = d["ids"] % V
    _, mask = ogspo.eos_mask(ids, V - 1)
    lp_true, _ = ologps.token_logps(hidden, weight, ids.view(-1))
    lp_true = lp_true.view(N, Tc)
    ref = lp_true + (d["ref"] - d["logp"])
    old = None if d["old"] is None else lp_true + (d["old"] - d["logp"])
    return hidden.view(N, Tc, H), weight, ids, ref, mask, d["rewards_per_func"], old


@pytest.mark.parametrize("cta", [1, 2])
@pytest.mark.parametrize("off_policy", [False, True])
@pytest.mark.parametrize("N,Tc,G,H,V,chunk", [(8, 64, 4, 128, 1024, 128), (8, 100, 4, 256, 5000, 300),
                                              (4, 512, 4, 512, 152064 // 16, 1 << 20)])
def test_fused_step_matches_oracle(N, Tc, G, H, V, chunk, off_policy, cta):
    from open_o3_video_b200 import _lib, logprob
    _lib.set_tunable("cta_pair", cta)
    try:
        hidden, weight, ids, ref, mask, rpf, old = _inputs(N, Tc, G, H, V, off_policy)
        exp, lp_ref, dH_ref, dW_ref = _oracle_step(hidden, weight, ids, ref, mask, rpf, G, 0.04, old)
        cu = lambda t: None if t is None else t.cuda()
        out = logprob.fused_logprob_gspo(hidden.cuda().bfloat16(), weight.cuda().bfloat16(), ids.cuda(), cu(ref),
                                         cu(mask), cu(rpf), G, 0.04, 0.2, 0.2, True, cu(old), chunk_tokens=chunk,
                                         fuse_dlogits=(cta == 2 and off_policy))   # both backward variants (fused: pair tiles only)
        torch.cuda.synchronize()
    finally:
        _lib.set_tunable("cta_pair_fwd", 1)
        _lib.set_tunable("cta_pair_bwd", 2)
    lp = out["per_token_logps"].cpu()
    assert ((lp - lp_ref).abs() / lp_ref.abs().clamp(min=1e-6)).max() < 1e-3
    np.testing.assert_allclose(out["loss"].item(), exp["loss"].item(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(out["mean_kl"].item(), exp["mean_kl"].item(), rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(out["advantages"].cpu().numpy(), exp["advantages"].numpy(), rtol=2e-5, atol=2e-6)
    assert torch.equal(out["completion_length"].cpu(), exp["completion_length"].to(torch.int32))
    dH = out["d_hidden"].float().cpu().view_as(dH_ref)
    dW = out["d_weight"].cpu()
    # gradients flow through bf16 logits (as in the reference's bf16 path): 1% in norm
    assert (dH - dH_ref).norm() / dH_ref.norm() < 1e-2
    assert (dW - dW_ref).norm() / dW_ref.norm() < 1e-2
    assert (dH - dH_ref).abs().max() < 3e-2 * dH_ref.abs().max()
    assert (dW - dW_ref).abs().max() < 3e-2 * dW_ref.abs().max()
    # masked-out tokens get exactly zero gradient
    assert (dH.view(N, Tc, H)[mask == 0] == 0).all()


@pytest.mark.parametrize("N,Tc,chunk", [(8, 64, 128), (9, 100, 300), (6, 256, 256)])
def test_pipelined_dlogits_is_bit_identical(N, Tc, chunk):
    """overlap_dlogits only reorders launches across streams: every output must be bit-identical."""
    from open_o3_video_b200 import logprob
    G = 3 if N == 9 else 2
    hidden, weight, ids, ref, mask, rpf, old = _inputs(N, Tc, G, 256, 5000, True, seed=5)
    args = (hidden.cuda().bfloat16(), weight.cuda().bfloat16(), ids.cuda(), ref.cuda(), mask.cuda(), rpf.cuda(), G, 0.04,
            0.2, 0.2, True, old.cuda())
    a = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk, overlap_dlogits=False, fuse_dlogits=False)
    for _ in range(3):                                        # repeated: buffer reuse across steps and streams
        b = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk, overlap_dlogits=True, fuse_dlogits=False)
    torch.cuda.synchronize()
    for k in ("loss", "per_token_logps", "advantages", "mean_kl", "d_hidden", "d_weight"):
        assert torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("N,Tc,chunk", [(8, 64, 128), (9, 100, 300)])
def test_fused_and_separate_softmax_backward_agree(N, Tc, chunk):
    """The whole step with the softmax backward inside the GEMMs (default) vs the separate dlogits pass: forward
    outputs bit-identical, gradients equal up to bf16 rounding of P (ex2.approx vs exp2f)."""
    from open_o3_video_b200 import _lib, logprob
    G = 3 if N == 9 else 2
    hidden, weight, ids, ref, mask, rpf, old = _inputs(N, Tc, G, 256, 5000, True, seed=6)
    args = (hidden.cuda().bfloat16(), weight.cuda().bfloat16(), ids.cuda(), ref.cuda(), mask.cuda(), rpf.cuda(), G, 0.04,
            0.2, 0.2, True, old.cuda())
    t = _lib.Trace()
    _lib.trace = t
    try:
        a = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk, fuse_dlogits=True)
    finally:
        _lib.trace = None
    assert not any(n == "o3v_lmhead_dlogits" for n, _ in t.calls)     # no elementwise pass over the logits chunk
    b = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk, fuse_dlogits=False)
    for k in ("loss", "per_token_logps", "advantages", "mean_kl"):
        assert torch.equal(a[k], b[k]), k
    for k in ("d_hidden", "d_weight"):
        x, y = a[k].float(), b[k].float()
        assert (x - y).norm() <= 2e-3 * y.norm(), k
    assert (a["d_hidden"][mask.cuda() == 0] == 0).all()


@pytest.mark.parametrize("N,Tc,chunk", [(8, 64, 128), (9, 100, 300)])
def test_exp_store_and_dlogits_backward_agree(N, Tc, chunk):
    """The default exp-store backward vs the round-1 dlogits pass: forward outputs bit-identical (K1's statistics do not
    depend on what it stores), gradients equal up to the bf16 rounding of the stored operand; no launch touches the
    [T, V] chunk between K1 and the two GEMMs."""
    from open_o3_video_b200 import _lib, logprob
    G = 3 if N == 9 else 2
    hidden, weight, ids, ref, mask, rpf, old = _inputs(N, Tc, G, 256, 5000, True, seed=7)
    args = (hidden.cuda().bfloat16(), weight.cuda().bfloat16(), ids.cuda(), ref.cuda(), mask.cuda(), rpf.cuda(), G, 0.04,
            0.2, 0.2, True, old.cuda())
    t = _lib.Trace()
    _lib.trace = t
    try:
        a = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk)
    finally:
        _lib.trace = None
    names = [n for n, _ in t.calls]
    assert "o3v_lmhead_dlogits" not in names and "o3v_lmhead_bwd_dhidden_exp" in names and "o3v_lmhead_bwd_dweight_exp" in names
    b = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk, backward="dlogits")
    for k in ("loss", "per_token_logps", "advantages", "mean_kl"):
        assert torch.equal(a[k], b[k]), k
    for k in ("d_hidden", "d_weight"):
        x, y = a[k].float(), b[k].float()
        assert (x - y).norm() <= 5e-3 * y.norm(), k
    assert (a["d_hidden"][mask.cuda() == 0] == 0).all()
    c = logprob.fused_logprob_gspo(*args, chunk_tokens=chunk)         # run to run bit-stable
    assert torch.equal(a["d_hidden"], c["d_hidden"]) and torch.equal(a["d_weight"], c["d_weight"])


def test_c1_shape_known_answer():
    """BASELINE config 1 / SURVEY Appendix B: 7B head, 1 x 4 x 512 tokens, on-policy,
    ref = logp + 0.1, rewards [0.5, 2.0, 1.25, 3.0], full mask -> loss = 0.000207."""
    from open_o3_video_b200 import logprob
    H, V, N, Tc = 3584, 152064, 4, 512
    hidden, weight, targets = synth.head_inputs(N * Tc, H, V)
    h, w, t = hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda()
    del hidden, weight
    lp = logprob.fused_logprob(h, w, t).view(N, Tc)
    rpf = torch.tensor([[0.5], [2.0], [1.25], [3.0]], device="cuda")
    mask = torch.ones(N, Tc, dtype=torch.int32, device="cuda")
    out = logprob.fused_logprob_gspo(h.view(N, Tc, H), w, t.view(N, Tc), lp + 0.1, mask, rpf, 4, 0.04)
    assert abs(out["loss"].item() - 0.000207) < 2e-6
    assert torch.equal(out["per_token_logps"], lp)                     # deterministic across calls
    # reference arithmetic on the same inputs, on the GPU in fp32 (too slow for CI on the host)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.empty(N * Tc, device="cuda")
    for s in range(0, N * Tc, 512):
        z = h[s:s + 512].float() @ w.float().T
        ref[s:s + 512] = z.log_softmax(-1).gather(1, t[s:s + 512, None])[:, 0]
    assert ((lp.view(-1) - ref).abs() / ref.abs()).max().item() < 1e-3
    # ... and the BACKWARD at the full head against fp32 on the GPU: d loss / d logp from the oracle's autograd on the
    # library's log-probs, P = g (onehot - softmax) from fp32 logits, dHidden of every token, 512 columns of dW
    lpc = lp.detach().cpu().requires_grad_(True)
    exp = ogspo.gspo_step(lpc, (lp + 0.1).cpu(), mask.cpu(), rpf.cpu(), 4, 0.04)
    exp["loss"].backward()
    g = lpc.grad.cuda().view(-1)
    wf = w.float()
    cols = torch.arange(70000, 70512, device="cuda")
    dH_ref = torch.empty(N * Tc, H, device="cuda")
    dW_ref = torch.zeros(512, H, device="cuda")
    for s in range(0, N * Tc, 512):
        z = h[s:s + 512].float() @ wf.T
        P = torch.softmax(z, -1).mul_(-g[s:s + 512, None])
        P[torch.arange(512, device="cuda"), t[s:s + 512]] += g[s:s + 512]
        dH_ref[s:s + 512] = P @ wf
        dW_ref += P[:, cols].T @ h[s:s + 512].float()
        del z, P
    dH = out["d_hidden"].float().view(N * Tc, H)
    assert ((dH - dH_ref).norm() / dH_ref.norm()).item() < 1e-2
    assert ((out["d_weight"][cols] - dW_ref).norm() / dW_ref.norm()).item() < 1e-2


def test_trainer_mixin_keeps_reference_contract():
    """`_get_per_token_logps(self, model, input_ids, **kwargs) -> [B, L-1]` (grpo_trainer.py:371) on a
    fake model whose backbone returns fixed hidden states, against the oracle; plus the
    prompt-skipping variant and the loss/metrics block."""
    import types
    from open_o3_video_b200.trainer import O3VB200TrainerMixin
    B, L, H, V, G = 4, 40, 128, 3000, 4
    hidden, weight, _ = synth.head_inputs(B * L, H, V, seed=31)
    ids = torch.randint(0, V, (B, L), generator=torch.Generator().manual_seed(32))
    exp = ologps.per_token_logps(hidden.view(B, L, H), weight, ids)

    class Backbone(torch.nn.Module):          # text-only, as at the transformers commit the reference pins
        def forward(self, inputs_embeds=None):
            return types.SimpleNamespace(last_hidden_state=inputs_embeds)

    class Model(torch.nn.Module):             # ...ForConditionalGeneration: merges vision features, then lm_head
        def __init__(self):
            super().__init__()
            self.model = Backbone()
            self.lm_head = torch.nn.Linear(H, V, bias=False).cuda().bfloat16()
            self.lm_head.weight.data.copy_(weight)
            self.seen = None

        def forward(self, input_ids, pixel_values_videos=None, video_grid_thw=None):
            self.seen = (pixel_values_videos, video_grid_thw)
            h = self.model(inputs_embeds=hidden.view(B, L, H).cuda().bfloat16()).last_hidden_state
            logits = self.lm_head(h)
            return types.SimpleNamespace(logits=logits.float())

    class T(O3VB200TrainerMixin):
        num_generations, beta, epsilon_low, epsilon_high, gspo = G, 0.04, 0.2, 0.2, True
        reward_funcs = []

    t, model = T(), torch.nn.parallel.DataParallel(Model()) if False else Model()
    lp = t._get_per_token_logps(model, ids.cuda(), pixel_values_videos="pix", video_grid_thw="thw")
    assert model.seen == ("pix", "thw")                               # the reference's kwargs reach model.forward
    assert isinstance(model.lm_head, torch.nn.Linear)                 # the head is back in place
    assert lp.shape == (B, L - 1)
    assert ((lp.cpu() - exp).abs() / exp.abs().clamp(min=1e-2)).max() < 1e-3
    t.o3v_prompt_length = 25
    lp2 = t._get_per_token_logps(model, ids.cuda())
    assert lp2.shape == (B, L - 1) and torch.equal(lp2[:, 24:], lp[:, 24:]) and (lp2[:, :24] == 0).all()
    # loss + metrics keys of grpo_trainer.py:711-738
    comp = lp[:, 24:]
    mask = torch.ones_like(comp, dtype=torch.int32)
    rpf = torch.rand(B, 2, device="cuda")
    loss = t.compute_policy_loss(comp.detach().requires_grad_(True), comp.detach() + 0.1, mask, rpf)
    ref = ogspo.gspo_step(comp.cpu(), comp.cpu() + 0.1, mask.cpu(), rpf.cpu(), G, 0.04)
    assert abs(loss.item() - ref["loss"].item()) < 1e-6
    assert set(t._metrics) >= {"completion_length", "all_wrong", "all_correct", "reward", "reward_std", "kl"}
    assert abs(t._metrics["kl"][0] - ref["mean_kl"].item()) < 1e-6


@pytest.mark.parametrize("num_items", [None, 37])
def test_sft_cross_entropy_matches_oracle(num_items):
    """the SFT loss on the fused head, with -100 labels (pad / visual tokens)."""
    from open_o3_video_b200 import logprob
    from oracle import sft as osft
    B, L, H, V = 3, 50, 128, 3000
    hidden, weight, _ = synth.head_inputs(B * L, H, V, seed=41)
    g = torch.Generator().manual_seed(42)
    labels = torch.randint(0, V, (B, L), generator=g)
    labels[torch.rand(B, L, generator=g) < 0.4] = -100
    labels[1, 1:] = -100                                       # a fully ignored sequence
    h_ref = hidden.view(B, L, H).clone().requires_grad_(True)
    w_ref = weight.clone().requires_grad_(True)
    exp = osft.causal_lm_loss(h_ref, w_ref, labels, num_items_in_batch=num_items)
    exp.backward()
    h = hidden.view(B, L, H).cuda().bfloat16().requires_grad_(True)
    w = weight.cuda().bfloat16().requires_grad_(True)
    loss = logprob.sft_cross_entropy(h, w, labels.cuda(), num_items_in_batch=num_items)
    loss.backward()
    assert abs(loss.item() - exp.item()) <= 1e-3 * abs(exp.item())
    assert (h.grad.float().cpu() - h_ref.grad).norm() / h_ref.grad.norm() < 1e-2
    assert (w.grad.float().cpu() - w_ref.grad).norm() / w_ref.grad.norm() < 1e-2
    assert (h.grad[1] == 0).all() and (h.grad[:, -1] == 0).all()
