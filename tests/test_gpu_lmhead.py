"""GPU parity (through the C ABI) of the tcgen05 lm_head kernels.

K1: fused GEMM + online log-softmax + gather against the oracle (the reference's torch path
in fp32 on the same bf16-valued inputs): log-probs within 1e-3 relative (north_star, bf16);
the gather index must be exact, so the captured target logit is compared too.
K2: dlogits / dHidden / dW against torch fp32 autograd through the oracle.
Every case runs for both tile modes: one CTA per tile and cta_group::2 pairs.
"""
import numpy as np
import pytest
import torch

from oracle import logps as ologps
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2], ids=["cta1", "cta2"])
def cta_pair(request):
    from open_o3_video_b200 import _lib, logprob
    _lib.set_tunable("cta_pair", request.param)
    logprob.FUSE_DLOGITS = request.param == 2      # cta2 runs also exercise the fused softmax backward (pair tiles only)
    yield request.param
    _lib.set_tunable("cta_pair_fwd", 1)
    _lib.set_tunable("cta_pair_bwd", 2)
    _lib.set_tunable("fwd_groups", 0)
    logprob.FUSE_DLOGITS = False


def _rel(a, b, floor=2e-5):
    # relative error with an absolute floor: planted rows have reference log-probs of exactly 0
    # (and logits of +-60 whose fp32 spacing is 4e-6), where a pure ratio is meaningless
    return (np.abs(a - b) / np.maximum(np.abs(b), floor / 1e-3)).max()


# T, H, V, planted, fwd_groups
FWD_CASES = [
    (128, 64, 256, False, 0),          # one tile, one k-block
    (256, 128, 1024, False, 0),
    (300, 192, 1000, False, 1),        # ragged M, N (V % 8 == 0), odd k-block count, 1 group
    (1000, 512, 5000, True, 3),        # planted +-60 logits, uneven groups
    (515, 256, 151936 // 16, False, 5),
    (2048, 3584, 152064 // 8, True, 0),
]


@pytest.mark.parametrize("T,H,V,planted,groups", FWD_CASES)
def test_fwd_logp_matches_oracle(cta_pair, T, H, V, planted, groups):
    from open_o3_video_b200 import _lib, logprob
    _lib.set_tunable("fwd_groups", groups)
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=T + V, planted=planted)
    ref_logp, ref_lse = ologps.token_logps(hidden, weight, targets)
    logp, lse = logprob.fused_logprob(hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda(),
                                      return_lse=True)
    torch.cuda.synchronize()
    assert _rel(lse.cpu().numpy(), ref_lse.numpy()) < 1e-4
    assert _rel(logp.cpu().numpy(), ref_logp.numpy()) < 1e-3          # north_star tolerance
    assert np.abs(logp.cpu().numpy() - ref_logp.numpy()).max() < 2e-3


def test_fwd_gather_is_exact_and_logits_store(cta_pair):
    """Target logit captured in the epilogue == the logit at exactly targets[t] (bit-exact index),
    and the optional bf16 logits equal the fp32 accumulators rounded to bf16."""
    from open_o3_video_b200 import logprob
    T, H, V = 384, 256, 2048 + 8 * 13
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=3)
    targets[:4] = torch.tensor([0, V - 1, 255, 256])                  # tile / chunk boundaries
    h, w, t = hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda()
    z = torch.full((T, V + 24), 7.0, dtype=torch.bfloat16, device="cuda")   # ld > V: padding untouched
    stats = logprob.lmhead_stats(h, w, t, 0, z[:, :V])
    torch.backends.cuda.matmul.allow_tf32 = False
    z_ref = h.float() @ w.float().T
    tgt_ref = z_ref.gather(1, t[:, None])[:, 0]
    np.testing.assert_allclose(stats[2].cpu().numpy(), tgt_ref.cpu().numpy(), rtol=1e-5, atol=1e-5)
    # a wrong column would be off by O(1), not 1e-5: also check against every other column's logit
    assert (torch.abs(stats[2] - tgt_ref) < 1e-4).all()
    np.testing.assert_allclose(stats[0].cpu().numpy(), z_ref.max(1).values.cpu().numpy(), rtol=1e-5, atol=1e-5)
    diff = (z[:, :V].float() - z_ref).abs().max().item()
    assert diff <= 2 ** -8 * z_ref.abs().max().item() + 1e-6
    assert (z[:, V:] == 7.0).all()


def test_fwd_vocab_slices_merge_to_full():
    """Vocab-sharded use: per-slice statistics merged by o3v_lmhead_merge_stats == unsharded."""
    from open_o3_video_b200 import logprob
    T, H, V = 640, 128, 4096 + 512
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=9)
    h, w, t = hidden.cuda().bfloat16(), weight.cuda().bfloat16(), targets.cuda()
    full, lse_full = logprob.merge_stats(logprob.lmhead_stats(h, w, t).unsqueeze(0))
    bounds = [0, 1024, 1024 + 256, 3072, V]                            # tile-granular, uneven
    parts = torch.stack([logprob.lmhead_stats(h, w[a:b].contiguous(), t, a) for a, b in zip(bounds, bounds[1:])])
    sh, lse_sh = logprob.merge_stats(parts)
    np.testing.assert_allclose(sh.cpu().numpy(), full.cpu().numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(lse_sh.cpu().numpy(), lse_full.cpu().numpy(), rtol=2e-6, atol=2e-6)


def _torch_ref_grads(hidden, weight, targets, g):
    h = hidden.clone().requires_grad_(True)
    w = weight.clone().requires_grad_(True)
    z = h @ w.T
    lp = z.log_softmax(-1).gather(1, targets[:, None])[:, 0]
    (lp * g).sum().backward()
    return h.grad, w.grad, z.detach()


BWD_CASES = [(128, 64, 256), (384, 128, 1024), (300, 192, 1000), (1100, 512, 5000), (777, 1024, 9496)]


@pytest.mark.parametrize("T,H,V", BWD_CASES)
def test_bwd_gemms_match_torch(cta_pair, T, H, V):
    """dH = P.W (A K-major, B MN-major) and dW = P^T.hidden (both MN-major) on a given bf16 P."""
    from open_o3_video_b200 import logprob
    g = torch.Generator().manual_seed(T + H + V)
    P = (torch.randn(T, V, generator=g) * 0.05).bfloat16()
    W = (torch.randn(V, H, generator=g) * 0.02).bfloat16()
    Hd = torch.randn(T, H, generator=g).bfloat16()
    Pc, Wc, Hc = P.cuda(), W.cuda(), Hd.cuda()
    torch.backends.cuda.matmul.allow_tf32 = False
    dH_ref = Pc.float() @ Wc.float()
    dW_ref = Pc.float().T @ Hc.float()
    dH32 = logprob.bwd_dhidden(Pc, Wc, fp32=True)
    np.testing.assert_allclose(dH32.cpu().numpy(), dH_ref.cpu().numpy(), rtol=1e-4, atol=1e-4 * dH_ref.abs().max().item())
    dH16 = logprob.bwd_dhidden(Pc, Wc)
    assert dH16.dtype == torch.bfloat16
    assert (dH16.float() - dH_ref).abs().max().item() <= 2 ** -8 * dH_ref.abs().max().item() * 1.01 + 1e-6
    dW = torch.full((V, H), float("nan"), device="cuda")
    logprob.bwd_dweight(Pc, Hc, dW, accumulate=False)
    np.testing.assert_allclose(dW.cpu().numpy(), dW_ref.cpu().numpy(), rtol=1e-4, atol=1e-4 * dW_ref.abs().max().item())
    logprob.bwd_dweight(Pc, Hc, dW, accumulate=True)                  # fp32 read-modify-write
    np.testing.assert_allclose(dW.cpu().numpy(), 2 * dW_ref.cpu().numpy(), rtol=1e-4,
                               atol=2e-4 * dW_ref.abs().max().item())


def test_dlogits_in_place():
    from open_o3_video_b200 import logprob
    T, V = 50, 1000
    g0 = torch.Generator().manual_seed(1)
    z = (torch.randn(T, V, generator=g0) * 2).bfloat16()
    targets = torch.randint(0, V, (T,), generator=g0)
    grad = torch.randn(T, generator=g0) * 0.01
    grad[3] = 0.0
    lse = torch.logsumexp(z.float(), -1)
    ref = grad[:, None] * (torch.nn.functional.one_hot(targets, V).float() - torch.exp(z.float() - lse[:, None]))
    buf = torch.zeros(T, V + 8, dtype=torch.bfloat16, device="cuda")
    buf[:, :V] = z.cuda()
    logprob.dlogits_(buf[:, :V], lse.cuda(), grad.cuda(), targets.cuda(), 0)
    out = buf[:, :V].float().cpu()
    assert (out - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-7
    assert (out[3] == 0).all() and (buf[:, V:] == 0).all()
    # vocab slice: the one-hot lands only in the owner slice
    buf2 = z[:, 256:512].cuda().contiguous()
    logprob.dlogits_(buf2, lse.cuda(), grad.cuda(), targets.cuda(), 256)
    assert (buf2.float().cpu() - ref[:, 256:512]).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-7


@pytest.mark.parametrize("T,H,V", [(384, 128, 1024), (1000, 512, 5000)])
def test_autograd_function_backward(cta_pair, T, H, V):
    """fused_logprob as a torch.autograd.Function: gradients w.r.t. hidden and weight."""
    from open_o3_video_b200 import logprob
    hidden, weight, targets = synth.head_inputs(T, H, V, seed=T)
    g = torch.randn(T, generator=torch.Generator().manual_seed(2)) * 0.01
    dH_ref, dW_ref, _ = _torch_ref_grads(hidden, weight, targets, g)
    for save_bytes in (1 << 40, 0):                                   # keep logits / recompute per chunk
        logprob.SAVE_LOGITS_BYTES = save_bytes
        h = hidden.cuda().bfloat16().requires_grad_(True)
        w = weight.cuda().bfloat16().requires_grad_(True)
        lp = logprob.fused_logprob(h, w, targets.cuda(), chunk_tokens=256)
        (lp * g.cuda()).sum().backward()
        for got, ref in ((h.grad, dH_ref), (w.grad, dW_ref)):
            err = (got.float().cpu() - ref).norm() / ref.norm()
            assert err < 1e-2, err
            assert (got.float().cpu() - ref).abs().max() < 3e-2 * ref.abs().max()
    logprob.SAVE_LOGITS_BYTES = 24 << 30


FUSED_CASES = [(128, 64, 256, 0), (300, 192, 1000, 0), (1100, 512, 5000, 0), (777, 1024, 9496, 0), (640, 256, 2376, 4752),
               (2900, 1792, 9496, 9496)]


@pytest.mark.parametrize("T,H,V,v_off", FUSED_CASES)
def test_fused_softmax_backward_gemms(T, H, V, v_off):
    """o3v_lmhead_bwd_{dhidden,dweight}_fused: the softmax backward applied to the A tiles in shared memory.
    Against (a) torch fp32 from the exact formula and (b) the separate in-place dlogits pass + plain GEMMs;
    ragged shapes, a vocab slice (v_off > 0: targets outside the slice get no one-hot), zero-gradient rows."""
    from open_o3_video_b200 import _lib, logprob
    g0 = torch.Generator().manual_seed(T + V)
    z = (torch.randn(T, V, generator=g0) * 1.5).bfloat16().cuda()
    W = (torch.randn(V, H, generator=g0) * 0.02).bfloat16().cuda()
    Hd = torch.randn(T, H, generator=g0).bfloat16().cuda()
    targets = torch.randint(0, V + 2 * v_off, (T,), generator=g0).cuda()          # global ids, some outside the slice
    grad = (torch.randn(T, generator=g0) * 0.01).cuda()
    grad[torch.rand(T, generator=g0) < 0.3] = 0.0
    grad[T // 2:T // 2 + 40] = 0.0                                               # a whole warp of masked rows
    lse = torch.logsumexp(z.float(), -1) + 0.3                                   # as if other slices held mass too
    torch.backends.cuda.matmul.allow_tf32 = False
    col = targets - v_off
    onehot = torch.zeros(T, V, device="cuda")
    inside = (col >= 0) & (col < V)
    onehot[inside.nonzero()[:, 0], col[inside]] = 1.0
    P = grad[:, None] * (onehot - torch.exp(z.float() - lse[:, None]))
    dH_ref, dW_ref = P @ W.float(), P.T @ Hd.float()
    sb = (lse, grad, targets, v_off)
    t = _lib.Trace()
    _lib.trace = t
    try:
        dH = logprob.bwd_dhidden(z, W, fp32=True, softmax_bwd=sb)
        dW = torch.full((V, H), float("nan"), device="cuda")
        logprob.bwd_dweight(z, Hd, dW, accumulate=False, softmax_bwd=sb)
    finally:
        _lib.trace = None
    assert [n for n, _ in t.calls] == ["o3v_lmhead_softmax_bwd_rows", "o3v_lmhead_bwd_dhidden_fused", "o3v_lmhead_softmax_bwd_rows", "o3v_lmhead_bwd_dweight_fused"]
    # (a) P is rounded to bf16 before the MMAs, as on the unfused path: 2^-9 relative per element
    assert (dH - dH_ref).norm() <= 4e-3 * dH_ref.norm() and (dW - dW_ref).norm() <= 4e-3 * dW_ref.norm()
    assert (dH[grad == 0] == 0).all()
    # (b) the separate pass (exp2f instead of ex2.approx: at most one bf16 ulp apart on a few elements)
    z2 = z.clone()
    logprob.dlogits_(z2, lse, grad, targets, v_off)
    dH2 = logprob.bwd_dhidden(z2, W, fp32=True)
    dW2 = torch.empty(V, H, device="cuda")
    logprob.bwd_dweight(z2, Hd, dW2, accumulate=False)
    assert (dH - dH2).norm() <= 1e-3 * dH2.norm() and (dW - dW2).norm() <= 1e-3 * dW2.norm()
    # accumulate, and the logits buffer is left untouched
    logprob.bwd_dweight(z, Hd, dW, accumulate=True, softmax_bwd=sb)
    assert (dW - 2 * dW_ref).norm() <= 4e-3 * (2 * dW_ref).norm()
    assert torch.equal(z, (torch.randn(T, V, generator=torch.Generator().manual_seed(T + V)) * 1.5).bfloat16().cuda())


def test_fused_softmax_backward_needs_the_default_tile_mode():
    from open_o3_video_b200 import _lib, logprob
    z = torch.zeros(128, 256, dtype=torch.bfloat16, device="cuda")
    W = torch.zeros(256, 64, dtype=torch.bfloat16, device="cuda")
    sb = (torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda"), torch.zeros(128, dtype=torch.int64, device="cuda"), 0)
    _lib.set_tunable("cta_pair_bwd", 1)
    try:
        with pytest.raises(_lib.O3VError) as e:
            logprob.bwd_dhidden(z, W, softmax_bwd=sb)
        assert e.value.code == -7
    finally:
        _lib.set_tunable("cta_pair_bwd", 2)


EXP_CASES = [(128, 64, 256, 0, 0.0), (300, 192, 1000, 0, 3.0), (1100, 512, 5000, 0, -20.0), (640, 256, 2376, 4752, 40.0),
             (2900, 1792, 9496, 9496, 0.0),
             (8256, 512, 19096, 0, 0.0)]      # 75 vocabulary blocks x 1 column tile = one wave of 74 + 1: K-split tail (ragged)


@pytest.mark.parametrize("T,H,V,v_off,shift", EXP_CASES)
def test_exp_store_backward(T, H, V, v_off, shift):
    """The exp-store path end to end on one vocabulary slice: K1 storing E = exp(z - ref) (statistics bit-identical to the
    plain K1), the softmax backward in the K2a epilogue, pre-scaled hidden + K2b + deterministic one-hot scatter, against
    torch fp32 from the exact formulas.  `shift` moves every logit (rows far from 0: the reference must follow),
    duplicate targets exercise the scatter runs, rows with g = 0 / row_keep = 0 must come out exactly zero."""
    from open_o3_video_b200 import logprob
    g0 = torch.Generator().manual_seed(T + V)
    Hd = torch.randn(T, H, generator=g0).bfloat16().cuda()
    W = (torch.randn(V, H, generator=g0) * 0.05).bfloat16().cuda()
    if shift:
        W[:, 0] = 0.0
        Hd[:, 0] = 1.0
        W[:, 0] = (torch.full((V,), shift)).bfloat16().cuda()           # z += shift for every column
    targets = torch.randint(0, V + 2 * v_off, (T,), generator=g0).cuda()
    targets[5:25] = v_off + 7                                          # a run of duplicates (one dW row, 20 tokens)
    targets[T - 3:] = v_off + V - 1
    grad = (torch.randn(T, generator=g0) * 0.01).cuda()
    keep = (torch.rand(T, generator=g0) > 0.3).to(torch.int32).cuda()
    keep[T // 2:T // 2 + 40] = 0
    grad = grad * keep
    torch.backends.cuda.matmul.allow_tf32 = False
    z = Hd.float() @ W.float().T
    lse = torch.logsumexp(z, -1) + 0.3                                 # as if other slices held mass too
    col = targets - v_off
    inside = (col >= 0) & (col < V)
    onehot = torch.zeros(T, V, device="cuda")
    onehot[inside.nonzero()[:, 0], col[inside]] = 1.0
    P = grad[:, None] * (onehot - torch.exp(z - lse[:, None]))
    dH_ref, dW_ref = P @ W.float(), P.T @ Hd.float()
    # K1: same statistics with and without the store, E = exp(z - ref) * keep
    ref = logprob.row_reference(Hd, logprob.sample_weight(W), targets)
    assert (ref >= z.max(1).values - 40).all() and (ref <= z.max(1).values + logprob.ROW_REF_MARGIN + 1e-3).all()
    E = torch.full((T, V + 8), 7.0, dtype=torch.bfloat16, device="cuda")
    st = logprob.lmhead_stats(Hd, W, targets, v_off, E[:, :V], row_ref=ref, row_keep=keep)
    assert torch.equal(st, logprob.lmhead_stats(Hd, W, targets, v_off))
    E_ref = torch.exp(z - ref[:, None]) * keep[:, None]
    assert ((E[:, :V].float() - E_ref).abs() <= 2 ** -7 * E_ref + 1e-30).all() and (E[:, V:] == 7.0).all()
    # backward
    rows, order = logprob.softmax_rows(lse, grad, targets, ref, v_off, V)
    dH = logprob.bwd_dhidden_exp(E[:, :V], rows, W, fp32=True)
    dW = torch.full((V, H), float("nan"), device="cuda")
    logprob.bwd_dweight_exp(E[:, :V], rows, order, Hd, dW, accumulate=False)
    assert (dH - dH_ref).norm() <= 4e-3 * dH_ref.norm() and (dW - dW_ref).norm() <= 4e-3 * dW_ref.norm()
    assert (dH[grad == 0] == 0).all()
    dH16 = logprob.bwd_dhidden_exp(E[:, :V], rows, W)
    assert dH16.dtype == torch.bfloat16 and (dH16.float() - dH_ref).norm() <= 6e-3 * dH_ref.norm()
    # deterministic (scatter runs are summed sequentially) and accumulating
    dW2 = torch.full((V, H), float("nan"), device="cuda")
    logprob.bwd_dweight_exp(E[:, :V], rows, order, Hd, dW2, accumulate=False)
    assert torch.equal(dW, dW2)
    logprob.bwd_dweight_exp(E[:, :V], rows, order, Hd, dW, accumulate=True)
    assert (dW - 2 * dW_ref).norm() <= 4e-3 * (2 * dW_ref).norm()


def test_dweight_tail_split_matches_the_single_launch(monkeypatch):
    """K2b with the K range of the last vocabulary rows split over the idle CTA pairs (logprob._dw_tail_plan) against the
    single launch: same sums up to the fp32 order of the K partials, deterministic, accumulating; ragged last block."""
    from open_o3_video_b200 import logprob
    T, H, V = 12352, 1024, 2 * 37 * 256 + 200            # 75 blocks x 2 column tiles = 2 waves of 74 + 2; 3 K splits
    assert logprob._dw_tail_plan(V, H, T, "cuda") == (74 * 256, 3)
    g0 = torch.Generator().manual_seed(5)
    Hd = torch.randn(T, H, generator=g0).bfloat16().cuda()
    E = (torch.rand(T, V, generator=g0) * 1e-3).bfloat16().cuda()
    lse = torch.zeros(T, device="cuda")
    ref = torch.zeros(T, device="cuda")
    grad = (torch.randn(T, generator=g0) * 0.01).cuda()
    grad[100:400] = 0
    targets = torch.randint(0, V, (T,), generator=g0).cuda()
    targets[:50] = V - 3                                  # a scatter run inside the tail rows
    rows, order = logprob.softmax_rows(lse, grad, targets, ref, 0, V)
    out = {}
    for split in (True, False):
        monkeypatch.setattr(logprob, "DW_TAIL_SPLIT", split)
        dW = torch.full((V, H), float("nan"), device="cuda")
        logprob.bwd_dweight_exp(E, rows, order, Hd, dW, accumulate=False)
        dW2 = torch.full((V, H), float("nan"), device="cuda")
        logprob.bwd_dweight_exp(E, rows, order, Hd, dW2, accumulate=False)
        assert torch.equal(dW, dW2)
        logprob.bwd_dweight_exp(E, rows, order, Hd, dW2, accumulate=True)
        out[split] = (dW, dW2)
    torch.cuda.synchronize()
    a, b = out[True][0], out[False][0]
    assert torch.equal(a[:74 * 256], b[:74 * 256])        # rows of the main launch: same kernel, same order
    assert (a - b).abs().max() <= 1e-5 * b.abs().max()
    assert (out[True][1] - 2 * b).abs().max() <= 2e-5 * b.abs().max()
    # against torch fp32 from the definition: dW = E^T (a * hidden) + scatter(g * hidden)
    torch.backends.cuda.matmul.allow_tf32 = False
    a_t = -grad * torch.exp(ref - lse)
    dW_ref = E.float().T @ (a_t[:, None] * Hd.float()).bfloat16().float()
    dW_ref.index_add_(0, targets, grad[:, None] * Hd.float())
    assert (a - dW_ref).norm() <= 2e-3 * dW_ref.norm()
