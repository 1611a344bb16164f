"""torchrun worker for tests/test_gpu_multi.py: vocab-sharded fused step on WORLD_SIZE GPUs,
checked on rank 0 against the single-GPU result of the same library."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_o3_video_b200 import logprob, sharded  # noqa: E402
from oracle import gspo as ogspo, synth  # noqa: E402


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    N, Tc, G, H, V = 8, 96, 4, 256, 151936 // 16          # ragged vocab (9496 = 37.09 tiles)
    hidden, weight, _ = synth.head_inputs(N * Tc, H, V, seed=21)
    d = synth.gspo_inputs(N, Tc, G, vocab=V + 1000, eos_id=V - 1, seed=22, off_policy=True)
    ids = (d["ids"] % V).to(dev)
    _, mask = ogspo.eos_mask(ids.cpu(), V - 1)
    h = hidden.to(dev).bfloat16().view(N, Tc, H)
    w = weight.to(dev).bfloat16()
    ref = (d["ref"]).to(dev) - 6.0
    old = (d["old"]).to(dev) - 6.0
    rpf = d["rewards_per_func"].to(dev)
    args = (ref, mask.to(dev), rpf, G, 0.04, 0.2, 0.2, True, old)
    w_local, v0 = sharded.shard_weight(w, rank, world)
    out = logprob.fused_logprob_gspo(h, w_local, ids, *args, v_offset=v0, group=dist.group.WORLD, chunk_tokens=2 * Tc)
    # autograd path, sharded
    h2 = h.view(-1, H).clone().requires_grad_(True)
    lp2 = logprob.fused_logprob(h2, w_local, ids.view(-1), v_offset=v0, group=dist.group.WORLD)
    lp2.sum().backward()
    torch.cuda.synchronize()
    # fused peer-memory exchange (symmetric memory + merge kernel reading every rank over NVLink)
    peer_err = 0.0
    ex = sharded.PeerExchange(dist.group.WORLD, capacity_tokens=2 * Tc)
    outp = logprob.fused_logprob_gspo(h, w_local, ids, *args, v_offset=v0, group=ex, chunk_tokens=2 * Tc)
    # + peer-mapped dHidden with the one-shot P2P all-reduce overlapped with the dW GEMM
    ex2 = sharded.PeerExchange(dist.group.WORLD, capacity_tokens=N * Tc, hidden_size=H)
    outq = logprob.fused_logprob_gspo(h, w_local, ids, *args, v_offset=v0, group=ex2, chunk_tokens=2 * Tc)
    torch.cuda.synchronize()
    dh_q = outq["d_hidden"].float().clone()
    lp3 = logprob.fused_logprob(h.view(-1, H)[: Tc + 7], w_local, ids.view(-1)[: Tc + 7], v_offset=v0, group=ex)   # T < capacity
    lp4 = logprob.fused_logprob(h.view(-1, H)[: Tc + 7], w_local, ids.view(-1)[: Tc + 7], v_offset=v0, group=dist.group.WORLD)
    torch.cuda.synchronize()
    peer_err = max((outp["per_token_logps"] - out["per_token_logps"]).abs().max().item(),
                   (outp["d_hidden"].float() - out["d_hidden"].float()).abs().max().item(),
                   abs(outp["loss"].item() - out["loss"].item()), (lp3 - lp4).abs().max().item(),
                   (outq["d_weight"] - out["d_weight"]).abs().max().item())
    # fp32 sum of the bf16 partials in rank order vs NCCL's reduction order: one bf16 rounding apart at most
    ar_err = ((dh_q - out["d_hidden"].float()).norm() / out["d_hidden"].float().norm()).item()
    if ar_err > 4e-3:
        peer_err = max(peer_err, ar_err)
    ok = True
    if rank == 0:
        one = logprob.fused_logprob_gspo(h, w, ids, *args, chunk_tokens=2 * Tc)
        e_lp = (out["per_token_logps"] - one["per_token_logps"]).abs().max().item()
        e_loss = abs(out["loss"].item() - one["loss"].item())
        e_dh = ((out["d_hidden"].float() - one["d_hidden"].float()).norm() / one["d_hidden"].float().norm()).item()
        v_a, v_b = sharded.vocab_slices(V, world)[0]
        e_dw = ((out["d_weight"] - one["d_weight"][v_a:v_b]).norm() / one["d_weight"][v_a:v_b].norm()).item()
        h3 = h.view(-1, H).clone().requires_grad_(True)
        logprob.fused_logprob(h3, w, ids.view(-1)).sum().backward()
        e_ag = ((h2.grad.float() - h3.grad.float()).norm() / h3.grad.float().norm()).item()
        print("multi-gpu world=%d: dlogp %.2e dloss %.2e dH %.2e dW %.2e autograd-dH %.2e peer-vs-nccl %.2e" %
              (world, e_lp, e_loss, e_dh, e_dw, e_ag, peer_err), flush=True)
        # log-probs: same fp32 arithmetic, different summation tree; dH: bf16 partial sums per slice
        ok = e_lp < 5e-5 and e_loss < 1e-6 and e_dh < 1e-2 and e_dw < 1e-3 and e_ag < 1e-2 and peer_err == 0.0
    if peer_err != 0.0:
        ok = False                      # same arithmetic in the same rank order: must be bit-identical
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
