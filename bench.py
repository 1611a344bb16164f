#!/usr/bin/env python
"""bench.py -- completion tokens/s of the fused logprob + GSPO fwd+bwd step (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c5]

A "step" is one pass of the hot path over one batch of synthetic rollouts: lm_head ->
log-softmax -> gather -> KL / advantages / GSPO loss, and the backward to dHidden and
dW_lm_head.  Default workload = BASELINE config 2 (Qwen2.5-VL-7B head, 8 prompts x G=8 x 2048
completion tokens, bf16, 1 B200).  N > 1 (launched by torchrun, one rank per GPU): the
vocabulary is sharded over the ranks (whole 256-column tiles), every rank sees all tokens,
only (max, sum-exp, target-logit) triples are all-gathered in the forward and dHidden is
all-reduced once per step -> strong scaling, T fixed.

`value`   : device-resident inputs, CUDA events around the K timed steps, max over ranks.
`e2e`     : same metric through the public API with HOST (pinned) inputs: H2D of hidden
            states / ids / ref log-probs / mask / rewards (each rank copies its 1/N of the
            sequences and the ranks all-gather over NVLink) and D2H of the loss inside the
            timed region (lm_head.weight is model state and stays on the device).
            `h2d_bytes_per_step` is per rank.
`roofline`: dominant kernel (the K1 tcgen05 GEMM with fused softmax-stats epilogue), timed
            live with CUDA events around its launches inside the timed region.
`cpu_baseline`: the oracle port of the reference's torch path (fp32) on the host cores, on
            a bounded sample of the same workload (one group of G rollouts at the same head,
            2048 tokens per step).
`--impl reference`: that CPU path alone, same JSON shape.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "c2": dict(head="qwen2.5-vl-7b", H=3584, V=152064, prompts=8, G=8, Tc=2048),
    "c3": dict(head="qwen3-vl-8b", H=4096, V=151936, prompts=16, G=8, Tc=4096),
    "c5": dict(head="qwen3-vl-8b", H=4096, V=151936, prompts=1, G=16, Tc=16384),
    "c1": dict(head="qwen2.5-vl-7b", H=3584, V=152064, prompts=1, G=4, Tc=512),
    # diagnostics: ONE rank's share of c2 / c3 at 8 GPUs (its 1/8 vocabulary slice, every token) on a single GPU, i.e.
    # the per-rank kernel shapes of the 8-GPU run without any communication
    "c2s8": dict(head="qwen2.5-vl-7b, 1/8 vocab slice", H=3584, V=19008, prompts=8, G=8, Tc=2048),
    "c3s8": dict(head="qwen3-vl-8b, 1/8 vocab slice", H=4096, V=18992, prompts=16, G=8, Tc=4096),
    "c3s4a": dict(head="qwen3-vl-8b, 149 of 594 vocab tiles (ranks 0-1 of 4)", H=4096, V=38144, prompts=16, G=8, Tc=4096),
    "c3s4b": dict(head="qwen3-vl-8b, 148 of 594 vocab tiles (ranks 2-3 of 4)", H=4096, V=37888, prompts=16, G=8, Tc=4096),
    "tiny": dict(head="tiny", H=256, V=8192, prompts=2, G=4, Tc=128),
    # what ONE rank of the reference's own launch sees per step (per-device batch of 1 prompt x 8 generations)
    "dev": dict(head="qwen2.5-vl-7b", H=3584, V=152064, prompts=1, G=8, Tc=2048),
}
METRIC = "completion tokens/s, fused logprob+GSPO fwd+bwd; % bf16 tensor-pipe peak"
BETA, EPS = 0.04, 0.2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(burst=p["bf16_tflops"], sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7 and r[0].isdigit()]
        if not rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        sm = sorted(int(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[3 + j].lower().startswith("active") for r in rows)]
        pw = sorted(float(r[2]) for r in rows)
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=int(rows[0][1]), reasons=reasons, samples=len(rows),
                    power_w_max=pw[-1], power_w_median=pw[len(pw) // 2])


# ----------------------------------------------------------------------------- CPU reference arm
def workload_name(name, cfg):
    """config.workload: the SAME string in both arms (the reference arm times a bounded sample of this workload)."""
    return "%s: %s head (H=%d, V=%d), %d prompts x G=%d x %d completion tokens = %d tokens/step, fused logprob+GSPO fwd+bwd" % (
        name, cfg["head"], cfg["H"], cfg["V"], cfg["prompts"], cfg["G"], cfg["Tc"], cfg["prompts"] * cfg["G"] * cfg["Tc"])


CPU_SAMPLE_TOKENS = 2048


def cpu_reference(steps, warmup, head_cfg):
    """The reference's torch path on the host: oracle/logps.py (grpo_trainer.py:371-384 with the
    lm_head) + oracle/gspo.py (the inline loss block) + loss.backward() to hidden and W, fp32,
    all host threads.  Bounded sample of the benchmarked workload: ONE group of its G rollouts at its head shape,
    completions cut to 2048 / G tokens (2048 tokens per step = config 1's size; the per-token cost of the path does
    not depend on the number of groups or on the completion length)."""
    from oracle import gspo as ogspo, logps as ologps, synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    H, V, G = head_cfg["H"], head_cfg["V"], head_cfg["G"]
    N, Tc = G, max(1, CPU_SAMPLE_TOKENS // G)
    hidden, weight, _ = synth.head_inputs(N * Tc, H, V)
    d = synth.gspo_inputs(N, Tc, G, vocab=V, eos_id=V - 1)
    ids = d["ids"] % V
    _, mask = ogspo.eos_mask(ids, V - 1)

    def step():
        h = hidden.view(N, Tc, H).clone().requires_grad_(True)
        w = weight.clone().requires_grad_(True)
        full_ids = torch.cat([ids[:, :1], ids], 1)          # [N, Tc+1]: logits[:, :-1] pair with ids[:, 1:]
        hh = torch.cat([h, h[:, -1:]], 1)
        lp = ologps.per_token_logps(hh, w, full_ids)
        out = ogspo.gspo_step(lp, lp.detach() + 0.1, mask, d["rewards_per_func"], G, BETA, EPS, EPS)
        out["loss"].backward()
        return float(out["loss"])

    for _ in range(warmup):
        step()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    best = min(times)
    return dict(value=N * Tc / best, unit="tokens/s", cores=cores, kind="port",
                sample="%d steps of 1 prompt x G=%d x %d tokens (= %d tokens) of this workload, %s head, fp32, fwd+bwd to hidden "
                       "and W, best step %.2f s" % (steps, G, Tc, N * Tc, head_cfg["head"], best)), sum(times) / len(times)


def run_reference(args):
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 3))
    base, mean_s = cpu_reference(steps, min(args.warmup, 1), cfg)
    line = dict(metric=METRIC, value=base["value"], unit="tokens/s", n_gpus=args.gpus, steps=steps,
                warmup=min(args.warmup, 1), ms_per_step=mean_s * 1e3, higher_is_better=True, scaling="strong",
                vs_baseline=None, dtype="fp32", data="synthetic", impl="reference",
                config=dict(workload=workload_name(args.config, cfg), parallelism="host CPU, %d threads" % base["cores"],
                            sample=base["sample"]),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- self-check (outside the timed regions)
def parity_check(out, hidden, ids, ref, mask, rpf, G, H, V, Tc, weight, v_off, rank, world, dev):
    """The step that was just timed against plain torch fp32 formulas (no library code, no oracle import) on a
    subset: log-probs of the first and the last sequence against a full-vocabulary fp32 recompute (the weight slices
    are all-gathered for it), the loss and d loss / d logp against an autograd restatement of grpo_trainer.py:635-706
    fed with the library's log-probs, dHidden of the first sequence (rank 0 owns those rows in either collective
    mode) and 256 columns of rank 0's dW slice.  Every entry is error / tolerance (tolerances of DESIGN.md section 6:
    1e-3 relative for log-probs and loss, 1e-2 in norm for the gradients)."""
    import torch.distributed as dist
    N = hidden.shape[0]
    torch.backends.cuda.matmul.allow_tf32 = False
    if world > 1:
        sizes = [0] * world
        dist.all_gather_object(sizes, int(weight.shape[0]))
        parts = [torch.empty(n, H, dtype=weight.dtype, device=dev) for n in sizes]
        dist.all_gather(parts, weight.contiguous())
        w_full = torch.cat(parts, 0)
        del parts
    else:
        w_full = weight
    lp = out["per_token_logps"]
    res = {}
    if rank == 0:
        wf = w_full.float()
        z_keep = None
        errs = []
        for n in (0, N - 1):
            z = hidden[n].float() @ wf.T                                                  # [Tc, V] fp32
            lse = torch.logsumexp(z, -1)
            lp_ref = z.gather(1, ids[n][:, None])[:, 0] - lse
            errs.append(((lp[n] - lp_ref).abs() / lp_ref.abs().clamp(min=1e-6)).max().item())
            if n == 0:
                z_keep, lse0 = z, lse
        res["logp_max_rel"] = max(errs)
        # loss + gradient w.r.t. the log-probs: autograd through the reference's formulas (on-policy: old = detach)
        x = lp.detach().clone().requires_grad_(True)
        m = mask.float()
        r = rpf.sum(1)
        mean_g = r.view(-1, G).mean(1).repeat_interleave(G)
        std_g = r.view(-1, G).std(1).repeat_interleave(G)
        adv = (r - mean_g) / (std_g + 1e-4)
        xc = torch.clamp(ref - x, -10, 10)
        kl = torch.exp(xc) - xc - 1
        lr = x - x.detach()
        c1 = torch.exp((lr * m).sum(-1) / m.sum(-1).clamp(min=1.0)).unsqueeze(-1)
        c2 = torch.clamp(c1, 1 - EPS, 1 + EPS)
        ptl = -torch.min(c1 * adv.unsqueeze(1), c2 * adv.unsqueeze(1)) + BETA * kl
        loss = ((ptl * m).sum(-1) / m.sum(-1).clamp(min=1.0)).mean()
        loss.backward()
        res["loss_rel"] = abs(out["loss"].item() - loss.item()) / max(abs(loss.item()), 1e-12)
        g0 = x.grad[0]                                                                   # [Tc]
        P0 = torch.exp(z_keep - lse0[:, None]).mul_(-g0[:, None])
        P0[torch.arange(Tc, device=dev), ids[0]] += g0
        dh_ref = P0 @ wf
        dh = out["d_hidden"].reshape(-1, H)[:Tc].float()                                  # rows 0..Tc: owned by rank 0
        res["dh_rel_first_seq"] = ((dh - dh_ref).norm() / dh_ref.norm().clamp(min=1e-30)).item()
        del P0, z_keep
        # dW: 256 columns of rank 0's slice need P[:, cols] of ALL tokens: z[:, cols] and lse from the library's own
        # log-probs are not available, so recompute lse by chunks over the full vocabulary
        cols = torch.arange(v_off + 1024, v_off + 1280, device=dev)
        gall = x.grad.view(-1)
        hs = hidden.view(-1, H)
        dW_ref = torch.zeros(256, H, device=dev)
        T = hs.shape[0]
        for s in range(0, T, 4096):
            e = min(T, s + 4096)
            if not bool((gall[s:e] != 0).any()):
                continue
            zc = hs[s:e].float() @ wf.T
            lsec = torch.logsumexp(zc, -1)
            Pc = torch.exp(zc[:, cols] - lsec[:, None]).mul_(-gall[s:e, None])
            hit = (ids.view(-1)[s:e, None] == cols[None, :])
            Pc += hit.float() * gall[s:e, None]
            dW_ref += Pc.T @ hs[s:e].float()
            del zc
        dw = out["d_weight"][1024:1280]
        res["dw_rel_256_cols"] = ((dw - dW_ref).norm() / dW_ref.norm().clamp(min=1e-30)).item()
        res["max_err_over_tol"] = max(res["logp_max_rel"] / 1e-3, res["loss_rel"] / 1e-3, res["dh_rel_first_seq"] / 1e-2,
                                      res["dw_rel_256_cols"] / 1e-2)
        res["reference"] = "torch fp32 on the GPU from the formulas of grpo_trainer.py:371-384, 635-706 (no library code)"
    if world > 1:
        dist.barrier()
    return res


# ----------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from open_o3_video_b200 import _lib, gspo, logprob, sharded
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = pg = dist.group.WORLD
    _lib.check(_lib.load().o3v_check_device(), "o3v_check_device")
    if args.cta_pair:
        _lib.set_tunable("cta_pair", args.cta_pair)
    if args.fwd_groups:
        _lib.set_tunable("fwd_groups", args.fwd_groups)
    for kv in args.tunable:
        k, v = kv.split("=")
        _lib.set_tunable(k, int(v))

    if args.no_chunk_plan:
        logprob.PLAN_CHUNKS = False
    H, V, G, Tc = cfg["H"], cfg["V"], cfg["G"], cfg["Tc"]
    N = cfg["prompts"] * G
    T = N * Tc
    # synthetic rollouts (seeded, identical on every rank): hidden ~ N(0,1), W ~ N(0, 0.02^2)
    g = torch.Generator(device=dev).manual_seed(20261018)
    hidden = torch.randn(N, Tc, H, device=dev, generator=g, dtype=torch.float32).bfloat16()
    w_full = (torch.randn(V, H, device=dev, generator=g, dtype=torch.float32) * 0.02).bfloat16()
    ids = torch.randint(0, V - 1, (N, Tc), device=dev, generator=g)
    lens = torch.randint(Tc // 4, Tc + 1, (N,), device=dev, generator=g)
    eos_id = V - 1
    ids[torch.arange(N, device=dev), lens - 1] = eos_id      # planted EOS drives the mask
    _, mask = gspo.eos_mask(ids, eos_id)
    weight, v_off = sharded.shard_weight(w_full, rank, world)
    if world > 1 and args.exchange == "peer":
        group = sharded.PeerExchange(dist.group.WORLD, capacity_tokens=T, hidden_size=H if args.overlap_allreduce else 0,
                                     dh_mode=args.dh_collective)
    # chunking: 0 = the library's own plan (a 10 GB logits buffer cut into whole sequences, logprob.plan_chunks)
    call_chunk_tokens = args.chunk_tokens
    v_plan = logprob._plan_vocab(weight.shape[0], pg, dev)
    seq_plan = logprob.plan_chunks(N, Tc, H, weight.shape[0], args.chunk_tokens or logprob.auto_chunk_tokens(v_plan), dev,
                                   tune=world == 1)
    args.chunk_tokens = max(seq_plan) * Tc
    del w_full
    # reference-model log-probs = policy log-probs + N(0, 0.1^2) (forward-only pass, untimed)
    ref = logprob.fused_logprob(hidden.view(T, H), weight, ids.view(T), v_offset=v_off, group=group).view(N, Tc)
    ref = ref + torch.randn(N, Tc, device=dev, generator=g) * 0.1
    rpf = torch.rand(N, 3, device=dev, generator=g)
    d_weight = torch.zeros(weight.shape[0], H, dtype=torch.float32, device=dev)
    torch.cuda.synchronize()

    def step(h, i, r, m, rw):
        return logprob.fused_logprob_gspo(h, weight, i, r, m, rw, G, BETA, EPS, EPS, True, None, v_offset=v_off,
                                          group=group, chunk_tokens=call_chunk_tokens, d_weight_out=None,
                                          overlap_dlogits=bool(args.overlap_dlogits),
                                          **({"backward": "exp"} if args.backward == "exp" else
                                             {"fuse_dlogits": args.backward == "smem"}))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    for _ in range(args.warmup):
        out = step(hidden, ids, ref, mask, rpf)
    barrier()

    # ---- timed region 1: device-resident inputs
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    trace = _lib.Trace(events=True)
    _lib.trace = trace
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step(hidden, ids, ref, mask, rpf)
    e1.record()
    barrier()
    _lib.trace = None
    clocks = sampler.stop() if rank == 0 else None
    ms_dev = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    loss = float(out["loss"])
    durs = trace.durations_ms()
    launches = trace.launches

    # ---- timed region 2: end to end from pinned host buffers
    # Every step's inputs start in pinned host memory and its loss ends there.  The copy of
    # step i+1's inputs runs on a side stream while step i computes (double-buffered device
    # staging), as a training loop's prefetcher would; all copies are inside the timed region.
    # Vocab-sharded ranks all need every token row: each rank copies only ITS 1/world slice of the
    # sequences over PCIe and the slices are all-gathered over NVLink (the real data-parallel layout).
    full = (hidden, ids, ref, mask, rpf)
    seq_lo, seq_hi = rank * N // world, (rank + 1) * N // world
    host = [t[seq_lo:seq_hi].cpu().pin_memory() for t in full]
    h2d = sum(t.numel() * t.element_size() for t in host)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [[torch.empty_like(t, device=dev) for t in full] for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    # Gather of the per-rank slices.  "p2p": `sharded.PeerGather`, every rank PUSHES its slice into all peers'
    # peer-mapped buffers with copy-engine transfers over NVLink, so no SM is taken from the persistent GEMM kernels
    # of the step that is running; "nccl": all-gather kernels (fallback).
    even = (N % world == 0)
    gather, pgather = "nccl" if world > 1 else "none", None
    if world > 1 and args.e2e_gather == "p2p":
        try:
            pgather = sharded.PeerGather(pg, [(tuple(t.shape), t.dtype) for t in full], device=dev, slots=2)
            staging, gather = pgather.bufs, "p2p"
        except Exception as exc:                              # keep the bench alive on a box without P2P
            if rank == 0:
                print("e2e gather: symmetric memory unavailable (%s), using NCCL all-gather" % exc, file=sys.stderr)
            pgather = None

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the step that last used this slot is done
            if gather == "p2p":
                dev_local = [d[seq_lo:seq_hi] for d in staging[slot]]
                for d, h in zip(dev_local, host):            # H2D of this rank's sequences (PCIe), in place in my buffer
                    d.copy_(h, non_blocking=True)
                pgather.gather(slot, dev_local, seq_lo)     # ... then pushed to every peer over NVLink
            else:
                for d, h in zip(staging[slot], host):
                    d[seq_lo:seq_hi].copy_(h, non_blocking=True)
                if world > 1:
                    for d in staging[slot]:
                        if even:
                            dist.all_gather_into_tensor(d, d[seq_lo:seq_hi], group=pg)
                        else:
                            parts = [d[r * N // world:(r + 1) * N // world] for r in range(world)]
                            dist.all_gather(parts, d[seq_lo:seq_hi], group=pg)
            ready[slot].record(copy_stream)

    def e2e_run(n):
        cur = torch.cuda.current_stream()
        for s_ in range(2):
            consumed[s_].record(cur)
        prefetch(0)
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                prefetch(slot ^ 1)
            cur.wait_event(ready[slot])
            o = step(*staging[slot])
            consumed[slot].record(cur)
            loss_host.copy_(o["loss"], non_blocking=True)

    e2e_run(min(args.warmup, 2))
    barrier()
    e0.record()
    e2e_run(args.steps)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps

    parity = dict(max_err_over_tol=None, skipped="--no-parity") if args.no_parity else \
        parity_check(out, hidden, ids, ref, mask, rpf, G, H, V, Tc, weight, v_off, rank, world, dev)
    shares_ranks = None
    if world > 1:                                           # every rank's per-kernel milliseconds (load balance / skew)
        mine = {k: sum(v) / args.steps for k, v in durs.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        shares_ranks = {k: [round(g.get(k, 0.0), 3) for g in gathered] for k in mine}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    flop_tok = 6.0 * H * V
    value = T / (ms_dev * 1e-3)
    # dominant kernel: K1 (tcgen05 GEMM + fused softmax statistics + bf16 logits store); the call
    # also contains the tiny partial-merge kernel, timed with it
    k1 = durs.get("o3v_lmhead_fwd+store", [])
    n_chunks = max(1, len(k1) // args.steps)
    tok_per_launch = T / n_chunks
    k1_ms = sum(k1) / len(k1) if k1 else float("nan")
    v_local = weight.shape[0]
    ach = 2.0 * tok_per_launch * H * v_local / (k1_ms * 1e-3) / 1e12
    shares = {k: sum(v) / args.steps for k, v in durs.items()}
    traffic, traffic_detail = None, None                   # DRAM bytes per K1 launch from the committed ncu capture
    tpath = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if world == 1 and os.path.isfile(tpath):
        with open(tpath) as f:
            tj = json.load(f).get("%s:%d" % (args.config, args.chunk_tokens))
        if tj:
            traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
            traffic_detail = dict(algorithmic_bytes=tj["algorithmic_bytes"], unit="bytes per launch",
                                  source="profiles/k1_traffic.json (ncu --set full, same command)")
    line = dict(
        metric=METRIC, value=value, unit="tokens/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
        ms_per_step=ms_dev, higher_is_better=True, scaling="strong", vs_baseline=None,
        dtype="bf16", data="synthetic",
        config=dict(workload=workload_name(args.config, cfg),
                    parallelism="vocab-sharded x%d (%s exchange of softmax triples, %s of dHidden)"
                                % (world, "fused NVLink peer-memory" if args.exchange == "peer" else "NCCL all-gather",
                                   {"reduce_scatter": "reduce-scatter by peer-memory pull (owners load their token rows over NVLink) beside the dW GEMM",
                                    "reduce_scatter_fused": "reduce-scatter fused into the K2a epilogue (NVLink stores to the token owners) + local slot sum",
                                    "all_reduce": "one-shot P2P all-reduce beside the dW GEMM"}[args.dh_collective]
                                   if (args.exchange == "peer" and args.overlap_allreduce) else "NCCL all-reduce")
                    if world > 1 else "single GPU", chunk_tokens=args.chunk_tokens, chunk_sequences=seq_plan,
                    cache="inputs (%.1f GB) and per-chunk logits are far larger than the 126 MB L2; no flush needed"
                          % ((hidden.numel() * 2 + weight.numel() * 2) / 1e9), loss=loss),
        frac_of_bf16_peak=dict(burst=flop_tok * value / world / 1e12 / pk["burst"],
                               sustained=flop_tok * value / world / 1e12 / pk["sustained"], source=pk["source"]),
        roofline=dict(bound="tensor", kernel="lmhead_gemm_kernel<EPI_STATS> (K1 + logits store), per token chunk",
                      achieved=ach, peak=pk["sustained"], unit="TFLOP/s", frac=ach / pk["sustained"],
                      frac_of_burst=ach / pk["burst"], peak_source=pk["source"] + " (sustained: kernel timed inside a long step)",
                      ms_per_launch=k1_ms, traffic=traffic, traffic_detail=traffic_detail),
        kernel_ms_per_step=shares, **({"kernel_ms_per_step_ranks": shares_ranks} if shares_ranks else {}),
        e2e=dict(value=T / (ms_e2e * 1e-3), unit="tokens/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=4,
                 ms_per_step=ms_e2e, gather=gather),
        gpu_launches=launches, clocks=clocks, parity=parity, parity_max_err=parity["max_err_over_tol"])
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"], _ = cpu_reference(1, 0, cfg)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--e2e-gather", choices=("p2p", "nccl"), default="p2p",
                    help="N > 1, end-to-end arm: how the per-rank input slices reach every rank")
    ap.add_argument("--overlap-dlogits", type=int, default=0,
                    help="1: run the dlogits pass of chunk c on a side stream beside K1 of chunk c+1")
    ap.add_argument("--backward", default="exp", choices=["exp", "dlogits", "smem"],
                    help="exp: K1 stores exp(z - ref), softmax backward folded into the GEMM epilogue / operands "
                         "(default); dlogits: separate in-place elementwise pass (round 1); smem: shared-memory "
                         "transform of the A tiles inside the GEMMs")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--chunk-tokens", type=int, default=0, help="0 = size the chunk from a 10 GB logits buffer")
    ap.add_argument("--cta-pair", type=int, default=0)
    ap.add_argument("--fwd-groups", type=int, default=0)
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: forward exchange of the softmax triples (peer = fused NVLink merge kernel)")
    ap.add_argument("--overlap-allreduce", type=int, default=1,
                    help="N > 1 with --exchange peer: one-shot P2P all-reduce of dHidden beside the dW GEMM (0 = NCCL)")
    ap.add_argument("--dh-collective", default="reduce_scatter", choices=["reduce_scatter", "reduce_scatter_fused", "all_reduce"],
                    help="N > 1: every rank gets ITS token rows of dHidden (data-parallel layout, SURVEY 8e): owners pull "
                         "their rows from the peers' partial buffers beside the dW GEMM (default), or the K2a epilogue "
                         "stores tiles at their owners (fused); or the full dHidden (one-shot P2P all-reduce)")
    ap.add_argument("--tunable", action="append", default=[], help="name=value for o3v_set_tunable (diagnostics)")
    ap.add_argument("--no-chunk-plan", action="store_true", help="even chunks only (logprob.PLAN_CHUNKS = False)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the torch fp32 self-check after the timed regions (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
