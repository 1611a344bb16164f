"""Roofline of the HBM-bound kernels (K3 GSPO, K3a EOS mask, K4 rewards, dlogits, merge):
achieved GB/s of algorithmic bytes against the measured copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_o3_video_b200 import gspo, logprob, rewards  # noqa: E402
from oracle import synth  # noqa: E402

dev = "cuda"
peak = 6532.2
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(p):
    peak = json.load(open(p))["hbm_gbs"]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                               # inputs here are smaller than the 126 MB L2: flush between iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def line(name, nbytes, fn):
    best, mean = timeit(fn)
    print(json.dumps(dict(kernel=name, algorithmic_bytes=nbytes, best_us=best * 1e3, mean_us=mean * 1e3,
                          achieved_gbs=nbytes / best / 1e6, peak_gbs=peak, frac=nbytes / best / 1e6 / peak)), flush=True)


g = torch.Generator(device=dev).manual_seed(0)
for name, N, Tc, G in (("c2", 64, 2048, 8), ("c3", 128, 4096, 8), ("c5", 16, 16384, 16)):
    lp = -torch.rand(N, Tc, device=dev, generator=g) * 5
    ref = lp + 0.1
    old = lp + 0.01
    ids = torch.randint(0, 1000, (N, Tc), device=dev, generator=g)
    rpf = torch.rand(N, 3, device=dev, generator=g)
    _, mask = gspo.eos_mask(ids, 7)
    line("K3a eos_mask %s" % name, N * Tc * 12, lambda: gspo.eos_mask(ids, 7))
    # logp, ref, mask read twice (two passes) + old + grad written: count each tensor once = 20 B/token
    line("K3 gspo fwd+bwd %s" % name, N * Tc * 20, lambda: gspo.gspo_raw(lp, ref, mask, rpf, G, 0.04, 0.2, 0.2, True, old, want_kl=False))
# K4 at BASELINE config 4: 65536 rollouts x 16 predictions
ro = synth.rollouts(8192, 8, seed=4)
arrays, dims = rewards.pack_rollouts(ro, 8)
dev_arrays = rewards.to_device(arrays, dev)
out = torch.empty(dims["R"], 5, dtype=torch.float64, device=dev)
nbytes = rewards.soa_bytes(arrays) + out.numel() * 8
line("K4 grounded rewards c4 (65536 x 16, %.0f B/rollout)" % (nbytes / dims["R"]), nbytes,
     lambda: rewards.grounded_rewards_device(dev_arrays, dims, out))
# K6 at the same scale: the 65536 rollouts rendered to completion text, scanned on the device
import time  # noqa: E402
from oracle import parse as oparse, rewards as orw  # noqa: E402
texts = [orw.render(r)[0] for r in ro]
text, offsets = rewards.encode_completions(texts)
d_text, d_off = text.to(dev), offsets.to(dev)
d_task = dev_arrays["task"]
rows, caps = rewards.parse_completions_device(d_text, d_off, d_task, 8)
n_text = int(offsets[-1])
written = int(dims["R"] * 20 + (rows["n_times"].sum() * 8 + rows["n_claims"].sum() * 16).item()
              + (rows["claim_nbox"] * (torch.arange(caps["C"], device=dev)[None, :] < rows["n_claims"][:, None])).sum().item() * 32)
line("K6 parse completions c4 (65536 rollouts, %.0f B text/rollout, rows P=%d C=%d Bc=%d)" %
     (n_text / dims["R"], caps["P"], caps["C"], caps["Bc"]), n_text + written,
     lambda: rewards.parse_completions_device(d_text, d_off, d_task, 8, caps, sync=False))
# end to end through the host API: Python strings in, [R, 5] rewards on the host out (UTF-8 encode, H2D, K6, overflow
# read-back, GT pack, K4, D2H), next to the reference's whole CPU path (regex + numerics) on a sample
gts = [ro[q * 8] for q in range(len(ro) // 8)]
rewards.rewards_from_text(texts[:4096], gts[:512], 8, dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
e2e_out = rewards.rewards_from_text(texts, gts, 8, dev).cpu()
dt_e2e = time.perf_counter() - t0
kws = [orw.render(r)[1] for r in ro[:2048]]
t0 = time.perf_counter()
ref_out = [orw.rewards_for_rollout(oparse.rollout_from_text(tx, kw)) for tx, kw in zip(texts[:2048], kws)]
dt_ref = time.perf_counter() - t0
assert float((e2e_out[:2048] - torch.tensor(ref_out, dtype=torch.float64)).abs().max()) < 1e-6
print(json.dumps(dict(kernel="text -> rewards end to end (rewards_from_text: encode + H2D + K6 + K4 + D2H), 65536 rollouts",
                      seconds=dt_e2e, rollouts_per_s=len(texts) / dt_e2e,
                      cpu_reference_rollouts_per_s=2048 / dt_ref, cpu_sample="2048 rollouts, oracle port of the whole "
                      "reward path (regex + numerics), 1 core", speedup_vs_one_core=(len(texts) / dt_e2e) / (2048 / dt_ref))),
      flush=True)
t0 = time.perf_counter()
for tx, r in zip(texts[:4096], ro[:4096]):
    oparse.parse_text(tx, r["task"])
dt = time.perf_counter() - t0
print(json.dumps(dict(kernel="K6 cpu baseline: reference regex/json/float path (oracle port, 1 core)",
                      sample="4096 of the 65536 rollouts", rollouts_per_s=4096 / dt,
                      text_mb_per_s=sum(len(t) for t in texts[:4096]) / dt / 1e6)), flush=True)
# dlogits / merge at the bench chunk
T, V = 8192, 152064
z = torch.randn(T, V, device=dev, generator=g, dtype=torch.bfloat16)
lse = torch.full((T,), 12.0, device=dev); gl = torch.full((T,), 1e-3, device=dev)
tg = torch.randint(0, V, (T,), device=dev, generator=g)
line("dlogits (T=8192, V=152064)", T * V * 4, lambda: logprob.dlogits_(z, lse, gl, tg))
parts = torch.randn(8, 3, 131072, device=dev, generator=g)
line("merge_stats (P=8, T=131072)", 8 * 3 * 131072 * 4 + 2 * 131072 * 4, lambda: logprob.merge_stats(parts))
