"""Sustained probe of K1 (+store) variants: tile mode x vocab groups."""
import os, subprocess, sys, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib, logprob  # noqa: E402

T, H, V = 32768, 3584, 152064
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
hidden = torch.randn(T, H, device=dev, generator=g).bfloat16()
weight = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
targets = torch.randint(0, V, (T,), device=dev, generator=g)
z = torch.empty(T, V, dtype=torch.bfloat16, device=dev)
flops = 2.0 * T * H * V


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            rows.append([float(x) for x in line.split(",")])
    p.terminate()


def run(name, fn, seconds=2.0):
    fn(); torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows)); th.start()
    time.sleep(0.3)
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < seconds:
        for _ in range(4):
            fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    rows = rows[5:] or rows
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    print("%-34s %7.2f ms %7.1f TF | sm %4.0f MHz | TF/GHz %6.1f" % (name, ms, flops / ms / 1e9, clk, flops / ms / 1e9 / (clk / 1000)), flush=True)


run("cublas", lambda: torch.matmul(hidden, weight.T, out=z))
for cta in (1, 2):
    _lib.set_tunable("cta_pair_fwd", cta)
    for groups in (4, 6, 8):
        _lib.set_tunable("fwd_groups", groups)
        run("K1 stats       cta%d g%d" % (cta, groups), lambda: logprob.lmhead_stats(hidden, weight, targets))
        run("K1 stats+store cta%d g%d" % (cta, groups), lambda: logprob.lmhead_stats(hidden, weight, targets, 0, z))
run("cublas again", lambda: torch.matmul(hidden, weight.T, out=z))
