"""Run one kernel back to back for ~3 s while sampling nvidia-smi: is a slow kernel slow because
its SM clock is lower (power) or at the same clock (pipeline / memory)?"""
import os
import subprocess
import sys
import threading
import time

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
dW = torch.zeros(V, H, device=dev)
flops = 2.0 * T * H * V


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks.mem,temperature.gpu",
                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            rows.append([float(x) for x in line.split(",")])
    p.terminate()


def run(name, fn, seconds=3.0):
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
    pw = sum(r[1] for r in rows) / len(rows)
    print("%-28s %7.2f ms %7.1f TF | sm %4.0f MHz  %4.0f W  mem %4.0f MHz  %2.0f C  | TF per GHz %6.1f" %
          (name, ms, flops / ms / 1e9, clk, pw, rows[-1][2], rows[-1][3], flops / ms / 1e9 / (clk / 1000)), flush=True)


z.normal_(0, 0.01)
run("cublas", lambda: torch.matmul(hidden, weight.T, out=z))
run("K1 stats cta1", lambda: logprob.lmhead_stats(hidden, weight, targets))
run("K1 stats+store cta1", lambda: logprob.lmhead_stats(hidden, weight, targets, 0, z))
z.normal_(0, 0.01)
for cta, wide in ((1, 0), (2, 0), (2, 1)):
    _lib.set_tunable("cta_pair_bwd", cta)
    _lib.set_tunable("bwd_wide", wide)
    run("K2a dH cta%d wide%d" % (cta, wide), lambda: logprob.bwd_dhidden(z, weight))
    run("K2b dW cta%d wide%d" % (cta, wide), lambda: logprob.bwd_dweight(z, hidden, dW, True))
z.zero_()
run("K2a dH cta2 wide (P = 0)", lambda: logprob.bwd_dhidden(z, weight))
run("cublas again", lambda: torch.matmul(hidden, weight.T, out=z))
