"""Kernel-level timing sweep on a B200 (CUDA events, L2 flushed between iterations).
Usage: python tools/sweep.py [T] [head]   -> one line per (kernel, tile mode, knob)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib, logprob  # noqa: E402


def timeit(fn, iters=3, warmup=1):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    head = sys.argv[2] if len(sys.argv) > 2 else "7b"
    H, V = (3584, 152064) if head == "7b" else (4096, 151936)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    hidden = torch.randn(T, H, device=dev, generator=g).bfloat16()
    weight = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
    targets = torch.randint(0, V, (T,), device=dev, generator=g)
    z = torch.empty(T, V, dtype=torch.bfloat16, device=dev)
    dW = torch.zeros(V, H, device=dev)
    flops = 2.0 * T * H * V
    print("T=%d H=%d V=%d  (2THV = %.3e flop)" % (T, H, V, flops), flush=True)

    def report(name, fn):
        try:
            best, mean = timeit(fn)
            print("%-34s best %8.3f ms  mean %8.3f ms  %7.1f TFLOP/s" % (name, best, mean, flops / best / 1e9), flush=True)
        except Exception as e:  # keep sweeping
            print("%-34s FAILED: %s" % (name, e), flush=True)

    # library reference point (not on the product path): cuBLAS bf16 GEMM of the same shape
    report("cublas hidden@W^T -> bf16", lambda: torch.matmul(hidden, weight.T, out=z))
    for cta in (1, 2):
        _lib.set_tunable("cta_pair", cta)
        for groups in (0, 1, 2, 4, 8, 16):
            _lib.set_tunable("fwd_groups", groups)
            report("K1 fwd stats      cta%d groups=%d" % (cta, groups), lambda: logprob.lmhead_stats(hidden, weight, targets))
        _lib.set_tunable("fwd_groups", 0)
        report("K1 fwd stats+store cta%d" % cta, lambda: logprob.lmhead_stats(hidden, weight, targets, 0, z))
        lse = torch.zeros(T, device=dev); gl = torch.full((T,), 1e-3, device=dev)
        report("dlogits (elementwise)", lambda: logprob.dlogits_(z, lse, gl, targets))
        z.normal_(0, 0.01)
        report("K2a dH = P.W        cta%d" % cta, lambda: logprob.bwd_dhidden(z, weight))
        report("K2b dW += P^T.h     cta%d" % cta, lambda: logprob.bwd_dweight(z, hidden, dW, True))
    _lib.set_tunable("cta_pair", 1)


if __name__ == "__main__":
    main()
