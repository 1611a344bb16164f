"""Kernel-level timing sweep on a B200 (CUDA events).
Usage: python tools/sweep.py [T] [head] [sustain_iters]
Each kernel is timed back to back `sustain_iters` times (default 8) after a warm-up so that the
power cap / clock droop of a long step is part of the number; min and mean are printed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib, logprob  # noqa: E402


def timeit(fn, iters):
    for _ in range(2):
        fn()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    evs[0].record()
    for i in range(iters):
        fn()
        evs[i + 1].record()
    torch.cuda.synchronize()
    ts = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
    head = sys.argv[2] if len(sys.argv) > 2 else "7b"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    H, V = (3584, 152064) if head == "7b" else (4096, 151936)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    hidden = torch.randn(T, H, device=dev, generator=g).bfloat16()
    weight = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
    targets = torch.randint(0, V, (T,), device=dev, generator=g)
    z = torch.empty(T, V, dtype=torch.bfloat16, device=dev)
    dW = torch.zeros(V, H, device=dev)
    flops = 2.0 * T * H * V
    print("T=%d H=%d V=%d  (2THV = %.3e flop), %d back-to-back iterations" % (T, H, V, flops, iters), flush=True)

    def report(name, fn):
        try:
            best, mean = timeit(fn, iters)
            print("%-36s best %8.3f ms %7.1f TF | mean %8.3f ms %7.1f TF" %
                  (name, best, flops / best / 1e9, mean, flops / mean / 1e9), flush=True)
        except Exception as e:  # keep sweeping
            print("%-36s FAILED: %s" % (name, e), flush=True)

    # library reference point (not on the product path): cuBLAS bf16 GEMM of the same shape
    report("cublas hidden@W^T -> bf16", lambda: torch.matmul(hidden, weight.T, out=z))
    lse = torch.zeros(T, device=dev)
    gl = torch.full((T,), 1e-3, device=dev)
    for cta in (1, 2):
        _lib.set_tunable("cta_pair", cta)
        for groups in (4, 6):
            _lib.set_tunable("fwd_groups", groups)
            report("K1 fwd stats       cta%d groups=%d" % (cta, groups), lambda: logprob.lmhead_stats(hidden, weight, targets))
        for groups in (4,):
            _lib.set_tunable("fwd_groups", groups)
            report("K1 fwd stats+store cta%d groups=%d" % (cta, groups), lambda: logprob.lmhead_stats(hidden, weight, targets, 0, z))
        _lib.set_tunable("fwd_groups", 0)
        z.normal_(0, 1.0)
        report("dlogits (elementwise)", lambda: logprob.dlogits_(z, lse, gl, targets))
        z.normal_(0, 0.01)
        report("K2a dH = P.W        cta%d" % cta, lambda: logprob.bwd_dhidden(z, weight))
        report("K2b dW += P^T.h     cta%d" % cta, lambda: logprob.bwd_dweight(z, hidden, dW, True))
    # experiment: the same dH / dW problems with pre-transposed (K-major) operands
    import ctypes
    lib = _lib.load()
    P = lambda t: ctypes.c_void_p(t.data_ptr())
    st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    Wt = weight.T.contiguous()            # [H, V]
    dH = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
    def gemm(A, lda, B, ldb, M, N, K, amn, bmn, out, ldo, f32, acc):
        _lib.check(lib.o3v_debug_gemm(P(A), lda, P(B), ldb, M, N, K, amn, bmn, P(out), ldo, f32, acc, st()), "debug_gemm")
    for cta in (1, 2):
        _lib.set_tunable("cta_pair", cta)
        report("dH  A=P K-maj, B=W  MN-maj cta%d" % cta, lambda: gemm(z, V, weight, H, T, H, V, 0, 1, dH, H, 0, 0))
        report("dH  A=P K-maj, B=Wt K-maj  cta%d" % cta, lambda: gemm(z, V, Wt, V, T, H, V, 0, 0, dH, H, 0, 0))
    zt = z.T.contiguous()                 # [V, T]
    ht = hidden.T.contiguous()            # [H, T]
    for cta in (1, 2):
        _lib.set_tunable("cta_pair", cta)
        report("dW  A=P^T MN, B=h MN        cta%d" % cta, lambda: gemm(z, V, hidden, H, V, H, T, 1, 1, dW, H, 1, 1))
        report("dW  A=Pt K-maj, B=ht K-maj  cta%d" % cta, lambda: gemm(zt, T, ht, T, V, H, T, 0, 0, dW, H, 1, 1))
    report("cublas hidden@W^T -> bf16 (again)", lambda: torch.matmul(hidden, weight.T, out=z))
    _lib.set_tunable("cta_pair", 1)


if __name__ == "__main__":
    main()
