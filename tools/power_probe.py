"""Operand-layout vs HBM-traffic power experiment (see profiles/r1_notes.md section 5)."""
import ctypes
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib  # noqa: E402

dev = "cuda"
lib = _lib.load()
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = lambda: ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def sample(stop, rows):
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    while not stop.is_set():
        line = p.stdout.readline()
        if line:
            rows.append([float(x) for x in line.split(",")])
    p.terminate()


def run(name, fn, flops, seconds=2.5, inner=16):
    fn(); torch.cuda.synchronize()
    stop, rows = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, rows)); th.start()
    time.sleep(0.3)
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time(); e0.record()
    while time.time() - t0 < seconds:
        for _ in range(inner):
            fn(); n += 1
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    ms = e0.elapsed_time(e1) / n
    rows = rows[5:] or rows
    clk = sorted(r[0] for r in rows)[len(rows) // 2]
    pw = sum(r[1] for r in rows) / len(rows)
    print("%-44s %8.3f ms %7.1f TF | sm %4.0f MHz %4.0f W | TF/GHz %6.1f" %
          (name, ms, flops / ms / 1e9, clk, pw, flops / ms / 1e9 / (clk / 1000)), flush=True)


def gemm(A, lda, B, ldb, M, N, K, amn, bmn, out, ldo, f32, acc):
    _lib.check(lib.o3v_debug_gemm(P(A), lda, P(B), ldb, M, N, K, amn, bmn, P(out), ldo, f32, acc, st()), "debug_gemm")


g = torch.Generator(device=dev).manual_seed(0)
for cta in (1, 2):
    _lib.set_tunable("cta_pair", cta)
    # ---- L2-resident problem: M=8192, N=3584, K=3584 (A 58 MB, B 26 MB)
    M, N, K = 8192, 3584, 3584
    fl = 2.0 * M * N * K
    A_k = torch.randn(M, K, device=dev, generator=g).bfloat16()          # [M,K] K-major
    A_m = A_k.T.contiguous()                                             # [K,M] MN-major
    B_k = (torch.randn(N, K, device=dev, generator=g) * 0.02).bfloat16()  # [N,K] K-major
    B_m = B_k.T.contiguous()                                             # [K,N] MN-major
    out16 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    out32 = torch.zeros(M, N, dtype=torch.float32, device=dev)
    run("cta%d L2-res  A K / B K   store" % cta, lambda: gemm(A_k, K, B_k, K, M, N, K, 0, 0, out16, N, 0, 0), fl)
    run("cta%d L2-res  A K / B MN  store" % cta, lambda: gemm(A_k, K, B_m, N, M, N, K, 0, 1, out16, N, 0, 0), fl)
    run("cta%d L2-res  A K / B K   accum" % cta, lambda: gemm(A_k, K, B_k, K, M, N, K, 0, 0, out32, N, 1, 1), fl)
    run("cta%d L2-res  A MN / B MN accum" % cta, lambda: gemm(A_m, M, B_m, N, M, N, K, 1, 1, out32, N, 1, 1), fl)
    del A_k, A_m, B_k, B_m, out16, out32
    # ---- HBM-streaming problem with the dH shape but K-major everywhere: A [T, V] 10 GB
    T, H, V = 32768, 3584, 152064
    fl = 2.0 * T * H * V
    Pm = (torch.randn(T, V, device=dev, generator=g, dtype=torch.bfloat16) * 0.01)
    Wt = (torch.randn(H, V, device=dev, generator=g, dtype=torch.bfloat16) * 0.02)      # [N=H, K=V] K-major
    Wm = Wt.T.contiguous()                                                              # [K=V, N=H] MN-major
    dH = torch.empty(T, H, dtype=torch.bfloat16, device=dev)
    run("cta%d dH-shape  A K / B K (Wt)" % cta, lambda: gemm(Pm, V, Wt, V, T, H, V, 0, 0, dH, H, 0, 0), fl, inner=4)
    run("cta%d dH-shape  A K / B MN (W)" % cta, lambda: gemm(Pm, V, Wm, H, T, H, V, 0, 1, dH, H, 0, 0), fl, inner=4)
    del Pm, Wt, Wm, dH
    torch.cuda.empty_cache()
_lib.set_tunable("cta_pair_fwd", 1); _lib.set_tunable("cta_pair_bwd", 2)
