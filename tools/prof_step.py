"""Small fused step for ncu: one wave of work items per kernel (T = 37 x 128 x k tokens)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib, gspo, logprob  # noqa: E402

T_mult = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cta = int(sys.argv[2]) if len(sys.argv) > 2 else 1
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
_lib.set_tunable("cta_pair", cta)
H, V, N, G = 3584, 152064, 2, 2
Tc = 37 * 128 * T_mult // N
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
hidden = torch.randn(N, Tc, H, device=dev, generator=g).bfloat16()
weight = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
ids = torch.randint(0, V, (N, Tc), device=dev, generator=g)
mask = torch.ones(N, Tc, dtype=torch.int32, device=dev)
ref = torch.full((N, Tc), -12.0, device=dev)
rpf = torch.rand(N, 3, device=dev, generator=g)
for _ in range(steps):
    out = logprob.fused_logprob_gspo(hidden, weight, ids, ref, mask, rpf, G, 0.04)
torch.cuda.synchronize()
print("loss", float(out["loss"]), "tokens", N * Tc)
