"""Launch the backward GEMMs in a few schedule variants; run under
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
to see how tile order / size changes DRAM traffic."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from open_o3_video_b200 import _lib, logprob  # noqa: E402

H, V = 3584, 152064
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
weight = (torch.randn(V, H, device=dev, generator=g) * 0.02).bfloat16()
for T in (32768,):                     # 10 m-blocks of 256 rows = one wave of 70 pair tiles; bench chunk
    hidden = torch.randn(T, H, device=dev, generator=g).bfloat16()
    z = (torch.randn(T, V, device=dev, generator=g, dtype=torch.bfloat16) * 0.01)
    dW = torch.zeros(V, H, device=dev)
    for wide in (1, 0):
        _lib.set_tunable("bwd_wide", wide)
        for sync in (1, 0):
            _lib.set_tunable("bwd_sync", sync)
            logprob.bwd_dhidden(z, weight)
            logprob.bwd_dweight(z, hidden, dW, True)
            torch.cuda.synchronize()
            print("T=%d wide=%d sync=%d: dH, dW launched" % (T, wide, sync), flush=True)
    del hidden, z, dW
