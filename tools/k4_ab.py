"""K4 at BASELINE config 4 (65536 rollouts x 16): kernel time with L2 flushed between iterations (one process per
O3V_REWARDS_CTA setting (threads per CTA); run under gpurun)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_o3_video_b200 import rewards  # noqa: E402
from oracle import synth  # noqa: E402
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
ro = synth.rollouts(8192, 8, seed=4)
arrays, dims = rewards.pack_rollouts(ro, 8)
dev_arrays = rewards.to_device(arrays, dev)
out = torch.empty(dims["R"], 5, dtype=torch.float64, device=dev)
nbytes = rewards.soa_bytes(arrays) + out.numel() * 8
ts = []
for i in range(13):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); rewards.grounded_rewards_device(dev_arrays, dims, out); b.record(); torch.cuda.synchronize()
    if i >= 3:
        ts.append(a.elapsed_time(b) * 1e3)
print(json.dumps(dict(kernel="K4 c4", cta_threads=os.environ.get("O3V_REWARDS_CTA", "64"), best_us=min(ts),
                      mean_us=sum(ts) / len(ts), soa_bytes=nbytes, frac_of_hbm=nbytes / min(ts) / 1e3 / 6532.2)))
