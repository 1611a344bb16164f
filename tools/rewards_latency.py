"""Latency of the reference-named reward callables at a training-sized batch (64 rollouts), next to the
reference's CPU path (oracle port) on the same inputs."""
import json
import os
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from open_o3_video_b200 import rewards  # noqa: E402
from oracle import parse as oparse, rewards as orw  # noqa: E402

warnings.simplefilter("ignore")
FILLER = "The person then walks across the room and picks up the object on the table. "
for task, pad in (("temporal-spatial free-form QA", 0), ("visual QA", 0), ("temporal QA", 0),
                  ("temporal-spatial free-form QA", 8000), ("temporal-spatial free-form QA", 64000)):
    pool = [c for c in oparse.text_cases(900, 77) if c[1]["task"] == task]
    if pad:     # realistic completion lengths (2048 tokens ~ 8 KB): prose between the grounded claims
        pool = [(t.replace(" then ", " " + FILLER * (pad // len(FILLER) // max(1, t.count(" then "))) + " then "), kw)
                for t, kw in pool]
    # the trainer's layout: 8 prompts x G = 8 rollouts, each prompt's kwargs repeated G times (same objects)
    cases = [(pool[8 + q * 8 + g][0], pool[q][1]) for q in range(8) for g in range(8)]
    kw0 = cases[0][1]
    completions = [[{"role": "assistant", "content": t}] for t, _ in cases]
    kwargs = {k: [kw[k] for _, kw in cases] for k in kw0}
    kwargs["task"] = [task] * len(cases)
    kwargs["step_percent"] = [kw0["step_percent"]] * len(cases)
    fns = [rewards.reward_funcs_registry[n] for n in rewards.REWARD_NAMES]
    for _ in range(3):
        rewards._cache["key"] = None
        out = [f(prompts=None, completions=completions, **kwargs) for f in fns]
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        rewards._cache["key"] = None
        t0 = time.perf_counter()
        out = [f(prompts=None, completions=completions, **kwargs) for f in fns]
        ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter()
    for _ in range(5):
        ref = [orw.rewards_for_rollout(oparse.rollout_from_text(t, dict(kw, task=task, step_percent=kw0["step_percent"])))
               for t, kw in cases]
    t_ref = (time.perf_counter() - t0) / 5
    err = max(abs(out[j][i] - ref[i][j]) for i in range(len(cases)) for j in range(5))
    print(json.dumps(dict(task=task, rollouts=len(cases), text_bytes=sum(len(t) for t, _ in cases),
                          five_callables_ms=dict(best=min(ts) * 1e3, median=sorted(ts)[len(ts) // 2] * 1e3),
                          cpu_reference_ms=t_ref * 1e3, max_abs_err=err)), flush=True)
