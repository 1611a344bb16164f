"""Seeded synthetic inputs shared by tests, smoke() and bench.py.

Test infrastructure (see oracle/__init__.py).  Everything is built on the CPU
from `torch.Generator().manual_seed(seed)` / `random.Random(seed)` so that the
oracle and the CUDA path see identical bits.
"""
import random

import torch

SEED = 20261018

HEADS = {
    "qwen2.5-vl-7b": dict(H=3584, V=152064),
    "qwen3-vl-8b": dict(H=4096, V=151936),
}


def head_inputs(T, H, V, seed=SEED, w_scale=0.02, planted=False):
    """hidden ~ N(0,1) [T,H], W ~ N(0, w_scale^2) [V,H], both rounded to bf16 values (returned
    as fp32 holding bf16-representable numbers), targets ~ U{0..V-1} int64.
    planted=True adds a few +-60 logits (overflow stress for the online softmax)."""
    g = torch.Generator().manual_seed(seed)
    hidden = torch.randn(T, H, generator=g).bfloat16().float()
    weight = (torch.randn(V, H, generator=g) * w_scale).bfloat16().float()
    targets = torch.randint(0, V, (T,), generator=g, dtype=torch.int64)
    if planted:
        # make row v of W parallel to hidden[t] so that logit(t, v) ~ +-60
        for i in range(min(8, T)):
            t = (i * 7919) % T
            v = (i * 104729 + 13) % V
            sign = 1.0 if i % 2 == 0 else -1.0
            h = hidden[t]
            weight[v] = (sign * 60.0 * h / (h * h).sum()).bfloat16().float()
            if i % 4 == 0:
                targets[t] = v
    return hidden, weight, targets


def gspo_inputs(N, Tc, G, seed=SEED, eos_id=151645, vocab=152064, off_policy=False, F=3):
    """Completion ids with one planted EOS per sequence (ids after it random, some extra
    EOS later, one sequence without any EOS, one with EOS at position 0), plausible
    log-probs, ref = logp + N(0, 0.1^2), optional old = logp + N(0, 0.05^2), and per-function
    rewards U[0,1) with one all-equal group (std = 0)."""
    g = torch.Generator().manual_seed(seed + 1)
    ids = torch.randint(0, vocab - 1000, (N, Tc), generator=g, dtype=torch.int64)
    ids[ids == eos_id] = 0
    lens = torch.randint(max(Tc // 4, 1), Tc + 1, (N,), generator=g)
    for n in range(N):
        L = int(lens[n])
        if n == 1 % N:
            continue                      # no EOS at all -> eos_idx = Tc
        if n == 2 % N:
            L = 1                         # EOS at position 0
        ids[n, L - 1] = eos_id
        if L + 3 < Tc:
            ids[n, L + 3] = eos_id        # a later EOS must not matter
    logp = -(torch.rand(N, Tc, generator=g) * 6.0 + 0.01)
    ref = logp + torch.randn(N, Tc, generator=g) * 0.1
    ref[0, 0] = logp[0, 0] + 12.0         # exercise both clamp sides of the KL
    ref[0, min(1, Tc - 1)] = logp[0, min(1, Tc - 1)] - 12.0
    old = logp + torch.randn(N, Tc, generator=g) * 0.05 if off_policy else None
    if off_policy and N >= 4:
        old[2] = logp[2] - 0.5            # ratio > 1 + eps
        old[3] = logp[3] + 0.5            # ratio < 1 - eps
    rewards_per_func = torch.rand(N, F, generator=g)
    if N >= 2 * G:
        rewards_per_func[G:2 * G] = rewards_per_func[G]      # all-equal group
    return dict(ids=ids, eos_id=eos_id, logp=logp, ref=ref, old=old,
                rewards_per_func=rewards_per_func)


_SIZES = [(448, 252), (364, 364), (640, 360), (500, 500)]
_TASKS = ("visual QA", "temporal QA", "temporal QA (MCQ)",
          "temporal-spatial free-form QA", "General video QA MCQ", "General video QA Free-form")
_STEPS = (0.0, 0.25, 0.74, 0.75, 1.0)


def _r2(x):
    return round(float(x), 2)


def rollouts(n_prompts, G, P=16, K=8, O=4, Gb=1, Bc=2, seed=SEED, tasks=None):
    """Structured rollouts (see oracle/rewards.py `Rollout`), n_prompts x G, GT shared per
    prompt.  5% inverted / degenerate boxes, 10% duplicated key-frame times, malformed
    boxes (wrong arity), empty predictions, missing think/answer."""
    rng = random.Random(seed + 2)
    out = []
    for q in range(n_prompts):
        task = (tasks or _TASKS)[q % len(tasks or _TASKS)]
        W, H = _SIZES[q % len(_SIZES)]
        refine = _SIZES[(q + 1) % len(_SIZES)]
        dur = _r2(rng.uniform(10, 120))
        nk = rng.randint(1, K)
        times = sorted(_r2(rng.uniform(0, dur)) for _ in range(nk))
        for i in range(1, nk):
            if rng.random() < 0.10:
                times[i] = times[i - 1]
        key_frames = [{"idx": 3 * i + 1, "time": times[i]} for i in range(nk)]
        key_items = {}
        for f in key_frames:
            objs = {}
            for o in range(rng.randint(1, O)):
                boxes = []
                for _ in range(rng.randint(1, Gb)):
                    x0, x1 = sorted((_r2(rng.random()), _r2(rng.random())))
                    y0, y1 = sorted((_r2(rng.random()), _r2(rng.random())))
                    boxes.append([x0, y0, x1, y1])
                objs["obj%d" % o] = boxes
            key_items[str(f["idx"])] = objs
        a, b = sorted((_r2(rng.uniform(0, dur)), _r2(rng.uniform(0, dur))))
        gt_seg = [a, b]
        gx0, gx1 = sorted((_r2(rng.uniform(0, W)), _r2(rng.uniform(0, W))))
        gy0, gy1 = sorted((_r2(rng.uniform(0, H)), _r2(rng.uniform(0, H))))
        gt_vbox = [gx0, gy0, gx1, gy1] if rng.random() > 0.05 else None
        step = _STEPS[q % len(_STEPS)]
        for _ in range(G):
            def pbox():
                x0, x1 = sorted((_r2(rng.uniform(0, W)), _r2(rng.uniform(0, W))))
                y0, y1 = sorted((_r2(rng.uniform(0, H)), _r2(rng.uniform(0, H))))
                u = rng.random()
                if u < 0.025:
                    return [x1, y0, x0, y1]                 # inverted
                if u < 0.05:
                    return [x0, y0, x0, y1]                 # zero area
                if u < 0.06:
                    return [x0, y0, x1]                     # wrong arity -> IoU 0 (reward_func.py:361)
                if u < 0.07:
                    return [int(x0), int(y0), int(x1) + 1, int(y1) + 1]   # json ints
                return [x0, y0, x1, y1]
            n_claims = rng.randint(0, P)
            claims = []
            for _c in range(n_claims):
                if rng.random() < 0.5 and nk > 0:
                    t = max(0.0, _r2(rng.choice(times) + rng.uniform(-2.0, 2.0)))
                else:
                    t = _r2(rng.uniform(0, dur))
                claims.append((t, [pbox() for _ in range(rng.randint(1, Bc))]))
            extra = [_r2(rng.uniform(0, dur)) for _ in range(rng.randint(0, max(P - n_claims, 0)))]
            think_times = [c[0] for c in claims] + extra
            s, e = _r2(rng.uniform(0, dur)), _r2(rng.uniform(0, dur))
            if rng.random() < 0.8 and e < s:
                s, e = e, s
            u = rng.random()
            ans_seg = None if u < 0.1 else ([s, s] if u < 0.15 else [s, e])
            r = dict(task=task, has_think=rng.random() > 0.03, has_answer=rng.random() > 0.03,
                     ans_seg=ans_seg, ans_box=pbox() if rng.random() > 0.1 else None,
                     think_times=think_times,
                     think_boxes=[pbox() for _ in range(rng.randint(0, 4))] if task == "visual QA" else [],
                     claims=claims if task != "visual QA" else [],
                     gt_seg=gt_seg, gt_vbox=gt_vbox, key_frames=key_frames, key_items=key_items,
                     image_size=(W, H), image_size_refine=refine, step_percent=step)
            if task == "visual QA":
                r["think_times"] = extra
            if not r["has_answer"]:          # nothing to parse from a missing <answer>
                r["ans_seg"], r["ans_box"] = None, None
            if not r["has_think"]:
                r["think_times"], r["think_boxes"], r["claims"] = [], [], []
            out.append(r)
    return out


def vstar_items(n, F=12, Pb=3, seed=SEED):
    """Synthetic V-STAR result items in the JSON shape eval_vstar.py reads (:213-312): GT segment
    `timestamps`, GT `bboxes` (one per annotated second), two answer chains with a temporal range and
    a {frame_id: box | [boxes]} spatial answer; 10 % missing / malformed answers."""
    rng = random.Random(seed + 3)
    items = []
    for i in range(n):
        W, H = _SIZES[i % len(_SIZES)]
        dur = _r2(rng.uniform(5, 200))
        a, b = sorted((_r2(rng.uniform(0, dur)), _r2(rng.uniform(0, dur))))
        nf = rng.randint(0 if i % 17 == 0 else 1, F)
        stamps = sorted(rng.sample(range(0, 400), nf))
        bboxes = []
        for t in stamps:
            x0, x1 = sorted((rng.randint(0, W), rng.randint(0, W)))
            y0, y1 = sorted((rng.randint(0, H), rng.randint(0, H)))
            bboxes.append({"timestamp": t, "xmin": x0, "ymin": y0, "xmax": x1, "ymax": y1})
        item = dict(timestamps=[a, b], bboxes=bboxes, width=W, height=H)
        for suffix in ("", "_2"):
            u = rng.random()
            if u < 0.08:
                at = []
            elif u < 0.12:
                at = [_r2(rng.uniform(0, dur))]                      # wrong arity -> 0
            else:
                s_, e_ = _r2(rng.uniform(0, dur)), _r2(rng.uniform(0, dur))
                at = [min(s_, e_), max(s_, e_)] if rng.random() < 0.85 else [max(s_, e_), min(s_, e_)]
            item["answer_temporal" + suffix] = at
            sp = {}
            for bx in bboxes:
                if rng.random() < 0.75:
                    def pb():
                        x0, x1 = sorted((rng.randint(0, W), rng.randint(0, W)))
                        y0, y1 = sorted((rng.randint(0, H), rng.randint(0, H)))
                        if rng.random() < 0.3:        # near the GT box
                            x0, y0, x1, y1 = bx["xmin"] + rng.randint(-9, 9), bx["ymin"] + rng.randint(-9, 9), \
                                bx["xmax"] + rng.randint(-9, 9), bx["ymax"] + rng.randint(-9, 9)
                        v = rng.random()
                        if v < 0.04:
                            return [x0, y0, x1]       # malformed -> IoU 0
                        return [x0, y0, x1, y1]
                    k = rng.randint(1, Pb)
                    sp[str(bx["timestamp"])] = pb() if (k == 1 and rng.random() < 0.5) else [pb() for _ in range(k)]
            if rng.random() < 0.05:
                sp = {}
            item["answer_spatial" + suffix] = sp
        items.append(item)
    return items
