"""GPU parity of K6 (o3v_parse_completions, csrc/parse.cu) through the C ABI: bit-exact against the oracle
(oracle/parse.py = the reference's own re / json / float calls) and, end to end with K4, against the golden
rewards the live reference produced from the same completion text."""
import json
import os
import warnings

import numpy as np
import pytest
import torch

from oracle import parse as op
from oracle import rewards as orw

import scan_host
from test_parse_cpu import EDGE_TEXTS

pytestmark = pytest.mark.gpu


def _device_parse(texts, tasks, G=1, caps=None, sync=True):
    from open_o3_video_b200 import rewards
    text, offsets = rewards.encode_completions(texts)
    task = torch.tensor([rewards.TASK_IDS[t] for t in tasks[::G]], dtype=torch.int32).cuda()
    rows, caps = rewards.parse_completions_device(text.cuda(), offsets.cuda(), task, G, caps, sync)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in rows.items()}, caps


def _check(texts, tasks, G=1, caps=None):
    got, caps = _device_parse(texts, tasks, G, caps)
    exp = op.pack([op.parse_text(t, k) for t, k in zip(texts, tasks)], caps["P"], caps["C"], caps["Bc"], caps["Tb"])
    bad = scan_host.mismatches(got, exp, op.used_mask(exp))
    assert not bad, [(r, k, texts[r][:100]) for r, k in sorted(bad)[:5]]
    return got, exp, caps


@pytest.mark.parametrize("seed", [1, 2])
def test_parse_bit_exact_on_wild_text(seed):
    texts, tasks = op.synth_batch(8000, seed)
    got, exp, caps = _check(texts, tasks)
    assert (exp["n_claims"] > 0).sum() > 400 and (exp["n_times"] > 0).sum() > 1500


@pytest.mark.parametrize("seed", [101, 107])
def test_parse_bit_exact_on_mutated_text(seed):
    texts, tasks = op.mutate_batch(8000, seed)
    _check(texts, tasks)


@pytest.mark.parametrize("task", op.TASKS)
def test_parse_edge_cases(task):
    _check(EDGE_TEXTS, [task] * len(EDGE_TEXTS))


def test_parse_alignment_and_long_text():
    texts = []
    for pad in list(range(0, 40)) + [495, 496, 505, 511, 512, 513, 1023, 1024, 1030, 70000]:
        texts.append("p" * pad + "<think>" + "q" * (pad % 7) + "<t>%d.5</t>s" % pad + "</think>" + "r" * pad
                     + "<answer>From <t>1</t>s to <t>%d</t>s</answer>" % pad)
    got, exp, _ = _check(texts, ["temporal QA"] * len(texts))
    assert (got["n_times"] == 1).all()


def test_parse_long_completions_cross_the_mask_cache():
    """Completions longer than the 8 cached 512-byte blocks: padding inserted before / inside / after <think>, so
    that spans and claim chains start inside the shared-memory cache and end outside it (register-cached blocks,
    sequential chain) or lie entirely outside."""
    import random
    rng = random.Random(3)
    base, tasks = op.synth_batch(1500, 21)
    texts = []
    for t in base:
        pad = "x" * rng.choice([0, 500, 3000, 3600, 4090, 4100, 6000, 9000])
        where = rng.random()
        if where < 0.3:
            t = pad + t
        elif where < 0.6:
            t = t.replace("<think>", "<think>" + pad, 1)
        elif where < 0.8:
            i = rng.randrange(len(t) + 1)
            t = t[:i] + pad + t[i:]
        else:
            t = t.replace("</think>", pad + "</think>", 1)
        texts.append(t)
    _check(texts, tasks)


def test_parse_overflow_relaunch_and_report():
    texts, tasks = op.synth_batch(2000, 9)
    small = dict(P=2, C=1, Bc=1, Tb=1)
    # sync=False: one pass; the report names capacities that fit
    got, _ = _device_parse(texts, tasks, caps=small, sync=False)
    full = op.pack([op.parse_text(t, k) for t, k in zip(texts, tasks)], 64, 64, 32, 32)
    ov = got["overflow"]
    assert ov[0] >= full["n_times"].max() and ov[1] >= full["n_claims"].max() and ov[3] >= full["n_tboxes"].max()
    # sync=True grows the rows until everything fits, then the result is exact
    got, exp, caps = _check(texts, tasks, caps=small)
    assert caps["P"] >= exp["n_times"].max() and caps["C"] >= exp["n_claims"].max()
    assert not got["overflow"].any()


def test_parse_empty_and_ragged():
    from open_o3_video_b200 import rewards
    got, _ = _device_parse([], [])
    assert got["flags"].shape == (0,)
    got, exp, _ = _check(["", "", "<think></think>", ""], ["visual QA"] * 4)
    assert got["flags"].tolist() == [0, 0, 1, 0]
    with pytest.raises(RuntimeError):
        rewards.parse_completions_device(torch.zeros(16, dtype=torch.uint8), torch.zeros(1, dtype=torch.int64),
                                         torch.zeros(1, dtype=torch.int32))


def test_parse_c4_scale_properties():
    """65 536 rollouts (BASELINE config 4 scale): oracle on a subset, permutation equivariance on the rest."""
    base, tasks = op.synth_batch(4096, 33, wild=False)
    reps = 16
    texts, tk = base * reps, tasks * reps
    got, caps = _device_parse(texts, tk)
    exp = op.pack([op.parse_text(t, k) for t, k in zip(base, tasks)], caps["P"], caps["C"], caps["Bc"], caps["Tb"])
    m = op.used_mask(exp)
    for rep in (0, 7, 15):
        sl = {k: got[k][rep * 4096:(rep + 1) * 4096] for k in exp}
        assert not scan_host.mismatches(sl, exp, m)
    perm = np.random.RandomState(0).permutation(len(texts))
    got_p, _ = _device_parse([texts[i] for i in perm], [tk[i] for i in perm], caps=caps)
    for k in ("flags", "n_times", "n_claims", "n_tboxes"):
        assert np.array_equal(got_p[k], got[k][perm])


def test_text_to_rewards_matches_golden_reference(golden_dir):
    """K6 + K4 through the reference-named callables == the live reference's rewards on the same text."""
    from open_o3_video_b200 import rewards
    g = json.load(open(os.path.join(golden_dir, "parse_cases.json")))
    cases = op.text_cases(g["n"], g["seed"])
    exp = np.array([[float(x) for x in row] for row in g["expected"]])
    got = np.zeros_like(exp)
    fns = [rewards.reward_funcs_registry[n] for n in g["names"]]
    for i, (text, kw) in enumerate(cases[:300]):              # one rollout per call, as the golden was made
        completions = [[{"role": "assistant", "content": text}]]
        kwargs = {k: [v] for k, v in kw.items()}
        for j, fn in enumerate(fns):
            got[i, j] = fn(prompts=None, completions=completions, **kwargs)[0]
    np.testing.assert_allclose(got[:300], exp[:300], rtol=0, atol=1e-6)     # north_star: rewards within 1e-6
    assert np.array_equal(got[:300, [0, 1, 2, 4]], exp[:300, [0, 1, 2, 4]])  # IoU / ratio columns are exact
    # batched: all cases of one task in a single call (GT differs per rollout, G = 1)
    from collections import defaultdict
    by_task = defaultdict(list)
    for i, (text, kw) in enumerate(cases):
        by_task[(kw["task"], kw["step_percent"])].append(i)
    for (task, step), idx in by_task.items():
        contents = [cases[i][0] for i in idx]
        gts = []
        for i in idx:
            kw = cases[i][1]
            seg, vbox = rewards.parse_gt_answer(task, kw["answer"])
            gts.append(dict(task=task, step_percent=step, gt_seg=seg, gt_vbox=vbox, key_frames=kw["key_frames"],
                            key_items=kw["key_items"], image_size=kw["image_size"],
                            image_size_refine=kw["image_size_refine"]))
        out = rewards.rewards_from_text(contents, gts, 1).cpu().numpy()
        np.testing.assert_allclose(out, exp[idx], rtol=0, atol=1e-6)


def test_reward_callables_share_ground_truth_within_a_group():
    """The trainer repeats each prompt's kwargs G times (same objects): the ground truth is packed once per prompt
    and the result equals the per-rollout oracle."""
    from open_o3_video_b200 import _lib, rewards
    import warnings
    G = 4
    for task in ("temporal-spatial free-form QA", "visual QA", "temporal QA (MCQ)"):
        cases = [c for c in op.text_cases(900, 31) if c[1]["task"] == task]
        prompts = cases[:6]
        texts, kws = [], []
        for q, (_, kw) in enumerate(prompts):
            for g in range(G):
                texts.append(cases[6 + q * G + g][0])          # a different completion for every rollout
                kws.append(kw)                                 # the SAME ground-truth objects G times
        completions = [[{"role": "assistant", "content": t}] for t in texts]
        kwargs = {k: [kw[k] for kw in kws] for k in kws[0]}
        kwargs["step_percent"] = [kws[0]["step_percent"]] * len(kws)
        assert rewards._shared_gt_group(kwargs, len(texts)) == G
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = np.array([orw.rewards_for_rollout(op.rollout_from_text(t, dict(kw, step_percent=kws[0]["step_percent"])))
                             for t, kw in zip(texts, kws)])
        launches = []
        cols = []
        for n in rewards.REWARD_NAMES:                         # as grpo_trainer.py:646-656: kwargs lists rebuilt per callable
            rebuilt = {k: list(v) for k, v in kwargs.items()}
            t = _lib.Trace()
            _lib.trace = t
            try:
                cols.append(rewards.reward_funcs_registry[n](prompts=None, completions=completions, **rebuilt))
            finally:
                _lib.trace = None
            launches.append(t.launches)
        np.testing.assert_allclose(np.array(cols).T, want, rtol=0, atol=1e-6)
        assert launches[0] > 0 and not any(launches[1:])       # one GPU pass serves all five callables


def test_more_than_32_boxes_score_like_the_reference():
    """Round-1 advisor finding: > 32 boxes in one claim / think block used to raise ValueError out of the reward
    callable and abort the step; the reference scores such rollouts normally."""
    from open_o3_video_b200 import rewards
    from test_parse_cpu import many_box_texts
    import warnings
    cases = many_box_texts()
    got_rows, exp_rows, caps = _check([t for t, _ in cases], [k for _, k in cases])
    assert caps["Bc"] >= 70 and caps["Tb"] >= 45
    kf = [{"idx": 3, "time": 5.4, "path": "a"}, {"idx": 9, "time": 12.0, "path": "b"}]
    ki = {"3": {"dog": [[.05, .1, .5, .6]], "x": [[.2, .2, .9, .9], [.0, .0, .1, .1]]}, "9": {"cat": [[.02, .03, .1, .14]]}}
    for text, task in cases:
        kw = dict(task=task, answer="<box>[120, 90, 310, 300]</box>" if task == "visual QA" else "free", key_frames=kf,
                  key_items=ki, image_size=(500, 400), image_size_refine=(448, 364), step_percent=0.3)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = orw.rewards_for_rollout(op.rollout_from_text(text, kw))
        completions = [[{"role": "assistant", "content": text}]]
        got = [rewards.reward_funcs_registry[n](prompts=None, completions=completions, **{k: [v] for k, v in kw.items()})[0]
               for n in rewards.REWARD_NAMES]
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
        assert got[4] > 0.0                                    # the spatial reward really depends on the boxes
    # the parsed-rollout route (pack_rollouts) takes them too
    ro = [op.rollout_from_text(t, dict(task=k, answer="<box>[120, 90, 310, 300]</box>" if k == "visual QA" else "free",
                                       key_frames=kf, key_items=ki, image_size=(500, 400), image_size_refine=(448, 364),
                                       step_percent=0.3)) for t, k in cases]
    for r in ro:
        out = rewards.rewards_from_rollouts([r], 1).cpu().numpy()[0]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            np.testing.assert_allclose(out, orw.rewards_for_rollout(r), rtol=0, atol=1e-6)
