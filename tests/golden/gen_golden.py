"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/gen_golden.py
Outputs (committed):
    logps_small.npz      reference _get_per_token_logps (grpo_trainer.py:371-384) on seeded inputs
    rewards_small.json   reference reward_func.py functions on rendered synthetic rollouts
    rewards_kat.json     the known-answer cases of SURVEY.md Appendix B, re-run here
    parse_cases.json     reference reward callables on seeded completion TEXT (well-formed and malformed):
                         pins the text extraction (regex / json / float) end to end
    gspo_small.npz       the reference's inline loss block (grpo_trainer.py:590-596, 635-636, 658, 675-681,
                         691-706, 737) executed from the reference file's own source lines
    compute_loss_small.json  the reference's whole compute_loss on the fake model / processor of
                         oracle/host_trainer.py (loss, metrics, gradient norms)
Inputs are regenerated from the seed by the tests (oracle/synth.py); only outputs
(and, for rewards, the small structured inputs) are stored.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import, synth, rewards as orw  # noqa: E402


def gen_logps():
    Trainer = ref_import.load_trainer_class()
    out = {}
    for name, (B, L, H, V, planted) in {
        "a": (2, 9, 64, 1000, False),
        "b": (3, 33, 128, 5003, True),
        "c": (1, 17, 256, 151936 // 16, False),
    }.items():
        hidden, weight, _ = synth.head_inputs(B * L, H, V, seed=synth.SEED + ord(name), planted=planted)
        g = torch.Generator().manual_seed(synth.SEED + 100 + ord(name))
        ids = torch.randint(0, V, (B, L), generator=g, dtype=torch.int64)
        model = ref_import.FakeLMHeadModel(hidden.view(B, L, H), weight)
        with torch.no_grad():
            lp = Trainer._get_per_token_logps(None, model, ids)
        out["logp_" + name] = lp.numpy()
        out["shape_" + name] = np.array([B, L, H, V, int(planted)])
    np.savez_compressed(os.path.join(HERE, "logps_small.npz"), **out)
    print("logps_small.npz", {k: v.shape for k, v in out.items()})


def gen_rewards():
    rf = ref_import.load_reward_func()
    rollouts = synth.rollouts(n_prompts=60, G=4, P=16, K=8, O=4, Gb=2, Bc=2)
    ref = orw.reference_rewards(rf, rollouts)
    with open(os.path.join(HERE, "rewards_small.json"), "w") as f:
        json.dump(dict(seed=synth.SEED, gen=dict(n_prompts=60, G=4, P=16, K=8, O=4, Gb=2, Bc=2),
                       names=list(orw.REWARD_NAMES),
                       expected=[[repr(float(x)) for x in row] for row in ref]), f)
    print("rewards_small.json", ref.shape, "nonzero per column", (ref != 0).sum(0))

    # Appendix B known answers, recomputed from the live reference
    base = dict(has_think=True, has_answer=True, ans_seg=None, ans_box=None, think_boxes=[],
                gt_seg=[10.0, 20.0], gt_vbox=None,
                key_frames=[{"idx": 3, "time": 5.4}, {"idx": 9, "time": 12.0}],
                key_items={"3": {"dog": [[.2, .2, .6, .6]]},
                           "9": {"cat": [[.4, .4, .8, .8]], "mouse": [[0, 0, .1, .1]]}},
                image_size=(500, 500), image_size_refine=(500, 500), step_percent=0.0)
    cases = []

    def add(name, **kw):
        r = dict(base); r.update(kw); cases.append((name, r))
    claims3 = [(5.0, [[100, 100, 300, 300]]), (12.5, [[10, 10, 50, 50], [200, 200, 400, 400]]),
               (0.5, [[0, 0, 10, 10]])]
    for sp in (0.0, 0.25, 0.74, 0.75, 1.0):
        add("tsfree_sp%s" % sp, task="temporal-spatial free-form QA", claims=claims3,
            think_times=[5.0, 12.5, 0.5], step_percent=sp)
    add("single_t99", task="temporal-spatial free-form QA", claims=[(99.0, [[100, 100, 300, 300]])], think_times=[99.0])
    add("single_t1", task="temporal-spatial free-form QA", claims=[(1.0, [[100, 100, 300, 300]])], think_times=[1.0])
    add("tiou_fwd", task="temporal QA", claims=[], think_times=[12.0, 25.5], ans_seg=[15.0, 30.0])
    add("tiou_rev", task="temporal QA", claims=[], think_times=[], ans_seg=[30.0, 15.0])
    add("tiou_mcq", task="temporal QA (MCQ)", claims=[], think_times=[10.0, 20.0, 20.01], ans_seg=[12.0, 18.0])
    add("vqa", task="visual QA", claims=[], think_times=[], think_boxes=[[10, 10, 60, 60], [100, 120, 300, 310]],
        ans_box=[100, 100, 300, 300], gt_vbox=[110.0, 90.0, 320.0, 300.0], image_size=(640, 360),
        image_size_refine=(448, 252))
    ref = orw.reference_rewards(rf, [c[1] for c in cases])
    with open(os.path.join(HERE, "rewards_kat.json"), "w") as f:
        json.dump([dict(name=n, rollout=r, expected=[repr(float(x)) for x in row])
                   for (n, r), row in zip(cases, ref)], f, indent=1)
    for (n, _), row in zip(cases, ref):
        print(n, row.tolist())
    iou = [rf.calculate_iou([0, 0, 1, 1], [0, 0, 1, 1]), rf.calculate_iou([0, 0, 1, 1], [2, 2, 3, 3]),
           rf.calculate_iou([1, 1, 1, 1], [1, 1, 1, 1]), rf.calculate_iou([0, 0, 1, 1], [0, 0, 1])]
    print("calculate_iou KAT", iou, type(iou[0]))


def gen_vstar():
    """Reference eval/test/eval_vstar.py numeric functions on synthetic result items."""
    rf = ref_import.load_vstar_functions()
    items = synth.vstar_items(150)
    rows = []
    for it in items:
        row = []
        for suffix in ("", "_2"):
            at = it.get("answer_temporal" + suffix)
            t = rf.calculate_temporal_iou(it["timestamps"], at) if at else 0.0
            sp = it.get("answer_spatial" + suffix)
            aps, miou = rf.calculate_spatial_metrics(it["bboxes"], sp) if sp else ([0.0] * 5, 0.0)
            row += [t, miou] + list(aps)
        rows.append([repr(float(x)) for x in row])
    with open(os.path.join(HERE, "vstar_small.json"), "w") as f:
        json.dump(dict(seed=synth.SEED, n=150, expected=rows), f)
    arr = np.array([[float(x) for x in r] for r in rows])
    print("vstar_small.json", arr.shape, "means", arr.mean(0).round(4).tolist())


def gen_parse():
    """Reference reward callables on (completion text, kwargs) cases from oracle/parse.text_cases, and the
    reference's own claim parser on think blocks."""
    import random
    import warnings
    from oracle import parse as op
    warnings.simplefilter("ignore")
    rf = ref_import.load_reward_func()
    n, seed = 1200, synth.SEED + 60
    cases = op.text_cases(n, seed)
    ref = op.reference_rewards_from_text(rf, cases)
    rng = random.Random(seed + 1)
    thinks = [op.synth_completion(rng, "temporal-spatial free-form QA", True) for _ in range(400)]
    claims = [[[c["timestamp"], c["bboxes"]] for c in rf.parse_temporal_spatial_reasoning_process(t)] for t in thinks]
    with open(os.path.join(HERE, "parse_cases.json"), "w") as f:
        json.dump(dict(seed=seed, n=n, names=list(orw.REWARD_NAMES),
                       expected=[[repr(float(x)) for x in row] for row in ref],
                       claims_seed=seed + 1, claims_n=400, claims=[repr(c) for c in claims]), f)
    print("parse_cases.json", ref.shape, "nonzero per column", (ref != 0).sum(0),
          "claims", sum(len(c) for c in claims))


GSPO_CASES = {
    # name: (N, Tc, G, off_policy, gspo, empty_row)
    "on_gspo": (8, 48, 4, False, True, None),
    "on_grpo": (8, 48, 4, False, False, None),
    "off_gspo": (8, 48, 4, True, True, None),
    "off_grpo": (8, 48, 4, True, False, None),
    "off_gspo_empty_seq": (8, 48, 4, True, True, 5),     # a sequence with an all-zero mask (mean_kl becomes NaN, :737)
    "off_gspo_g8": (16, 33, 8, True, True, None),
}


def gspo_case_inputs(name):
    """Seeded inputs of one GSPO golden case (shared with tests/test_oracle_golden.py)."""
    N, Tc, G, off, gs, empty = GSPO_CASES[name]
    d = synth.gspo_inputs(N, Tc, G, seed=synth.SEED + 7 * len(name), off_policy=off)
    return d, (N, Tc, G, off, gs, empty)


def gen_gspo():
    """The reference's INLINE loss block (grpo_trainer.py:590-596, 635-636, 658, 675-681, 691-706, 737), executed
    from the reference file's own source lines (oracle/ref_import.load_loss_block)."""
    mask_block, loss_block = ref_import.load_loss_block()
    out = {}
    for name in GSPO_CASES:
        d, (N, Tc, G, off, gs, empty) = gspo_case_inputs(name)
        m = mask_block(d["ids"], d["eos_id"])
        mask = m["completion_mask"].clone()
        if empty is not None:
            mask[empty] = 0
        lp = d["logp"].clone().requires_grad_(True)
        r = loss_block(lp, d["ref"], mask, d["rewards_per_func"], G, 0.04, 0.2, 0.2, gs, d["old"] if off else None)
        r["loss"].backward()
        out[name + "/eos_idx"] = m["eos_idx"].numpy()
        out[name + "/completion_mask"] = m["completion_mask"].numpy()
        for k in ("per_token_kl", "rewards", "advantages", "std_grouped_rewards", "loss", "mean_kl"):
            out[name + "/" + k] = r[k].detach().numpy()
        out[name + "/grad"] = lp.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "gspo_small.npz"), **out)
    print("gspo_small.npz", {k: float(out[k + "/loss"]) for k in GSPO_CASES})


def gen_compute_loss():
    """The reference's whole compute_loss (grpo_trainer.py:402-738) on the fakes of oracle/host_trainer.py."""
    from oracle import host_trainer as ht
    cls = ref_import.load_trainer_class()
    ht.install_reference_fakes(sys.modules[cls.__module__])
    res = {}
    for key, gs in (("gspo", True), ("grpo", False)):
        model, ref = ht.FakeVLModel(seed=11), ht.FakeVLModel(seed=12).eval()
        t = ht.configure(cls.__new__(cls), model, ref, gspo=gs)
        loss = t.compute_loss(model, [ht.make_example()])
        loss.backward()
        res[key] = dict(loss=repr(loss.item()), metrics={k: repr(float(v[0])) for k, v in t._metrics.items()},
                        g_head=repr(model.lm_head.weight.grad.norm().item()),
                        g_embed=repr(model.embed.weight.grad.norm().item()))
    with open(os.path.join(HERE, "compute_loss_small.json"), "w") as f:
        json.dump(res, f, indent=1)
    print("compute_loss_small.json", res["gspo"]["loss"], res["grpo"]["loss"])


def reward_failure_cases():
    """Batches (B prompts x G rollouts, kwargs repeated G times as grpo_trainer.py:650-654 builds them) in which
    the GROUND TRUTH of one prompt is unusable.  Shared with the tests."""
    seg = lambda a, b: "<think>at <t>%s</t>s and <t>%s</t>s</think><answer>From <t>%s</t>s to <t>%s</t>s</answer>" % (a, b, a, b)
    nomatch = "<think>hm <t>3.0</t>s</think><answer>no idea</answer>"
    kf = [{"idx": 3, "time": 5.0, "path": "a.jpg"}]
    ki = {"3": {"dog": [[0.1, 0.1, 0.5, 0.5]]}}
    def batch(task, answers, texts, G, key_frames=None):
        B = len(answers)
        rep = lambda xs: [x for x in xs for _ in range(G)]
        return dict(task=task, G=G, texts=texts,
                    kwargs=dict(task=rep([task] * B), answer=rep(answers), step_percent=rep([0.3] * B),
                                key_frames=rep(key_frames if key_frames is not None else [kf] * B),
                                key_items=rep([ki] * B), image_size=rep([(640, 360)] * B),
                                image_size_refine=rep([(448, 252)] * B)))
    six = [seg(12, 18), nomatch, seg(1, 2), seg(11.5, 30), seg(6, 8), nomatch]
    return {
        "tiou_bad_literal_middle": batch("temporal QA", ["[10.0, 20.0]", "oops(", "[5.0, 9.0]"], six, 2),
        "tiou_bad_literal_first": batch("temporal QA", ["[", "[10.0, 20.0]", "[5.0, 9.0]"], six, 2),
        "tiou_mcq_no_second_line": batch("temporal QA (MCQ)", ["B\n[10.0, 20.0]", "B", "C\n[5.0, 9.0]"], six, 2),
        "tiou_three_numbers": batch("temporal QA", ["[10.0, 20.0]", "[1, 2, 3]", "[5.0, 9.0]"], six, 2),
        "point_empty_key_frames": batch("temporal-spatial free-form QA", ["x", "y"], six[:4], 2, key_frames=[kf, []]),
        "point_empty_key_frames_no_times": batch("temporal-spatial free-form QA", ["x", "y"],
                                                 [six[0], six[1], "<think>none</think><answer>a</answer>", "plain"], 2,
                                                 key_frames=[kf, []]),
    }


def gen_reward_failures():
    """The live reference's five numeric reward callables on `reward_failure_cases()`: values, or the exception type
    a callable raises out of."""
    import contextlib
    import io
    rf = ref_import.load_reward_func()
    res = {}
    for name, c in reward_failure_cases().items():
        comps = [[{"role": "assistant", "content": t}] for t in c["texts"]]
        row = {}
        for fn in orw.REWARD_NAMES:
            try:
                with contextlib.redirect_stdout(io.StringIO()):
                    row[fn] = [repr(float(x)) for x in getattr(rf, fn)(completions=comps, prompts=None, **c["kwargs"])]
            except Exception as exc:                       # noqa: BLE001
                row[fn] = {"raises": type(exc).__name__}
        res[name] = row
        print(name, row)
    with open(os.path.join(HERE, "reward_failures.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 1:                     # e.g. `gen_golden.py gspo compute_loss`
        for name in sys.argv[1:]:
            globals()["gen_" + name]()
        sys.exit(0)
    gen_gspo()
    gen_compute_loss()
    gen_reward_failures()
    gen_parse()
    gen_logps()
    gen_rewards()
    gen_vstar()
