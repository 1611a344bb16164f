"""K6 (completion text -> struct of arrays) on the CPU box:
  * the oracle (oracle/parse.py) against the golden vectors made by the live reference, and against the live
    reference itself when /root/reference is present;
  * the scanner's logic (csrc/scan_core.cuh, compiled for the host by tests/hostbuild) against the oracle,
    bit-exact, on seeded well-formed and malformed text; float() and json.loads acceptance and rounding.
The GPU kernel itself is checked in tests/test_gpu_parse.py."""
import json
import math
import os
import random
import struct
import warnings

import numpy as np
import pytest

from oracle import parse as op
from oracle import ref_import
from oracle import rewards as orw

import scan_host


def _bits(x):
    return struct.pack("d", x)


def _same(a, b):
    if a is None or b is None:
        return a is b
    return _bits(a) == _bits(b) or (a != a and b != b)


# ------------------------------------------------------------------------------- oracle vs reference
def test_oracle_matches_golden_reference_rewards(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "parse_cases.json")))
    cases = op.text_cases(g["n"], g["seed"])
    exp = np.array([[float(x) for x in row] for row in g["expected"]])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = np.array([orw.rewards_for_rollout(op.rollout_from_text(t, kw)) for t, kw in cases])
    assert got.shape == exp.shape == (g["n"], 5)
    assert (exp != 0).sum(0).min() > 20                      # every reward column is exercised
    assert np.array_equal(got.view(np.uint64), exp.view(np.uint64))
    rng = random.Random(g["claims_seed"])
    thinks = [op.synth_completion(rng, "temporal-spatial free-form QA", True) for _ in range(g["claims_n"])]
    for t, want in zip(thinks, g["claims"]):
        assert repr([[ts, bx] for ts, bx in op.parse_claims(t)]) == want


@pytest.mark.skipif(not ref_import.reference_available(), reason="live reference only in the build container")
def test_oracle_matches_live_reference():
    rf = ref_import.load_reward_func()
    cases = op.text_cases(400, 4242)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = op.reference_rewards_from_text(rf, cases)
        got = np.array([orw.rewards_for_rollout(op.rollout_from_text(t, kw)) for t, kw in cases])
    assert np.array_equal(got.view(np.uint64), ref.view(np.uint64))


# ------------------------------------------------------------------------------- scanner logic vs oracle
def _compare(texts, tasks, P=16, C=16, Bc=4, Tb=8, G=1, grow=True):
    """Host build of the scanner vs the oracle.  Like the product wrapper, rows are grown and the scan repeated
    while the overflow report is non-zero (grow=False: single pass, for the report itself)."""
    buf, off = op.encode(texts)
    task_ids = [op.TASKS.index(t) for t in tasks[::G]]
    while True:
        got, ov = scan_host.parse(buf, off, task_ids, G, P, C, Bc, Tb)
        if not grow or not ov.any():
            break
        P, C, Bc, Tb = (max(c, int(o)) for c, o in zip((P, C, Bc, Tb), ov))
    exp = op.pack([op.parse_text(t, k) for t, k in zip(texts, tasks)], P, C, Bc, Tb)
    return scan_host.mismatches(got, exp, op.used_mask(exp)), got, exp, ov


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_scanner_bit_exact_on_wild_text(seed):
    texts, tasks = op.synth_batch(6000, seed)
    bad, got, exp, ov = _compare(texts, tasks)
    assert not bad, sorted(bad)[:5]
    assert (exp["n_claims"] > 0).sum() > 300 and (exp["n_times"] > 0).sum() > 1000 and (exp["n_tboxes"] > 0).sum() > 100


@pytest.mark.parametrize("seed", [101, 102])
def test_scanner_bit_exact_on_mutated_text(seed):
    texts, tasks = op.mutate_batch(8000, seed)
    bad, got, exp, ov = _compare(texts, tasks)
    assert not bad, [(r, k, texts[r][:120]) for r, k in sorted(bad)[:5]]


def test_scanner_capacity_overflow_report():
    """Too-small rows: the report names a capacity that fits, rollouts that fitted are already exact."""
    texts, tasks = op.synth_batch(3000, 9)
    bad, got, exp, ov = _compare(texts, tasks, P=2, C=1, Bc=1, Tb=1, grow=False)
    full = op.pack([op.parse_text(t, k) for t, k in zip(texts, tasks)], 64, 64, 32, 32)
    nb = np.where(np.arange(64)[None, :] < full["n_claims"][:, None], full["claim_nbox"], 0).max(1)
    assert ov[0] >= full["n_times"].max() and ov[1] >= full["n_claims"].max() and ov[3] >= full["n_tboxes"].max()
    assert ov[2] > 1
    fitted = (full["n_times"] <= 2) & (full["n_claims"] <= 1) & (nb <= 1) & (full["n_tboxes"] <= 1)
    # candidates are an upper bound of the matches: a rollout can be reported although it would have fitted,
    # but one that is NOT reported must be exact
    sure = fitted & np.array([t.count("<t>") <= 2 and t.count("<obj>") <= 1 and t.count("<box>[") <= 1 for t in texts])
    assert sure.sum() > 500
    assert not {(r, k) for r, k in bad if sure[r]}
    bad, got, exp, ov = _compare(texts, tasks, P=2, C=1, Bc=1, Tb=1)      # grown until it fits
    assert not bad and not ov.any()


EDGE_TEXTS = [
    "", "<think>", "<think></think>", "</think><think>", "<answer></answer>", "<think><t>5</t>s</think>",
    "<think><t>5</t>s<t>1.2.3</t>s</think><answer>x</answer>",                     # one bad float empties the list
    "<think><t></t>s<t>.</t>s<t>5.</t>s</think>",
    "<think>a</think><think><t>9</t>s</think>",                                      # first think span only
    "<answer>From <t>3</t>s to <t>7.5</t>s</answer>", "<answer> From <t>3</t>s  to <t>7</t>s</answer>",
    "<answer>From <t>.5</t>s to <t>7</t>s <t>1</t>s to <t>2.</t>s</answer>",
    "<answer>\n<t>١٢</t>s to <t>１３.５</t>s </answer>",
    "<think><obj>a</obj><box>[1,2,3,4]</box>at<t>5</t>s</think><answer>B</answer>",
    "<think><obj>a</obj> junk <obj>b</obj><box>[1,2,3,4]</box>at<t> 5.0\n</t>s</think>",
    "<think><obj>a</obj><box>[1]</box>xx<box>[2,3,4,5]</box>at<t>1e1</t>s<obj>b</obj><box>[1,\n2]</box>at<t>2</t>s</think>",
    "<think><obj>a</obj><box>[1,2,3,4]</box>at<t>5</t>s<obj>b</obj><box>[1,2,3,4]</box>at<t>nope</t>s"
    "<obj>c</obj><box>[1,2,3,4}</box>at<t>6</t>s<obj>d</obj><box>[5,6,7,8]</box><box>[1,2,3]</box>at<t>\x1f7　</t>s</think>",
    "<think><obj>a</obj><box>[1,2,3,4]</box>at<t>5</think><answer></t>s</answer>",  # match must end inside <think>
    "<think>see <box>[1, 2, 3, 4]</box> and <box>[1, 2\n, 3, 4]</box> <box>[5,6,7,8] x [9]</box><box>[[1,2],[3,4]]</box></think>"
    "<answer><box>[1,2] [3]</box> <box>[10, 20, 30, 40]</box></answer>",
    "<think><box>[true, false, null, NaN]</box><box>[Infinity, -Infinity, 1e400, -0]</box><box>[\"1\", \" 2 \", \"1_0\", \"nan\"]</box>"
    "<box>[\"a\",1,2,3]</box><box>[{},1,2,3]</box><box>[{\"k\":[1,{\"z\":null}]},1,2,3]</box><box>[1,2,3,4,]</box></think><answer><box>[1.5,2.5,3.5,4.5]</box></answer>",
    "<think>" + "<t>1</t>s" * 40 + "</think>",
    "<think>" + "x" * 5000 + "<t>77.25</t>s" + "y" * 3000 + "</think><answer>" + "z" * 700 + "</answer>",
]


@pytest.mark.parametrize("task", op.TASKS)
def test_scanner_edge_cases(task):
    texts = EDGE_TEXTS * 1
    bad, got, exp, ov = _compare(texts, [task] * len(texts), P=64, C=8, Bc=4, Tb=8)
    assert not bad, [(r, k, texts[r][:80]) for r, k in sorted(bad)[:5]]


def test_scanner_literal_search_alignment():
    """Tags at every offset relative to the 16-byte / 512-byte steps of the warp-wide search."""
    texts = []
    for pad in list(range(0, 40)) + [495, 496, 505, 511, 512, 513, 1023, 1024, 1030]:
        texts.append("p" * pad + "<think>" + "q" * (pad % 7) + "<t>%d.5</t>s" % pad + "</think>" + "r" * pad
                     + "<answer>From <t>1</t>s to <t>%d</t>s</answer>" % pad)
    bad, got, exp, ov = _compare(texts, ["temporal QA"] * len(texts))
    assert not bad
    assert (exp["n_times"] == 1).all() and ((exp["flags"] & op.RF_ANS_SEG) != 0).all()


def test_scanner_groups_share_task():
    texts, tasks = op.synth_batch(600, 17, tasks=("visual QA", "temporal QA"))
    tasks = [t for t in tasks[:150] for _ in range(4)]          # G = 4 rollouts per prompt
    bad, got, exp, ov = _compare(texts, tasks, G=4)
    assert not bad


# ------------------------------------------------------------------------------- float() and json.loads
def _ref_float(s, strip=True):
    try:
        return float(s.strip() if strip else s)
    except ValueError:
        return None


FLOAT_CASES = ["0", "12.5", ".5", "5.", ".", "", "1e5", "1E-5", "1e", "1e+", "-1", "+1", "--1", "inf", "-inf", "Infinity",
               "INFINITY", "nan", "-nan", "+NaN", "infinit", "1_0", "1__0", "_1", "1_", "1_.0", "1._0", "1e_5", "1e5_0",
               "1_0.0_1e1_0", "0x10", "1 2", " 12 ", " 12　", "\x1f12", "١٢.٥", "１２", "1٢", "²", "1e٢", "1.5.2",
               "1e400", "1e-400", "4.9e-324", "2.4703282292062327e-324", "2.4703282292062328e-324",
               "1.7976931348623157e308", "1.7976931348623158e308", "1.7976931348623159e308", "9007199254740993",
               "9007199254740992.5", "9007199254740993.0000000000000000000000001", "123456789012345678901234567890",
               "1" * 400, "0." + "0" * 400 + "1", "1" + "0" * 3000 + "e-3000", "0e99999999999999999999",
               "1e99999999999999999999", "1e-99999999999999999999", "0" * 500 + "." + "0" * 500 + "5e501"]


def test_python_float_known_cases():
    for s in FLOAT_CASES:
        for strip in (True, False):
            assert _same(scan_host.python_float(s, strip), _ref_float(s, strip)), (s[:60], strip)


def test_python_float_random_and_halfway():
    rng = random.Random(3)
    for _ in range(30000):
        k = rng.choice([1, 2, 3, 5, 8, 15, 17, 19, 20, 25, 40])
        s = "".join(rng.choice("0123456789") for _ in range(rng.randint(0, k))) + rng.choice(["", ".", "."]) + \
            "".join(rng.choice("0123456789") for _ in range(rng.randint(0, k)))
        if rng.random() < 0.3:
            s += rng.choice("eE") + rng.choice(["", "+", "-"]) + str(rng.randint(0, rng.choice([5, 30, 330])))
        assert _same(scan_host.python_float(s), _ref_float(s)), s
    from decimal import Decimal, getcontext
    from fractions import Fraction
    getcontext().prec = 1200
    for _ in range(1500):                                     # exact midpoints between adjacent doubles, +- one digit
        bits = rng.getrandbits(64) & 0x7FEFFFFFFFFFFFFF
        x, y = struct.unpack("d", struct.pack("Q", bits))[0], struct.unpack("d", struct.pack("Q", bits + 1))[0]
        if math.isinf(y):
            continue
        mid = (Fraction(x) + Fraction(y)) / 2
        s = format(Decimal(mid.numerator) / Decimal(mid.denominator), "e")
        m, e = s.split("e")
        for cand in (s, m + "1e" + e, m[:-1] + "e" + e if m[-1] != "." else s):
            assert _same(scan_host.python_float(cand), _ref_float(cand)), cand[:50]


def _ref_box(s):
    try:
        v = json.loads(s)
    except json.JSONDecodeError:
        return None
    a = op.box_ok(v) if len(v) == 4 else None
    try:
        numeric = np.array(v, dtype=float).ndim == 1
    except (ValueError, TypeError):
        numeric = False
    return len(v), numeric, (list(a) if a is not None else None)


def test_json_box_random():
    rng = random.Random(7)
    atoms = ["0", "1", "-1", "12.5", "1e5", "1E-3", "-0", "-0.0", "123456789012345678901234567890", "1e400", "-1e400",
             "1e-400", "true", "false", "null", "NaN", "Infinity", "-Infinity", '"a"', '"12"', '" 3.5 "', '""', '"\\n"',
             '"é"', '"1_0"', '"inf"', "{}", '{"a":1}', '{"a":[1,2],"b":{"c":null}}', "[]", "[1]", "[[1,2],[3]]", '{"a" 1}',
             "{1:2}", '{"a":1,}', "[1,]", "01", "1.", ".5", "+1", "- 1", "1e", "1e+", "tru", "nul", "nan", "inf", "0x1",
             "1 2", '"unterminated', "'a'", '"\t"', '"\x1f"', '"\\x"', '"\\u00e9"', '"\\u00g9"']
    ws = ["", " ", "  ", "\t", "\n", "\r\n", "\x0b", " "]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for _ in range(40000):
            n = rng.choice([0, 1, 2, 3, 4, 4, 4, 4, 5, 6])
            s = "[" + ",".join(rng.choice(ws[:6] if rng.random() < 0.9 else ws) + rng.choice(atoms) +
                               rng.choice(ws[:6] if rng.random() < 0.9 else ws) for _ in range(n)) + "]"
            u = rng.random()
            if u < 0.15 and len(s) > 2:
                i = rng.randrange(1, len(s))
                s = s[:i] + s[i + 1:]
            elif u < 0.3:
                i = rng.randrange(1, len(s) + 1)
                s = s[:i] + rng.choice(',[]{}": -e.0x\n') + s[i:]
            got, ref = scan_host.json_box(s), _ref_box(s)
            assert (got is None) == (ref is None), s
            if got is None:
                continue
            assert got[0] == ref[0], s
            escaped_numeric = ref[1] and any(isinstance(x, str) and "\\" in json.dumps(x) for x in json.loads(s))
            if not escaped_numeric:                           # documented deviation, see test below
                assert got[1] == ref[1], s
                if ref[2] is not None:
                    assert all(_same(a, b) for a, b in zip(got[2], ref[2])), s


def test_documented_deviations():
    """Inputs on which the reference itself raises or leaves the float domain."""
    # a JSON string with a backslash escape that unescapes to a number: numpy parses it, the scanner does not try
    assert scan_host.json_box('["\\u0031", 2, 3, 4]') == (4, False, [0.0, 2.0, 3.0, 4.0])
    # integers beyond float range: np.array raises OverflowError (uncaught in the reference); here +inf
    n, numeric, v = scan_host.json_box("[1" + "0" * 400 + ", 1, 2, 3]")
    assert (n, numeric) == (4, True) and math.isinf(v[0])
    # homogeneous nested lists become a 2-D array in the reference (then an ambiguous truth value): here the box scores 0
    assert scan_host.json_box("[[1],[2],[3],[4]]")[:2] == (4, False)


# ------------------------------------------------------------------------------- more than 32 boxes (round-1 advisor finding)
def many_box_texts():
    """Degenerate repetition loops: 40-70 boxes in one claim / one visual-QA think block, invalid ones (wrong
    length, strings, non-JSON) on both sides of the 32-bit validity mask."""
    def boxes(n, bad=(), drop=()):
        out = []
        for b in range(n):
            if b in drop:
                out.append("<box>[1,2,3,4}</box>")                 # not JSON: claims are dropped, think boxes skipped
            elif b in bad:
                out.append("<box>[%d,2,3]</box>" % b if b % 2 else "<box>[\"a\",%d,3,4]</box>" % b)
            else:
                out.append("<box>[%d,%d,%d,%d]</box>" % (10 + b, 20 + b, 200 + 3 * b, 220 + 2 * b))
        return "".join(out)
    ts = "<think><obj>dog</obj>%sat<t>5.0</t>s then <obj>cat</obj><box>[10,10,50,50]</box>at<t>12.5</t>s</think><answer>B</answer>"
    vq = "<think>%s</think><answer><box>[100,100,300,300]</box></answer>"
    return [
        (ts % boxes(40), "temporal-spatial free-form QA"),
        (ts % boxes(40, bad=(0, 5, 31, 32, 33, 39)), "temporal-spatial free-form QA"),
        (ts % boxes(70, bad=tuple(range(30, 70, 3))), "temporal-spatial free-form QA"),
        (ts % boxes(33, bad=(32,)), "temporal-spatial free-form QA"),
        (vq % boxes(40), "visual QA"),
        (vq % boxes(45, bad=(2, 31, 32, 40), drop=(1, 30, 33, 44)), "visual QA"),
        (vq % boxes(64, bad=tuple(range(0, 64, 2))), "visual QA"),
    ]


def test_scanner_more_than_32_boxes_marks_validity_in_the_slot():
    texts, tasks = zip(*many_box_texts())
    bad, got, exp, ov = _compare(list(texts), list(tasks))
    assert not bad, sorted(bad)
    assert exp["claim_nbox"].max() == 70 and exp["n_tboxes"].max() == 64
    # the invalid markers are exact bit patterns (mismatches() treats every NaN as equal)
    marker = np.uint64(0x7FF8B0B0DEADBEEF)
    used = op.used_mask(exp)
    for name in ("claim_box", "think_box"):
        g, e = got[name][..., 0].view(np.uint64), exp[name][..., 0].view(np.uint64)
        u = used[name][..., 0]
        assert np.array_equal((g == marker) & u, (e == marker) & u) and ((e == marker) & u).sum() > 3
        assert not (u[..., :32] & (e[..., :32] == marker)).any()                     # below 32 the mask rules
