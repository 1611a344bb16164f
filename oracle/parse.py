"""Oracle: completion text -> parsed rollout -> rollout-side struct of arrays.

Test infrastructure (see oracle/__init__.py).  This is the text extraction of the reference's
  src/r1-v/src/open_r1/reward_func.py
restated with Python's own `re`, `json` and `float` -- the very engines the reference calls, so
regex backtracking, JSON acceptance and decimal rounding are the reference's by construction.
Every statement cites the reference line it repeats.  Pinned in tests/test_oracle_golden.py
against the live reference: `parse_temporal_spatial_reasoning_process` is called directly, the
other extractions are inline in the reward callables and are pinned end to end (oracle parse +
oracle numerics == reference reward on the same text).  Fixtures: tests/golden/parse_cases.json.

It is the checker for kernel K6 (csrc/parse.cu); the product never imports it.
"""
import ast
import json
import random
import re

import numpy as np

TASKS = ("visual QA", "temporal QA", "temporal QA (MCQ)",
         "temporal-spatial free-form QA", "General video QA MCQ", "General video QA Free-form")
RF_HAS_THINK, RF_HAS_ANSWER, RF_ANS_SEG, RF_ANS_BOX = 1, 2, 4, 8


def extract_answer(text):
    """reward_func.py:90-95 (and :188-193)."""
    pattern = r'<answer>\s*(.*?)\s*</answer>'
    match = re.search(pattern, text, re.DOTALL)
    if match:
        return match.group(1).strip()
    return ""


def parse_claims(think_content):
    """reward_func.py:308-335 -> [(timestamp, bboxes)] (object names are never used by a reward)."""
    pattern = r"<obj>(.*?)</obj>((?:<box>\[.*?\]</box>)+)at<t>(.*?)</t>s"           # :310
    parsed = []
    for match in re.finditer(pattern, think_content, re.DOTALL):                     # :314
        try:
            all_boxes_str = match.group(2)                                            # :317
            timestamp = float(match.group(3).strip())                                 # :318-319
            individual_box_strs = re.findall(r'\[.*?\]', all_boxes_str)               # :321
            bboxes = [json.loads(b_str) for b_str in individual_box_strs]             # :322
            parsed.append((timestamp, bboxes))
        except (json.JSONDecodeError, ValueError, IndexError):                        # :332
            continue
    return parsed


def parse_text(content, task):
    """What the five numeric reward callables extract from one completion.

    Gating by task mirrors which callable reads what: the answer segment only for the temporal
    tasks (:99-103), boxes of the answer / think only for visual QA (:196, :490), claims for the
    tasks that reach :535."""
    think_match = re.search(r"<think>(.*?)</think>", content, re.DOTALL)             # :394, :437, :481
    answer_match = re.search(r"<answer>(.*?)</answer>", content, re.DOTALL)          # :482
    r = dict(has_think=bool(think_match), has_answer=bool(answer_match), ans_seg=None, ans_box=None,
             think_times=[], think_boxes=[], claims=[])
    output_ans = extract_answer(content)                                              # :113, :207
    if task in ("temporal QA", "temporal QA (MCQ)"):
        pattern = r"<t>(\d+\.?\d*)</t>s to <t>(\d+\.?\d*)</t>s"                      # :119, :149
        match = re.search(pattern, output_ans)
        if match:
            r["ans_seg"] = [float(match.group(1)), float(match.group(2))]             # :123-126
    if think_match:
        think_content = think_match.group(1)
        try:
            r["think_times"] = [float(m) for m in re.findall(r'<t>([\d.]+)</t>s', think_content)]   # :405-412
        except ValueError:
            r["think_times"] = []                                                     # :413-415, :450-451
    if task == "visual QA":
        pattern = r"<box>(\[.*?\])</box>"                                             # :211, :492
        match_pred = re.search(pattern, output_ans)                                   # :220
        if match_pred:
            try:
                r["ans_box"] = json.loads(match_pred.group(1))                        # :222-223
            except Exception:                                                         # outer except :231 -> reward 0
                r["ans_box"] = None
        if think_match:
            for bbox in re.findall(pattern, think_match.group(1)):                    # :505
                try:
                    r["think_boxes"].append(json.loads(bbox))                         # :511
                except Exception:                                                     # :512-513
                    pass
    elif think_match:
        r["claims"] = parse_claims(think_match.group(1))                              # :535
    return r


INVALID_BOX = np.array([0x7FF8B0B0DEADBEEF], np.uint64).view(np.float64)[0]   # O3V_INVALID_BOX_BITS (include/o3v.h)


def box_ok(b):
    """calculate_iou's acceptance of a prediction (reward_func.py:361-367)."""
    if not (isinstance(b, list) and len(b) == 4):
        return None
    try:
        a = np.array(b, dtype=float)
    except (ValueError, TypeError, IndexError):
        return None
    return a if a.shape == (4,) else None


def pack(parsed, P, C, Bc, Tb):
    """[parse_text(...)] -> rollout-side arrays in the o3v_rewards_soa layout (zero filled)."""
    R = len(parsed)
    a = dict(flags=np.zeros(R, np.int32), ans_seg=np.zeros((R, 2)), ans_box=np.zeros((R, 4)),
             n_times=np.zeros(R, np.int32), think_times=np.zeros((R, P)), n_claims=np.zeros(R, np.int32),
             claim_t=np.zeros((R, C)), claim_nbox=np.zeros((R, C), np.int32), claim_valid=np.zeros((R, C), np.uint32),
             claim_box=np.zeros((R, C, Bc, 4)), n_tboxes=np.zeros(R, np.int32), tbox_valid=np.zeros(R, np.uint32),
             think_box=np.zeros((R, Tb, 4)))
    for i, r in enumerate(parsed):
        f = (RF_HAS_THINK if r["has_think"] else 0) | (RF_HAS_ANSWER if r["has_answer"] else 0)
        if r["ans_seg"] is not None:
            f |= RF_ANS_SEG
            a["ans_seg"][i] = r["ans_seg"]
        if r["ans_box"] is not None and box_ok(r["ans_box"]) is not None:
            f |= RF_ANS_BOX
            a["ans_box"][i] = box_ok(r["ans_box"])
        a["flags"][i] = f
        a["n_times"][i] = len(r["think_times"])
        a["think_times"][i, :min(P, len(r["think_times"]))] = r["think_times"][:P]
        a["n_claims"][i] = len(r["claims"])
        for c, (t, boxes) in enumerate(r["claims"][:C]):
            a["claim_t"][i, c] = t
            a["claim_nbox"][i, c] = len(boxes)
            for b, box in enumerate(boxes):
                v = box_ok(box)
                if v is not None:
                    if b < 32:
                        a["claim_valid"][i, c] |= np.uint32(1 << b)
                    if b < Bc:
                        a["claim_box"][i, c, b] = v
                elif 32 <= b < Bc:                      # include/o3v.h: beyond the mask, validity lives in the slot
                    a["claim_box"][i, c, b, 0] = INVALID_BOX
        a["n_tboxes"][i] = len(r["think_boxes"])
        for b, box in enumerate(r["think_boxes"]):
            v = box_ok(box)
            if v is not None:
                if b < 32:
                    a["tbox_valid"][i] |= np.uint32(1 << b)
                if b < Tb:
                    a["think_box"][i, b] = v
            elif 32 <= b < Tb:
                a["think_box"][i, b, 0] = INVALID_BOX
    return a


def used_mask(a):
    """name -> boolean mask of the entries a parser must define (rows up to the counts)."""
    R = a["flags"].shape[0]
    P, C = a["think_times"].shape[1], a["claim_t"].shape[1]
    Bc, Tb = a["claim_box"].shape[2], a["think_box"].shape[1]
    m = {k: np.ones(a[k].shape, bool) for k in ("flags", "n_times", "n_claims", "n_tboxes", "tbox_valid")}
    m["ans_seg"] = np.repeat(((a["flags"] & RF_ANS_SEG) != 0)[:, None], 2, 1)
    m["ans_box"] = np.repeat(((a["flags"] & RF_ANS_BOX) != 0)[:, None], 4, 1)
    m["think_times"] = np.arange(P)[None, :] < a["n_times"][:, None]
    cm = np.arange(C)[None, :] < a["n_claims"][:, None]
    m["claim_t"] = cm
    m["claim_nbox"] = cm
    m["claim_valid"] = cm
    def box_mask(valid, boxes, exists):
        """valid [..] uint32 masks, boxes [.., B, 4], exists [.., B] (slot < count) -> [.., B, 4] entries to compare:
        all four numbers of a valid box; beyond bit 31 the first number always (value or the invalid marker) and the
        rest when the slot is not marked invalid."""
        B = boxes.shape[-2]
        idx = np.arange(B)
        low = ((valid[..., None] >> np.minimum(idx, 31).astype(np.uint32)) & 1) != 0
        high_valid = boxes[..., 0].view(np.uint64) != np.uint64(0x7FF8B0B0DEADBEEF)
        is_valid = np.where(idx < 32, low, high_valid) & exists
        m4 = np.repeat(is_valid[..., None], 4, -1)
        m4[..., 0] |= exists & (idx >= 32)
        return m4
    exists_c = cm[:, :, None] & (np.arange(Bc)[None, None, :] < a["claim_nbox"][:, :, None])
    m["claim_box"] = box_mask(a["claim_valid"], a["claim_box"], exists_c)
    exists_t = np.arange(Tb)[None, :] < a["n_tboxes"][:, None]
    m["think_box"] = box_mask(a["tbox_valid"], a["think_box"], exists_t)
    return m


# --------------------------------------------------------------------------------------
# Seeded text generator: well-formed completions in the reference's output format, plus
# every malformation the parsers have a branch for.
# --------------------------------------------------------------------------------------
_UNI_DIGITS = ["٠١٢٣٤٥٦٧٨٩", "०१२३४५६७८९", "０１２３４５６７８９", "𝟎𝟏𝟐𝟑𝟒𝟓𝟔𝟕𝟖𝟗"]
_SPACES = [" ", "\t", "\n", " ", " ", "　", "\x1f", "\x0b", "", " "]


def _number(rng, wild):
    kind = rng.random()
    if kind < 0.55 or not wild:
        s = "%d" % rng.randint(0, 200)
        if rng.random() < 0.7:
            s += "." + "".join(rng.choice("0123456789") for _ in range(rng.randint(0, 3)))
        return s
    if kind < 0.65:
        return "".join(rng.choice("0123456789") for _ in range(rng.randint(15, 60))) + rng.choice(["", ".", ".5", ".%s" % ("3" * 30)])
    if kind < 0.72:
        tbl = rng.choice(_UNI_DIGITS)
        return "".join(tbl[int(c)] if c.isdigit() else c for c in "%.2f" % (rng.random() * 100))
    if kind < 0.80:
        return rng.choice(["1.2.3", ".", "..", "5.", ".5", "1..", "007", "0.0", "00.50"])
    if kind < 0.86:
        return rng.choice(["1e2", "1E-2", "-3.5", "+4", "inf", "nan", "-inf", "Infinity", "1_0", "1__0", "_1", "0x10",
                           "1e", "١٢", "1e400", "4.9e-324", "2.4703282292062327e-324", "9007199254740993",
                           "179769313486231580793728971405303415079934132710037826936173778980444968292764750946649017977"
                           "587207096330286416692887910946555547851940402630657488671505820681908902000708383676273854845"
                           "817711531764475730270069855571366959622842914819860834936475292719074168444365510704342711559"
                           "699508093042880177904174497792"])
    if kind < 0.93:
        return rng.choice(_SPACES) * rng.randint(1, 2) + "%.1f" % (rng.random() * 50) + rng.choice(_SPACES) * rng.randint(0, 2)
    return rng.choice(["", "abc", "12 s", "1 2", "١٢.٥", "１２", "12\x00"])


def _box_payload(rng, wild):
    def coord():
        k = rng.random()
        if k < 0.6 or not wild:
            return "%d" % rng.randint(0, 640)
        if k < 0.75:
            return "%.3f" % (rng.random() * 640)
        if k < 0.8:
            return rng.choice(["1e2", "1.5E+1", "-12", "-0", "-0.0", "0.5e-3", "1E400", "12345678901234567890123"])
        if k < 0.86:
            return rng.choice(["true", "false", "null", "NaN", "Infinity", "-Infinity"])
        if k < 0.92:
            return rng.choice(['"12"', '" 7.5 "', '"a"', '""', '"1_0"', '"nan"', '"x\\ny"', '"\x1f3"', '" 4"'])
        return rng.choice(["{}", '{"a": 1}', '{"a": {"b": [1, 2]}}', "[1, 2]", "[]", "01", "1.", ".5", "+1", "1e", "tru", "-", "0x1"])
    n = 4 if (rng.random() < 0.8 or not wild) else rng.choice([0, 1, 3, 5, 8])
    sep = rng.choice([", ", ",", " , ", ",\t"]) if wild else ", "
    body = sep.join(coord() for _ in range(n))
    if wild and rng.random() < 0.08:
        body += rng.choice([",", " x", "]]", "[", "\n", " \n ,1", '"', "}"])
    if wild and rng.random() < 0.05:
        body = rng.choice([" ", "\n", "\r\n"]) + body + rng.choice([" ", "\t"])
    return "[" + body + "]"


def _junk(rng, wild):
    words = ["the", "person", "walks", "to", "door", "then", "at", "<", ">", "[", "]", "\n", "s", "</t>", "<t>", "<box>",
             "</box>", "<obj>", "</obj>", "at<t>", "</t>s", "]</box>", "日本語", "é", "</think>", "<think>", "<answer>",
             "</answer>", "<box>[", "]</box>at<t>", "</obj><box>["]
    n = rng.randint(0, 6)
    pool = words if (wild and rng.random() < 0.3) else words[:6]
    return " ".join(rng.choice(pool) for _ in range(n))


def synth_completion(rng, task, wild=True):
    """One completion in the reference's `<think>..</think><answer>..</answer>` format
    (the system prompt's format, grpo_trainer / data_loader), optionally malformed."""
    parts = []
    for _ in range(rng.randint(0, 6)):
        k = rng.random()
        if k < 0.45:
            nb = 1 if rng.random() < 0.7 else rng.randint(2, 4)
            boxes = "".join("<box>%s</box>" % _box_payload(rng, wild) for _ in range(nb))
            if wild and rng.random() < 0.05:
                boxes = boxes.replace("</box><box>", "</box> <box>", 1)
            parts.append("<obj>%s</obj>%sat<t>%s</t>s" % (rng.choice(["man", "dog", "red car", ""]), boxes, _number(rng, wild)))
        elif k < 0.75:
            parts.append("around <t>%s</t>s" % _number(rng, wild))
        elif k < 0.9:
            parts.append("see <box>%s</box>" % _box_payload(rng, wild))
        else:
            parts.append(_junk(rng, wild))
    think = (" " + _junk(rng, wild) + " ").join(parts)
    if task in ("temporal QA", "temporal QA (MCQ)"):
        ans = "From <t>%s</t>s to <t>%s</t>s" % (_number(rng, wild), _number(rng, wild))
        if wild and rng.random() < 0.1:
            ans = ans.replace(" to ", rng.choice(["  to ", " to", "-", " to <t>1</t>s to "]))
    elif task == "visual QA":
        ans = "It is at <box>%s</box>" % _box_payload(rng, wild)
        if wild and rng.random() < 0.1:
            ans = "first <box>[1, 2\n, 3]</box> then " + ans
    else:
        ans = rng.choice(["B", "A person opens the door.", ""])
    pad = lambda: rng.choice(["", " ", "\n", " "]) if wild else ""
    text = "<think>" + think + "</think>" + pad() + "<answer>" + pad() + ans + pad() + "</answer>"
    if wild:
        k = rng.random()
        if k < 0.04:
            text = text.replace("</think>", "", 1)
        elif k < 0.08:
            text = text.replace("<think>", "", 1)
        elif k < 0.12:
            text = text.replace("</answer>", "", 1)
        elif k < 0.16:
            text = text.replace("<answer>", "", 1)
        elif k < 0.20:
            text = text + text
        elif k < 0.24:
            text = "<answer>early</answer>" + text
        elif k < 0.27:
            text = text.replace("<think>", "<think><think>", 1).replace("</think>", "</think></think>", 1)
        elif k < 0.30:
            text = ""
        elif k < 0.33:
            text = _junk(rng, True) + text + _junk(rng, True)
    return text


def synth_batch(n, seed, tasks=None, wild=True):
    """-> (texts, task names), one task per text."""
    rng = random.Random(seed)
    tasks = tasks or TASKS
    out, tk = [], []
    for i in range(n):
        task = tasks[i % len(tasks)]
        out.append(synth_completion(rng, task, wild))
        tk.append(task)
    return out, tk


_MUT_ALPHABET = list("<>/[]\n tsobjxhinkawer.,0123456789") + [
    "<t>", "</t>s", "<box>[", "]</box>", "<obj>", "</obj>", "at<t>", "</obj><box>[", "]</box>at<t>", "<think>", "</think>",
    "<answer>", "</answer>", " to ", "\u2003", "\u0663"]


def mutate_batch(n, seed, tasks=None):
    """synth_batch followed by random character-level edits (insert a tag fragment, delete, duplicate, move or
    reverse a span): overlapping, nested and truncated tags that the token-level generator never emits."""
    rng = random.Random(seed)
    base, tk = synth_batch(n, seed, tasks)
    out = []
    for t in base:
        for _ in range(rng.choice([0, 1, 1, 2, 3, 5, 8])):
            if not t:
                break
            i = rng.randrange(len(t) + 1)
            r = rng.random()
            if r < 0.4:
                t = t[:i] + rng.choice(_MUT_ALPHABET) + t[i:]
            elif r < 0.7:
                t = t[:i] + t[min(len(t), i + rng.randint(1, 6)):]
            elif r < 0.85:
                j = min(len(t), i + rng.randint(1, 20))
                t = t[:i] + t[i:j] * 2 + t[j:]
            else:
                j = rng.randrange(len(t) + 1)
                a, b = min(i, j), max(i, j)
                t = t[:a] + t[b:] + t[a:b] if rng.random() < 0.5 else t[:a] + t[a:b][::-1] + t[b:]
        out.append(t)
    return out, tk


def encode(texts):
    """list[str] -> (uint8 buffer padded to 16 bytes, int64 offsets [R+1])."""
    blobs = [t.encode("utf-8", "surrogatepass") for t in texts]
    offsets = np.zeros(len(blobs) + 1, np.int64)
    np.cumsum([len(b) for b in blobs], out=offsets[1:])
    total = int(offsets[-1])
    buf = np.zeros((total + 15) // 16 * 16 + 16, np.uint8)
    buf[:total] = np.frombuffer(b"".join(blobs), np.uint8)
    return buf, offsets


# --------------------------------------------------------------------------------------
# End to end: (completion text, reward kwargs) -> the five rewards, and seeded cases.
# --------------------------------------------------------------------------------------
def parse_gt(task, answer):
    """Ground-truth side of kwargs['answer'] as the reward callables read it."""
    gt_seg, gt_vbox = [0.0, 0.0], None
    if task in ("temporal QA", "temporal QA (MCQ)"):
        gt_ans = answer
        if task == "temporal QA (MCQ)":
            gt_ans = gt_ans.split("\n")[1]                                            # :146, :408
        gt_seg = list(ast.literal_eval(gt_ans))                                       # :118, :147, :409
    if task == "visual QA":
        match_gt = re.search(r"<box>(\[.*?\])</box>", '<answer>%s</answer>' % answer)  # :88, :186, :213, :493
        if match_gt:
            try:
                gt_vbox = json.loads(match_gt.group(1))                               # :216, :498
            except Exception:
                gt_vbox = None
    return gt_seg, gt_vbox


def rollout_from_text(text, kw):
    """(completion, one rollout's kwargs) -> the parsed-rollout dict of oracle/rewards.py."""
    r = parse_text(text, kw["task"])
    r["task"] = kw["task"]
    r["gt_seg"], r["gt_vbox"] = parse_gt(kw["task"], kw["answer"])
    for k in ("key_frames", "key_items", "image_size", "image_size_refine", "step_percent"):
        r[k] = kw[k]
    return r


def text_cases(n, seed):
    """n (text, kwargs) pairs: ground truth from oracle/synth.rollouts; completions half rendered from
    the structured rollout (so the rewards are non-trivial), half from the wild generator."""
    from . import rewards as orw, synth
    ro = synth.rollouts(n_prompts=n, G=1, P=6, K=5, O=3, Gb=2, Bc=2, seed=seed)
    rng = random.Random(seed + 7)
    cases = []
    for r in ro:
        text, kw = orw.render(r)
        u = rng.random()
        if u < 0.5:
            text = synth_completion(rng, r["task"], wild=True)
        elif u < 0.6:
            text = synth_completion(rng, r["task"], wild=False)
        cases.append((text, kw))
    return cases


def reference_rewards_from_text(reward_func_module, cases):
    """Run the LIVE reference reward callables on (text, kwargs) cases, one rollout per call."""
    import contextlib
    import io
    from .rewards import REWARD_NAMES
    out = np.zeros((len(cases), 5), dtype=np.float64)
    fns = [getattr(reward_func_module, n) for n in REWARD_NAMES]
    for i, (text, kw) in enumerate(cases):
        completions = [[{"role": "assistant", "content": text}]]
        kwargs = {k: [v] for k, v in kw.items()}
        with contextlib.redirect_stdout(io.StringIO()):
            for j, fn in enumerate(fns):
                out[i, j] = fn(prompts=None, completions=completions, **kwargs)[0]
    return out
