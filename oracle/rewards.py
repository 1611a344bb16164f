"""Oracle: numeric cores of the spatio-temporal rewards, on PARSED rollouts.

Test infrastructure (see oracle/__init__.py).  Pure-Python / numpy float64
restatement of the arithmetic in
  src/r1-v/src/open_r1/reward_func.py
    :86-181   ans_tiou_reward            (temporal IoU, :128-143 dup :156-171)
    :184-236  ans_viou_reward            (:210-226)
    :337-354  convert_coord_format, convert_coord_format_gqa
    :356-386  calculate_iou
    :388-426  thk_temporal_segment_reward (:417-421)
    :429-472  thk_temporal_point_reward   (adaptive temporal proximity, :453-467)
    :475-605  thk_spatial_reward          (temporal gating + bbox IoU, :542-603;
                                           visual-QA branch :490-525)
The regex / json / ast text extraction is out of scope: a rollout
here is the structure the reference has AFTER parsing, see `Rollout` below.
Pinned against the reference run on rendered text in
tests/test_oracle_golden.py (fixtures: tests/golden/rewards_*.json).

A rollout is a dict:
  task            str  (kwargs['task'][0])
  has_think       bool   re.search(r"<think>(.*?)</think>") matched
  has_answer      bool   re.search(r"<answer>(.*?)</answer>") matched
  ans_seg         [s, e] or None : floats of the answer's "<t>s</t>s to <t>e</t>s" match
  ans_box         list or None   : json of the first <box> in the answer
  think_times     [float]        : every <t>..</t>s inside <think>
  think_boxes     [list]         : every <box>[..]</box> inside <think> (visual QA branch)
  claims          [(t, [box, ...])] : parse_temporal_spatial_reasoning_process output
  gt_seg          [s, e]         : ast.literal_eval(answer) for the temporal tasks
  gt_vbox         list or None   : json of the <box> in the GT answer (visual QA)
  key_frames      [{"idx": int, "time": float}]
  key_items       {str(idx): {obj: [[x0,y0,x1,y1] normalised, ...]}}
  image_size, image_size_refine : (W, H)
  step_percent    float
"""
import math

import numpy as np

TASKS = ("visual QA", "temporal QA", "temporal QA (MCQ)",
         "temporal-spatial free-form QA", "General video QA MCQ", "General video QA Free-form")
REWARD_NAMES = ("ans_tiou_reward", "ans_viou_reward", "thk_temporal_segment_reward",
                "thk_temporal_point_reward", "thk_spatial_reward")


def convert_coord_format(bbox, image_size):
    """reward_func.py:337-346."""
    nx_min, ny_min, nx_max, ny_max = bbox
    width, height = image_size
    return [nx_min * width, ny_min * height, nx_max * width, ny_max * height]


def convert_coord_format_gqa(bbox, image_size, image_size_refine):
    """reward_func.py:349-354 (the reference mutates in place; we return a new list)."""
    return [bbox[0] * image_size_refine[0] / image_size[0],
            bbox[1] * image_size_refine[1] / image_size[1],
            bbox[2] * image_size_refine[0] / image_size[0],
            bbox[3] * image_size_refine[1] / image_size[1]]


def calculate_iou(boxA, boxB):
    """reward_func.py:356-386.  boxA: GT, boxB: pred."""
    try:
        if not (isinstance(boxB, list) and len(boxB) == 4):                # :361
            return 0.0
        a = np.array(boxA, dtype=float)                                    # :364
        b = np.array(boxB, dtype=float)                                    # :365
    except (ValueError, TypeError, IndexError):
        return 0.0
    xA = max(a[0], b[0]); yA = max(a[1], b[1])                             # :370-371
    xB = min(a[2], b[2]); yB = min(a[3], b[3])                             # :372-373
    inter_area = max(0, xB - xA) * max(0, yB - yA)                         # :376
    boxA_area = (a[2] - a[0]) * (a[3] - a[1])                              # :378
    boxB_area = (b[2] - b[0]) * (b[3] - b[1])                              # :379
    union_area = boxA_area + boxB_area - inter_area                        # :382
    return inter_area / union_area if union_area > 0 else 0.0              # :385


def temporal_iou(pred, gt):
    """reward_func.py:128-143.  pred = [s, e] already float-parsed, or None."""
    if pred is None:
        return 0.0
    start_time, end_time = pred
    if end_time < start_time:                                              # :128
        return 0.0
    start1, end1 = start_time, end_time
    start2, end2 = gt
    intersection_start = max(start1, start2)                               # :138
    intersection_end = min(end1, end2)                                     # :139
    intersection_length = max(0, intersection_end - intersection_start)    # :140
    union_length = max(end1, end2) - min(start1, start2)                   # :141
    return intersection_length / union_length if union_length != 0 else 0  # :142


def ans_tiou(r):
    """reward_func.py:86-181 after parsing."""
    if r["task"] in ("temporal QA", "temporal QA (MCQ)"):                  # :99-103
        return float(temporal_iou(r["ans_seg"], r["gt_seg"]))
    return 0.0                                                             # :173


def ans_viou(r):
    """reward_func.py:184-236 after parsing."""
    if r["task"] != "visual QA":                                           # :196
        return 0.0
    if r["ans_box"] is None or r["gt_vbox"] is None:                       # :212-224
        return 0.0
    gt = convert_coord_format_gqa(r["gt_vbox"], r["image_size"], r["image_size_refine"])  # :225
    return float(calculate_iou(gt, r["ans_box"]))                          # :226


def thk_temporal_segment(r):
    """reward_func.py:388-426 after parsing."""
    t = r["task"]
    if (not r["has_think"]) or t == "visual QA" or t == "temporal-spatial free-form QA" \
            or "General video QA" in t:                                    # :396
        return 0.0
    times = r["think_times"]
    reward = 0.0
    if len(times) > 0:                                                     # :416
        for pred_time in times:
            if r["gt_seg"][0] <= pred_time <= r["gt_seg"][1]:              # :418
                reward += 1.0
        reward = reward / len(times)                                       # :420
    return reward


def thk_temporal_point(r):
    """reward_func.py:429-472 after parsing (adaptive temporal proximity)."""
    t = r["task"]
    if (not r["has_think"]) or t in ("visual QA", "temporal QA", "temporal QA (MCQ)") \
            or "General video QA" in t:                                    # :439
        return 0.0
    step_percent = r["step_percent"]
    pred_times = r["think_times"]
    if len(pred_times) == 0:                                               # :452, :469
        return 0.0
    gt_times = [f["time"] for f in r["key_frames"]]                        # :453
    total = 0.0
    for time in pred_times:
        time_diff = min([abs(time - g) for g in gt_times])                 # :457
        if step_percent < 3 / 4:                                           # :459
            sigma = 4 * (1 - step_percent)
        else:
            sigma = 1
        total += np.exp(-(time_diff ** 2) / (2 * sigma ** 2))              # :463
    return float(total / len(pred_times))                                  # :467


def thk_spatial(r):
    """reward_func.py:475-605 after parsing (temporal gating + bbox IoU)."""
    if (not r["has_think"]) or (not r["has_answer"]):                      # :484
        return 0.0
    t = r["task"]
    if t == "visual QA":                                                   # :490
        boxes = r["think_boxes"]
        if len(boxes) > 0 and r["gt_vbox"] is not None:                    # :513
            max_iou = 0.0
            gt = convert_coord_format_gqa(r["gt_vbox"], r["image_size"], r["image_size_refine"])
            for b in boxes:
                max_iou = max(max_iou, calculate_iou(gt, b))               # :517-518
            return float(max_iou)
        return 0.0
    if t == "temporal QA" or t == "temporal QA (MCQ)" or "General video QA" in t:   # :528
        return 0.0
    claims = r["claims"]
    if not claims:                                                         # :537
        return 0.0
    gt_items = r["key_items"]
    gt_times = [f["time"] for f in r["key_frames"]]
    total = 0.0
    for pred_time, bboxes in claims:
        closest_time = -1
        min_time_diff = float("inf")
        threshold = 1.0
        for ii in range(len(gt_times)):
            if gt_times[ii] - pred_time < threshold:                       # :556 (signed, one-sided)
                time_diff = abs(gt_times[ii] - pred_time)
                if time_diff < min_time_diff:                              # :558 (strict: first wins)
                    min_time_diff = time_diff
                    closest_time = gt_times[ii]
        if closest_time == -1:                                             # :561
            continue
        key_frame = None
        for f in r["key_frames"]:
            if f["time"] == closest_time:                                  # :567
                key_frame = f
                break
        if bboxes is not None and isinstance(bboxes, list) and key_frame is not None:
            objects = gt_items[str(key_frame["idx"])]                      # :572
            max_iou = 0.0
            for obj in objects.keys():
                claim_boxes = bboxes
                gt_boxes = objects[obj]
                try:
                    is_claim_originally_multiple = isinstance(claim_boxes[0], list)   # :579
                except Exception:                                          # :580-582 (no boxes survived findall)
                    continue
                if not is_claim_originally_multiple:                       # :584-585
                    claim_boxes = [claim_boxes]
                lst = []
                for gt_box in gt_boxes:
                    g = convert_coord_format(gt_box, r["image_size"])      # :591
                    ious = [calculate_iou(g, c) for c in claim_boxes]      # :592
                    lst.append(max(ious) if ious else 0.0)                 # :593
                if lst:
                    iou = sum(lst) / len(lst)                              # :597
                    if iou > max_iou:
                        max_iou = iou
            total += max_iou                                               # :601
    return float(total / len(claims))                                      # :603


def rewards_for_rollout(r):
    """[ans_tiou, ans_viou, thk_temporal_segment, thk_temporal_point, thk_spatial] (float64)."""
    return [ans_tiou(r), ans_viou(r), thk_temporal_segment(r), thk_temporal_point(r), thk_spatial(r)]


def rewards_for_rollouts(rollouts):
    return np.array([rewards_for_rollout(r) for r in rollouts], dtype=np.float64).reshape(-1, 5)


# --------------------------------------------------------------------------------------
# Rendering a parsed rollout back to (completion text, kwargs) for the live reference.
# --------------------------------------------------------------------------------------
def _num(x):
    """repr() round-trips a Python float exactly; require the plain-decimal form the
    reference's regexes accept (`[\\d.]+`, `\\d+\\.?\\d*`)."""
    s = repr(float(x)) if not isinstance(x, int) else repr(x)
    assert "e" not in s and "-" not in s and "n" not in s, s
    return s


def _box(b):
    return "[" + ", ".join(repr(v) for v in b) + "]"


def render(r):
    """-> (completion, kwargs_for_one_rollout) such that the reference's own parsers
    recover exactly the structure in `r`."""
    parts = []
    n_claim_times = 0
    for t, boxes in r["claims"]:
        parts.append("<obj>thing</obj>" + "".join("<box>%s</box>" % _box(b) for b in boxes)
                     + "at<t>%s</t>s" % _num(t))
        n_claim_times += 1
    if r["task"] == "visual QA":
        for b in r["think_boxes"]:
            parts.append("see <box>%s</box>" % _box(b))
    for t in r["think_times"][n_claim_times:]:
        parts.append("around <t>%s</t>s" % _num(t))
    think = " then ".join(parts)
    if r["task"] in ("temporal QA", "temporal QA (MCQ)") and r["ans_seg"] is not None:
        ans = "From <t>%s</t>s to <t>%s</t>s" % (_num(r["ans_seg"][0]), _num(r["ans_seg"][1]))
    elif r["task"] == "visual QA" and r["ans_box"] is not None:
        ans = "It is at <box>%s</box>" % _box(r["ans_box"])
    else:
        ans = "B"
    text = ""
    if r["has_think"]:
        text += "<think>" + think + "</think>"
    if r["has_answer"]:
        text += "<answer>" + ans + "</answer>"
    if r["task"] == "temporal QA":
        answer = _box(r["gt_seg"])
    elif r["task"] == "temporal QA (MCQ)":
        answer = "A\n" + _box(r["gt_seg"])
    elif r["task"] == "visual QA":
        answer = "object <box>%s</box>" % _box(r["gt_vbox"]) if r["gt_vbox"] is not None else "object"
    else:
        answer = "B"
    kwargs = dict(task=r["task"], answer=answer, key_frames=r["key_frames"], key_items=r["key_items"],
                  image_size=tuple(r["image_size"]), image_size_refine=tuple(r["image_size_refine"]),
                  step_percent=r["step_percent"])
    return text, kwargs


def reference_rewards(reward_func_module, rollouts):
    """Run the LIVE reference reward functions on rendered rollouts, one rollout per call
    (the reference reads kwargs['task'][0] / ['step_percent'][0] for the whole batch)."""
    import contextlib, io
    out = np.zeros((len(rollouts), 5), dtype=np.float64)
    fns = [getattr(reward_func_module, n) for n in REWARD_NAMES]
    for i, r in enumerate(rollouts):
        text, kw = render(r)
        completions = [[{"role": "assistant", "content": text}]]
        kwargs = {k: [v] for k, v in kw.items()}
        with contextlib.redirect_stdout(io.StringIO()):
            for j, fn in enumerate(fns):
                out[i, j] = fn(prompts=None, completions=completions, **kwargs)[0]
    return out
