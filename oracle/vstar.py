"""Oracle: V-STAR scorer numerics.

Test infrastructure (see oracle/__init__.py).  Restates eval/test/eval_vstar.py of the reference:
  :90-109   calculate_temporal_iou
  :112-133  compute_iou
  :135-146  calculate_bbox_iou
  :148-178  calculate_spatial_metrics (mIoU + AP@{0.1,0.3,0.5,0.7,0.9})
  :180-198  calculate_spatial_random
  :362-410  the aggregate statistics of print_stats (R1@IoU, means, AM / LGM, mAM / mLGM)
The LLM-judge VQA score (:43-73) is an INPUT here.  The per-item functions are pinned against
the reference's own code (oracle/ref_import.load_vstar_functions) in tests/golden/vstar_small.json;
the aggregate block is nested inside evaluate_json and cannot be executed: restated line for line.
"""
import math

import numpy as np

IOU_THRESHOLDS = [0.1, 0.3, 0.5, 0.7, 0.9]


def calculate_temporal_iou(gt_range, pred_range):
    """eval_vstar.py:90-109 (pred already literal_eval'ed when it was a string)."""
    if not pred_range:
        return 0.0
    if not isinstance(pred_range, (list, tuple)) or len(pred_range) != 2 or \
            not all(isinstance(x, (int, float)) for x in pred_range):
        return 0.0
    gt_start, gt_end = gt_range
    pred_start, pred_end = pred_range
    intersection = max(0, min(gt_end, pred_end) - max(gt_start, pred_start))      # :106
    union = max(gt_end, pred_end) - min(gt_start, pred_start)                      # :107
    return intersection / union if union > 0 else 0.0                              # :108


def compute_iou(gt_bbox, pred_bbox):
    """eval_vstar.py:112-133."""
    if not isinstance(pred_bbox, (list, tuple)) or len(pred_bbox) != 4:
        return 0.0
    gx0, gy0, gx1, gy1 = gt_bbox["xmin"], gt_bbox["ymin"], gt_bbox["xmax"], gt_bbox["ymax"]
    px0, py0, px1, py1 = pred_bbox
    x1 = max(gx0, px0); y1 = max(gy0, py0); x2 = min(gx1, px1); y2 = min(gy1, py1)
    intersection = max(0, x2 - x1) * max(0, y2 - y1)                               # :124
    gt_area = (gx1 - gx0) * (gy1 - gy0)
    pred_area = (px1 - px0) * (py1 - py0)
    union = gt_area + pred_area - intersection                                     # :129
    return intersection / union if union > 0 else 0.0


def calculate_bbox_iou(gt_bbox, pred_bboxes):
    """eval_vstar.py:135-146."""
    try:
        if not pred_bboxes:
            return 0.0
        if isinstance(pred_bboxes[0], (int, float)) and len(pred_bboxes) == 4:
            pred_bboxes = [pred_bboxes]
        return max([compute_iou(gt_bbox, pb) for pb in pred_bboxes])
    except Exception:
        return 0.0


def calculate_spatial_metrics(gt_bboxes, pred_bboxes):
    """eval_vstar.py:148-178 -> (aps[5], mIoU)."""
    if not pred_bboxes:
        return [0.0] * 5, 0.0
    ious = []
    for box in gt_bboxes:
        frame_id = str(box["timestamp"])
        if isinstance(pred_bboxes, dict) and frame_id in pred_bboxes:
            ious.append(calculate_bbox_iou(box, pred_bboxes[frame_id]))
        else:
            ious.append(0.0)
    mIoU = np.mean(ious) if ious else 0.0
    aps = [np.mean([1 if iou >= t else 0 for iou in ious]) if len(ious) > 0 else 0.0 for t in IOU_THRESHOLDS]
    return aps, mIoU


def item_scores(item):
    """[tIoU1, mIoU1, AP1 x5, tIoU2, mIoU2, AP2 x5] for one result item (eval_vstar.py:262-312)."""
    out = []
    for suffix in ("", "_2"):
        at = item.get("answer_temporal" + suffix)
        t = calculate_temporal_iou(item["timestamps"], at) if at else 0.0            # :262-265, :277-280
        sp = item.get("answer_spatial" + suffix)
        aps, miou = calculate_spatial_metrics(item["bboxes"], sp) if sp else ([0.0] * 5, 0.0)   # :293-296
        out += [t, miou] + list(aps)
    return [float(x) for x in out]


def aggregate(scores, vqa_scores):
    """print_stats (eval_vstar.py:362-410) on per-item score rows [I, 14] and judge scores [I]."""
    scores = np.asarray(scores, dtype=np.float64)
    vqa = np.asarray(vqa_scores)
    total = len(vqa)
    acc_vqa = float((vqa >= 2).sum()) / total                                        # :365
    res = dict(acc_vqa=acc_vqa)
    for c, off in ((1, 0), (2, 7)):
        tiou, miou = scores[:, off], scores[:, off + 1]
        res["r1_iou30_%d" % c] = float(np.mean(tiou >= 0.3))                         # :367-369
        res["r1_iou50_%d" % c] = float(np.mean(tiou >= 0.5))
        res["r1_iou70_%d" % c] = float(np.mean(tiou >= 0.7))
        res["mean_tiou_%d" % c] = float(np.mean(tiou))                               # :370
        res["mean_aps_%d" % c] = [float(np.mean(scores[:, off + 2 + k])) for k in range(5)]   # :377
        res["mean_miou_%d" % c] = float(np.mean(miou))                               # :378
        res["AM%d" % c] = (acc_vqa + res["mean_tiou_%d" % c] + res["mean_miou_%d" % c]) / 3   # :404
        res["LGM%d" % c] = -(math.log(1 - acc_vqa) + math.log(1 - res["mean_tiou_%d" % c])
                             + math.log(1 - res["mean_miou_%d" % c])) / 3            # :408
        res["vqa_temp_%d" % c] = float(((vqa >= 2) & (tiou >= 0.3)).sum()) / total   # :331, :384
        res["vqa_spat_%d" % c] = float(((vqa >= 2) & (miou >= 0.1)).sum()) / total   # :339
        res["temp_spat_%d" % c] = float(((tiou >= 0.3) & (miou >= 0.1)).sum()) / total   # :347
        res["vqa_temp_spat_%d" % c] = float(((vqa >= 2) & (tiou >= 0.3) & (miou >= 0.1)).sum()) / total   # :355
    res["mAM"] = (res["AM1"] + res["AM2"]) / 2                                       # :406
    res["mLGM"] = (res["LGM1"] + res["LGM2"]) / 2                                    # :410
    return res
