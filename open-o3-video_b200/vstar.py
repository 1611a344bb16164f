"""V-STAR scorer numerics on the GPU (K5), host side.

Offline counterpart of the reward numerics: the per-item temporal IoU,
spatial mIoU and AP@{0.1..0.9} of the reference's eval/test/eval_vstar.py:90-178 for both answer
chains in one launch, and the aggregate statistics of its print_stats (:362-410: R1@IoU, means,
AM / LGM, mAM / mLGM).  The LLM-judge VQA score (:43-73) is an input.  Items are the reference's
result-JSON dicts (`timestamps`, `bboxes`, `answer_temporal[_2]`, `answer_spatial[_2]`).
"""
import ast
import ctypes
import math
from typing import Sequence

import numpy as np
import torch

from . import _lib
from .rewards import to_device

COLUMNS = ("tIoU", "mIoU", "AP@0.1", "AP@0.3", "AP@0.5", "AP@0.7", "AP@0.9")


def _num(x):
    return isinstance(x, (int, float))


def _temporal(ans):
    """eval_vstar.py:92-104: falsy -> none; a string is literal_eval'ed; must be 2 numbers."""
    if not ans:
        return None
    if isinstance(ans, str):
        try:
            ans = ast.literal_eval(ans)
        except (ValueError, SyntaxError):
            return None
    if not isinstance(ans, (list, tuple)) or len(ans) != 2 or not all(_num(x) for x in ans):
        return None
    return [float(ans[0]), float(ans[1])]


def _frame_boxes(pred):
    """calculate_bbox_iou (:135-146): -> list of candidate boxes (possibly malformed), [] if none."""
    try:
        if not pred:
            return []
        if _num(pred[0]) and len(pred) == 4:
            return [pred]
        return list(pred)
    except Exception:
        return []


def _box_ok(b):
    return isinstance(b, (list, tuple)) and len(b) == 4 and all(_num(x) for x in b)


def pack_items(items: Sequence[dict]):
    I = len(items)
    F = max([len(it["bboxes"]) for it in items] + [1])
    frames = [[[_frame_boxes((it.get("answer_spatial" + sfx) or {}).get(str(b["timestamp"]))
                             if isinstance(it.get("answer_spatial" + sfx), dict) else None)
                for b in it["bboxes"]] for sfx in ("", "_2")] for it in items]
    Pb = max([len(fb) for it in frames for ch in it for fb in ch] + [1])
    if F > 64 or Pb > 32:
        raise ValueError("more than 64 annotated frames or 32 predicted boxes per frame")
    a = dict(t_valid=np.zeros((I, 2), np.int32), gt_seg=np.zeros((I, 2)), pred_seg=np.zeros((I, 2, 2)),
             sp_valid=np.zeros((I, 2), np.int32), n_frames=np.zeros(I, np.int32), gt_box=np.zeros((I, F, 4)),
             n_pb=np.zeros((I, 2, F), np.int32), pb_valid=np.zeros((I, 2, F), np.uint32), pb=np.zeros((I, 2, F, Pb, 4)))
    for i, it in enumerate(items):
        a["gt_seg"][i] = it["timestamps"]
        a["n_frames"][i] = len(it["bboxes"])
        for f, b in enumerate(it["bboxes"]):
            a["gt_box"][i, f] = [b["xmin"], b["ymin"], b["xmax"], b["ymax"]]
        for c, sfx in enumerate(("", "_2")):
            t = _temporal(it.get("answer_temporal" + sfx))
            if t is not None:
                a["t_valid"][i, c] = 1
                a["pred_seg"][i, c] = t
            a["sp_valid"][i, c] = 1 if it.get("answer_spatial" + sfx) else 0
            for f, boxes in enumerate(frames[i][c]):
                a["n_pb"][i, c, f] = len(boxes)
                for k, box in enumerate(boxes):
                    if _box_ok(box):
                        a["pb_valid"][i, c, f] |= np.uint32(1 << k)
                        a["pb"][i, c, f, k] = box
    return a, dict(I=I, F=F, Pb=Pb)


def score_items(items: Sequence[dict], device="cuda") -> torch.Tensor:
    """[I, 14] float64 on the device: chain 1 (tIoU, mIoU, AP x5), chain 2 (same)."""
    if len(items) == 0:
        return torch.empty(0, 14, dtype=torch.float64, device=device)
    arrays, dims = pack_items(items)
    dev = to_device(arrays, device)
    out = torch.empty(dims["I"], 14, dtype=torch.float64, device=device)
    soa = _lib.VstarSoA()
    soa.I, soa.F, soa.Pb = dims["I"], dims["F"], dims["Pb"]
    for name, _ in _lib.VstarSoA._fields_[3:]:
        setattr(soa, name, dev[name].data_ptr())
    with torch.cuda.device(out.device):
        _lib.call("o3v_vstar_scores", 1, _lib.load().o3v_vstar_scores, ctypes.byref(soa),
                  ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def evaluate(items: Sequence[dict], vqa_scores: Sequence[float]) -> dict:
    """The statistics eval_vstar.py prints for "Overall" (:362-410) from the GPU per-item scores."""
    sc = score_items(items).cpu().numpy()
    vqa = np.asarray(vqa_scores)
    total = len(vqa)
    acc = float((vqa >= 2).sum()) / total
    res = dict(acc_vqa=acc, per_item=sc)
    for c, off in ((1, 0), (2, 7)):
        tiou, miou = sc[:, off], sc[:, off + 1]
        res["r1_iou30_%d" % c] = float(np.mean(tiou >= 0.3))
        res["r1_iou50_%d" % c] = float(np.mean(tiou >= 0.5))
        res["r1_iou70_%d" % c] = float(np.mean(tiou >= 0.7))
        res["mean_tiou_%d" % c] = float(np.mean(tiou))
        res["mean_aps_%d" % c] = [float(np.mean(sc[:, off + 2 + k])) for k in range(5)]
        res["mean_miou_%d" % c] = float(np.mean(miou))
        res["AM%d" % c] = (acc + res["mean_tiou_%d" % c] + res["mean_miou_%d" % c]) / 3
        res["LGM%d" % c] = -(math.log(1 - acc) + math.log(1 - res["mean_tiou_%d" % c])
                             + math.log(1 - res["mean_miou_%d" % c])) / 3
        res["vqa_temp_%d" % c] = float(((vqa >= 2) & (tiou >= 0.3)).sum()) / total
        res["vqa_spat_%d" % c] = float(((vqa >= 2) & (miou >= 0.1)).sum()) / total
        res["temp_spat_%d" % c] = float(((tiou >= 0.3) & (miou >= 0.1)).sum()) / total
        res["vqa_temp_spat_%d" % c] = float(((vqa >= 2) & (tiou >= 0.3) & (miou >= 0.1)).sum()) / total
    res["mAM"] = (res["AM1"] + res["AM2"]) / 2
    res["mLGM"] = (res["LGM1"] + res["LGM2"]) / 2
    return res
