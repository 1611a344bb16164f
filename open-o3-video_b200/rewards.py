"""Spatio-temporal rewards: completion text scanned on the GPU (K6), numerics on the GPU (K4).

Drop-in for the numeric reward callables of the reference
(src/r1-v/src/open_r1/reward_func.py): same names, same `f(completions, **kwargs) ->
list[float]` signature (grpo_trainer.py:655), same task gating.  The completions are shipped to
the device as UTF-8 bytes and `o3v_parse_completions` (K6) extracts what the reference's regex /
json / float() calls extract, bit for bit; the ground-truth side of the kwargs (a few numbers per
prompt, already structured) is packed on the host exactly as the reference reads it (cited per
function); temporal IoU, in-segment ratio, adaptive temporal proximity, temporal gating + bbox
IoU and the visual-QA IoUs then run in one more launch over the struct-of-arrays batch
(include/o3v.h `o3v_rewards_soa`, K4).  `parse_rollout` / `pack_rollouts` keep the host-side
route for callers that already hold parsed rollouts.  `ans_acc_reward` and `format_reward` are
pure string rewards and are not part of this path.
"""
import ast
import ctypes
import json
import re
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib

TASK_IDS = {"visual QA": 0, "temporal QA": 1, "temporal QA (MCQ)": 2,
            "temporal-spatial free-form QA": 3, "General video QA MCQ": 4, "General video QA Free-form": 5}
REWARD_NAMES = ("ans_tiou_reward", "ans_viou_reward", "thk_temporal_segment_reward",
                "thk_temporal_point_reward", "thk_spatial_reward")
RF_HAS_THINK, RF_HAS_ANSWER, RF_ANS_SEG, RF_ANS_BOX, GF_VBOX = 1, 2, 4, 8, 1
# O3V_INVALID_BOX_BITS (include/o3v.h): marks, in its own slot, a box with index >= 32 that is not a list of 4 numbers
INVALID_BOX = np.array([0x7FF8B0B0DEADBEEF], np.uint64).view(np.float64)[0]


# ----------------------------------------------------------------------------- parsing
def _box_ok(b) -> bool:
    """What calculate_iou accepts as a prediction (reward_func.py:361-368): a list of 4 numbers."""
    if not (isinstance(b, list) and len(b) == 4):
        return False
    try:
        np.array(b, dtype=float)
        return True
    except (ValueError, TypeError):
        return False


def parse_claims(think_content: str):
    """reward_func.py:308-335 (parse_temporal_spatial_reasoning_process) -> [(t, [box, ...])]."""
    pattern = r"<obj>(.*?)</obj>((?:<box>\[.*?\]</box>)+)at<t>(.*?)</t>s"
    claims = []
    for match in re.finditer(pattern, think_content, re.DOTALL):
        try:
            timestamp = float(match.group(3).strip())
            boxes = [json.loads(b) for b in re.findall(r"\[.*?\]", match.group(2))]
            claims.append((timestamp, boxes))
        except (json.JSONDecodeError, ValueError, IndexError):
            continue
    return claims


def parse_gt_answer(task: str, answer: str):
    """Ground-truth side of kwargs['answer'] -> (gt_seg, gt_vbox), as the reference reads it."""
    gt_seg, gt_vbox = [0.0, 0.0], None
    if task in ("temporal QA", "temporal QA (MCQ)"):
        gt = answer.split("\n")[1] if task == "temporal QA (MCQ)" else answer       # :116, :146
        gt_seg = [float(x) for x in ast.literal_eval(gt)]                           # :118, :147, :409
    if task == "visual QA":
        mg = re.search(r"<box>(\[.*?\])</box>", "<answer>%s</answer>" % answer)      # :204, :493
        if mg:
            try:
                gt_vbox = json.loads(mg.group(1))                                   # :216, :498
            except Exception:
                gt_vbox = None
    return gt_seg, gt_vbox


def parse_rollout(content: str, task: str, answer: str, key_frames=None, key_items=None, image_size=None,
                  image_size_refine=None, step_percent: float = 0.0) -> dict:
    """One completion + its reward kwargs -> the parsed structure the kernels consume."""
    think_match = re.search(r"<think>(.*?)</think>", content, re.DOTALL)            # :392
    answer_match = re.search(r"<answer>(.*?)</answer>", content, re.DOTALL)         # :482
    r = dict(task=task, has_think=bool(think_match), has_answer=bool(answer_match), ans_seg=None, ans_box=None,
             think_times=[], think_boxes=[], claims=[], gt_seg=[0.0, 0.0], gt_vbox=None,
             key_frames=key_frames or [], key_items=key_items or {}, image_size=image_size or (1, 1),
             image_size_refine=image_size_refine or (1, 1), step_percent=step_percent)
    m = re.search(r"<answer>\s*(.*?)\s*</answer>", content, re.DOTALL)              # :91 extract_answer
    output_ans = m.group(1).strip() if m else ""
    think = think_match.group(1) if think_match else ""
    r["gt_seg"], r["gt_vbox"] = parse_gt_answer(task, answer)
    if task in ("temporal QA", "temporal QA (MCQ)"):
        mm = re.search(r"<t>(\d+\.?\d*)</t>s to <t>(\d+\.?\d*)</t>s", output_ans)   # :119
        if mm:
            r["ans_seg"] = [float(mm.group(1)), float(mm.group(2))]
    if think_match:
        try:
            r["think_times"] = [float(x) for x in re.findall(r"<t>([\d.]+)</t>s", think)]   # :405, :447
        except ValueError:
            r["think_times"] = []
    if task == "visual QA":
        pat = r"<box>(\[.*?\])</box>"
        mp = re.search(pat, output_ans)                                             # :212
        if mp:
            try:
                r["ans_box"] = json.loads(mp.group(1))
            except Exception:
                r["ans_box"] = None
        for b in re.findall(pat, think):                                            # :505-512
            try:
                r["think_boxes"].append(json.loads(b))
            except Exception:
                pass
    elif think_match:
        r["claims"] = parse_claims(think)                                           # :535 (used by the ungated tasks)
    return r


# ----------------------------------------------------------------------------- packing
def pack_gt(gts: Sequence[dict]):
    """Per-prompt ground truth (dicts with task, step_percent, gt_seg, gt_vbox, key_frames, key_items,
    image_size, image_size_refine) -> the per-prompt arrays of o3v_rewards_soa, plus dims K, O, Gb.
    Ragged lists are padded in Python and converted with one np.array call per field (element-wise numpy
    stores cost more than the kernels at BASELINE config 4's 8192 prompts)."""
    Q = len(gts)
    frames = [g["key_frames"] for g in gts]
    objs = [[list(g["key_items"].get(str(f["idx"]), {}).values()) for f in fr] for g, fr in zip(gts, frames)]
    K = max([len(fr) for fr in frames] + [1])
    O = max([len(ob) for fo in objs for ob in fo] + [1])
    Gb = max([len(bx) for fo in objs for ob in fo for bx in ob] + [1])
    for g in gts:
        if g["task"] not in TASK_IDS:
            raise ValueError("Unknown task: %s" % g["task"])            # data_loader.py:33
    zbox = [0.0, 0.0, 0.0, 0.0]
    # existing entries as flat lists (comprehensions), their (prompt, frame, object, box) indices from the counts
    def within(counts):                                   # 0,1,..,c0-1, 0,1,..,c1-1, ...
        counts = np.asarray(counts, np.int64)
        starts = np.cumsum(counts) - counts
        return np.arange(int(counts.sum())) - np.repeat(starts, counts)
    n_fr = np.array([len(fr) for fr in frames], np.int64)
    n_ob = np.array([len(ob) for fo in objs for ob in fo], np.int64)              # per frame
    n_bx = np.array([len(bx) for fo in objs for ob in fo for bx in ob], np.int64)  # per object
    fq, fk = np.repeat(np.arange(Q), n_fr), within(n_fr)
    of = np.repeat(np.arange(n_ob.size), n_ob)                                     # object -> frame
    oq, ok, oo = fq[of], fk[of], within(n_ob)
    bo_ = np.repeat(np.arange(n_bx.size), n_bx)                                    # box -> object
    bq, bk, bo, bg = oq[bo_], ok[bo_], oo[bo_], within(n_bx)
    ft = [f["time"] for fr in frames for f in fr]
    fn, on = n_ob, n_bx
    bv = [box for fo in objs for ob in fo for bx in ob for box in bx]
    a = dict(
        task=np.array([TASK_IDS[g["task"]] for g in gts], np.int32).reshape(Q),
        step_percent=np.array([g["step_percent"] for g in gts], np.float64).reshape(Q),
        gt_flags=np.array([GF_VBOX if g["gt_vbox"] is not None else 0 for g in gts], np.int32).reshape(Q),
        gt_seg=np.array([g["gt_seg"] for g in gts], np.float64).reshape(Q, 2),
        gt_vbox=np.array([g["gt_vbox"] if g["gt_vbox"] is not None else zbox for g in gts], np.float64).reshape(Q, 4),
        image_size=np.array([g["image_size"] for g in gts], np.float64).reshape(Q, 2),
        image_refine=np.array([g["image_size_refine"] for g in gts], np.float64).reshape(Q, 2),
        n_kf=np.array([len(fr) for fr in frames], np.int32).reshape(Q),
        kf_time=np.zeros((Q, K)), n_obj=np.zeros((Q, K), np.int32), n_gtbox=np.zeros((Q, K, O), np.int32),
        gt_box=np.zeros((Q, K, O, Gb, 4)))
    if len(ft):
        a["kf_time"][fq, fk] = ft
        a["n_obj"][fq, fk] = fn
    if on.size:
        a["n_gtbox"][oq, ok, oo] = on
    if len(bv):
        a["gt_box"][bq, bk, bo, bg] = np.array(bv, np.float64).reshape(len(bv), 4)
    return a, dict(K=K, O=O, Gb=Gb)


def pack_rollouts(rollouts: Sequence[dict], G: int = 1):
    """Parsed rollouts (GT identical within each block of G) -> dict of numpy arrays in the
    o3v_rewards_soa layout, plus the dims."""
    R = len(rollouts)
    assert R % G == 0
    Q = R // G
    P = max([len(r["think_times"]) for r in rollouts] + [1])
    C = max([len(r["claims"]) for r in rollouts] + [1])
    Bc = max([len(b) for r in rollouts for _, b in r["claims"]] + [1])
    Tb = max([len(r["think_boxes"]) for r in rollouts] + [1])
    gts = [rollouts[q * G] for q in range(Q)]
    gt_arrays, gt_dims = pack_gt(gts)
    a = dict(
        flags=np.zeros(R, np.int32), ans_seg=np.zeros((R, 2)), ans_box=np.zeros((R, 4)),
        n_times=np.zeros(R, np.int32), think_times=np.zeros((R, P)), n_claims=np.zeros(R, np.int32),
        claim_t=np.zeros((R, C)), claim_nbox=np.zeros((R, C), np.int32), claim_valid=np.zeros((R, C), np.uint32),
        claim_box=np.zeros((R, C, Bc, 4)), n_tboxes=np.zeros(R, np.int32), tbox_valid=np.zeros(R, np.uint32),
        think_box=np.zeros((R, Tb, 4)))
    for i, r in enumerate(rollouts):
        f = (RF_HAS_THINK if r["has_think"] else 0) | (RF_HAS_ANSWER if r["has_answer"] else 0)
        if r["ans_seg"] is not None:
            f |= RF_ANS_SEG
            a["ans_seg"][i] = r["ans_seg"]
        if r["ans_box"] is not None and _box_ok(r["ans_box"]):
            f |= RF_ANS_BOX
            a["ans_box"][i] = r["ans_box"]
        a["flags"][i] = f
        n = len(r["think_times"])
        a["n_times"][i] = n
        a["think_times"][i, :n] = r["think_times"]
        a["n_claims"][i] = len(r["claims"])
        for c, (t, boxes) in enumerate(r["claims"]):
            a["claim_t"][i, c] = t
            a["claim_nbox"][i, c] = len(boxes)
            for b, box in enumerate(boxes):
                if _box_ok(box):
                    if b < 32:
                        a["claim_valid"][i, c] |= np.uint32(1 << b)
                    a["claim_box"][i, c, b] = box
                elif b >= 32:                                  # beyond the mask: validity lives in the slot
                    a["claim_box"][i, c, b, 0] = INVALID_BOX
        a["n_tboxes"][i] = len(r["think_boxes"])
        for b, box in enumerate(r["think_boxes"]):
            if _box_ok(box):
                if b < 32:
                    a["tbox_valid"][i] |= np.uint32(1 << b)
                a["think_box"][i, b] = box
            elif b >= 32:
                a["think_box"][i, b, 0] = INVALID_BOX
    a.update(gt_arrays)
    dims = dict(R=R, G=G, P=P, C=C, Bc=Bc, Tb=Tb, **gt_dims)
    return a, dims


def soa_bytes(arrays) -> int:
    return int(sum(v.nbytes for v in arrays.values()))


_pinned = {}
_pinned_busy = {}


def _pinned_buffer(name: str, nbytes: int) -> torch.Tensor:
    """A reusable pinned staging buffer (grown geometrically): pinning per call costs more than the kernels.
    Waits for the H2D copy that last read the buffer (see `_staged`)."""
    ev = _pinned_busy.pop(name, None)
    if ev is not None:
        ev.synchronize()
    buf = _pinned.get(name)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1 << 16, 2 * (buf.numel() if buf is not None else 0)), dtype=torch.uint8,
                          pin_memory=torch.cuda.is_available())
        _pinned[name] = buf
    return buf


def _staged(name: str):
    """Call after enqueueing an H2D copy out of staging buffer `name` on the current stream."""
    ev = torch.cuda.Event()
    ev.record()
    _pinned_busy[name] = ev


def to_device(arrays, device):
    """Host SoA -> device tensors: the arrays are packed back to back (16-byte aligned) into one pinned
    staging buffer and cross PCIe in ONE copy; the result tensors are views of the device copy."""
    items = [(k, np.ascontiguousarray(v)) for k, v in arrays.items()]
    offs, total = [], 0
    for _, v in items:
        offs.append(total)
        total += (v.nbytes + 15) // 16 * 16
    total = max(total, 16)
    stage = _pinned_buffer("soa", total)[:total]
    sn = stage.numpy()
    for (k, v), o in zip(items, offs):
        sn[o:o + v.nbytes] = v.reshape(-1).view(np.uint8)
    with torch.cuda.device(device):
        dev = stage.to(device, non_blocking=True)
        _staged("soa")
    out = {}
    for (k, v), o in zip(items, offs):
        dt = torch.int32 if v.dtype == np.uint32 else torch.from_numpy(np.empty(0, v.dtype)).dtype
        out[k] = dev[o:o + v.nbytes].view(dt).view(v.shape)
    return out


def grounded_rewards_device(dev_arrays, dims, out: Optional[torch.Tensor] = None):
    """K4 launch on device-resident SoA -> [R, 5] float64 (device)."""
    any_t = dev_arrays["flags"]
    if not any_t.is_cuda:
        raise RuntimeError("open-o3-video_b200 ops take CUDA tensors only (no CPU fallback)")
    R = dims["R"]
    if out is None:
        out = torch.empty(R, 5, dtype=torch.float64, device=any_t.device)
    soa = _lib.RewardsSoA()
    soa.R, soa.G = R, dims["G"]
    for k in ("P", "C", "Bc", "Tb", "K", "O", "Gb"):
        setattr(soa, k, dims[k])
    for name, _ in _lib.RewardsSoA._fields_[10:]:
        setattr(soa, name, dev_arrays[name].data_ptr())
    with torch.cuda.device(any_t.device):
        _lib.call("o3v_grounded_rewards", 1, _lib.load().o3v_grounded_rewards, ctypes.byref(soa),
                  ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    return out


def rewards_from_rollouts(rollouts: Sequence[dict], G: int = 1, device="cuda") -> torch.Tensor:
    if len(rollouts) == 0:
        return torch.empty(0, 5, dtype=torch.float64, device=device)
    arrays, dims = pack_rollouts(rollouts, G)
    return grounded_rewards_device(to_device(arrays, device), dims)


# ----------------------------------------------------------------------------- K6: text scanned on the device
ROLLOUT_ROWS = (("flags", torch.int32, ()), ("ans_seg", torch.float64, (2,)), ("ans_box", torch.float64, (4,)),
                ("n_times", torch.int32, ()), ("think_times", torch.float64, ("P",)), ("n_claims", torch.int32, ()),
                ("claim_t", torch.float64, ("C",)), ("claim_nbox", torch.int32, ("C",)),
                ("claim_valid", torch.int32, ("C",)), ("claim_box", torch.float64, ("C", "Bc", 4)),
                ("n_tboxes", torch.int32, ()), ("tbox_valid", torch.int32, ()), ("think_box", torch.float64, ("Tb", 4)))
DEFAULT_CAPS = dict(P=16, C=16, Bc=4, Tb=8)


def encode_completions(contents: Sequence[str]):
    """list[str] -> (uint8 tensor padded as o3v_parse_args.text requires, int64 offsets [R+1]), both in pinned
    host memory when a GPU is present.  The returned text aliases a staging buffer that the next call reuses."""
    joined = "".join(contents)
    if joined.isascii():                                    # byte offsets = character offsets
        blob = joined.encode("ascii")
        lengths = np.fromiter(map(len, contents), np.int64, len(contents))
    else:
        blobs = [c.encode("utf-8", "surrogatepass") for c in contents]
        blob = b"".join(blobs)
        lengths = np.fromiter(map(len, blobs), np.int64, len(blobs))
    total = len(blob)
    padded = (total + 15) // 16 * 16 + 16
    text = _pinned_buffer("text", padded)[:padded]
    tn = text.numpy()
    tn[:total] = np.frombuffer(blob, np.uint8)
    tn[total:] = 0
    offsets = _pinned_buffer("offsets", 8 * (len(contents) + 1))[:8 * (len(contents) + 1)].view(torch.int64)
    on = offsets.numpy()
    on[0] = 0
    np.cumsum(lengths, out=on[1:])
    return text, offsets


def parse_completions_device(text: torch.Tensor, offsets: torch.Tensor, task: torch.Tensor, G: int = 1,
                             caps: Optional[dict] = None, sync: bool = True):
    """K6 launch: UTF-8 bytes on the device -> (rollout-side arrays of o3v_rewards_soa, dims P/C/Bc/Tb).

    text uint8 (padded, see include/o3v.h), offsets int64 [R+1], task int32 [R/G], all CUDA tensors.
    With sync=True the overflow report is read back and the launch is repeated with larger rows
    when some rollout had more timestamps / claims / boxes than the capacities; with sync=False the
    caller gets the report tensor under key "overflow" and checks it itself."""
    if not (text.is_cuda and offsets.is_cuda and task.is_cuda):
        raise RuntimeError("open-o3-video_b200 ops take CUDA tensors only (no CPU fallback)")
    lib = _lib.load()
    R = offsets.numel() - 1
    caps = dict(DEFAULT_CAPS if caps is None else caps)
    dev = text.device
    with torch.cuda.device(dev):
        while True:
            # one allocation for the workspace, the overflow report and every output row (views, 16-byte aligned)
            shapes = [("overflow", torch.int32, (4,))] + [
                (name, dt, (R,) + tuple(caps[d] if isinstance(d, str) else d for d in shape)) for name, dt, shape in ROLLOUT_ROWS]
            ws_bytes = (lib.o3v_parse_workspace_bytes(R, caps["P"], caps["C"], caps["Tb"]) + 15) // 16 * 16
            sizes = [(int(np.prod(sh)) * (8 if dt == torch.float64 else 4) + 15) // 16 * 16 for _, dt, sh in shapes]
            slab = torch.empty(max(16, ws_bytes + sum(sizes)), dtype=torch.uint8, device=dev)
            ws = slab[:max(ws_bytes, 8)].view(torch.int64)
            out, o = {}, ws_bytes
            for (name, dt, sh), nb in zip(shapes, sizes):
                n = int(np.prod(sh)) * (8 if dt == torch.float64 else 4)
                out[name] = slab[o:o + n].view(dt).view(sh)
                o += nb
            a = _lib.ParseArgs()
            a.R, a.G = R, G
            for k in ("P", "C", "Bc", "Tb"):
                setattr(a, k, caps[k])
            a.text, a.offsets, a.task = text.data_ptr(), offsets.data_ptr(), task.data_ptr()
            for name in out:
                setattr(a, name, out[name].data_ptr())
            _lib.call("o3v_parse_completions", 3, lib.o3v_parse_completions, ctypes.byref(a),
                      ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel() * 8),
                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
            if not sync:
                break
            over = out["overflow"].tolist()
            if not any(over):
                break
            for k, v in zip(("P", "C", "Bc", "Tb"), over):
                if v:
                    caps[k] = v
    return out, caps


def rewards_from_text(contents: Sequence[str], gts: Sequence[dict], G: int = 1, device="cuda",
                      caps: Optional[dict] = None, to_host: bool = False):
    """Completion strings + per-prompt ground truth (see pack_gt) -> [R, 5] float64: K6 (scan, three launches)
    then K4 (numerics).  Device tensor, or (to_host=True) a numpy array.

    K4 is launched right behind K6 without waiting for the overflow report; the report and the rewards come back
    in one synchronisation and the pair is repeated with larger rows only if something did not fit."""
    R = len(contents)
    if R == 0:
        out = torch.empty(0, 5, dtype=torch.float64, device=device)
        return out.cpu().numpy() if to_host else out
    assert R == len(gts) * G
    gt_arrays, gt_dims = pack_gt(gts)
    text, offsets = encode_completions(contents)
    dev_gt = to_device(gt_arrays, device)
    with torch.cuda.device(device):
        d_text, d_off = text.to(device, non_blocking=True), offsets.to(device, non_blocking=True)
        _staged("text")
        _staged("offsets")
        caps = dict(DEFAULT_CAPS if caps is None else caps)
        while True:
            rows, caps = parse_completions_device(d_text, d_off, dev_gt["task"], G, caps, sync=False)
            over_dev = rows.pop("overflow")
            rows.update(dev_gt)
            out = grounded_rewards_device(rows, dict(R=R, G=G, **caps, **gt_dims))
            host = _pinned_buffer("result", R * 40 + 16)
            h_out = host[:R * 40].view(torch.float64).view(R, 5)
            h_over = host[R * 40:R * 40 + 16].view(torch.int32)
            h_out.copy_(out, non_blocking=True)
            h_over.copy_(over_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            over = h_over.tolist()
            if not any(over):
                break
            for k, v in zip(("P", "C", "Bc", "Tb"), over):
                if v:
                    caps[k] = v
    return h_out.numpy().copy() if to_host else out


# ----------------------------------------------------------------------------- reference-named callables
_cache = {"key": None, "val": None}


_GT_KEYS = ("answer", "key_frames", "key_items", "image_size", "image_size_refine")


def _shared_gt_group(kwargs, R: int) -> int:
    """Rollouts per prompt when the batch repeats each prompt's ground truth G times in a row, as the trainer
    builds it (`reward_kwargs[key].extend([example[key]] * num_generations)`, grpo_trainer.py:651-653): the
    repeats are the SAME objects, so identity (or equal answer strings) is enough to find G.  1 if the batch has
    no such structure.  The ground truth is then parsed and packed once per prompt."""
    cols = [kwargs[k] for k in _GT_KEYS if kwargs.get(k) is not None]
    if R < 2 or not cols:
        return 1
    same = lambda i: all(c[i] is c[i - 1] or (isinstance(c[i], str) and c[i] == c[i - 1]) for c in cols)
    G = 1
    while G < R and same(G):
        G += 1
    if G == 1 or R % G:
        return 1
    for i in range(G, R):
        if (i % G != 0) != same(i):                           # every block: G identical, then a change
            if i % G != 0:
                return 1
    return G


class _Batch:
    """The five reward columns of one batch plus what the per-callable failure semantics need."""
    __slots__ = ("val", "G", "task", "gt_err", "seg_unpack_ok", "no_key_frames", "contents")


def _gt_of_prompt(task, answer):
    """-> (gt_seg, gt_vbox, error, unpack_ok).  `error`: what `ast.literal_eval` (or the MCQ split) raises on this
    ground truth in the reference (reward_func.py:116-118, :146-147, :407-409); `unpack_ok`: the literal is a
    sequence of exactly two numbers, so `start2, end2 = gt_ans` (:137, :165) works."""
    try:
        seg, vbox = parse_gt_answer(task, answer)
        if len(seg) == 2:
            return seg, vbox, None, True
        if len(seg) > 2:                                  # evaluates, but `start2, end2 = gt_ans` cannot unpack it
            return seg[:2], vbox, None, False
        err = IndexError("list index out of range")       # gt_ans[1] at :419
    except Exception as exc:                              # noqa: BLE001 - the reference catches Exception (:175)
        err = exc
    # literal_eval itself may have worked (e.g. three numbers): thk_temporal_segment_reward only reads [0] and [1]
    try:
        gt = answer.split("\n")[1] if task == "temporal QA (MCQ)" else answer
        lit = ast.literal_eval(gt)
        seg = [float(lit[0]), float(lit[1])]
        return seg, None, None, False
    except Exception:                                     # noqa: BLE001
        return [0.0, 0.0], None, err, False


def _batch_rewards(completions, kwargs) -> _Batch:
    contents = [c[0]["content"] for c in completions]
    task = kwargs["task"][0]                                   # the reference reads element 0 for the batch
    step = kwargs["step_percent"][0] if "step_percent" in kwargs else 0.0
    # the trainer calls every reward callable with the same batch (grpo_trainer.py:646-656) but REBUILDS the kwargs
    # lists for each of them (:650-654) out of the same example objects: compute the five columns once per batch.
    # The key holds the completion texts and answers by VALUE (immutable strings) and the structured ground truth
    # by the identity of its per-rollout elements; a value snapshot of the nested key frames / items for every one
    # of the five calls cost more than the whole GPU path at training batch sizes.
    ident = lambda name: tuple(map(id, kwargs[name])) if kwargs.get(name) is not None else None
    ans = kwargs.get("answer")
    key = (tuple(contents), task, step, tuple(ans) if ans is not None else None, ident("key_frames"), ident("key_items"),
           ident("image_size"), ident("image_size_refine"))
    if _cache["key"] == key:
        return _cache["val"]
    get = lambda name, i: (kwargs[name][i] if name in kwargs and kwargs[name] is not None else None)
    R = len(contents)
    G = _shared_gt_group(kwargs, R)
    b = _Batch()
    b.G, b.task, b.contents = G, task, contents
    b.gt_err, b.seg_unpack_ok, b.no_key_frames = [], [], []
    gts = []
    for i in range(0, R, G):
        gt_seg, gt_vbox, err, unpack_ok = _gt_of_prompt(task, get("answer", i) or "")
        b.gt_err.append(err)
        b.seg_unpack_ok.append(unpack_ok)
        b.no_key_frames.append(not (get("key_frames", i) or []))
        gts.append(dict(task=task, step_percent=step, gt_seg=gt_seg, gt_vbox=gt_vbox,
                        key_frames=get("key_frames", i) or [], key_items=get("key_items", i) or {},
                        image_size=get("image_size", i) or (1, 1),
                        image_size_refine=get("image_size_refine", i) or (1, 1)))
    b.val = rewards_from_text(contents, gts, G, to_host=True)
    _cache["key"], _cache["val"] = key, b
    return b


def grounded_rewards(completions, **kwargs) -> np.ndarray:
    """All five numeric rewards for a batch: [len(completions), 5] float64 (host), before the per-callable
    failure semantics of `apply_failure_semantics`."""
    return _batch_rewards(completions, kwargs).val


_THINK_RE = re.compile(r"<think>(.*?)</think>", re.DOTALL)
_TIME_RE = re.compile(r"<t>([\d.]+)</t>s")


def apply_failure_semantics(j: int, col: np.ndarray, task: str, G: int, gt_err, seg_unpack_ok, no_key_frames,
                            contents) -> List[float]:
    """What the reference does when the GROUND TRUTH of a prompt is unusable, per callable (`col` = that callable's
    column computed with a neutral ground truth for the broken prompts; prompts are blocks of G rollouts).

    0 ans_tiou_reward: the per-rollout `try/except Exception -> 0.0` (reward_func.py:175-177) swallows the error,
      but `idx += 1` sits INSIDE the try (:174): after the first rollout whose `ast.literal_eval(answer[idx])`
      raises, idx never advances again, every later rollout of the batch re-reads that same broken answer and
      scores 0.0 as well.  Reproduced: zeros from the first rollout of the first broken prompt to the end of the batch.
      A literal that evaluates but is not two numbers only fails at `start2, end2 = gt_ans` for rollouts whose
      answer matched: those score 0.0 here too (the idx drift that follows is not reproduced).
    1 ans_viou_reward: GT box that is not JSON -> swallowed, 0.0 (:230-232).
    2 thk_temporal_segment_reward: `ast.literal_eval` is OUTSIDE any try (:409): the reference raises out of the
      callable for the first rollout that reaches it (a <think> block, task temporal QA / MCQ).  Same exception here.
    3 thk_temporal_point_reward: `min([])` over an empty key-frame list raises ValueError (:457) for the first
      rollout with a parsable timestamp in <think>.  Same here.
    4 thk_spatial_reward: no ground-truth parsing that can fail before the arithmetic; unchanged."""
    out = [float(x) for x in col]
    temporal = task in ("temporal QA", "temporal QA (MCQ)")
    if j == 0 and temporal:
        for q, err in enumerate(gt_err):
            if err is not None:
                print("Error in reward_fn for question_type '%s': %s" % ("TG" if task == "temporal QA" else "TG_MCQ", err))
                for i in range(q * G, len(out)):
                    out[i] = 0.0
                break
            if not seg_unpack_ok[q]:
                for i in range(q * G, (q + 1) * G):
                    out[i] = 0.0
    elif j == 2 and temporal:
        for q, err in enumerate(gt_err):
            if err is not None and any(_THINK_RE.search(c) for c in contents[q * G:(q + 1) * G]):
                raise err
    elif j == 3 and not (task in ("visual QA", "temporal QA", "temporal QA (MCQ)") or "General video QA" in task):
        for q, empty in enumerate(no_key_frames):
            if not empty:
                continue
            for c in contents[q * G:(q + 1) * G]:
                m = _THINK_RE.search(c)
                if not m:
                    continue
                try:
                    times = [float(x) for x in _TIME_RE.findall(m.group(1))]
                except ValueError:
                    times = []
                if times:
                    raise ValueError("min() iterable argument is empty")      # what `min([])` says (:457)
    return out


def _column(j):
    def f(completions, **kwargs) -> List[float]:
        b = _batch_rewards(completions, kwargs)
        if not any(e is not None for e in b.gt_err) and all(b.seg_unpack_ok) and not any(b.no_key_frames):
            return [float(x) for x in b.val[:, j]]
        return apply_failure_semantics(j, b.val[:, j], b.task, b.G, b.gt_err, b.seg_unpack_ok, b.no_key_frames, b.contents)
    return f


ans_tiou_reward = _column(0)
ans_viou_reward = _column(1)
thk_temporal_segment_reward = _column(2)
thk_temporal_point_reward = _column(3)
thk_spatial_reward = _column(4)
for _j, _n in enumerate(REWARD_NAMES):
    globals()[_n].__name__ = _n          # `reward_func.__name__` names the metric (grpo_trainer.py:718)
    globals()[_n].__doc__ = "B200 drop-in for reward_func.%s (numeric core on the GPU)." % _n

# The numeric subset of the reference's registry under ITS keys (grpo.py:58-66: "ans_tiou": ans_tiou_reward, ...), so
# that `reward_funcs_registry.update(open_o3_video_b200.rewards.reward_funcs_registry)` swaps exactly those entries
# ("ans_acc" and "format" stay the reference's string rewards).  The function names are accepted as keys too.
reward_funcs_registry = {n[:-len("_reward")]: globals()[n] for n in REWARD_NAMES}
reward_funcs_registry.update({n: globals()[n] for n in REWARD_NAMES})
