"""ctypes binding of lib/libo3v.so (C ABI: include/o3v.h).

This is the reference-side FFI stub of INTEGRATION.md.  It fails loudly: a missing library,
a missing symbol, a non-B200 device or any non-zero return code raises; nothing falls back
to the CPU or to eager PyTorch.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libo3v.so")


class RewardsSoA(ctypes.Structure):
    """struct o3v_rewards_soa (include/o3v.h)."""
    _fields_ = [("R", c_int64), ("G", c_int64),
                ("P", c_int32), ("C", c_int32), ("Bc", c_int32), ("Tb", c_int32),
                ("K", c_int32), ("O", c_int32), ("Gb", c_int32), ("pad_", c_int32)] + [(n, c_void_p) for n in (
                    "flags", "ans_seg", "ans_box", "n_times", "think_times", "n_claims", "claim_t",
                    "claim_nbox", "claim_valid", "claim_box", "n_tboxes", "tbox_valid", "think_box",
                    "task", "step_percent", "gt_flags", "gt_seg", "gt_vbox", "image_size", "image_refine", "n_kf",
                    "kf_time", "n_obj", "n_gtbox", "gt_box")]


class VstarSoA(ctypes.Structure):
    """struct o3v_vstar_soa (include/o3v.h)."""
    _fields_ = [("I", c_int64), ("F", c_int32), ("Pb", c_int32)] + [(n, c_void_p) for n in (
        "t_valid", "gt_seg", "pred_seg", "sp_valid", "n_frames", "gt_box", "n_pb", "pb_valid", "pb")]


class ParseArgs(ctypes.Structure):
    """struct o3v_parse_args (include/o3v.h)."""
    _fields_ = [("R", c_int64), ("G", c_int64),
                ("P", c_int32), ("C", c_int32), ("Bc", c_int32), ("Tb", c_int32)] + [(n, c_void_p) for n in (
                    "text", "offsets", "task", "flags", "ans_seg", "ans_box", "n_times", "think_times", "n_claims",
                    "claim_t", "claim_nbox", "claim_valid", "claim_box", "n_tboxes", "tbox_valid", "think_box",
                    "overflow")]


# name -> (restype, argtypes); must list EVERY function include/o3v.h declares
SIGNATURES = {
    "o3v_version": (c_int, []),
    "o3v_strerror": (c_char_p, [c_int]),
    "o3v_check_device": (c_int, []),
    "o3v_set_tunable": (c_int, [c_char_p, c_int]),
    "o3v_debug_occupy_sms": (c_int, [c_int32, c_int32, c_int64, c_void_p]),
    "o3v_debug_gemm": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32,
                               c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "o3v_eos_mask": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "o3v_lmhead_fwd_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "o3v_lmhead_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                               c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "o3v_lmhead_merge_stats": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "o3v_lmhead_merge_stats_peers": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "o3v_allreduce_bf16_peers": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int32, c_void_p]),
    "o3v_add_slabs_f32": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "o3v_reduce_scatter_bf16_peers": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p]),
    "o3v_lmhead_dlogits": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_void_p]),
    "o3v_lmhead_bwd_dhidden": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                       c_void_p, c_int32, c_void_p]),
    "o3v_lmhead_bwd_dweight": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                       c_void_p, c_int32, c_void_p]),
    "o3v_lmhead_fwd_exp": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64,
                                   c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "o3v_lmhead_softmax_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                        c_void_p, c_void_p]),
    "o3v_lmhead_bwd_dhidden_exp": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p,
                                           c_int32, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "o3v_lmhead_bwd_dweight_exp": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                           c_void_p, c_int32, c_void_p, c_void_p]),
    "o3v_lmhead_bwd_dhidden_scatter": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64,
                                               c_int64, c_int64, c_int64, c_int64, c_void_p]),
    "o3v_sum_slots_bf16": (c_int, [c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p, c_int32, c_void_p]),
    "o3v_lmhead_softmax_bwd_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p]),
    "o3v_lmhead_bwd_dhidden_fused": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                             c_void_p, c_int32, c_void_p]),
    "o3v_lmhead_bwd_dweight_fused": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                             c_void_p, c_int32, c_void_p]),
    "o3v_gspo_workspace_bytes": (c_size_t, [c_int64]),
    "o3v_gspo_fwd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int64, c_int64, c_int64, c_int64, c_int64, c_int64,
                                 c_float, c_float, c_float, c_int32,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_size_t, c_void_p]),
    "o3v_grounded_rewards": (c_int, [ctypes.POINTER(RewardsSoA), c_void_p, c_void_p]),
    "o3v_vstar_scores": (c_int, [ctypes.POINTER(VstarSoA), c_void_p, c_void_p]),
    "o3v_parse_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32, c_int32]),
    "o3v_parse_completions": (c_int, [ctypes.POINTER(ParseArgs), c_void_p, c_size_t, c_void_p]),
}

_lib = None


def load() -> ctypes.CDLL:
    """dlopen lib/libo3v.so and bind every symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            "libo3v.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU / PyTorch fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


class O3VError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = load().o3v_strerror(code)
        super().__init__("%s failed with code %d: %s" % (where, code, msg.decode() if msg else "?"))


def check(code: int, where: str) -> None:
    if code != 0:
        raise O3VError(code, where)


def set_tunable(name: str, value: int) -> None:
    check(load().o3v_set_tunable(name.encode(), int(value)), "o3v_set_tunable(%s)" % name)


class Trace:
    """Optional per-call instrumentation used by bench.py: counts the kernels this library
    launches and (events=True) brackets every C-ABI call with CUDA events on the launching
    stream, so per-kernel durations are measured live inside the timed region."""

    def __init__(self, events: bool = False):
        self.events = events
        self.launches = 0
        self.records = {}
        self.calls = []          # (entry point, its integer arguments in order): e.g. o3v_lmhead_fwd -> (T, V, H, ...)

    def durations_ms(self):
        """name -> list of elapsed ms (call after a device synchronize)."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.records.items()}


trace = None   # set to a Trace() to instrument


def call(name: str, n_kernels: int, fn, *args) -> None:
    """Invoke one C-ABI entry point (which enqueues `n_kernels` kernels) and raise on error."""
    t = trace
    if t is not None and t.events:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(fn(*args), name)
        b.record()
        t.records.setdefault(name, []).append((a, b))
    else:
        check(fn(*args), name)
    if t is not None:
        t.launches += n_kernels
        t.calls.append((name, tuple(a for a in args if isinstance(a, int) and not isinstance(a, bool))))
