"""Test helper: builds tests/hostbuild/libscan_host.so (csrc/scan_core.cuh compiled by g++) and wraps it.

Only the CPU test-suite uses this; the product path is the CUDA kernel (csrc/parse.cu)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostbuild", "scan_host.cpp")
LIB = os.path.join(HERE, "hostbuild", "libscan_host.so")
CORE = os.path.join(os.path.dirname(HERE), "open-o3-video_b200", "csrc", "scan_core.cuh")

ROW_NAMES = ("flags", "ans_seg", "ans_box", "n_times", "think_times", "n_claims", "claim_t", "claim_nbox",
             "claim_valid", "claim_box", "n_tboxes", "tbox_valid", "think_box")


class Args(ctypes.Structure):
    _fields_ = [("R", ctypes.c_int64), ("G", ctypes.c_int64), ("P", ctypes.c_int32), ("C", ctypes.c_int32),
                ("Bc", ctypes.c_int32), ("Tb", ctypes.c_int32)] + [
        (n, ctypes.c_void_p) for n in ("text", "offsets", "task") + ROW_NAMES + ("overflow",)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if (not os.path.isfile(LIB)) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", SRC, "-o", LIB], check=True)
    lib = ctypes.CDLL(LIB)
    lib.scan_host_python_float.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    lib.scan_host_json_box.argtypes = [ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int),
                                       ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def python_float(s: str, strip=True):
    b = s.encode("utf-8", "surrogatepass")
    out = ctypes.c_double()
    ok = load().scan_host_python_float(b, len(b), 1 if strip else 0, ctypes.byref(out))
    return out.value if ok else None


def json_box(s: str):
    """-> None (invalid JSON) or (len, numeric, first four values)."""
    b = s.encode("utf-8", "surrogatepass")
    n, num, out = ctypes.c_int(), ctypes.c_int(), (ctypes.c_double * 4)()
    if load().scan_host_json_box(b, len(b), ctypes.byref(n), ctypes.byref(num), out) == 0:
        return None
    return n.value, bool(num.value), list(out)


def empty_rows(R, P, C, Bc, Tb, fill=True):
    shapes = dict(flags=(np.int32, ()), ans_seg=(np.float64, (2,)), ans_box=(np.float64, (4,)), n_times=(np.int32, ()),
                  think_times=(np.float64, (P,)), n_claims=(np.int32, ()), claim_t=(np.float64, (C,)),
                  claim_nbox=(np.int32, (C,)), claim_valid=(np.uint32, (C,)), claim_box=(np.float64, (C, Bc, 4)),
                  n_tboxes=(np.int32, ()), tbox_valid=(np.uint32, ()), think_box=(np.float64, (Tb, 4)))
    return {k: np.full((R,) + sh, -777 if dt == np.float64 else 0x5A5A5A5A, dt) for k, (dt, sh) in shapes.items()}


def parse(text_buf, offsets, task_ids, G, P, C, Bc, Tb):
    """Host build of the scanner on an encoded batch -> (rows, overflow[4])."""
    R = len(offsets) - 1
    out = empty_rows(R, P, C, Bc, Tb)
    ov = np.zeros(4, np.int32)
    task = np.ascontiguousarray(task_ids, np.int32)
    a = Args(R=R, G=G, P=P, C=C, Bc=Bc, Tb=Tb)
    a.text, a.offsets, a.task, a.overflow = text_buf.ctypes.data, offsets.ctypes.data, task.ctypes.data, ov.ctypes.data
    for k in out:
        setattr(a, k, out[k].ctypes.data)
    load().scan_host_parse(ctypes.byref(a))
    return out, ov


def mismatches(got, exp, mask):
    """Rows where `got` differs from `exp` on the entries `mask` marks as defined (doubles compared as bits,
    any NaN == any NaN)."""
    bad = set()
    for k in exp:
        g, e = np.asarray(got[k]), np.asarray(exp[k])
        if e.dtype == np.float64:
            ne = (g.view(np.uint64) != e.view(np.uint64)) & ~(np.isnan(g) & np.isnan(e))
        else:
            ne = g.view(e.dtype) != e
        ne &= mask[k]
        bad |= {(int(r), k) for r in np.argwhere(ne)[:, 0]}
    return bad
