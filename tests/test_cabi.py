"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/o3v.h declares; argument validation works without a GPU; the
product package never imports the oracle."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from open_o3_video_b200 import _build, _lib
    _build.build()
    return _lib.load()


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "o3v.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(o3v_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(lib):
    from open_o3_video_b200 import _lib
    declared = _declared_functions()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(lib, name), "libo3v.so does not export %s" % name
    assert sorted(_lib.SIGNATURES) == declared, "ctypes binding and include/o3v.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (o3v_[a-z0-9_]+)", out))
    assert exported == set(declared), exported ^ set(declared)


def test_version_and_strerror(lib):
    assert lib.o3v_version() == 100
    assert b"sm_100" in lib.o3v_strerror(-3)
    assert lib.o3v_strerror(0) == b"ok"
    assert b"invalid argument" in lib.o3v_strerror(-1)


def test_argument_validation_without_gpu(lib):
    """Bad arguments are rejected before any CUDA call: error codes, never a crash."""
    null = ctypes.c_void_p(0)
    one = ctypes.c_void_p(16)   # never dereferenced: validation fails first
    assert lib.o3v_eos_mask(null, 1, 1, 0, null, null, null) == -1
    assert lib.o3v_lmhead_fwd(null, null, null, 1, 1, 64, 0, null, null, 0, null, 0, null) == -1
    assert lib.o3v_lmhead_fwd(one, one, one, 128, 256, 100, 0, one, null, 0, one, 1 << 30, null) == -6  # H % 64
    assert lib.o3v_gspo_fwd_bwd(one, null, one, one, one, 6, 8, 1, 4, 0, 6, 0.04, 0.2, 0.2, 1,
                                one, null, null, null, null, null, null, one, 1 << 20, null) == -1   # N % G
    assert lib.o3v_gspo_fwd_bwd(one, null, one, one, one, 8, 8, 1, 4, 0, 8, 0.04, 0.2, 0.2, 1,
                                one, null, null, null, null, null, null, one, 8, null) == -4         # workspace
    assert lib.o3v_grounded_rewards(None, null, null) == -1
    assert lib.o3v_parse_completions(None, null, 0, null) == -1
    from open_o3_video_b200 import _lib
    a = _lib.ParseArgs()
    a.R, a.G, a.P, a.C, a.Bc, a.Tb = 4, 3, 1, 1, 1, 1                    # R % G
    assert lib.o3v_parse_completions(ctypes.byref(a), one, 1 << 20, null) == -1
    a.G, a.Bc = 1, 33                                                    # more than 32 boxes per claim
    assert lib.o3v_parse_completions(ctypes.byref(a), one, 1 << 20, null) == -1
    a.Bc = 1
    for name, _ in _lib.ParseArgs._fields_[6:]:
        setattr(a, name, 16)
    a.text = 24                                                          # text must be 16-byte aligned
    assert lib.o3v_parse_completions(ctypes.byref(a), one, 1 << 20, null) == -2
    a.text = 16
    assert lib.o3v_parse_completions(ctypes.byref(a), one, 8, null) == -4    # workspace
    assert lib.o3v_parse_workspace_bytes(4, 2, 3, 1) == 4 * 40 + 16 + 4 * (2 + 3 + 1) * 4
    assert lib.o3v_set_tunable(b"nope", 1) == -1
    assert lib.o3v_lmhead_fwd_workspace_bytes(128, 1024, 64) == 4 * 3 * 128 * 4
    # K1 vocab-group count (csrc/lmhead.cu:fwd_groups; 148 SMs assumed without a device): 4 when the work fills
    # its waves, another count only for a >= 4 % saving by the wave arithmetic, the smaller one on near-ties
    groups = lambda T, V: lib.o3v_lmhead_fwd_workspace_bytes(T, V, 3584) // (12 * T)
    assert groups(32768, 152064) == 4 and groups(131072, 19200) == 4 and groups(131072, 38144) == 4
    assert groups(34816, 152064) == 7 and groups(26624, 152064) == 7
    assert lib.o3v_gspo_workspace_bytes(8) == (3 * 8 + 4 + 8 * 32 * 4) * 4


def test_sass_is_blackwell_native():
    """tcgen05.mma / tcgen05.ld / TMA must be in the shipped binary (B200_PROFILING.md)."""
    from open_o3_video_b200 import _lib
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16" not in sass      # no legacy mma.sync path


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "open-o3-video_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
                assert "/root/reference" not in src, f


def test_ops_refuse_cpu_tensors():
    import torch
    from open_o3_video_b200 import gspo, logprob
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        gspo.eos_mask(torch.zeros(2, 4, dtype=torch.int64), 1)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        logprob.fused_logprob(torch.zeros(4, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16),
                              torch.zeros(4, dtype=torch.int64))


def test_reward_registry_uses_the_reference_keys():
    """grpo.py:58-66 registers the callables under short keys; `.update()` with ours must replace exactly the
    numeric ones and keep `__name__` (the metric name, grpo_trainer.py:718)."""
    from open_o3_video_b200 import rewards
    ref_keys = {"ans_tiou", "ans_viou", "thk_temporal_point", "thk_temporal_segment", "thk_spatial"}
    assert ref_keys <= set(rewards.reward_funcs_registry)
    for k in ref_keys:
        f = rewards.reward_funcs_registry[k]
        assert f.__name__ == k + "_reward" and rewards.reward_funcs_registry[k + "_reward"] is f
    assert "ans_acc" not in rewards.reward_funcs_registry and "format" not in rewards.reward_funcs_registry
