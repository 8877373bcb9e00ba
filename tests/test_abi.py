"""The C-ABI library loads, exports every symbol include/tapclip.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "tapclip.h")).read()
    return sorted(set(re.findall(r"TAPCLIP_API\s+[\w\s\*]+?\b(tapclip_\w+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from tapclip_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 24
    lib = _lib.load()
    for s in syms:
        assert hasattr(lib, s), f"libtapclip.so does not export {s}"
    assert set(_lib.PROTOTYPES) == set(syms), "ctypes prototypes and include/tapclip.h disagree"
    assert b"sm_100a" in lib.tapclip_version()


def test_header_cites_the_reference_interface():
    txt = open(os.path.join(ROOT, "include", "tapclip.h")).read()
    for cite in ("clip_wrapper.py:46-47", "model_wrapper.py", "attribution_monitor.py", "prompt_adjustor.py", "prompt_learner.py",
                 "train.py:65-67", "eval_metrics.py:19-29"):
        assert cite in txt


def test_no_silent_cpu_fallback():
    """Without a CUDA device the engine refuses to be created (there is no CPU path to fall back to)."""
    if torch.cuda.is_available():
        return
    from tapclip_b200 import _lib
    from tapclip_b200.configs import get_model_config
    from tapclip_b200.engine import Engine
    try:
        Engine(get_model_config("mini-16"))
        raised = False
    except _lib.TapclipError as e:
        raised = "no CPU fallback" in str(e)
    assert raised
    lib = _lib.load()
    cfg = _lib.TapclipConfig(64, 16, 256, 2, 4, 256, 2, 4, 256, 77, 0, 1)
    h = C.c_void_p()
    assert lib.tapclip_create(C.byref(cfg), C.byref(h)) != 0
    assert b"no CUDA device" in lib.tapclip_last_error() or b"CUDA" in lib.tapclip_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tapclip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} reaches into oracle/"
