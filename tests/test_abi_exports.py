"""The C-ABI library loads on a machine without a GPU, exports every symbol include/rspt_gpu.h
declares, and refuses to compute (no CPU fallback).  No compute calls are made here."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    from rspt_b200.build import build
    build()
    from rspt_b200 import _lib
    L = _lib.lib()
    hdr = open(os.path.join(ROOT, "include", "rspt_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(rspt_gpu_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in rspt_gpu.h but not exported"
    assert sorted(_lib.EXPORTS) == declared


def test_refuses_without_gpu():
    import torch
    if torch.cuda.is_available():
        return
    from rspt_b200 import _lib
    h = ctypes.c_void_p()
    assert _lib.lib().rspt_gpu_create(0, 3, 12, 8192, 3, 0, None, 4, ctypes.byref(h)) == -5
    assert not h.value


def test_product_never_touches_the_oracle():
    """The product package must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "rspt_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, fn)).read()
                assert "oracle" not in text.lower() or fn == "__init__.py" and False, f"{fn} mentions the oracle"
