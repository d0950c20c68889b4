"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol declared
in include/fmb200.h, fails loudly without a GPU, and the product never touches oracle/."""
import ctypes
import os
import re

import pytest

from _util import ROOT

PKG = os.path.join(ROOT, "fm_for_online_recommendation_b200")


def header_functions():
    src = open(os.path.join(ROOT, "include", "fmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fmb_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    from fm_for_online_recommendation_b200 import build
    so = build.build()
    lib = ctypes.CDLL(so)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fmb200.h but not exported"


def test_ctypes_table_covers_the_header():
    from fm_for_online_recommendation_b200 import _lib
    assert sorted(_lib._SIGS) == header_functions()


def test_no_cpu_fallback_and_loud_failure():
    import torch
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.load()
    assert lib.fmb_version() >= 100
    if not torch.cuda.is_available():
        with pytest.raises(pkg.FmbError):
            pkg.require_cuda()
        with pytest.raises(pkg.FmbError):
            pkg.FMAdam([3, 4], embedding_size=4)


def test_argument_errors_are_reported_not_thrown():
    import fm_for_online_recommendation_b200 as pkg
    lib = pkg.load()
    rc = lib.fmb_fm_forward(None, None, None, None, 4, 2, 4, None, None, None, None, None, None, 0, None, None, None)
    assert rc == -1 and b"fmb_fm_forward" in lib.fmb_last_error()
    assert lib.fmb_rowp(10) == 16 and lib.fmb_rowp(64) == 80 and lib.fmb_kp4(10) == 12


def test_product_never_imports_the_oracle():
    bad = []
    for dp, _, fns in os.walk(PKG):
        if os.path.basename(dp) in ("build", "lib", "__pycache__"):
            continue
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+\"[^\"]*oracle", txt, flags=re.M):
                    bad.append(fn)
    assert not bad, bad
