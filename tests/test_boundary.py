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


def test_dropin_import_paths_mirror_the_reference():
    """`from models.models_online_deep.fm_adam import FMAdam` (main_experiment.py:9-13) resolves to the
    B200 classes once fm_for_online_recommendation_b200/dropin is first on sys.path."""
    import importlib
    import sys
    dropin = os.path.join(PKG, "dropin")
    saved = {k: v for k, v in sys.modules.items() if k == "models" or k.startswith("models.")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, dropin)
    try:
        import fm_for_online_recommendation_b200.classical as cl
        import fm_for_online_recommendation_b200.deep as dp
        for mod, name, home in [("models.models_online_deep.fm_adam", "FMAdam", dp),
                                ("models.models_online_deep.deepfm_adam", "DeepFMAdam", dp),
                                ("models.models_online_deep.nfm_adam", "NFMAdam", dp),
                                ("models.models_online_deep.deepfm_onn", "DeepFMOnn", dp),
                                ("models.models_online_deep.nfm_onn", "NFMOnn", dp),
                                ("models.models_online.FM_FTRL", "FM_FTRL", cl),
                                ("models.models_online.SFTRL_CCFM", "SFTRL_CCFM", cl),
                                ("models.models_online.SFTRL_Vanila", "SFTRL_Vanila", cl)]:
            assert getattr(importlib.import_module(mod), name) is getattr(home, name)
    finally:
        sys.path.remove(dropin)
        for k in [k for k in sys.modules if k == "models" or k.startswith("models.")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_reference_constructor_signatures_are_kept():
    """positional order and defaults of the five deep constructors (note NFMOnn's
    (num_classes, batch_size) order, nfm_onn.py:14-15)."""
    import inspect
    import fm_for_online_recommendation_b200.deep as dp
    want = {
        "FMAdam": ["feature_sizes", "embedding_size", "num_classes", "b", "n", "use_cuda"],
        "DeepFMAdam": ["feature_sizes", "embedding_size", "num_hidden_layers", "neuron_per_hidden_layer", "batch_size",
                       "num_classes", "b", "n", "use_cuda"],
        "NFMAdam": ["feature_sizes", "embedding_size", "num_hidden_layers", "neuron_per_hidden_layer", "num_classes", "b",
                    "n", "use_cuda"],
        "DeepFMOnn": ["feature_sizes", "embedding_size", "num_hidden_layers", "neuron_per_hidden_layer", "batch_size",
                      "num_classes", "b", "n", "s", "use_cuda"],
        "NFMOnn": ["feature_sizes", "embedding_size", "num_hidden_layers", "neuron_per_hidden_layer", "num_classes",
                   "batch_size", "b", "n", "s", "use_cuda"],
    }
    for name, params in want.items():
        got = [p for p in inspect.signature(getattr(dp, name).__init__).parameters if p != "self"]
        assert got[:len(params)] == params, (name, got)
        sig = inspect.signature(getattr(dp, name).__init__)
        assert sig.parameters["embedding_size"].default == 4 and sig.parameters["n"].default == 0.01
