"""Opt-in NVTX ranges around the reference-facing methods ("Tracing / profiling").

The reference only brackets its loops with `time()` (fm_adam.py:96,118; FM_FTRL.py:48,89-90).  Here a profiler capture
(`ncu --nvtx`, Nsight Systems) can attribute kernels to the reference call they belong to:

    FMB_NVTX=1 ncu --nvtx --nvtx-include "DeepFMAdam.update_embedding/" ... python bench.py

Off by default and free when off: the methods are wrapped once, at import, only when FMB_NVTX=1 (or `enable()` is called);
no per-call check is left on the hot path otherwise.
"""
import functools
import os

_METHODS = ("first_order", "second_order", "forward_fm", "forward", "update_embedding", "fit", "predict", "predict_proba",
            "run_experiment", "online_learning", "update_embedding_fused", "update_embedding_peers",
            "update_embedding_pipelined")
_wrapped = set()


def _nvtx():
    import torch
    return torch.cuda.nvtx


def wrap(cls, methods=_METHODS, nvtx=None):
    """Wrap the methods `cls` itself defines in a range named 'Class.method'; idempotent.  Returns the wrapped names."""
    done = []
    for name in methods:
        fn = cls.__dict__.get(name)
        if fn is None or not callable(fn) or (cls, name) in _wrapped:
            continue
        label = f"{cls.__name__}.{name}"

        def ranged(self, *a, __fn=fn, __label=label, **kw):
            nv = nvtx or _nvtx()
            nv.range_push(__label)
            try:
                return __fn(self, *a, **kw)
            finally:
                nv.range_pop()

        setattr(cls, name, functools.wraps(fn)(ranged))
        _wrapped.add((cls, name))
        done.append(name)
    return done


def enable(nvtx=None):
    """Wrap every reference-facing class of the package (base classes included: the methods live there)."""
    from . import classical, deep, sharded
    seen = []
    for mod in (deep, classical, sharded):
        for obj in vars(mod).values():
            if isinstance(obj, type) and obj.__module__ == mod.__name__:
                if wrap(obj, nvtx=nvtx):
                    seen.append(obj.__name__)
    return seen


def enabled_by_env():
    return os.environ.get("FMB_NVTX", "0") not in ("", "0")
