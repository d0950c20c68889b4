"""B200-native (sm_100a) FM-family training hot path behind the reference's Python model classes.

Public surface mirrors haan6/fm-for-online-recommendation:
  deep family      : FMAdam, DeepFMAdam, NFMAdam, DeepFMOnn, NFMOnn, AFMAdam   (models/models_online_deep/*.py)
  classical family : FM_FTRL, SFTRL_CCFM, SFTRL_Vanila, RRF_Online    (models/models_online/*.py)
All arithmetic runs in lib/libfmb200.so (hand-written CUDA, include/fmb200.h); there is no CPU path.
"""
from ._lib import FmbError, load, require_cuda  # noqa: F401

__all__ = ["FmbError", "load", "require_cuda"]


_nvtx_done = False


def _maybe_trace():
    """FMB_NVTX=1: NVTX ranges around the reference-facing methods (tracing.py); nothing happens otherwise"""
    global _nvtx_done
    if not _nvtx_done:
        _nvtx_done = True
        from . import tracing
        if tracing.enabled_by_env():
            tracing.enable()


def __getattr__(name):  # lazy: importing the package must not need torch.cuda
    if name in ("FMAdam", "DeepFMAdam", "NFMAdam", "DeepFMOnn", "NFMOnn", "AFMAdam", "EncodedBatch"):
        from . import deep
        _maybe_trace()
        return getattr(deep, name)
    if name in ("FM_FTRL", "SFTRL_CCFM", "SFTRL_Vanila", "RRF_Online"):
        from . import classical
        _maybe_trace()
        return getattr(classical, name)
    raise AttributeError(name)
