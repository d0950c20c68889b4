"""Evaluation metrics on the device: drop-in versions of utils/metric_manager.py plus an exact AUC.

`regression_metric` / `classfication_metric` keep the reference's names (typo included), arguments and return shapes
(utils/metric_manager.py:7-29); inputs may be numpy arrays, lists or tensors.  The curves are produced by one kernel each
(csrc/metrics.cu); there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


def _dev64(a):
    if torch.is_tensor(a):
        return a.detach().to("cuda", torch.float64).reshape(-1).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1))).cuda()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def regression_metric(pred, real):
    """running mean squared error: [n+1, 1] fp64, first entry inf (metric_manager.py:7-15)"""
    lib = _lib.require_cuda()
    p, r = _dev64(pred), _dev64(real)
    out = torch.empty(p.numel() + 1, dtype=torch.float64, device="cuda")
    check(lib.fmb_metric_regression(ptr(p), ptr(r), p.numel(), ptr(out), _stream()), "fmb_metric_regression")
    return out.cpu().numpy().reshape(-1, 1)


def classfication_metric(pred, real):
    """(per-sample scaled log-loss [n,1], running accuracy [n,1]) (metric_manager.py:18-29)"""
    lib = _lib.require_cuda()
    p, r = _dev64(pred), _dev64(real)
    n = p.numel()
    metric = torch.empty(n, dtype=torch.float64, device="cuda")
    acc = torch.empty(n, dtype=torch.float64, device="cuda")
    check(lib.fmb_metric_classification(ptr(p), ptr(r), n, ptr(metric), ptr(acc), _stream()), "fmb_metric_classification")
    return metric.cpu().numpy().reshape(-1, 1), acc.cpu().numpy().reshape(-1, 1)


def auc(scores, labels):
    """exact ROC AUC (ties count one half): labels > 0 are positives.  Pair counting on the device, n <= 2^20."""
    lib = _lib.require_cuda()
    s = (scores.detach() if torch.is_tensor(scores) else torch.from_numpy(np.asarray(scores, np.float32))).to(
        "cuda", torch.float32).reshape(-1).contiguous()
    y = (labels.detach() if torch.is_tensor(labels) else torch.from_numpy(np.asarray(labels, np.float32))).to(
        "cuda", torch.float32).reshape(-1).contiguous()
    c = torch.zeros(4, dtype=torch.int64, device="cuda")
    check(lib.fmb_auc_pairs(ptr(s), ptr(y), s.numel(), ptr(c), _stream()), "fmb_auc_pairs")
    gt, eq, npos, nneg = (int(v) for v in c.cpu())
    if npos == 0 or nneg == 0:
        return float("nan")
    return (gt + 0.5 * eq) / (npos * nneg)


def confusion(pred, labels):
    """{'tp','fp','tn','fn'} of boolean predictions against 0/1 labels (fm_adam.py:101-111)"""
    lib = _lib.require_cuda()
    p = (pred.detach() if torch.is_tensor(pred) else torch.from_numpy(np.asarray(pred).astype(np.uint8))).to(
        "cuda", torch.uint8).reshape(-1).contiguous()
    y = (labels.detach() if torch.is_tensor(labels) else torch.from_numpy(np.asarray(labels, np.float32))).to(
        "cuda", torch.float32).reshape(-1).contiguous()
    c = torch.zeros(4, dtype=torch.int64, device="cuda")
    check(lib.fmb_confusion(ptr(p), ptr(y), p.numel(), ptr(c), _stream()), "fmb_confusion")
    tp, fp, tn, fn = (int(v) for v in c.cpu())
    return {"tp": tp, "fp": fp, "tn": tn, "fn": fn}
