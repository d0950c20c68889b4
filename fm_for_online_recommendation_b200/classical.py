"""Drop-in replacements for the reference's classical online learners (models/models_online/*.py):
`FM_FTRL` (FM_FTRL.py), `SFTRL_CCFM` (SFTRL_CCFM.py), `SFTRL_Vanila` (SFTRL_Vanila.py).

Same constructor `(inputs_matrix[N,d] DoubleTensor, outputs[N] DoubleTensor, task, learning_rate, m)`,
same `online_learning()` -> `(preds, reals, seconds)`, same state attributes afterwards
(`w1`, `W2` / `BT_P`, `BT_N`, `row_count_p`, `row_count_n` / `w`, `g_w`), same
`ValueError('Nan contained')`.  The whole fp64 per-example stream runs as ONE persistent CUDA kernel
(csrc/classical.cu); there is no CPU path.
"""
import ctypes as C
import time

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

Tensor_type = torch.DoubleTensor


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class FM_Base(torch.nn.Module):
    """FM_Base.py:15-66."""

    def __init__(self, inputs_matrix, outputs, task, learning_rate, feature_m):
        super().__init__()
        self._lib = _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())
        X = torch.as_tensor(inputs_matrix, dtype=torch.float64)
        self._X = X.to(self.device).contiguous()
        self.At = self._X.t()
        self.b = torch.as_tensor(outputs, dtype=torch.float64).reshape(-1).to(self.device).contiguous()
        self._thres = 1e-12
        self.num_data = X.shape[0]
        self.num_feature = X.shape[1]
        if task not in ("reg", "cls"):
            raise NotImplementedError
        self.task = task
        self.eta = learning_rate
        self.m = feature_m

    def _task_id(self):
        return 1 if self.task == "cls" else 0

    def _finish(self, preds, status, start):
        st = int(status.item())
        if st != 0:
            raise ValueError('Nan contained')  # FM_FTRL.py:64-65
        p = preds.cpu().numpy()
        real = self.b.cpu().numpy()
        for idx in range(0, self.num_data, 1000):
            print(' %d th : pred %f , real %f ' % (idx, p[idx], real[idx]))
        end = time.time()
        print('learning time : %f ' % (end - start))
        return p, real, (end - start)

    def online_learning(self, logger=None):
        raise NotImplementedError


class FM_FTRL(FM_Base):
    """FM_FTRL.py:25-92: w = -eta * (accumulated gradient); second-order gradient without the loss factor."""

    def __init__(self, inputs_matrix, outputs, task, learning_rate, num_feature):
        super().__init__(inputs_matrix, outputs, task, learning_rate, num_feature)
        self.model_name = "FM_FTRL"

    def _init_parameter(self):
        # same CPU RNG stream as FM_FTRL.py:42-43 (fp32 randn cast to double)
        self.w1 = torch.randn(self.num_feature, 1).type(Tensor_type).to(self.device)
        self.W2 = torch.randn(2 * self.m, self.num_feature - 1).type(Tensor_type).to(self.device)

    def online_learning(self):
        start = time.time()
        self._init_parameter()
        print(self.model_name + '_' + str(self.eta) + '_' + str(self.m) + '_start')
        N, d, m2 = self.num_data, self.num_feature, 2 * self.m
        self.w1 = self.w1.contiguous()
        self.W2 = self.W2.contiguous()
        g_w1 = torch.zeros(d, dtype=torch.float64, device=self.device)
        g_W2 = torch.zeros(m2, d - 1, dtype=torch.float64, device=self.device)
        preds = torch.empty(N, dtype=torch.float64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self._lib.fmb_ftrl_fm_run(ptr(self._X), ptr(self.b), N, d, m2, self._task_id(), float(self.eta),
                                        ptr(self.w1), ptr(self.W2), ptr(g_w1), ptr(g_W2), ptr(preds), ptr(status),
                                        _stream()), "fmb_ftrl_fm_run")
        return self._finish(preds, status, start)


class _SFTRL(FM_Base):
    _VANILA = 0

    def __init__(self, inputs_matrix, outputs, task, learning_rate, num_feature):
        super().__init__(inputs_matrix, outputs, task, learning_rate, num_feature)
        self.row_count_p = 0
        self.row_count_n = 0
        ds = self.num_feature - 1 if self._VANILA else self.num_feature
        self.BT_P = torch.zeros(ds, 2 * self.m, dtype=torch.float64, device=self.device)
        self.BT_N = torch.zeros(ds, 2 * self.m, dtype=torch.float64, device=self.device)
        if self._VANILA:
            self.w = torch.zeros(self.num_feature, 1, dtype=torch.float64, device=self.device)
            self.g_w = torch.zeros(self.num_feature, 1, dtype=torch.float64, device=self.device)

    def online_learning(self):
        start = time.time()
        print("==" * 20)
        print(self.model_name + '_' + str(self.eta) + '_' + str(self.m) + '_start')
        N, d, m = self.num_data, self.num_feature, self.m
        preds = torch.empty(N, dtype=torch.float64, device=self.device)
        status = torch.zeros(1, dtype=torch.int32, device=self.device)
        rc = torch.tensor([self.row_count_p, self.row_count_n], dtype=torch.int32, device=self.device)
        wsb = self._lib.fmb_sftrl_workspace_bytes(d, m)
        ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
        w = self.w if self._VANILA else None
        g_w = self.g_w if self._VANILA else None
        check(self._lib.fmb_sftrl_run(ptr(self._X), ptr(self.b), N, d, m, self._task_id(), self._VANILA,
                                      float(self.eta), ptr(self.BT_P), ptr(self.BT_N), ptr(rc), ptr(w), ptr(g_w),
                                      ptr(preds), ptr(status), ptr(ws), wsb, _stream()), "fmb_sftrl_run")
        out = self._finish(preds, status, start)
        self.row_count_p, self.row_count_n = (int(v) for v in rc.cpu())
        return out


class SFTRL_CCFM(_SFTRL):
    """SFTRL_CCFM.py:18-121."""

    def __init__(self, inputs_matrix, outputs, task, learning_rate, num_feature):
        super().__init__(inputs_matrix, outputs, task, learning_rate, num_feature)
        self.model_name = "SFTRL_CCFM"


class SFTRL_Vanila(_SFTRL):
    """SFTRL_Vanila.py:16-123."""
    _VANILA = 1

    def __init__(self, inputs_matrix, outputs, task, learning_rate, num_feature):
        super().__init__(inputs_matrix, outputs, task, learning_rate, num_feature)
        self.model_name = "SFTRL_Vanila"


class RRF_Online(torch.nn.Module):
    """RRF_Online.py:18-187: online reparameterised random Fourier features.  Same constructor and state
    attributes (`gamma` log-scale [d,1], `w` [2D], `eps` [d,D]), same RNG draws on the CPU in the same order
    (np.random.rand for gamma, torch.randn for w, torch.randn for eps), `online_learning()` -> (preds, reals, seconds)
    with only the non-NaN samples appended.  The stream runs as one persistent fp64 kernel (csrc/rrf.cu)."""

    def __init__(self, inputs_matrix, outputs, task, loss_type=None, gamma=None, w=None, num_sampled_spectral=10,
                 random_seed=100, lr_RRF_w=0.05, lr_RRF_gamma=0.05):
        super().__init__()
        self._lib = _lib.require_cuda()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.X = torch.as_tensor(inputs_matrix, dtype=torch.float64).to(self.device).contiguous()
        self.Y = torch.as_tensor(outputs, dtype=torch.float64).reshape(-1).to(self.device).contiguous()
        self.loss_type = loss_type
        self.num_feature = self.X.shape[1]
        self.model_name = "RRF_Online"
        self.task = task
        self.num_sampled_spectral = num_sampled_spectral
        self.lr_RRF_w = lr_RRF_w
        self.lr_RRF_gamma = lr_RRF_gamma
        self.random_seed = random_seed
        self._init_param(gamma, w, loss_type)

    def _init_param(self, gamma, w, loss_type):   # RRF_Online.py:46-67
        if self.task == 'cls':
            self.loss_type = 'logit' if loss_type is None else 'hinge'
        elif self.task == 'reg':
            self.loss_type = 'l2' if loss_type is None else 'l1'
        else:
            raise NotImplementedError('wrong task assigned')
        if gamma is None:
            g = Tensor_type(np.log(np.random.rand(self.num_feature, 1)))
        else:
            g = Tensor_type(np.log(gamma) * np.ones((self.num_feature, 1)))
        self.gamma = g.to(self.device)
        if w is None:
            self.w = (0.1 * torch.randn(2 * self.num_sampled_spectral).type(Tensor_type)).to(self.device)
        else:
            self.w = torch.as_tensor(w, dtype=torch.float64).to(self.device).clone()
        self.eps = torch.randn(self.num_feature, self.num_sampled_spectral).type(Tensor_type).to(self.device)

    def online_learning(self):
        if self.loss_type in ('hinge', 'l1'):
            raise NotImplementedError('wrong loss type in get_grad')   # RRF_Online.py:115-122
        start = time.time()
        print("==" * 20)
        N = self.X.shape[0]
        preds = torch.empty(N, dtype=torch.float64, device=self.device)
        nvalid = torch.zeros(1, dtype=torch.int32, device=self.device)
        gam = self.gamma.reshape(-1).contiguous()
        wv = self.w.contiguous()
        eps = self.eps.contiguous()
        check(self._lib.fmb_rrf_run(ptr(self.X), ptr(self.Y), N, self.num_feature, self.num_sampled_spectral,
                                    1 if self.task == 'cls' else 0, float(self.lr_RRF_w), float(self.lr_RRF_gamma), ptr(gam),
                                    ptr(wv), ptr(eps), ptr(preds), ptr(nvalid), _stream()), "fmb_rrf_run")
        self.gamma = gam.reshape(-1, 1)
        self.w = wv
        n = int(nvalid.item())
        p = preds[:n].cpu().numpy()
        real = self.Y.cpu().numpy()
        # the reference appends the label of every non-NaN sample; with no NaN that is the whole stream
        reals = real if n == N else real[:n]
        for t in range(0, min(n, N), 1000):
            print(' %d th : pred %f , real %f ' % (t, p[t], real[t]))
        end = time.time()
        print('learning time : %f ' % (end - start))
        return p, reals, (end - start)
