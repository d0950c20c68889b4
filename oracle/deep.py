"""ctypes front-end of oracle/fm_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``OracleDeep`` mirrors the method surface of the reference's deep family
(models/models_online_deep/*.py: forward / forward_fm / update_embedding / fit / predict /
run_experiment) on numpy arrays, with the tables of all fields concatenated (global row id =
field offset + local id).
"""
import ctypes as C

import numpy as np

from . import build

KINDS = {"FMAdam": 0, "DeepFMAdam": 1, "NFMAdam": 2, "DeepFMOnn": 3, "NFMOnn": 4}

_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int32)
_u8 = C.POINTER(C.c_uint8)


class _Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("F", C.c_int32), ("k", C.c_int32), ("L", C.c_int32), ("H", C.c_int32),
                ("R", C.c_int32), ("batch_size", C.c_int32), ("update_mode", C.c_int32),
                ("w1", _f), ("V", _f), ("mlp", _f), ("bias", _f), ("alpha", _f),
                ("lr", C.c_float), ("hb", C.c_float), ("hs", C.c_float),
                ("gA", _f), ("gB", _f), ("touched", _u8), ("shard_G", C.c_int32), ("rank_B", C.c_int32),
                ("fz_V", _f), ("fn_V", _f), ("fz_w1", _f), ("fn_w1", _f), ("f_bias", _f),
                ("f_beta", C.c_float), ("f_l1", C.c_float), ("f_l2", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        P = C.POINTER(_Model)
        L.orc_sum_aten.restype = C.c_float
        L.orc_sum_aten.argtypes = [_f, C.c_int64]
        L.orc_mlp_numel.restype = C.c_int64
        L.orc_mlp_numel.argtypes = [P]
        L.orc_fm_forward.restype = None
        L.orc_fm_forward.argtypes = [P, _i, _f, C.c_int, _f, _f, _f, _f, _f, _f]
        L.orc_forward.restype = None
        L.orc_forward.argtypes = [P, _i, _f, C.c_int, _f, _f]
        L.orc_predict.restype = None
        L.orc_predict.argtypes = [P, _i, _f, C.c_int, _u8]
        L.orc_update_embedding.restype = C.c_float
        L.orc_update_embedding.argtypes = [P, _i, _f, _f, C.c_int]
        L.orc_fit.restype = None
        L.orc_fit.argtypes = [P, _i, _f, _f, C.c_int]
        L.orc_run_experiment.restype = None
        L.orc_run_experiment.argtypes = [P, _i, _f, _f, C.c_int, C.POINTER(C.c_int64), _u8]
        L.orc_loss_delta.restype = C.c_float
        L.orc_loss_delta.argtypes = [C.c_int, _f, _f, C.c_int, _f]
        L.orc_update_dense.restype = None
        L.orc_update_dense.argtypes = [_f, _f, C.c_int64, C.c_float, C.c_int]
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def mlp_numel(k, L, H):
    return 0 if L == 0 else H * k + H + (L - 1) * (H * H + H)


class OracleDeep:
    """One deep-family learner. Parameters live in numpy arrays the caller may read or overwrite:
    ``w1`` [R], ``V`` [R,k], ``mlp`` flat (W0[H,k] c0[H] W1[H,H] c1[H] ...), ``bias`` [1], ``alpha`` [L]."""

    def __init__(self, kind, feature_sizes, k, L=0, H=0, lr=0.01, bias=0.99, batch_size=1, hb=0.99, hs=0.2,
                 update_mode=0, seed=0):
        self.kind = kind
        self.feature_sizes = list(feature_sizes)
        self.F = len(feature_sizes)
        self.offsets = np.concatenate([[0], np.cumsum(feature_sizes)]).astype(np.int64)
        self.R = int(self.offsets[-1])
        self.k, self.L, self.H = k, (0 if kind == "FMAdam" else L), (0 if kind == "FMAdam" else H)
        rng = np.random.RandomState(seed)
        self.w1 = rng.standard_normal(self.R).astype(np.float32)
        self.V = rng.standard_normal((self.R, k)).astype(np.float32)
        n = mlp_numel(k, self.L, self.H)
        self.mlp = (rng.uniform(-0.3, 0.3, size=max(n, 1))).astype(np.float32)
        self.bias = np.array([bias], dtype=np.float32)
        self.alpha = np.full(max(self.L, 1), 1.0 / (self.L + 1), dtype=np.float32)
        self.gA = np.zeros(self.R * (k + 1), dtype=np.float32)
        self.gB = np.zeros(self.R * (k + 1), dtype=np.float32)
        self.touched = np.zeros(self.R, dtype=np.uint8)
        self.m = _Model(KINDS[kind], self.F, k, self.L, self.H, self.R, batch_size, update_mode,
                        _fp(self.w1), _fp(self.V), _fp(self.mlp), _fp(self.bias), _fp(self.alpha),
                        lr, hb, hs, _fp(self.gA), _fp(self.gB), self.touched.ctypes.data_as(_u8), 0, 0,
                        None, None, None, None, None, 1.0, 0.0, 0.0)

    def enable_ftrl(self, beta=1.0, l1=0.0, l2=0.0):
        """update_mode 2: per-coordinate FTRL-Proximal with zero-initialised z/n state (FM-only steps)."""
        # n = 0 and z chosen so that the closed form w(z, n) reproduces the current weights (warm start)
        z0 = lambda w: (-(w * np.float32(np.float32(beta) / np.float32(self.m.lr) + np.float32(l2)))
                        - np.sign(w) * np.float32(l1)).astype(np.float32)
        self.fz_V = z0(self.V); self.fn_V = np.zeros_like(self.V)
        self.fz_w1 = z0(self.w1); self.fn_w1 = np.zeros_like(self.w1)
        self.f_bias = np.array([z0(self.bias)[0], 0.0], np.float32)
        m = self.m
        m.update_mode = 2
        m.fz_V, m.fn_V, m.fz_w1, m.fn_w1, m.f_bias = (_fp(self.fz_V), _fp(self.fn_V), _fp(self.fz_w1), _fp(self.fn_w1),
                                                      _fp(self.f_bias))
        m.f_beta, m.f_l1, m.f_l2 = beta, l1, l2

    def set_rank_partial_order(self, rank_B):
        """Sum duplicate rows in rank-partial order (csrc/shard2.cu): the batch is the concatenation of per-rank
        batches of rank_B samples; 0 = reference order."""
        self.m.rank_B = int(rank_B)

    def set_shard_order(self, G):
        """Evaluate forward sums in the G-way row-sharded order (csrc/sharded.cu); 0/1 = reference order."""
        self.m.shard_G = int(G)

    # -- helpers ---------------------------------------------------------
    def layer_views(self):
        """[(W_l [H,in], c_l [H])] views into the flat mlp buffer."""
        out, o = [], 0
        for l in range(self.L):
            nin = self.k if l == 0 else self.H
            W = self.mlp[o:o + self.H * nin].reshape(self.H, nin)
            o += self.H * nin
            c = self.mlp[o:o + self.H]
            o += self.H
            out.append((W, c))
        return out

    def global_ids(self, Xi):
        Xi = np.asarray(Xi, dtype=np.int64).reshape(-1, self.F)
        return np.ascontiguousarray(Xi + self.offsets[:-1][None, :]).astype(np.int32)

    def _prep(self, Xi, Xv):
        ids = self.global_ids(Xi)
        xv = np.ascontiguousarray(np.asarray(Xv, dtype=np.float32).reshape(-1, self.F))
        return ids, xv, ids.shape[0]

    # -- reference surface -------------------------------------------------
    def fm_parts(self, Xi, Xv):
        ids, xv, B = self._prep(Xi, Xv)
        first = np.empty((B, self.F), np.float32)
        S = np.empty((B, self.k), np.float32)
        bi = np.empty((B, self.k), np.float32)
        sf = np.empty(B, np.float32)
        sb = np.empty(B, np.float32)
        z = np.empty(B, np.float32)
        lib().orc_fm_forward(C.byref(self.m), _ip(ids), _fp(xv), B, _fp(first), _fp(S), _fp(bi), _fp(sf), _fp(sb),
                             _fp(z))
        return dict(first=first, S=S, bi=bi, sum_first=sf, sum_bi=sb, z_fm=z)

    def forward_fm(self, Xi, Xv):
        return self.fm_parts(Xi, Xv)["z_fm"]

    def forward(self, Xi, Xv):
        ids, xv, B = self._prep(Xi, Xv)
        z = np.empty(B, np.float32)
        pl = np.empty((max(self.L, 1), B), np.float32)
        lib().orc_forward(C.byref(self.m), _ip(ids), _fp(xv), B, _fp(z), _fp(pl))
        if self.kind.endswith("Onn"):
            return z, pl
        return z

    def predict(self, Xi, Xv):
        ids, xv, B = self._prep(Xi, Xv)
        p = np.empty(B, np.uint8)
        lib().orc_predict(C.byref(self.m), _ip(ids), _fp(xv), B, p.ctypes.data_as(_u8))
        return p.astype(bool)

    def update_embedding(self, Xi, Xv, Y):
        ids, xv, B = self._prep(Xi, Xv)
        y = np.ascontiguousarray(np.asarray(Y, dtype=np.float32).reshape(-1))
        return float(lib().orc_update_embedding(C.byref(self.m), _ip(ids), _fp(xv), _fp(y), B))

    def fit(self, Xi, Xv, Y):
        ids, xv, B = self._prep(Xi, Xv)
        y = np.ascontiguousarray(np.asarray(Y, dtype=np.float32).reshape(-1))
        lib().orc_fit(C.byref(self.m), _ip(ids), _fp(xv), _fp(y), B)

    def run_experiment(self, data_Xi, data_Xv, data_Y):
        ids, xv, N = self._prep(data_Xi, data_Xv)
        y = np.ascontiguousarray(np.asarray(data_Y, dtype=np.float32).reshape(-1))
        conf = (C.c_int64 * 4)()
        preds = np.empty(N, np.uint8)
        lib().orc_run_experiment(C.byref(self.m), _ip(ids), _fp(xv), _fp(y), N, conf, preds.ctypes.data_as(_u8))
        tp, fp, tn, fn = (int(c) for c in conf)
        return dict(tp=tp, fp=fp, tn=tn, fn=fn), preds.astype(bool)
