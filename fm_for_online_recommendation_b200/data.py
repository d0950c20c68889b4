"""Device-resident input pipeline: the data set is encoded ONCE on the device and the batch builders
of the reference's utils/data_preprocess.py hand out `EncodedBatch` objects instead of Python lists of lists.

What it replaces (file:line relative to the reference root):
  * the per-call `torch.LongTensor(Xi)` / `FloatTensor(Xv)` list conversions of every model method
    (models/models_online_deep/deepfm_adam.py:47-48,57-58) -- a `DeviceDataset` batch is already what the kernels read;
  * `_construct_batch_criteo_data` (utils/data_preprocess.py:154-180), `create_ten_iter` (:193-229), `create_dataset`
    (:232-262), `balance_criteo_data` / `balance_svm_data` (:46-82, :120-151): contiguous batches are zero-copy views, shuffled
    or re-balanced ones are ONE row-gather launch (`fmb_dataset_take`, csrc/dataset.cu) driven by the index list;
  * `read_svm_file`'s vocabulary build (:96-108; `list.index` per cell, O(N * vocabulary)) -- `fmb_dict_encode_first_seen`,
    a per-column hash table: O(N).

Index lists are shuffled on the host with Python's `random` module in exactly the reference's call order, so with the same
`random.seed` the batches hold the same samples in the same order as the reference's (tests/test_dataset.py, fixtures generated
from the reference's own functions).  There is no CPU path for the row movement or the encoding.
"""
import ctypes as C
import random as _random

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr
from .deep import EncodedBatch


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class DeviceDataset:
    """ids int32 [N,F] GLOBAL row ids (field offset + per-field id), xv fp32 [N,F] or None (all ones: what the Criteo loader
    produces, data_preprocess.py:41), y fp32 [N] or None -- all on the device; `feature_sizes` as in the reference's result
    dict.  `result['size']` of the reference is `len(ds)`."""

    def __init__(self, ids, xv, y, feature_sizes):
        self.ids, self.xv, self.y = ids, xv, y
        self.feature_sizes = tuple(int(s) for s in feature_sizes)
        self.device = ids.device
        self.field_size = ids.shape[1]
        self._labels_host = None

    def __len__(self):
        return self.ids.shape[0]

    @property
    def size(self):
        return len(self)

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_local_ids(cls, index, value, label, feature_sizes, device="cuda"):
        """`index` [N,F] per-field ids (list of lists / ndarray / tensor), `value` [N,F] or None, `label` [N] or None: the
        'index' / 'value' / 'label' entries of the reference's result dict (data_preprocess.py:28-45, :87-117).  Ids outside
        [0, feature_sizes[f]) raise IndexError like nn.Embedding in the reference."""
        lib = _lib.require_cuda()
        dev = torch.device(device)
        sizes = np.asarray(feature_sizes, dtype=np.int64).reshape(-1)
        F = sizes.size
        off = np.zeros(F + 1, np.int64)
        np.cumsum(sizes, out=off[1:])
        if off[-1] >= 2 ** 31:
            raise ValueError("more than 2^31 rows")
        if torch.is_tensor(index):
            loc = index.to(dev, torch.int64).reshape(-1, F).contiguous()
        else:
            loc = torch.from_numpy(np.ascontiguousarray(np.asarray(index, dtype=np.int64).reshape(-1, F))).to(dev)
        n = loc.shape[0]
        ids = torch.empty((n, F), dtype=torch.int32, device=dev)
        err = torch.zeros(1, dtype=torch.int32, device=dev)
        off_dev = torch.from_numpy(off.astype(np.int32)).to(dev)
        check(lib.fmb_dataset_encode_ids(ptr(loc), n, F, ptr(off_dev), ptr(ids), ptr(err), _stream()), "fmb_dataset_encode_ids")
        if int(err.item()):
            raise IndexError("index out of range in self")
        xv = None
        if value is not None:
            if torch.is_tensor(value):
                v = value.to(dev, torch.float32).reshape(-1, F).contiguous()
                if not bool((v == 1.0).all()):
                    xv = v
            else:
                v = np.asarray(value, dtype=np.float32).reshape(-1, F)
                if not np.all(v == 1.0):
                    xv = torch.from_numpy(np.ascontiguousarray(v)).to(dev)
        y = None
        if label is not None:
            if torch.is_tensor(label):
                y = label.to(dev, torch.float32).reshape(-1).contiguous()
            else:
                y = torch.from_numpy(np.asarray(label, dtype=np.float32).reshape(-1)).to(dev)
        return cls(ids, xv, y, sizes)

    @classmethod
    def from_result(cls, result, device="cuda"):
        """the dict returned by the reference's read_criteo_data / read_svm_file / balance_* functions"""
        return cls.from_local_ids(result["index"], result["value"], result["label"], result["feature_sizes"], device)

    @classmethod
    def from_svm_matrix(cls, X, y, device="cuda"):
        """read_svm_file after `load_svmlight_file` (data_preprocess.py:96-117): X [N,d] real values, y in {-1, +1} (or
        {0, 1}).  Every column is dictionary-encoded in order of first appearance on the device; 'value' keeps the raw
        features, labels -1 become 0."""
        codes, sizes = dict_encode_first_seen(X, device)
        dev = codes.device
        Xd = _as_f64(X, dev)
        yy = _as_f64(y, dev).reshape(-1).to(torch.int64)          # `.astype(int)` of the reference truncates
        yy = torch.where(yy == -1, torch.zeros_like(yy), yy)
        return cls.from_local_ids(codes, Xd.to(torch.float32), yy.to(torch.float32), sizes, dev)

    # ------------------------------------------------------------------ batches
    def _encoded(self, ids, xv, y):
        e = EncodedBatch(ids, xv, y)
        e.feature_sizes = self.feature_sizes
        return e

    def batch(self, lo, hi):
        """rows [lo, hi) as an EncodedBatch: views, no copy (the inner loop of _construct_batch_criteo_data,
        data_preprocess.py:165-168)"""
        if lo < 0 or hi > len(self) or lo > hi:
            raise IndexError("list index out of range")
        return self._encoded(self.ids[lo:hi], None if self.xv is None else self.xv[lo:hi], None if self.y is None else self.y[lo:hi])

    def take(self, indices, count_positives=False):
        """rows `indices` (host list / ndarray / device int64 tensor), in that order, as ONE gather launch"""
        lib = _lib.require_cuda()
        if torch.is_tensor(indices):
            idx = indices.to(self.device, torch.int64).reshape(-1).contiguous()
        else:
            idx = torch.from_numpy(np.asarray(indices, dtype=np.int64).reshape(-1)).to(self.device)
        n, F = idx.numel(), self.field_size
        ids = torch.empty((n, F), dtype=torch.int32, device=self.device)
        xv = None if self.xv is None else torch.empty((n, F), dtype=torch.float32, device=self.device)
        y = None if self.y is None else torch.empty(n, dtype=torch.float32, device=self.device)
        st = torch.zeros(2, dtype=torch.int32, device=self.device)      # [positives, error]
        check(lib.fmb_dataset_take(ptr(self.ids), ptr(self.xv), ptr(self.y), F, len(self), ptr(idx), n, ptr(ids), ptr(xv),
                                   ptr(y), ptr(st), ptr(st[1:]), _stream()), "fmb_dataset_take")
        pos, err = (int(v) for v in st.cpu())
        if err:
            raise IndexError("list index out of range")
        e = self._encoded(ids, xv, y)
        return (e, pos) if count_positives else e

    def labels_host(self):
        """labels as a host int array (N small integers: the one thing the index bookkeeping needs on the host)"""
        if self._labels_host is None:
            self._labels_host = self.y.cpu().numpy().astype(np.int64)
        return self._labels_host

    def select(self, indices):
        """a new DeviceDataset of rows `indices` (balance_* functions)"""
        e = self.take(indices)
        return DeviceDataset(e.ids, e.xv, e.y, self.feature_sizes)


def _as_f64(a, dev):
    if torch.is_tensor(a):
        return a.to(dev, torch.float64).contiguous()
    if hasattr(a, "toarray"):
        a = a.toarray()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(dev)


def dict_encode_first_seen(X, device="cuda"):
    """codes int32 [N,d] (device) and sizes int64 [d] (host): column c's code of X[i,c] is its index in the list of the
    column's distinct values in order of first appearance -- read_svm_file's `feature_sizes[...].index(emb)` loop
    (data_preprocess.py:100-108).  NaNs raise ValueError."""
    lib = _lib.require_cuda()
    dev = torch.device(device)
    Xd = _as_f64(X, dev)
    if Xd.dim() != 2:
        raise ValueError("X must be [N, d]")
    N, d = Xd.shape
    codes = torch.empty((N, d), dtype=torch.int32, device=dev)
    sizes = torch.zeros(d, dtype=torch.int32, device=dev)
    if N == 0:
        return codes, np.zeros(d, np.int64)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    wsb = lib.fmb_dict_encode_workspace_bytes(N, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    check(lib.fmb_dict_encode_first_seen(ptr(Xd), N, d, ptr(codes), ptr(sizes), ptr(err), ptr(ws), wsb, _stream()),
          "fmb_dict_encode_first_seen")
    if int(err.item()):
        raise ValueError("NaN in the feature matrix")
    return codes, sizes.cpu().numpy().astype(np.int64)


# ---------------------------------------------------------------------- the reference's batch builders
def _split_by_label(ds):
    """_find_pos_and_neg (data_preprocess.py:183-190): row numbers of the negatives and of the positives, ascending"""
    lab = ds.labels_host()
    return {"0": np.nonzero(lab == 0)[0].tolist(), "1": np.nonzero(lab == 1)[0].tolist()}


def construct_batch_criteo_data(ds, num_batchdata, num_batch):
    """_construct_batch_criteo_data (data_preprocess.py:154-180): `num_batch` consecutive batches of `num_batchdata` rows.
    Returns (Xi_list, Xv_list, Y_list, ratio_list): Xi_list[i] is the EncodedBatch (ids, values and labels on the device),
    Xv_list[i] the same object (model methods ignore Xv / Y when Xi is an EncodedBatch), Y_list[i] the labels as a host list."""
    if num_batch * num_batchdata > len(ds):
        raise IndexError("list index out of range")
    lab = ds.labels_host()
    Xi, Y, ratios = [], [], []
    for i in range(num_batch):
        lo, hi = i * num_batchdata, (i + 1) * num_batchdata
        Xi.append(ds.batch(lo, hi))
        yl = lab[lo:hi].tolist()
        pos = sum(yl)
        ratios.append((len(yl) - pos, pos))
        Y.append(yl)
    return Xi, list(Xi), Y, ratios


def _gathered(ds, index_lists):
    lab = ds.labels_host()
    Xi = [ds.take(ix) for ix in index_lists]
    Y = [lab[np.asarray(ix, dtype=np.int64)].tolist() if len(ix) else [] for ix in index_lists]
    return Xi, list(Xi), Y


def create_ten_iter(ds, num_batch, num_batchdata, rng=_random):
    """create_ten_iter (data_preprocess.py:193-229): batch i holds int(num_batchdata / num_batch * (i + 1)) positives (taken
    from the front of the remaining positives) and the rest negatives, shuffled with `random.shuffle`."""
    ratios = _split_by_label(ds)
    lists, ratio_list = [], []
    for i in range(num_batch):
        num_pos = int(num_batchdata / num_batch * (i + 1))
        num_neg = num_batchdata - num_pos
        ratio_list.append((num_neg, num_pos))
        indices = ratios["1"][:num_pos] + ratios["0"][:num_neg]
        ratios["1"] = ratios["1"][num_pos:]
        ratios["0"] = ratios["0"][num_neg:]
        rng.shuffle(indices)
        lists.append(indices)
    return (*_gathered(ds, lists), ratio_list)


def create_dataset(ds, batch_ratio, num_batch, num_batchdata, rng=_random):
    """create_dataset (data_preprocess.py:232-262): every batch holds int(num_batchdata / num_batch * batch_ratio) positives"""
    ratios = _split_by_label(ds)
    lists, ratio_list = [], []
    for _ in range(num_batch):
        ratio_list.append((batch_ratio, num_batch - batch_ratio))
        num_pos = int(num_batchdata / num_batch * batch_ratio)
        num_neg = num_batchdata - num_pos
        indices = ratios["1"][:num_pos] + ratios["0"][:num_neg]
        ratios["1"] = ratios["1"][num_pos:]
        ratios["0"] = ratios["0"][num_neg:]
        rng.shuffle(indices)
        lists.append(indices)
    return (*_gathered(ds, lists), ratio_list)


def balance(ds, rng=_random):
    """balance_criteo_data / balance_svm_data (data_preprocess.py:46-82, :120-151): as many negatives as positives, shuffled"""
    idx = _split_by_label(ds)
    rng.shuffle(idx["0"])
    idx["0"] = idx["0"][:len(idx["1"])]
    idx["0"].extend(idx["1"])
    rng.shuffle(idx["0"])
    return ds.select(idx["0"])


# ---------------------------------------------------------------------- file readers (host IO, then one upload)
def read_criteo_data(file_path, emb_file, device="cuda"):
    """read_criteo_data (data_preprocess.py:28-45) as a DeviceDataset: rows `label,idx_0,...,idx_38`; feature_sizes = number of
    categories per field in `emb_file` (rows `field,category,index`, load_criteo_category_index :15-26)."""
    cats = [set() for _ in range(39)]
    with open(emb_file, "r") as f:
        for line in f:
            d = line.strip().split(",")
            cats[int(d[0])].add(d[1])
    sizes = [len(c) for c in cats]
    a = np.loadtxt(file_path, delimiter=",", dtype=np.int64, ndmin=2)
    return DeviceDataset.from_local_ids(a[:, 1:], None, a[:, 0], sizes, device)


def read_svm_file(file_path, permutation=False, device="cuda"):
    """read_svm_file (data_preprocess.py:87-117) as a DeviceDataset (the reference fixes n_features=8: cod-rna)"""
    from sklearn.datasets import load_svmlight_file
    X, y = load_svmlight_file(file_path, n_features=8)
    X = X.toarray()
    if permutation:
        idx = np.random.permutation(X.shape[0])
        X, y = np.asarray(X[idx]), np.asarray(y[idx])
    return DeviceDataset.from_svm_matrix(X, y, device)
