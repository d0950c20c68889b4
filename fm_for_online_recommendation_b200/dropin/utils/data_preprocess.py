"""import shim: `from utils import data_preprocess` (main_experiment.py:1) resolves to the device-resident pipeline.

Same function names and arguments as the reference's utils/data_preprocess.py; the batch builders return the same 4-tuple
`(batch_train_Xi_list, batch_train_Xv_list, batch_train_Y_list, ratio_list)`, where every Xi entry is an `EncodedBatch`
(ids, values and labels already on the device; the model methods accept it in place of the list of lists), the Xv entry is the
same object and the Y entry is the host list of labels the scripts use for their accuracy print-outs
(main_experiment.py:100-105).  `read_*` return a dict with the reference's 'size', 'label', 'feature_sizes' entries plus
'dataset' (the DeviceDataset); the per-sample 'index' / 'value' lists of lists are what this pipeline exists to avoid.
"""
from fm_for_online_recommendation_b200 import data as _data


def _result(ds):
    return {"size": len(ds), "label": ds.labels_host().tolist(), "feature_sizes": list(ds.feature_sizes), "dataset": ds}


def read_criteo_data(file_path, emb_file):
    return _result(_data.read_criteo_data(file_path, emb_file))


def balance_criteo_data(file_path, emb_file):
    return _result(_data.balance(_data.read_criteo_data(file_path, emb_file)))


def read_svm_file(file_path, permutation=False):
    return _result(_data.read_svm_file(file_path, permutation))


def balance_svm_data(file_path):
    return _result(_data.balance(_data.read_svm_file(file_path)))


def _construct_batch_criteo_data(train_dict, num_batchdata, num_batch):
    ds = train_dict["dataset"] if isinstance(train_dict, dict) and "dataset" in train_dict else _data.DeviceDataset.from_result(train_dict)
    return _data.construct_batch_criteo_data(ds, num_batchdata, num_batch)


def create_ten_iter(file_path, emb_file, num_batch, num_batchdata):
    return _data.create_ten_iter(_data.read_criteo_data(file_path, emb_file), num_batch, num_batchdata)


def create_dataset(file_path, emb_file, batch_ratio, num_batch, num_batchdata):
    return _data.create_dataset(_data.read_criteo_data(file_path, emb_file), batch_ratio, num_batch, num_batchdata)
