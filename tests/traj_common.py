"""Long-horizon trajectory configurations shared by tests/golden/make_trajectory.py (which runs the live
reference) and tests/test_trajectory.py (which replays the same batches through the oracle and the CUDA path).

Batches are regenerated from seeds (np.random.RandomState is a frozen generator), so the fixtures only hold the
reference's checkpoints: losses of every step, full parameters where they are small, a digest plus a sample of
rows where they are not.
"""
import hashlib

import numpy as np

# main_experiment.py:56-58: 39 fields (13 dense + 26 categorical), 1 006 628 rows
CRITEO_TINY = [63, 113, 126, 51, 224, 148, 100, 79, 104, 9, 32, 57, 82, 1457, 555, 176373, 129683, 305, 19, 11887,
               632, 3, 41738, 5170, 175446, 3170, 27, 11356, 165602, 10, 4641, 2030, 4, 172761, 18, 15, 57903, 86,
               44549]

# name -> dict(kind, ctor kwargs of the reference class, feature_sizes, B, lr, scale, method, steps, checkpoints)
TRAJ = {
    # BASELINE.json configs[0] shape: 10 000 steps, the north_star horizon
    "cfg1_ue": dict(kind="FMAdam", kw=dict(embedding_size=10), sizes=[943, 1682], B=256, lr=1e-3, scale=0.2,
                    method="update_embedding", steps=10000, ckpt=[100, 1000, 3000, 10000], seed=11),
    # FMAdam.fit: BCE-with-logits on an already sigmoided output (fm_adam.py:80), B not a multiple of 32 so the
    # scalar tail of torch.sigmoid is exercised every step
    "cfg1_fit": dict(kind="FMAdam", kw=dict(embedding_size=10), sizes=[943, 1682], B=250, lr=1e-3, scale=0.2,
                     method="fit", steps=3000, ckpt=[100, 1000, 3000], seed=12),
    # raw N(0,1) initialisation and the scripts' learning-rate scale: saturated logits, denormal gradients
    "cfg1_raw": dict(kind="FMAdam", kw=dict(embedding_size=10), sizes=[943, 1682], B=256, lr=1e-2, scale=None,
                     method="update_embedding", steps=3000, ckpt=[100, 1000, 3000], seed=13),
    # configs[2] shape (Frappe fields, k=64, B=4096); NFM's update_embedding uses the double-sigmoid loss
    "cfg3_ue": dict(kind="NFMAdam", kw=dict(embedding_size=64, num_hidden_layers=1, neuron_per_hidden_layer=64),
                    sizes=[957, 4082, 7, 7, 2, 3, 2, 9, 80, 233], B=4096, lr=1e-3, scale=0.1,
                    method="update_embedding", steps=1000, ckpt=[100, 1000], seed=14),
    # configs[3] shape (Criteo-tiny fields of main_experiment.py:56-58, k=10, B=8192), FM step of DeepFMAdam
    "cfg4_ue": dict(kind="DeepFMAdam", kw=dict(embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=400),
                    sizes=CRITEO_TINY, B=8192, lr=1e-3, scale=0.05, method="update_embedding", steps=1000,
                    ckpt=[100, 1000], seed=15),
}


def sizes_of(cfg):
    return list(cfg["sizes"])


def batch(cfg, step):
    """(Xi[B,F] int64 local ids, Xv[B,F] fp32 ones, Y[B] fp32) of step `step`."""
    sizes = sizes_of(cfg)
    rng = np.random.RandomState(cfg["seed"] * 100003 + step)
    B = cfg["B"]
    Xi = np.stack([rng.randint(0, fs, size=B) for fs in sizes], 1).astype(np.int64)
    # labels follow a fixed hashed "teacher" so that AUC / RMSE of the learner mean something
    f = np.arange(len(sizes), dtype=np.int64)[None, :]
    t = (((Xi * 2654435761 + f * 40503) % 1000).astype(np.float64) / 500.0 - 1.0).sum(1) * (1.5 / np.sqrt(len(sizes)))
    Y = (rng.uniform(size=B) < 1.0 / (1.0 + np.exp(-(t - 0.8)))).astype(np.float32)
    return Xi, np.ones(Xi.shape, np.float32), Y


EVAL_STEP = 10 ** 6   # batch(cfg, EVAL_STEP) is the held-out evaluation batch of a trajectory


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def sample_rows(R, n=1024, seed=7):
    """row indices stored in full for the big tables"""
    if R <= n:
        return np.arange(R)
    return np.sort(np.random.RandomState(seed).choice(R, size=n, replace=False))


def init_tables(cfg):
    """Deterministic initial (w1 [R], V [R,k]) of a trajectory: N(0,1) * scale from a numpy seed, so that big tables
    need not be stored.  (The reference's own torch-RNG initialisation order is covered by the A0 tests.)"""
    sizes = sizes_of(cfg)
    R, k = int(sum(sizes)), cfg["kw"]["embedding_size"]
    rng = np.random.RandomState(cfg["seed"] + 7919)
    sc = np.float32(1.0 if cfg["scale"] is None else cfg["scale"])
    w1 = rng.standard_normal(R).astype(np.float32) * sc
    V = rng.standard_normal((R, k)).astype(np.float32) * sc
    return w1, V
