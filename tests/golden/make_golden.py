"""Generate golden fixtures from the REFERENCE itself (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports /root/reference read-only (haan6/fm-for-online-recommendation), runs its own classes on
small seeded synthetic inputs with use_cuda=False (CPU oracle, the bias is
trainable there) and stores inputs, initial parameters and every output in tests/golden/*.npz.
/root/reference does not exist on the GPU box, so tests only ever read the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.dont_write_bytecode = True
    if "matplotlib" not in sys.modules:  # models_online imports matplotlib only to call use('Agg')
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    sys.path.insert(0, REF)


def flat_params(model):
    w1 = np.concatenate([e.weight.detach().numpy()[:, 0] for e in model.first_order_embeddings]).astype(np.float32)
    V = np.concatenate([e.weight.detach().numpy() for e in model.second_order_embeddings]).astype(np.float32)
    parts = []
    for lin in getattr(model, "hidden_layers", []):
        parts.append(lin.weight.detach().numpy().reshape(-1))
        parts.append(lin.bias.detach().numpy().reshape(-1))
    mlp = np.concatenate(parts).astype(np.float32) if parts else np.zeros(1, np.float32)
    bias = np.array(model.bias.detach().numpy(), dtype=np.float32, copy=True).reshape(1)
    out = dict(w1=w1, V=V, mlp=mlp, bias=bias)
    if hasattr(model, "alpha"):
        out["alpha"] = np.array(model.alpha.detach().numpy(), dtype=np.float32, copy=True)
    return out


def synth(feature_sizes, n, seed, real_xv=False, zipf=False):
    rng = np.random.RandomState(seed)
    if zipf:
        Xi = np.stack([np.minimum(rng.zipf(1.3, size=n) - 1, fs - 1) for fs in feature_sizes], 1)
    else:
        Xi = np.stack([rng.randint(0, fs, size=n) for fs in feature_sizes], 1)
    Xv = rng.uniform(0.25, 2.0, size=Xi.shape).astype(np.float32) if real_xv else np.ones(Xi.shape, np.float32)
    Y = (rng.uniform(size=n) < 0.3).astype(np.float32)
    return Xi.astype(np.int64), Xv, Y


def record(name, cls, ctor_kwargs, feature_sizes, B, steps, seed, real_xv=False, zipf=False, online_n=0,
           scale=None):
    torch.manual_seed(seed)
    torch.set_num_threads(1)
    model = cls(feature_sizes, use_cuda=False, **ctor_kwargs)
    if scale is not None:  # shrink N(0,1) embeddings so logits do not saturate the sigmoid
        with torch.no_grad():
            for e in list(model.first_order_embeddings) + list(model.second_order_embeddings):
                e.weight.mul_(scale)
    out = {"feature_sizes": np.asarray(feature_sizes, np.int64), "B": B, "steps": steps, "seed": seed,
           "lr": np.float32(ctor_kwargs.get("n", 0.01))}
    for key, v in flat_params(model).items():
        out["init_" + key] = v
    is_onn = name.endswith("Onn") or "Onn" in cls.__name__
    Xi, Xv, Y = synth(feature_sizes, B, seed + 1, real_xv, zipf)
    out.update(Xi=Xi, Xv=Xv, Y=Y)
    with torch.no_grad():
        fwd = model.forward(Xi.tolist(), Xv.tolist())
        if is_onn:
            out["fwd0"] = fwd[0].numpy()
            out["fwd0_layers"] = fwd[1].numpy()
        else:
            out["fwd0"] = fwd.numpy()
        if hasattr(model, "forward_fm"):
            out["fwd_fm0"] = model.forward_fm(Xi.tolist(), Xv.tolist()).numpy()
            out["first0"] = model.first_order(Xi.tolist(), Xv.tolist()).numpy()
            out["second0"] = model.second_order(Xi.tolist(), Xv.tolist()).numpy()
    out["pred0"] = np.asarray(model.predict(Xi.tolist(), Xv.tolist()))
    # batch update_embedding steps (a fresh batch per step, like cfg1's mini-batches)
    losses = []
    ue_batches = []
    for s in range(steps):
        bXi, bXv, bY = synth(feature_sizes, B, seed + 100 + s, real_xv, zipf)
        ue_batches.append((bXi, bXv, bY))
        loss = model.update_embedding(bXi.tolist(), bXv.tolist(), bY.tolist())
        losses.append(float(loss.detach()))
    out["ue_Xi"] = np.stack([b[0] for b in ue_batches])
    out["ue_Xv"] = np.stack([b[1] for b in ue_batches])
    out["ue_Y"] = np.stack([b[2] for b in ue_batches])
    out["ue_loss"] = np.asarray(losses, np.float32)
    for key, v in flat_params(model).items():
        out["after_ue_" + key] = v
    # batch fit steps (Adam family: any B; ONN: B must equal batch_size)
    fitB = B if not is_onn else ctor_kwargs.get("batch_size", 1)
    fit_batches = []
    for s in range(steps):
        bXi, bXv, bY = synth(feature_sizes, fitB, seed + 200 + s, real_xv, zipf)
        fit_batches.append((bXi, bXv, bY))
        model.fit(bXi.tolist(), bXv.tolist(), bY.tolist())
    out["fit_Xi"] = np.stack([b[0] for b in fit_batches])
    out["fit_Xv"] = np.stack([b[1] for b in fit_batches])
    out["fit_Y"] = np.stack([b[2] for b in fit_batches])
    for key, v in flat_params(model).items():
        out["after_fit_" + key] = v
    with torch.no_grad():
        fwd = model.forward(Xi.tolist(), Xv.tolist())
        out["fwd1"] = (fwd[0] if is_onn else fwd).numpy()
    if online_n:
        oXi, oXv, oY = synth(feature_sizes, online_n, seed + 300, real_xv, zipf)
        _, acc, roc, conf = model.run_experiment(oXi.tolist(), oXv.tolist(), [int(v) for v in oY])
        out.update(on_Xi=oXi, on_Xv=oXv, on_Y=oY, on_acc=np.float64(acc),
                   on_conf=np.asarray([conf["tp"], conf["fp"], conf["tn"], conf["fn"]], np.int64))
        for key, v in flat_params(model).items():
            out["after_on_" + key] = v
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith("after_fit")})


def main():
    _import_reference()
    from models.models_online_deep.fm_adam import FMAdam
    from models.models_online_deep.deepfm_adam import DeepFMAdam
    from models.models_online_deep.nfm_adam import NFMAdam
    from models.models_online_deep.deepfm_onn import DeepFMOnn
    from models.models_online_deep.nfm_onn import NFMOnn

    small = [7, 5, 11, 3, 13, 4]
    frappe_s = [9, 40, 7, 7, 2, 3, 2, 9, 8, 23]
    # raw N(0,1) init (saturated logits, the reference's real regime) and a scaled variant (live gradients)
    record("fm_cfg1", FMAdam, dict(embedding_size=10, n=0.01), [943, 1682], 256, 4, 0)
    record("fm_small_scaled", FMAdam, dict(embedding_size=10, n=0.001), small, 64, 6, 1, real_xv=True, scale=0.2,
           online_n=40)
    record("deepfm_small", DeepFMAdam, dict(embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=16,
                                            n=0.001), small, 64, 4, 2, real_xv=True, scale=0.2, online_n=30)
    record("deepfm_raw", DeepFMAdam, dict(embedding_size=10, num_hidden_layers=2, neuron_per_hidden_layer=8,
                                          n=0.0001), small, 50, 3, 3, zipf=True)
    record("nfm_k64", NFMAdam, dict(embedding_size=64, num_hidden_layers=1, neuron_per_hidden_layer=64, n=0.001),
           frappe_s, 32, 4, 4, scale=0.1, online_n=20)
    record("deepfm_onn", DeepFMOnn, dict(embedding_size=10, num_hidden_layers=5, neuron_per_hidden_layer=10,
                                         n=0.0001, batch_size=1), small, 50, 4, 5, scale=0.2, online_n=40)
    record("nfm_onn", NFMOnn, dict(embedding_size=10, num_hidden_layers=5, neuron_per_hidden_layer=10, n=0.0001,
                                   batch_size=1), small, 50, 4, 6, scale=0.2, online_n=40)
    record("nfm_onn_b8", NFMOnn, dict(embedding_size=10, num_hidden_layers=3, neuron_per_hidden_layer=10, n=0.01,
                                      batch_size=8), small, 8, 4, 7, scale=0.2)


if __name__ == "__main__":
    main()
