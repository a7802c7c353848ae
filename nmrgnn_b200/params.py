"""Host-side parameter container for the GNN forward path.

Mirrors what the reference keeps in its Keras layers:
``EdgeFCBlock.edge_fc`` (nmrgnn/model.py:118-128), ``GNNModel.embed_layer`` /
``out_layer`` (model.py:239-241), ``MPLayer.w`` (nmrgnn/layers.py:11-18),
``FCBlock.fc`` (model.py:184-188), the RBF grid (layers.py:126-129) and the
peak standardisation constants (model.py:222-228, 242-243).
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

ACTIVATIONS = {"linear": 0, "softplus": 1, "relu": 2, "tanh": 3}

# Baked into the reference's traced graph as Const nodes gnn-model/mul_3/y and
# gnn-model/mul_4/y; the bit patterns are the float32 values.
_BASELINE_STD_BITS = {2: 0x4129A7C9, 3: 0x424BC3CE, 4: 0x40C14CF5}
_BASELINE_AVG_BITS = {2: 0x42FC0000, 3: 0x42EDE8F6, 4: 0x40B428F6}


def baseline_standards(num_elem: int = 10) -> Tuple[np.ndarray, np.ndarray]:
    std = np.zeros(num_elem, np.float32)
    avg = np.zeros(num_elem, np.float32)
    for k, bits in _BASELINE_STD_BITS.items():
        std[k] = np.array([bits], np.uint32).view(np.float32)[0]
    for k, bits in _BASELINE_AVG_BITS.items():
        avg[k] = np.array([bits], np.uint32).view(np.float32)[0]
    return std, avg


def rbf_centers(low: float, high: float, count: int) -> Tuple[np.ndarray, np.float32]:
    """``tf.cast(tf.linspace(low, high, count), tf.float32)`` evaluated the way TF
    does for float32 inputs (start + delta*i in float32) so the grid matches the
    Const baked in the SavedModel bit for bit; gap = centers[1]-centers[0]
    (nmrgnn/layers.py:126-129)."""
    lo = np.float32(low)
    hi = np.float32(high)
    if count == 1:
        c = np.array([lo], np.float32)
        return c, np.float32(0)
    step = np.float32((hi - lo) / np.float32(count - 1))
    c = (lo + step * np.arange(count, dtype=np.float32)).astype(np.float32)
    c[-1] = hi
    return c, np.float32(c[1] - c[0])


@dataclass
class GNNParams:
    edge_fc: List[Tuple[np.ndarray, np.ndarray]]          # [(W[in,out], b[out])], last is linear
    embed: np.ndarray                                      # W_e[C, F]
    mp_w: List[np.ndarray]                                 # each w[F, F, E]  (l, m, n)
    fc: List[Tuple[np.ndarray, np.ndarray]]                # [(W, b)], last maps F -> F//2
    out: Tuple[np.ndarray, np.ndarray]                     # W_o[F2, C], b_o[C]
    peak_std: np.ndarray                                   # [C]
    peak_avg: np.ndarray                                   # [C]
    rbf_low: float = 0.005
    rbf_high: float = 0.20
    mp_activation: str = "softplus"
    fc_activation: str = "softplus"
    meta: Dict[str, object] = field(default_factory=dict)

    # ----------------------------------------------------------------- shapes
    @property
    def num_elem(self) -> int:
        return int(self.embed.shape[0])

    @property
    def atom_feature_size(self) -> int:
        return int(self.embed.shape[1])

    @property
    def edge_feature_size(self) -> int:
        return int(self.edge_fc[-1][0].shape[1])

    @property
    def edge_hidden_size(self) -> int:
        return int(self.edge_fc[0][0].shape[0])

    @property
    def rbf_count(self) -> int:
        return self.edge_hidden_size

    def validate(self) -> None:
        F, E, H, C = self.atom_feature_size, self.edge_feature_size, self.edge_hidden_size, self.num_elem
        prev = H
        for i, (W, b) in enumerate(self.edge_fc):
            if W.shape[0] != prev or b.shape != (W.shape[1],):
                raise ValueError(f"edge_fc[{i}] has inconsistent shape {W.shape}/{b.shape}")
            prev = W.shape[1]
        for i, w in enumerate(self.mp_w):
            if w.shape != (F, F, E):
                raise ValueError(f"mp_w[{i}] shape {w.shape} != {(F, F, E)}")
        prev = F
        for i, (W, b) in enumerate(self.fc):
            if W.shape[0] != prev or b.shape != (W.shape[1],):
                raise ValueError(f"fc[{i}] has inconsistent shape {W.shape}/{b.shape}")
            if i < len(self.fc) - 1 and W.shape[1] != F:
                raise ValueError("residual FC layers must preserve the feature width")
            prev = W.shape[1]
        Wo, bo = self.out
        if Wo.shape != (prev, C) or bo.shape != (C,):
            raise ValueError(f"out layer shape {Wo.shape}/{bo.shape} inconsistent")
        if self.peak_std.shape != (C,) or self.peak_avg.shape != (C,):
            raise ValueError("peak standards must have one entry per element")
        for name in (self.mp_activation, self.fc_activation):
            if name not in ACTIVATIONS:
                raise ValueError(f"unknown activation {name!r}")

    def astype(self, dtype) -> "GNNParams":
        c = lambda a: np.asarray(a, dtype)
        return GNNParams([(c(W), c(b)) for W, b in self.edge_fc], c(self.embed), [c(w) for w in self.mp_w],
                         [(c(W), c(b)) for W, b in self.fc], (c(self.out[0]), c(self.out[1])),
                         c(self.peak_std), c(self.peak_avg), self.rbf_low, self.rbf_high,
                         self.mp_activation, self.fc_activation, dict(self.meta))

    # ------------------------------------------------------------------- I/O
    def save(self, path: str) -> None:
        arrs: Dict[str, np.ndarray] = {"embed": self.embed, "out_w": self.out[0], "out_b": self.out[1],
                                       "peak_std": self.peak_std, "peak_avg": self.peak_avg}
        for i, (W, b) in enumerate(self.edge_fc):
            arrs[f"edge_fc_{i}_w"], arrs[f"edge_fc_{i}_b"] = W, b
        for i, w in enumerate(self.mp_w):
            arrs[f"mp_{i}_w"] = w
        for i, (W, b) in enumerate(self.fc):
            arrs[f"fc_{i}_w"], arrs[f"fc_{i}_b"] = W, b
        meta = dict(self.meta)
        meta.update(format="nmrgnn_b200.params.v1", n_edge_fc=len(self.edge_fc), n_mp=len(self.mp_w),
                    n_fc=len(self.fc), rbf_low=self.rbf_low, rbf_high=self.rbf_high,
                    mp_activation=self.mp_activation, fc_activation=self.fc_activation)
        arrs["meta_json"] = np.frombuffer(json.dumps(meta, sort_keys=True).encode(), np.uint8)
        np.savez(path, **arrs)

    @classmethod
    def load(cls, path: str) -> "GNNParams":
        with np.load(path) as z:
            meta = json.loads(bytes(z["meta_json"]).decode())
            p = cls(
                edge_fc=[(z[f"edge_fc_{i}_w"], z[f"edge_fc_{i}_b"]) for i in range(meta["n_edge_fc"])],
                embed=z["embed"],
                mp_w=[z[f"mp_{i}_w"] for i in range(meta["n_mp"])],
                fc=[(z[f"fc_{i}_w"], z[f"fc_{i}_b"]) for i in range(meta["n_fc"])],
                out=(z["out_w"], z["out_b"]),
                peak_std=z["peak_std"], peak_avg=z["peak_avg"],
                rbf_low=meta["rbf_low"], rbf_high=meta["rbf_high"],
                mp_activation=meta["mp_activation"], fc_activation=meta["fc_activation"], meta=meta)
        p.validate()
        return p

    @classmethod
    def from_tf_checkpoint(cls, path: str, peak_std: Optional[np.ndarray] = None,
                           peak_avg: Optional[np.ndarray] = None, edge_fc_layers: Optional[int] = None,
                           rbf_low: Optional[float] = None, rbf_high: Optional[float] = None,
                           mp_activation: Optional[str] = None, fc_activation: Optional[str] = None) -> "GNNParams":
        """Build from a TensorBundle written by the reference (``model.save`` /
        ``ModelCheckpoint``; nmrgnn/main.py:63-68,82).  ``variables/<i>`` are the
        model's sub-layer weights in creation order: edge FC (kernel,bias)*, MP w*,
        FC (kernel,bias)* — the split is recovered from tensor ranks/shapes.

        What the weights do not say — activations, RBF range, per-element standards — is taken, in this order, from
        the keyword arguments, from the SavedModel's ``saved_model.pb`` (Keras layer metadata and the constants baked
        into the traced graph, as ``tf.keras.models.load_model`` restores them; savedmodel_meta.py), or, for the
        standards of a bare ``ModelCheckpoint`` only, from the packaged standards (the reference builds the model
        with ``nmrdata.load_standards()`` before ``load_weights``, nmrgnn/model.py:222-228).  Activations and RBF
        range of a bare checkpoint must be given: the reference needs the same ``hp`` to rebuild the model, and a
        silent default would predict wrong values for a relu / tanh model."""
        from .savedmodel_meta import read_savedmodel_meta, standards_from_constants
        meta = None
        for d in (path, os.path.dirname(os.path.normpath(path))):
            if os.path.isfile(os.path.join(d, "saved_model.pb")):
                meta = read_savedmodel_meta(d)
                break
        hp = (meta or {}).get("hypers") or {}
        rbf = (meta or {}).get("rbf") or {}
        if mp_activation is None:
            mp_activation = hp.get("mp_activation")
        if fc_activation is None:
            fc_activation = hp.get("fc_activation")
        if rbf_low is None:
            rbf_low = rbf.get("low", hp.get("rbf_low"))
        if rbf_high is None:
            rbf_high = rbf.get("high", hp.get("rbf_high"))
        missing = [k for k, v in (("mp_activation", mp_activation), ("fc_activation", fc_activation),
                                  ("rbf_low", rbf_low), ("rbf_high", rbf_high)) if v is None]
        if missing:
            raise ValueError(f"{path}: {', '.join(missing)} cannot be recovered "
                             f"({'saved_model.pb has no Keras metadata for them' if meta else 'no saved_model.pb next to the checkpoint'}); "
                             "pass them as keyword arguments (the hyper-parameters the model was trained with)")
        from .tensorbundle import load_gnn_variables

        v = load_gnn_variables(path)
        n = 0
        while f"variables/{n}" in v:
            n += 1
        seq = [v[f"variables/{i}"] for i in range(n)]
        mp_idx = [i for i, a in enumerate(seq) if a.ndim == 3]
        if not mp_idx or mp_idx != list(range(mp_idx[0], mp_idx[-1] + 1)):
            raise ValueError("checkpoint does not look like a GNNModel (no contiguous rank-3 MP weights)")
        first, last = mp_idx[0], mp_idx[-1]
        if first % 2 or (n - last - 1) % 2:
            raise ValueError("unexpected (kernel,bias) pairing in checkpoint")
        edge_fc = [(seq[i], seq[i + 1]) for i in range(0, first, 2)]
        fc = [(seq[i], seq[i + 1]) for i in range(last + 1, n, 2)]
        if edge_fc_layers is not None and len(edge_fc) != edge_fc_layers:
            raise ValueError("edge_fc_layers mismatch")
        C = v["embed_layer/kernel"].shape[0]
        if (peak_std is None or peak_avg is None) and meta is not None:
            baked = standards_from_constants(meta["constants"], C)
            if baked is None:
                raise ValueError(f"{path}: the per-element standards baked into the traced graph (gnn-model/mul_3/y, "
                                 "mul_4/y) were not found; pass peak_std / peak_avg")
            peak_std, peak_avg = baked
        if peak_std is None or peak_avg is None:
            peak_std, peak_avg = baseline_standards(C)
        if rbf.get("count") is not None and int(rbf["count"]) != edge_fc[0][0].shape[0]:
            raise ValueError("RBFExpansion count in saved_model.pb does not match the first edge layer's input width")
        p = cls(edge_fc=edge_fc, embed=v["embed_layer/kernel"], mp_w=seq[first:last + 1], fc=fc,
                out=(v["out_layer/kernel"], v["out_layer/bias"]),
                peak_std=np.asarray(peak_std, np.float32), peak_avg=np.asarray(peak_avg, np.float32),
                rbf_low=float(rbf_low), rbf_high=float(rbf_high), mp_activation=str(mp_activation),
                fc_activation=str(fc_activation), meta={"source": os.path.basename(os.path.normpath(path)),
                                                        "hypers_from": "saved_model.pb" if hp or rbf else "arguments"})
        p.validate()
        return p

    @classmethod
    def random(cls, num_elem: int = 16, atom_feature_size: int = 256, edge_feature_size: int = 3,
               edge_hidden_size: int = 128, mp_layers: int = 4, fc_layers: int = 4, edge_fc_layers: int = 4,
               mp_activation: str = "softplus", fc_activation: str = "softplus",
               rbf_low: float = 0.005, rbf_high: float = 0.20, seed: int = 0,
               peak_std: Optional[np.ndarray] = None, peak_avg: Optional[np.ndarray] = None) -> "GNNParams":
        """Freshly initialised model with the reference's hyper-parameter names and
        defaults (nmrgnn/model.py:22-36); Glorot-uniform kernels and zero biases as
        Keras Dense / add_weight do by default."""
        rng = np.random.default_rng(seed)

        def glorot(shape, fan_in, fan_out):
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            return rng.uniform(-lim, lim, size=shape).astype(np.float32)

        F, E, H, C = atom_feature_size, edge_feature_size, edge_hidden_size, num_elem
        edge_fc = []
        prev = H
        for _ in range(edge_fc_layers - 1):
            edge_fc.append((glorot((prev, H), prev, H), np.zeros(H, np.float32)))
            prev = H
        edge_fc.append((glorot((prev, E), prev, E), np.zeros(E, np.float32)))
        # Keras' glorot on a rank-3 shape: receptive field = prod(shape[:-2])
        mp_w = [glorot((F, F, E), F * F, F * E) for _ in range(mp_layers)]
        fc = [(glorot((F, F), F, F), np.zeros(F, np.float32)) for _ in range(fc_layers - 1)]
        fc.append((glorot((F, F // 2), F, F // 2), np.zeros(F // 2, np.float32)))
        if peak_std is None:
            peak_std = np.ones(C, np.float32)        # model.py:224
        if peak_avg is None:
            peak_avg = np.zeros(C, np.float32)       # model.py:225
        p = cls(edge_fc=edge_fc, embed=glorot((C, F), C, F), mp_w=mp_w, fc=fc,
                out=(glorot((F // 2, C), F // 2, C), np.zeros(C, np.float32)),
                peak_std=np.asarray(peak_std, np.float32), peak_avg=np.asarray(peak_avg, np.float32),
                rbf_low=rbf_low, rbf_high=rbf_high, mp_activation=mp_activation,
                fc_activation=fc_activation, meta={"source": f"random(seed={seed})"})
        p.validate()
        return p


def baseline_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "models", "baseline", "baseline.npz")
