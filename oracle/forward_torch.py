"""TEST INFRASTRUCTURE — not product code (see oracle/forward.py header).

torch-CPU restatement of the reference forward in the *reference's own op order*
(the three einsums exactly as lowered in the traced SavedModel graph, no algebraic
re-association), multi-threaded over the host cores.  This is the CPU baseline
bench.py times (`cpu_baseline`, `--impl reference`): TensorFlow/Keras cannot be
installed here, and torch's MKL/oneDNN kernels are the closest stand-in for TF's
Eigen/MKL CPU kernels.  Checked against the golden vectors in tests/test_oracle.py.
"""
from __future__ import annotations

import numpy as np
import torch

from nmrgnn_b200.params import GNNParams, rbf_centers


def _softplus(x: torch.Tensor) -> torch.Tensor:
    # tf.nn.softplus functor: thresholded log(exp(x)+1)  (tensorflow/core/kernels/softplus_op.h)
    eps = torch.finfo(x.dtype).eps
    thr = float(np.log(eps) + 2.0)
    ex = torch.exp(x)
    return torch.where(x > -thr, x, torch.where(x < thr, ex, torch.log(ex + 1)))


def _act(name):
    return {"softplus": _softplus, "relu": torch.relu, "tanh": torch.tanh, "linear": lambda x: x}[name]


class TorchReference:
    """model(inputs) like the revived Keras model; one graph per call (the reference
    has no batch dimension: nmrgnn/main.py:236-245)."""

    def __init__(self, params: GNNParams, dtype=torch.float32, reference_order: bool = True):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        self.p = params
        self.dtype = dtype
        self.reference_order = reference_order
        self.edge_fc = [(t(W), t(b)) for W, b in params.edge_fc]
        self.embed = t(params.embed)
        self.mp_w = [t(w) for w in params.mp_w]
        self.mp_wp = [t(np.transpose(w, (0, 2, 1)).reshape(-1, w.shape[1])) for w in params.mp_w]
        self.fc = [(t(W), t(b)) for W, b in params.fc]
        self.out = (t(params.out[0]), t(params.out[1]))
        self.peak_std, self.peak_avg = t(params.peak_std), t(params.peak_avg)
        c, gap = rbf_centers(params.rbf_low, params.rbf_high, params.rbf_count)
        self.centers, self.gap = t(c), float(gap)

    @torch.no_grad()
    def __call__(self, inputs) -> np.ndarray:
        atoms, nlist, edges, inv_degree = inputs
        dt = self.dtype
        atoms = torch.as_tensor(np.asarray(atoms)).to(dt)
        nlist = torch.as_tensor(np.asarray(nlist)).to(torch.int64)
        edges = torch.as_tensor(np.asarray(edges)).to(dt)
        inv_degree = torch.as_tensor(np.asarray(inv_degree)).to(dt).reshape(-1)
        fc_act, mp_act = _act(self.p.fc_activation), _act(self.p.mp_activation)
        m = (edges > 0).to(dt)[..., None]                                   # model.py:251
        x = torch.exp(-(edges[..., None] - self.centers) ** 2 / self.gap) * m   # layers.py:137-140, model.py:257
        for i, (W, b) in enumerate(self.edge_fc):                          # model.py:132-138
            x = x @ W + b
            if i < len(self.edge_fc) - 1:
                x = fc_act(x)
        e = x * m                                                           # model.py:261
        h = atoms @ self.embed                                              # model.py:262
        for w, wp in zip(self.mp_w, self.mp_wp):                            # model.py:158-169
            g = h[nlist]                                                    # layers.py:33
            if self.reference_order:
                A = torch.einsum("lmn,ijl->mnij", w, g)                     # traced lowering of layers.py:39-40
                B = torch.einsum("mnij,ijn->mi", A, e)
                r = torch.einsum("mi,i->im", B, inv_degree)
            else:
                T = torch.einsum("ijn,ijl->iln", e, g).reshape(g.shape[0], -1)
                r = (T @ wp) * inv_degree[:, None]
            h = mp_act(r) + h                                               # layers.py:42, model.py:167
        for W, b in self.fc[:-1]:                                           # model.py:191-196
            h = fc_act(h @ W + b) + h
        W, b = self.fc[-1]
        h = fc_act(h @ W + b)
        full = h @ self.out[0] + self.out[1]                                # model.py:268
        peaks = torch.sum(full * atoms * self.peak_std + atoms * self.peak_avg, dim=-1)   # model.py:272-273
        return peaks.numpy()

    def per_graph(self, batch) -> np.ndarray:
        atoms, nlist, edges, inv, offs = batch
        out = np.empty(atoms.shape[0], np.float32 if self.dtype == torch.float32 else np.float64)
        for g in range(len(offs) - 1):
            a, b = int(offs[g]), int(offs[g + 1])
            out[a:b] = self((atoms[a:b], np.asarray(nlist[a:b]).astype(np.int64) - a, edges[a:b], inv[a:b]))
        return out
