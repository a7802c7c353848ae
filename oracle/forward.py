"""TEST INFRASTRUCTURE — not product code.  Only tests/, tools/ fixture generators,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; nmrgnn_b200/ never does.

CPU restatement (NumPy) of the reference's GNN inference forward, written from
the reference sources and checked against the NumPy execution of the reference's
own traced SavedModel graph (oracle/savedmodel_interp.py -> tests/golden/*.npz).

Parity status: the reference's own tests hold no golden values for this path and
TensorFlow cannot be installed here, so no TensorFlow-executed output exists:
in the strict sense of the task rules PARITY IS UNPINNED.  What the oracle is
pinned to instead is the reference's traced SavedModel graph + pretrained
weights executed node by node with NumPy kernels (op order, activations, einsum
lowering and baked constants come from the reference's artefact; only the TF
kernels are restated).

Every function cites the reference lines it follows (paths relative to the
reference repo root).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np

from nmrgnn_b200.params import GNNParams, rbf_centers


# --------------------------------------------------------------------------- #
# activations                                                                 #
# --------------------------------------------------------------------------- #
def softplus(x: np.ndarray) -> np.ndarray:
    """tf.nn.softplus as TF's Eigen functor computes it (thresholded
    log(exp(x)+1)); used by every Dense(activation='softplus') and MPLayer
    (nmrgnn/model.py:33-36 defaults)."""
    eps = np.finfo(x.dtype).eps
    thr = np.log(x.dtype.type(eps)) + x.dtype.type(2)
    with np.errstate(over="ignore"):
        ex = np.exp(x)
    return np.where(x > -thr, x, np.where(x < thr, ex, np.log(ex + x.dtype.type(1)))).astype(x.dtype)


def activation(name: str):
    if name == "softplus":
        return softplus
    if name == "relu":
        return lambda x: np.maximum(x, x.dtype.type(0))
    if name == "tanh":
        return np.tanh
    if name in ("linear", None):
        return lambda x: x
    raise ValueError(name)


# --------------------------------------------------------------------------- #
# blocks                                                                      #
# --------------------------------------------------------------------------- #
def edge_mask(edges: np.ndarray) -> np.ndarray:
    """nmrgnn/model.py:251 — cast(edge_input > 0)[..., newaxis]."""
    return (edges > 0).astype(edges.dtype)[..., None]


def rbf_expansion(edges: np.ndarray, low: float, high: float, count: int) -> np.ndarray:
    """nmrgnn/layers.py:137-140 with the grid of layers.py:126-129:
    exp(-(d - mu)^2 / gap); note the division by gap, not gap^2."""
    centers, gap = rbf_centers(low, high, count)
    centers = centers.astype(edges.dtype)
    gap = edges.dtype.type(gap)
    return np.exp(-(edges[..., None] - centers) ** 2 / gap)


def edge_fc_block(x: np.ndarray, layers: Sequence[Tuple[np.ndarray, np.ndarray]], act: str) -> np.ndarray:
    """nmrgnn/model.py:132-138; Dense stack of model.py:118-128: all layers use
    fc_activation except the last (linear)."""
    f = activation(act)
    for i, (W, b) in enumerate(layers):
        x = x @ W + b
        if i < len(layers) - 1:
            x = f(x)
    return x


def mp_layer_reference_order(nodes, nlist, edges3, inv_degree, w, act: str) -> np.ndarray:
    """nmrgnn/layers.py:26-46 in the contraction order TF chose when tracing
    einsum('ijn,ijl,lmn,i->im') (SavedModel nodes mp-block/MPLayer/einsum*/Einsum,
    Einsum_1, Einsum_2): lmn,ijl->mnij ; mnij,ijn->mi ; mi,i->im."""
    g = nodes[nlist]                                         # layers.py:33
    A = np.einsum("lmn,ijl->mnij", w, g, optimize=True)
    B = np.einsum("mnij,ijn->mi", A, edges3, optimize=True)
    r = np.einsum("mi,i->im", B, inv_degree)
    return activation(act)(r)                                # layers.py:42


def mp_layer(nodes, nlist, edges3, inv_degree, w, act: str) -> np.ndarray:
    """Same function re-associated the way the CUDA path computes it:
    T[i,(l,n)] = sum_j e[i,j,n] h[nl[i,j],l]; r = inv_deg * (T @ W'), W'[(l,n),m]=w[l,m,n]."""
    N, K = nlist.shape
    F, _, E = w.shape
    g = nodes[nlist]                                         # [N,K,F]
    T = np.einsum("ijn,ijl->iln", edges3, g).reshape(N, F * E)
    Wp = np.transpose(w, (0, 2, 1)).reshape(F * E, F)
    r = (T @ Wp) * inv_degree[:, None]
    return activation(act)(r)


def mp_block(nodes, nlist, edges3, inv_degree, ws: Sequence[np.ndarray], act: str,
             reference_order: bool = False, collect: Optional[List[np.ndarray]] = None) -> np.ndarray:
    """nmrgnn/model.py:158-169: nodes = MPLayer(...) + nodes, every layer reads the
    previous layer's nodes."""
    layer = mp_layer_reference_order if reference_order else mp_layer
    for w in ws:
        nodes = layer(nodes, nlist, edges3, inv_degree, w, act) + nodes   # model.py:167
        if collect is not None:
            collect.append(nodes)
    return nodes


def fc_block(nodes: np.ndarray, layers: Sequence[Tuple[np.ndarray, np.ndarray]], act: str) -> np.ndarray:
    """nmrgnn/model.py:191-196: residual Dense layers, then Dense(F//2) without residual."""
    f = activation(act)
    for W, b in layers[:-1]:
        nodes = f(nodes @ W + b) + nodes                     # model.py:193
    W, b = layers[-1]
    return f(nodes @ W + b)                                  # model.py:194


def readout(nodes, atoms, out: Tuple[np.ndarray, np.ndarray], peak_std, peak_avg) -> np.ndarray:
    """nmrgnn/model.py:268,272-273."""
    full = nodes @ out[0] + out[1]
    return np.sum(full * atoms * peak_std + atoms * peak_avg, axis=-1)


def inv_degree_from_nlist(nlist: np.ndarray, dtype=np.float32) -> np.ndarray:
    """nmrgnn/library.py:115-116 / main.py:241: divide_no_nan(1, sum(nlist > 0)).
    A genuine neighbour with index 0 is not counted — preserved on purpose."""
    deg = np.sum(nlist > 0, axis=1).astype(dtype)
    out = np.zeros_like(deg)
    np.divide(dtype(1) if callable(dtype) else 1.0, deg, out=out, where=deg > 0)
    return out


# --------------------------------------------------------------------------- #
# whole forward                                                               #
# --------------------------------------------------------------------------- #
def forward(params: GNNParams, atoms, nlist, edges, inv_degree, dtype=np.float32,
            reference_order: bool = False, intermediates: Optional[dict] = None) -> np.ndarray:
    """nmrgnn/model.py:245-274 (training=False: GaussianNoise and Dropout are
    identities, model.py:253,266-267)."""
    dt = np.dtype(dtype)
    p = params.astype(dt)
    atoms = np.asarray(atoms, dt)
    edges = np.asarray(edges, dt)
    inv_degree = np.asarray(inv_degree, dt).reshape(-1)
    nlist = np.asarray(nlist).astype(np.int64)
    n = atoms.shape[0]
    if nlist.size and (nlist.min() < 0 or nlist.max() >= n):
        raise IndexError("nlist index out of range")         # TF CPU GatherV2 raises too

    m = edge_mask(edges)                                                         # model.py:251
    rbf = rbf_expansion(edges, p.rbf_low, p.rbf_high, p.rbf_count) * m           # model.py:254,257
    e3 = edge_fc_block(rbf, p.edge_fc, p.fc_activation) * m                      # model.py:258,261
    h = atoms @ p.embed                                                          # model.py:262
    hs: List[np.ndarray] = []
    h = mp_block(h, nlist, e3, inv_degree, p.mp_w, p.mp_activation, reference_order, hs)  # model.py:264
    z = fc_block(h, p.fc, p.fc_activation)                                       # model.py:265
    peaks = readout(z, atoms, p.out, p.peak_std, p.peak_avg)                     # model.py:268-273
    if intermediates is not None:
        intermediates.update(edge_features=e3, embed=atoms @ p.embed, mp_nodes=hs, fc_nodes=z)
    return peaks


def forward_per_graph(params: GNNParams, atoms, nlist, edges, inv_degree, graph_offsets,
                      dtype=np.float32, reference_order: bool = True) -> np.ndarray:
    """The reference has no batch dimension (one graph per call, main.py:236-245):
    run a concatenated batch graph by graph, undoing the per-graph nlist offset."""
    out = np.empty(atoms.shape[0], dtype)
    offs = np.asarray(graph_offsets, np.int64)
    for g in range(len(offs) - 1):
        a, b = int(offs[g]), int(offs[g + 1])
        nl = np.asarray(nlist[a:b]).astype(np.int64)
        ed = np.asarray(edges[a:b])   # batching adds `a` to every slot, padded ones included
        out[a:b] = forward(params, atoms[a:b], nl - a, ed, inv_degree[a:b], dtype, reference_order)
    return out
