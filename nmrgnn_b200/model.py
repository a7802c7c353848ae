"""Host-side mirror of the reference's model classes (nmrgnn/model.py,
nmrgnn/layers.py) on top of the C ABI.  Same names, argument order and error
behaviour as the Keras classes for the *inference* call; all arithmetic runs in
the CUDA kernels of libnmrgnn_b200.so — there is no CPU path here.

    model = nmrgnn_b200.load_model()
    peaks = model((atoms, nlist, edges, inv_degree))      # np.ndarray float32 [N]

Inputs may be NumPy arrays / nested lists (host) or torch CUDA tensors (device);
the result lives where the inputs live.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _capi
from .params import ACTIVATIONS, GNNParams

try:  # torch is plumbing only (device tensors, streams); host arrays work without it
    import torch
except Exception:  # pragma: no cover
    torch = None


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


def _on_device(*xs) -> bool:
    dev = [x for x in xs if _is_torch(x) and x.is_cuda]
    if dev and len(dev) != len(xs):
        raise ValueError("inputs must be all host arrays or all CUDA tensors")
    return bool(dev)


def _host(x, dtype) -> np.ndarray:
    if _is_torch(x):
        x = x.detach().cpu().numpy()
    elif hasattr(x, "numpy") and not isinstance(x, np.ndarray):  # tf.Tensor-like
        x = x.numpy()
    return np.ascontiguousarray(np.asarray(x), dtype=dtype)


def _dev(x, dtype, device):
    return x.to(device=device, dtype=dtype).contiguous()


def _stream_ptr(device) -> int:
    """cudaStream_t of torch's current stream; the legacy default stream (handle 0) is
    passed as cudaStreamLegacy (0x1) because NULL means "library-owned stream" in the ABI."""
    return int(torch.cuda.current_stream(device).cuda_stream) or 1


def _dims(p: GNNParams) -> _capi.Dims:
    return _capi.Dims(p.num_elem, p.atom_feature_size, p.edge_feature_size, p.edge_hidden_size,
                      len(p.edge_fc), len(p.mp_w), len(p.fc), ACTIVATIONS[p.mp_activation],
                      ACTIVATIONS[p.fc_activation], float(p.rbf_low), float(p.rbf_high))


def _weights(p: GNNParams) -> List[np.ndarray]:
    w: List[np.ndarray] = []
    for W, b in p.edge_fc:
        w += [W, b]
    w.append(p.embed)
    w += list(p.mp_w)
    for W, b in p.fc:
        w += [W, b]
    w += [p.out[0], p.out[1], p.peak_std, p.peak_avg]
    return w


class _Engine:
    """One C-ABI handle per (params, device); shared by the layer objects of a model."""

    def __init__(self, params: GNNParams, device: int = 0):
        params.validate()
        self.params = params
        self.device = int(device)
        self.handle = _capi.Handle(_dims(params), _weights(params), self.device)

    @property
    def torch_device(self):
        return torch.device("cuda", self.device)


class _Layer:
    def __init__(self, engine: _Engine, name: str):
        self._engine = engine
        self.name = name

    @property
    def hypers(self) -> Dict[str, object]:
        p = self._engine.params
        return dict(atom_feature_size=p.atom_feature_size, edge_feature_size=p.edge_feature_size,
                    edge_hidden_size=p.edge_hidden_size, mp_layers=len(p.mp_w), fc_layers=len(p.fc),
                    edge_fc_layers=len(p.edge_fc), mp_activation=p.mp_activation, fc_activation=p.fc_activation,
                    rbf_low=p.rbf_low, rbf_high=p.rbf_high)


class EdgeFCBlock(_Layer):
    """RBFExpansion -> mask -> EdgeFCBlock -> mask, fused (nmrgnn/model.py:251-261,
    110-144; nmrgnn/layers.py:102-140).  Call with the *distances* [N,K]; returns the
    trained edge features [N,K,E] (the RBF tensor is never materialised)."""

    def __init__(self, engine: _Engine):
        super().__init__(engine, "edge-fc-block")

    def __call__(self, edge_input):
        eng = self._engine
        E = eng.params.edge_feature_size
        if _on_device(edge_input):
            d = _dev(edge_input, torch.float32, eng.torch_device)
            out = torch.empty(tuple(d.shape) + (E,), dtype=torch.float32, device=d.device)
            eng.handle.edge_features(d, d.numel(), out, _capi.MEM_DEVICE, _stream_ptr(d.device))
            return out
        d = _host(edge_input, np.float32)
        out = np.empty(d.shape + (E,), np.float32)
        eng.handle.edge_features(d, d.size, out, _capi.MEM_HOST)
        return out


class RBFExpansion(EdgeFCBlock):
    """Kept as an importable name (nmrgnn.custom_objects); on this path the RBF
    expansion only exists fused inside the edge kernel."""


class MPLayer(_Layer):
    """One message-passing layer *including* the residual MPBlock adds
    (nmrgnn/layers.py:26-46 + nmrgnn/model.py:167):
    returns act(einsum('ijn,ijl,lmn,i->im', edges, nodes[nlist], w, inv_degree)) + nodes."""

    def __init__(self, engine: _Engine, index: int):
        super().__init__(engine, "MPLayer")
        self.index = index

    def __call__(self, inputs):
        nodes, nlist, edges, inv_degree = inputs
        eng = self._engine
        F, E = eng.params.atom_feature_size, eng.params.edge_feature_size
        if _on_device(nodes, nlist, edges, inv_degree):
            dev = eng.torch_device
            nodes = _dev(nodes, torch.float32, dev)
            nlist = _dev(nlist, torch.int32, dev)
            edges = _dev(edges, torch.float32, dev)
            inv_degree = _dev(inv_degree, torch.float32, dev).reshape(-1)
            n, k = nlist.shape
            _check_shapes(n, k, nodes.shape, edges.shape, inv_degree.shape, F, E)
            out = torch.empty_like(nodes)
            eng.handle.mp_layer(self.index, nodes, nlist, edges, inv_degree, n, k, out, _capi.MEM_DEVICE,
                                _stream_ptr(dev))
            return out
        nodes = _host(nodes, np.float32)
        nlist = _host(nlist, np.int32)
        edges = _host(edges, np.float32)
        inv_degree = _host(inv_degree, np.float32).reshape(-1)
        n, k = nlist.shape
        _check_shapes(n, k, nodes.shape, edges.shape, inv_degree.shape, F, E)
        out = np.empty_like(nodes)
        eng.handle.mp_layer(self.index, nodes, nlist, edges, inv_degree, n, k, out, _capi.MEM_HOST)
        return out


def _check_shapes(n, k, nodes_shape, edges_shape, inv_shape, F, E):
    if tuple(nodes_shape) != (n, F):
        raise ValueError(f"nodes must be [{n},{F}], got {tuple(nodes_shape)}")
    if tuple(edges_shape) != (n, k, E):
        raise ValueError(f"edge features must be [{n},{k},{E}], got {tuple(edges_shape)}")
    if tuple(inv_shape) != (n,):
        raise ValueError(f"inv_degree must be [{n}], got {tuple(inv_shape)}")


class MPBlock(_Layer):
    """nmrgnn/model.py:148-175."""

    def __init__(self, engine: _Engine):
        super().__init__(engine, "mp-block")
        self.mp = [MPLayer(engine, i) for i in range(len(engine.params.mp_w))]

    def __call__(self, inputs):
        nodes = inputs[0]
        for layer in self.mp:
            nodes = layer([nodes] + list(inputs[1:]))
        return nodes


class FCBlock(_Layer):
    """nmrgnn/model.py:179-202.  ``fcblock(nodes)`` returns the F/2-wide features;
    the fused readout is reached through ``GNNModel.readout``."""

    def __init__(self, engine: _Engine):
        super().__init__(engine, "fc-block")

    def _run(self, nodes, atoms):
        eng = self._engine
        p = eng.params
        F, F2, C = p.atom_feature_size, p.atom_feature_size // 2, p.num_elem
        if _on_device(nodes) or (atoms is not None and _on_device(atoms)):
            dev = eng.torch_device
            nodes = _dev(nodes, torch.float32, dev)
            n = nodes.shape[0]
            atoms = torch.zeros((n, C), dtype=torch.float32, device=dev) if atoms is None else _dev(atoms, torch.float32, dev)
            peaks = torch.empty(n, dtype=torch.float32, device=dev)
            fc = torch.empty((n, F2), dtype=torch.float32, device=dev)
            if tuple(nodes.shape) != (n, F) or tuple(atoms.shape) != (n, C):
                raise ValueError("bad nodes/atoms shape")
            eng.handle.fc_readout(nodes, atoms, n, peaks, fc, _capi.MEM_DEVICE, _stream_ptr(dev))
            return peaks, fc
        nodes = _host(nodes, np.float32)
        n = nodes.shape[0]
        atoms = np.zeros((n, C), np.float32) if atoms is None else _host(atoms, np.float32)
        if nodes.shape != (n, F) or atoms.shape != (n, C):
            raise ValueError("bad nodes/atoms shape")
        peaks = np.empty(n, np.float32)
        fc = np.empty((n, F2), np.float32)
        eng.handle.fc_readout(nodes, atoms, n, peaks, fc, _capi.MEM_HOST)
        return peaks, fc

    def __call__(self, nodes):
        return self._run(nodes, None)[1]


class GNNModel:
    """Inference twin of nmrgnn.model.GNNModel (nmrgnn/model.py:205-274)."""

    def __init__(self, params: GNNParams, device: int = 0, name: str = "gnn-model"):
        self.name = name
        self._engine = _Engine(params, device)
        self.edge_fc_block = EdgeFCBlock(self._engine)
        self.edge_rbf = RBFExpansion(self._engine)
        self.mp_block = MPBlock(self._engine)
        self.fc_block = FCBlock(self._engine)
        self.peak_std = params.peak_std
        self.peak_avg = params.peak_avg

    # -- reference-compatible surface -------------------------------------------------
    @property
    def hypers(self):
        return self.edge_fc_block.hypers

    @property
    def params(self) -> GNNParams:
        return self._engine.params

    @property
    def handle(self) -> _capi.Handle:
        return self._engine.handle

    @property
    def device(self) -> int:
        return self._engine.device

    def embed_layer(self, atoms):
        eng = self._engine
        F, C = eng.params.atom_feature_size, eng.params.num_elem
        if _on_device(atoms):
            a = _dev(atoms, torch.float32, eng.torch_device)
            if a.ndim != 2 or a.shape[1] != C:
                raise ValueError(f"atoms must be [N,{C}]")
            out = torch.empty((a.shape[0], F), dtype=torch.float32, device=a.device)
            eng.handle.embed(a, a.shape[0], out, _capi.MEM_DEVICE, _stream_ptr(a.device))
            return out
        a = _host(atoms, np.float32)
        if a.ndim != 2 or a.shape[1] != C:
            raise ValueError(f"atoms must be [N,{C}]")
        out = np.empty((a.shape[0], F), np.float32)
        eng.handle.embed(a, a.shape[0], out, _capi.MEM_HOST)
        return out

    def readout(self, nodes, atoms):
        """out_layer + peak standardisation after FCBlock (model.py:265-273)."""
        return self.fc_block._run(nodes, atoms)[0]

    def __call__(self, inputs, training: bool = False):
        if training:
            raise NotImplementedError("nmrgnn_b200 implements the inference forward only (training=False)")
        if not isinstance(inputs, (tuple, list)) or len(inputs) != 4:
            raise ValueError("inputs must be (atoms, nlist, edges, inv_degree)")
        atoms, nlist, edges, inv_degree = inputs
        eng = self._engine
        C = eng.params.num_elem
        if _on_device(atoms, nlist, edges, inv_degree):
            dev = eng.torch_device
            atoms = _dev(atoms, torch.float32, dev)
            nlist = _dev(nlist, torch.int32, dev)
            edges = _dev(edges, torch.float32, dev)
            inv_degree = _dev(inv_degree, torch.float32, dev).reshape(-1)
            self._check(atoms.shape, nlist.shape, edges.shape, inv_degree.shape, C)
            n, k = nlist.shape
            peaks = torch.empty(n, dtype=torch.float32, device=dev)
            eng.handle.forward(atoms, nlist, edges, inv_degree, n, k, peaks, _capi.MEM_DEVICE, _stream_ptr(dev))
            return peaks
        atoms = _host(atoms, np.float32)
        nlist_h = _host(nlist, np.int64)
        if nlist_h.size and (nlist_h.max() >= 2 ** 31 or nlist_h.min() < -2 ** 31):
            raise IndexError("nlist index out of int32 range")
        nlist = nlist_h.astype(np.int32)
        edges = _host(edges, np.float32)
        inv_degree = _host(inv_degree, np.float32).reshape(-1)
        self._check(atoms.shape, nlist.shape, edges.shape, inv_degree.shape, C)
        n, k = nlist.shape
        peaks = np.empty(n, np.float32)
        eng.handle.forward(atoms, nlist, edges, inv_degree, n, k, peaks, _capi.MEM_HOST)
        return peaks

    predict = __call__

    @staticmethod
    def _check(a_shape, nl_shape, e_shape, inv_shape, C):
        if len(nl_shape) != 2:
            raise ValueError("nlist must be [N,K]")
        n, k = nl_shape
        if tuple(a_shape) != (n, C):
            raise ValueError(f"atoms must be [{n},{C}], got {tuple(a_shape)}")
        if tuple(e_shape) != (n, k):
            raise ValueError(f"edges must be [{n},{k}], got {tuple(e_shape)}")
        if tuple(inv_shape) != (n,):
            raise ValueError(f"inv_degree must be [{n}], got {tuple(inv_shape)}")
        if n > 0 and k < 1:
            raise ValueError("need at least one neighbour slot")

    def synchronize(self):
        """Wait for device-tensor calls enqueued on torch's current stream and raise
        deferred device-side errors (IndexError for an out-of-range nlist entry)."""
        self._engine.handle.synchronize(_stream_ptr(self._engine.torch_device) if torch is not None else None)

    def close(self):
        self._engine.handle.close()


def build_GNNModel(hp: Optional[Dict[str, object]] = None, metrics: bool = False, loss_balance: float = 1.0,
                   num_elem: int = 16, seed: int = 0, device: int = 0, peak_std=None, peak_avg=None) -> GNNModel:
    """Fresh randomly initialised model with the reference's hyper-parameter names
    and defaults (nmrgnn/model.py:12-41).  Optimizer / loss / metric wiring
    (model.py:43-104) is training-only and out of scope."""
    hp = dict(hp or {})
    p = GNNParams.random(
        num_elem=num_elem,
        atom_feature_size=int(hp.get("atom_feature_size", 256)),
        edge_feature_size=int(hp.get("edge_feature_size", 3)),
        edge_hidden_size=int(hp.get("edge_hidden_size", 128)),
        mp_layers=int(hp.get("mp_layers", 4)), fc_layers=int(hp.get("fc_layers", 4)),
        edge_fc_layers=int(hp.get("edge_fc_layers", 4)),
        mp_activation=str(hp.get("mp_activation", "softplus")),
        fc_activation=str(hp.get("fc_activation", "softplus")),
        rbf_low=float(hp.get("rbf_low", 0.005)), rbf_high=float(hp.get("rbf_high", 0.20)), seed=seed,
        peak_std=peak_std, peak_avg=peak_avg)
    return GNNModel(p, device=device)
