"""Reference-compatible library API (nmrgnn/library.py): ``load_model``,
``universe2graph``, ``check_peaks`` plus the two ``nmrdata`` lookups they rely on
(``load_embeddings``, ``load_standards``), restated from the constants baked in
the pretrained SavedModel because ``nmrdata`` is not vendored by the reference."""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import numpy as np

from .graph import ELEMENT_INDEX, inv_degree_from_nlist, one_hot_elements
from .model import GNNModel
from .params import GNNParams, baseline_path, baseline_standards


def load_embeddings() -> Dict[str, Dict[str, int]]:
    """Element -> one-hot column (stand-in for nmrdata.load_embeddings()['atom'])."""
    return {"atom": dict(ELEMENT_INDEX)}


def load_standards(num_elem: int = 10) -> Dict[int, Tuple[str, float, float]]:
    """index -> (element, mean ppm, std ppm), the layout check_peaks indexes
    (nmrgnn/library.py:34-40: ps[1] = avg, ps[2] = std)."""
    std, avg = baseline_standards(num_elem)
    names = {v: k for k, v in ELEMENT_INDEX.items()}
    return {i: (names.get(i, "X"), float(avg[i]), float(std[i])) for i in range(num_elem)}


def load_baseline() -> str:
    """Path of the packaged pretrained weights (nmrgnn/library.py:22-27)."""
    return baseline_path()


def load_model(model_file: Optional[str] = None, device: int = 0) -> GNNModel:
    """Load the chemical-shift model; with no file the pretrained baseline is used
    (nmrgnn/library.py:92-103).  ``model_file`` may be this repo's ``.npz`` weights
    or a TensorFlow SavedModel / checkpoint directory written by the reference
    (``model.save`` / ``ModelCheckpoint``) — read without TensorFlow."""
    if model_file is None:
        model_file = load_baseline()
    model_file = os.fspath(model_file)
    if os.path.isfile(model_file) and model_file.endswith(".npz"):
        params = GNNParams.load(model_file)
    else:
        params = GNNParams.from_tf_checkpoint(model_file)
    return GNNModel(params, device=device)


def universe2graph(u, neighbor_number: int = 16, model: Optional[GNNModel] = None, num_elem: int = 10):
    """(atoms, nlist, edges, inv_degree) from a Universe-like object in Angstrom with
    explicit hydrogens (nmrgnn/library.py:106-117).  ``u`` needs ``u.atoms.positions``
    and ``u.atoms.elements`` (MDAnalysis Universe or nmrgnn_b200.graph.Universe).
    With ``model`` given the neighbour search runs on the GPU (nmrgnn_knn_graph);
    otherwise the host KD-tree builder is used."""
    pos_nm = np.ascontiguousarray(np.asarray(u.atoms.positions, np.float32) / np.float32(10.0))
    elements = getattr(u.atoms, "elements", None)
    if elements is None:
        elements = [str(n)[0] for n in u.atoms.names]
    atoms = one_hot_elements(elements, num_elem)
    n = pos_nm.shape[0]
    if model is not None:
        nlist = np.empty((n, neighbor_number), np.int32)
        edges = np.empty((n, neighbor_number), np.float32)
        inv_degree = np.empty(n, np.float32)
        offs = np.array([0, n], np.int64)
        from . import _capi
        model.handle.knn_graph(pos_nm, offs, n, 1, neighbor_number, 0.0, nlist, edges, inv_degree, _capi.MEM_HOST)
        return atoms, nlist, edges, inv_degree
    from .graph import knn_graph_host
    nlist, edges = knn_graph_host(pos_nm, neighbor_number)
    return atoms, nlist, edges, inv_degree_from_nlist(nlist)


def check_peaks(atoms, peaks, cutoff_sigma: float = 4, warn_sigma: float = 2.5) -> np.ndarray:
    """True where a peak is within ``warn_sigma`` standard deviations of its
    element's training mean; raises ``Warning`` if fewer than 75 % are
    (nmrgnn/library.py:30-47)."""
    atoms = np.asarray(atoms)
    peaks = np.asarray(peaks, np.float64).reshape(-1)
    standards = load_standards(atoms.shape[1])
    idx = np.argmax(atoms != 0, axis=1)
    avg = np.array([standards[int(i)][1] for i in idx])
    std = np.array([standards[int(i)][2] for i in idx])
    with np.errstate(divide="ignore", invalid="ignore"):
        confident = (std != 0) & ~((peaks - avg) ** 2 / std ** 2 > warn_sigma ** 2)
    if confident.shape[0] and confident.sum() / confident.shape[0] < 0.75:
        raise Warning("Your peaks look awful. Likely solvent or missing hydrogens or bad units. "
                      "Check README for suggestions")
    return confident
