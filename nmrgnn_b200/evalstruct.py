"""Streaming driver of the forward path: the reference's ``nmrgnn eval-struct`` command
(nmrgnn/main.py:192-278) without MDAnalysis / nmrdata / pandas.

For every frame of a structure or trajectory: build the k-nearest-neighbour graph on the GPU
(nmrgnn_knn_graph: what nmrdata.parse_universe + the inv_degree line do on the host, main.py:239-242),
run the model (main.py:245), flag implausible peaks (check_peaks, main.py:246 / library.py:30-47) and
append the rows ``index, residues, resids, names, peaks (2 dp), confident, time, frame``
(main.py:250-259).  Rows are accumulated in Python lists and written once (the reference's per-frame
``pd.concat``, main.py:264, is quadratic in the number of frames).  The three wall-clock buckets of the
reference ("MDAnalysis" graph build, "Model Inference", "Parsing"; main.py:231-267) are kept.
"""
from __future__ import annotations

import csv
import os
import time
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from . import _capi
from .graph import Universe, one_hot_elements, read_pdb
from .library import check_peaks, load_model
from .model import GNNModel

COLUMNS = ["index", "residues", "resids", "names", "peaks", "confident", "time", "frame"]


def _universe(src: Union[str, os.PathLike, Universe, Sequence[str]]) -> Universe:
    if isinstance(src, Universe) or hasattr(src, "trajectory"):
        return src
    if isinstance(src, (list, tuple)):
        if len(src) == 0:
            raise ValueError("Must pass at least one struct file")          # main.py:201-202
        if len(src) > 1:
            raise ValueError("topology + trajectory pairs need MDAnalysis; pass a (multi-MODEL) PDB file")
        src = src[0]
    return read_pdb(os.fspath(src))


def eval_struct(struct_files, output_csv: Optional[str] = None, model: Optional[GNNModel] = None,
                model_file: Optional[str] = None, neighbor_number: int = 16, stride: int = 1,
                device: int = 0, raise_on_bad_peaks: bool = False) -> Dict[str, List]:
    """Predict the chemical shifts of every ``stride``-th frame.  Returns the table as a dict of columns
    (and writes it to ``output_csv`` if given); ``result["timing"]`` holds the three time buckets in seconds
    and ``result["frames"]`` the number of frames evaluated."""
    u = _universe(struct_files)
    own_model = model is None
    if model is None:
        model = load_model(model_file, device=device)
    try:
        num_elem = model.params.num_elem
        elements = getattr(u.atoms, "elements", None)
        if elements is None:
            elements = [str(n)[0] for n in u.atoms.names]
        atoms = one_hot_elements(elements, num_elem)                       # constant over the trajectory
        n = atoms.shape[0]
        k = int(neighbor_number)
        nlist = np.empty((n, k), np.int32)
        edges = np.empty((n, k), np.float32)
        inv_degree = np.empty(n, np.float32)
        offs = np.array([0, n], np.int64)
        names = [str(x) for x in u.atoms.names]
        resnames = [str(x) for x in u.atoms.resnames]
        resids = [int(x) for x in u.atoms.resids]
        cols: Dict[str, List] = {c: [] for c in COLUMNS}
        timing = {"graph": 0.0, "inference": 0.0, "parsing": 0.0}
        n_eval = 0
        for ts in u.trajectory[::stride]:
            t0 = time.perf_counter()
            pos_nm = np.ascontiguousarray(np.asarray(u.atoms.positions, np.float32) / np.float32(10.0))
            model.handle.knn_graph(pos_nm, offs, n, 1, k, 0.0, nlist, edges, inv_degree, _capi.MEM_HOST)
            t1 = time.perf_counter()
            peaks = model((atoms, nlist, edges, inv_degree))
            try:
                confident = check_peaks(atoms, peaks)
            except Warning:
                if raise_on_bad_peaks:
                    raise
                confident = np.zeros(n, bool)
            t2 = time.perf_counter()
            cols["index"].extend(range(n))
            cols["residues"].extend(resnames)
            cols["resids"].extend(resids)
            cols["names"].extend(names)
            cols["peaks"].extend(np.round(peaks.astype(np.float64), 2).tolist())
            cols["confident"].extend(confident.tolist())
            cols["time"].extend([float(getattr(ts, "time", 0.0))] * n)
            cols["frame"].extend([int(getattr(ts, "frame", n_eval))] * n)
            t3 = time.perf_counter()
            timing["graph"] += t1 - t0
            timing["inference"] += t2 - t1
            timing["parsing"] += t3 - t2
            n_eval += 1
        if output_csv is not None:
            with open(output_csv, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(COLUMNS)
                w.writerows(zip(*[cols[c] for c in COLUMNS]))
        out: Dict[str, List] = dict(cols)
        out["timing"] = timing          # type: ignore[assignment]
        out["frames"] = n_eval          # type: ignore[assignment]
        return out
    finally:
        if own_model:
            model.close()
