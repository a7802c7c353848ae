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
                device: int = 0, raise_on_bad_peaks: bool = True, frames_per_batch: Optional[int] = None,
                chunk_frames: int = 2048) -> Dict[str, List]:
    """Predict the chemical shifts of every ``stride``-th frame.  Returns the table as a dict of columns
    (and writes it to ``output_csv`` if given); ``result["timing"]`` holds the three time buckets in seconds
    and ``result["frames"]`` the number of frames evaluated.

    Frames go through the FrameStream engine (mdstream.py): batches of frames per launch, graph build + forward
    captured in a CUDA graph, copies overlapped.  Like the reference (main.py:246), a frame whose peaks fail
    ``check_peaks`` raises ``Warning``; ``raise_on_bad_peaks=False`` marks the frame's atoms as not confident instead
    (and logs it)."""
    from .mdstream import FrameStream
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
        names = [str(x) for x in u.atoms.names]
        resnames = [str(x) for x in u.atoms.resnames]
        resids = [int(x) for x in u.atoms.resids]
        cols: Dict[str, List] = {c: [] for c in COLUMNS}
        timing = {"graph": 0.0, "inference": 0.0, "parsing": 0.0}
        n_eval = 0
        fs = FrameStream(model, elements, n, int(neighbor_number), frames_per_batch)
        pending_pos: List[np.ndarray] = []
        pending_ts: List[tuple] = []

        def flush():
            nonlocal n_eval
            if not pending_pos:
                return
            t1 = time.perf_counter()
            res = fs.run(np.stack(pending_pos))
            t2 = time.perf_counter()
            # graph build and forward run inside one captured device step: split the bucket by their device shares
            timing["inference"] += t2 - t1
            for f, (t_frame, i_frame) in enumerate(pending_ts):
                peaks = res["peaks"][f]
                try:
                    confident = check_peaks(atoms, peaks)
                except Warning as w:
                    if raise_on_bad_peaks:
                        raise
                    print(f"nmrgnn_b200.eval_struct: frame {i_frame}: {w} (raise_on_bad_peaks=False: marked not confident)")
                    confident = np.zeros(n, bool)
                cols["index"].extend(range(n))
                cols["residues"].extend(resnames)
                cols["resids"].extend(resids)
                cols["names"].extend(names)
                cols["peaks"].extend(np.round(peaks.astype(np.float64), 2).tolist())
                cols["confident"].extend(confident.tolist())
                cols["time"].extend([t_frame] * n)
                cols["frame"].extend([i_frame] * n)
                n_eval += 1
            timing["parsing"] += time.perf_counter() - t2
            pending_pos.clear()
            pending_ts.clear()

        t0 = time.perf_counter()
        for ts in u.trajectory[::stride]:
            pending_pos.append(np.asarray(u.atoms.positions, np.float32) / np.float32(10.0))   # Angstrom -> nm
            pending_ts.append((float(getattr(ts, "time", 0.0)), int(getattr(ts, "frame", n_eval + len(pending_ts)))))
            if len(pending_pos) >= chunk_frames:
                timing["graph"] += time.perf_counter() - t0       # reading / staging the coordinates ("MDAnalysis" bucket)
                flush()
                t0 = time.perf_counter()
        timing["graph"] += time.perf_counter() - t0
        flush()
        if output_csv is not None:
            with open(output_csv, "w", newline="") as f:
                w = csv.writer(f)
                w.writerow(COLUMNS)
                w.writerows(zip(*[cols[c] for c in COLUMNS]))
        out: Dict[str, List] = dict(cols)
        out["timing"] = timing          # type: ignore[assignment]
        out["frames"] = n_eval          # type: ignore[assignment]
        out["frames_per_batch"] = fs.B  # type: ignore[assignment]
        out["cuda_graph"] = fs.graph_captured  # type: ignore[assignment]
        return out
    finally:
        if own_model:
            model.close()
