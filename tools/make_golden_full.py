#!/usr/bin/env python
"""Full-size golden peaks for the BASELINE.json configurations themselves.

Executes the reference's traced SavedModel graph (oracle/savedmodel_interp.py), graph by graph, in
float32 and float64 on

* config[1]: the 64 synthetic protein graphs bench.py times (workloads.protein_batch(64), 164 105 atoms),
* config[2]: the 1 024 small molecules (workloads.small_molecule_batch(1024), K = 8),

and writes tests/golden/full_config{2,3}.npz holding ONLY the outputs (`peaks` f32, `peaks_f64`),
`graph_offsets` and a sha256 digest of every graph's inputs: the inputs regenerate from their seeds
(nmrgnn_b200/workloads.py), and the GPU tests check the digests before comparing, so a host whose
libm generates different inputs fails loudly instead of comparing against the wrong vectors.

Runs only in the build container (needs /root/reference).  ~8 minutes on 8 cores.
"""
import hashlib
import multiprocessing as mp
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nmrgnn_b200 import workloads  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
_IT = {}


def graph_digest(atoms, nlist, edges, inv) -> str:
    h = hashlib.sha256()
    for a, dt in ((atoms, np.float32), (nlist, np.int32), (edges, np.float32), (inv, np.float32)):
        h.update(np.ascontiguousarray(a, dt).tobytes())
    return h.hexdigest()[:16]


def _interp(dtype):
    from oracle.savedmodel_interp import SavedModelInterpreter, reference_dir
    if dtype not in _IT:
        _IT[dtype] = SavedModelInterpreter(reference_dir(), dtype)
    return _IT[dtype]


def _one(args):
    kind, seed = args
    g = workloads.protein_graph(seed) if kind == "protein" else workloads.small_molecule_graph(seed)
    p32 = _interp(np.float32)(*g).astype(np.float32)
    p64 = _interp(np.float64)(*g).astype(np.float64)
    return g[0].shape[0], graph_digest(*g), p32, p64


def make(name, kind, n_graphs, workers, note):
    t0 = time.time()
    with mp.get_context("fork").Pool(workers) as pool:
        res = pool.map(_one, [(kind, s) for s in range(n_graphs)], chunksize=1 if kind == "protein" else 16)
    sizes = np.array([r[0] for r in res], np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, peaks=np.concatenate([r[2] for r in res]),
                        peaks_f64=np.concatenate([r[3] for r in res]), graph_offsets=offs,
                        digests=np.array([r[1] for r in res]), note=np.frombuffer(note.encode(), np.uint8))
    print(f"{name}: {n_graphs} graphs, {offs[-1]} atoms -> {os.path.getsize(path)} bytes, {time.time() - t0:.0f} s")


if __name__ == "__main__":
    os.environ.setdefault("OMP_NUM_THREADS", "2")
    w = int(os.environ.get("WORKERS", "4"))
    make("full_config3", "small", 1024, w, "BASELINE configs[2]: workloads.small_molecule_batch(1024), K=8; "
         "traced SavedModel graph in NumPy, per graph, f32 and f64")
    make("full_config2", "protein", 64, w, "BASELINE configs[1]: workloads.protein_batch(64), K=16; "
         "traced SavedModel graph in NumPy, per graph, f32 and f64")
