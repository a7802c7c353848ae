"""GPU diagnostic: node MLP + readout, tensor-core kernel against the exact-FP32 kernel on random node features of a
2 482-atom graph; prints which rows (mod 128) and which of the 128 output features differ.  Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import nmrgnn_b200  # noqa: E402
from conftest import load_golden  # noqa: E402

m = nmrgnn_b200.load_model()
h = m.handle
g = load_golden("g108m")
h.set_option("tc_min_atoms", 0)
atoms = g["atoms"]
rng = np.random.default_rng(0)
nodes = (rng.normal(size=(atoms.shape[0], 256)) * 0.5 + 1).astype(np.float32)
y, z = m.fc_block._run(nodes, atoms)[:2]
h.set_option("force_ffma", 1)
y2, z2 = m.fc_block._run(nodes, atoms)[:2]
h.set_option("force_ffma", 0)
z, z2 = np.asarray(z), np.asarray(z2)
e = np.abs(z - z2) > 1e-3 * np.abs(z2) + 1e-3
rows = np.flatnonzero(e.any(1))
print("bad rows", len(rows), "of", len(y), "| mod 128:", sorted(set((rows % 128).tolist())))
if len(rows):
    r = int(rows[0])
    print("row", r, "bad features:", np.flatnonzero(e[r]).tolist())
