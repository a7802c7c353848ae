"""GPU diagnostic: which block of the tensor-core path costs accuracy on a fixture?  Composes the forward
from the per-block API with each block on either path (t = tcgen05, f = FFMA) and prints the tolerance
ratio against the fp64 golden.  Prints only."""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nmrgnn_b200  # noqa: E402
from conftest import load_golden, tol_ratio  # noqa: E402


def main():
    m = nmrgnn_b200.load_model()
    m.handle.set_option("tc_min_atoms", 0)
    names = sys.argv[1:] or ["smallmol12_k8", "prot3_batch", "prot300"]
    for name in names:
        g = load_golden(name)
        for comp in (1, 0):
            m.handle.set_option("tc_compensate", comp)
            for pe, pm, pf in itertools.product("tf", repeat=3):
                def use(c):
                    m.handle.set_option("force_ffma", 1 if c == "f" else 0)
                use(pe)
                e3 = m.edge_fc_block(g["edges"])
                use("f")
                h = m.embed_layer(g["atoms"])
                use(pm)
                h = m.mp_block([h, g["nlist"], e3, g["inv_degree"]])
                use(pf)
                y = m.readout(h, g["atoms"])
                err = np.abs(y - g["peaks_f64"]) / (1e-4 * np.abs(g["peaks_f64"]) + 1e-4)
                print(f"{name:14s} comp={comp} edge={pe} mp={pm} fc={pf}: tol_ratio {err.max():.3f} (atom {int(err.argmax())}, "
                      f"ref {g['peaks_f64'][int(err.argmax())]:.4f}), 2nd {np.sort(err)[-2]:.3f}")
    m.handle.set_option("force_ffma", 0)
    m.handle.set_option("tc_compensate", 1)


if __name__ == "__main__":
    main()
