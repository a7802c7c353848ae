"""GPU diagnostic: response of the tensor-core path's peak error (config 2, all 164 105 atoms, vs the fp64 golden peaks)
to the round-toward-zero compensation constants.  Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402


def stats(y, ref):
    e = np.abs(y - ref) / (1e-4 * np.abs(ref) + 1e-4)
    return f"max {e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} p99.9 {np.quantile(e, 0.999):.3f} rms {np.sqrt(np.mean(e * e)):.4f} >1: {int((e > 1).sum())}"


def main():
    m = nmrgnn_b200.load_model()
    z = np.load(os.path.join(ROOT, "tests", "golden", "full_config2.npz"))
    atoms, nlist, edges, inv, offs = workloads.protein_batch(64, first_seed=0)
    ref = z["peaks_f64"]
    g = (atoms, nlist, edges, inv)
    print("calibrated:", m.handle.tc_compensation())
    print("calibrated         ", stats(m(g).astype(np.float64), ref))
    for d in (-40, -20, -10, 10, 20, 40):
        m.handle.set_option("mp_comp_delta_x10", d)
        print(f"mp delta {d / 10:+.1f}      ", stats(m(g).astype(np.float64), ref))
        m.handle.set_option("mp_comp_delta_x10", -d)
    for c in (100, 120, 132, 145):
        m.handle.set_option("mp_comp_x10", c)
        print(f"mp all = {c / 10:.1f}      ", stats(m(g).astype(np.float64), ref))
    m.handle.set_option("mp_comp_x10", -1)
    print("recalibrated       ", stats(m(g).astype(np.float64), ref))
    for c in (0, 15, 29, 45, 60):
        m.handle.set_option("fc_comp_x10", c)
        print(f"fc = {c / 10:.1f}           ", stats(m(g).astype(np.float64), ref))


if __name__ == "__main__":
    main()
