"""GPU diagnostic: the single-accumulator MP kernel (option mp_single_acc) against the two-accumulator kernel: peak error
on the full BASELINE configs 2 / 3 (vs the fp64 golden peaks) as a function of the compensation slope, and the MP launch
time.  Usage: diag_single_acc.py [slope_x100 ...]   Prints only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads  # noqa: E402


def stats(y, ref):
    e = np.abs(y - ref) / (1e-4 * np.abs(ref) + 1e-4)
    return (f"max {e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} p99.9 {np.quantile(e, 0.999):.3f} "
            f"rms {np.sqrt(np.mean(e * e)):.4f} >1: {int((e > 1).sum())}")


def main():
    slopes = [int(a) for a in sys.argv[1:]] or [45, 55, 65]
    m = nmrgnn_b200.load_model()
    h = m.handle
    b2 = workloads.protein_batch(64, first_seed=0)
    b3 = workloads.small_molecule_batch(1024, first_seed=0)
    r2 = np.load(os.path.join(ROOT, "tests", "golden", "full_config2.npz"))["peaks_f64"]
    r3 = np.load(os.path.join(ROOT, "tests", "golden", "full_config3.npz"))["peaks_f64"]

    def run(label):
        comp = [round(c, 1) for c in h.tc_compensation()["mp_layers"]]
        y2 = m(b2[:4]).astype(np.float64)
        y3 = m(b3[:4]).astype(np.float64)
        h.set_option("profile", 1)
        ts = []
        for _ in range(5):
            m(b2[:4])
            ts.append(np.mean(h.stage_times()["mp_layers"]))
        h.set_option("profile", 0)
        print(f"{label} comp {comp} | cfg2 {stats(y2, r2)} | cfg3 {stats(y3, r3)} | mp ms {np.median(ts):.4f}", flush=True)

    h.set_option("mp_single_acc", 0)
    run("two accumulators          ")
    h.set_option("mp_single_acc", 1)
    for sl in slopes:
        h.set_option("mp_pos_comp1_x100", sl)
        run(f"single accumulator c'={sl / 100:.2f}")


if __name__ == "__main__":
    main()
