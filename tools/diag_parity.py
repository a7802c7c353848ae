"""GPU diagnostic (run under gpurun): accumulation behaviour of the tcgen05 path and
per-block / per-path errors on the golden fixtures.  Prints only; asserts nothing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import nmrgnn_b200  # noqa: E402
from conftest import load_golden, rel_err, scaled_err, tol_ratio  # noqa: E402


def tf32(x):
    u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def main():
    m = nmrgnn_b200.load_model()
    print("path:", m.handle.compute_path)
    print("tc compensation (x 2^-24):", m.handle.tc_compensation())
    rng = np.random.default_rng(0)
    # 1. accumulation: tf32-exact positive inputs -> every product exact, only the adds round
    A = rng.uniform(0.5, 1.5, size=(128, 64)).astype(np.float16).astype(np.float32)   # exact in fp16 and tf32
    W = rng.uniform(0.5, 1.5, size=(64, 128)).astype(np.float16).astype(np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64)
    for mode in (1, 0, 3, 2):
        d = m.handle.selftest_gemm(A, W, mode).astype(np.float64)
        ulp = np.spacing(ref.astype(np.float32)).astype(np.float64)
        e = (d - ref) / ulp
        print(f"selftest mode {mode} positive tf32-exact inputs: signed err in ulp mean {e.mean():+.3f} "
              f"std {e.std():.3f} min {e.min():+.2f} max {e.max():+.2f}")
    d32 = (A @ W).astype(np.float64)
    e = (d32 - ref) / np.spacing(ref.astype(np.float32))
    print(f"numpy fp32 matmul                       : signed err in ulp mean {e.mean():+.3f} std {e.std():.3f}")

    # 2. per-block errors, both paths
    for name in ["smallmol12_k8", "ring5_bonded", "prot300", "g108m", "edge_cases64", "prot3_batch"]:
        g = load_golden(name)
        gr = (g["atoms"], g["nlist"], g["edges"], g["inv_degree"])
        for path in ("tc", "ffma"):
            m.handle.set_option("force_ffma", 1 if path == "ffma" else 0)
            y = m(gr)
            line = f"{name:14s} {path:5s} peaks tol_ratio {tol_ratio(y, g['peaks_f64']):.3f} rel {rel_err(y, g['peaks_f64']):.2e}"
            if "edge_features" in g:
                e3 = m.edge_fc_block(g["edges"])
                line += f" | edge scaled_err {scaled_err(e3, g['edge_features']):.2e}"
                h = g["embed"]
                errs = []
                for l in range(4):
                    out = m.mp_block.mp[l]([h, g["nlist"], g["edge_features"], g["inv_degree"]])
                    errs.append(scaled_err(out, g[f"mp_nodes_{l}"]))
                    h = g[f"mp_nodes_{l}"]
                line += " | mp " + " ".join(f"{x:.1e}" for x in errs)
                fc = m.fc_block(g["mp_nodes_3"])
                line += f" | fc {scaled_err(fc, g['fc_nodes']):.1e}"
            print(line)
        print(f"{name:14s} traced-fp32 vs fp64: tol_ratio {tol_ratio(g['peaks'], g['peaks_f64']):.3f}")
    m.handle.set_option("force_ffma", 0)


if __name__ == "__main__":
    main()
