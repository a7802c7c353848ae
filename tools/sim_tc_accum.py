"""CPU experiment (NumPy bit-model, no GPU): the round-toward-zero accumulation of the MP layer's main tensor-core
chain (48 instructions of 16 fp16 x fp16 products per output) on real layer inputs from a golden fixture, and how much
of its error different compensation schemes remove:
  const : D * (1 + c), c fitted by least squares (what `calibrate_mp` measures on the GPU)
  pos   : D + sum_k gamma_k P_k with gamma_k = c' (n - k + 1) + b  -- the expected truncation of a product entering at
          instruction k (it is truncated with the accumulator n - k + 1 times); linear in the products, so it can be
          folded into the `lo` image of W' at create time
  fit48 : gamma_k free (least squares over the sample)
Model: align the 16 products and the accumulator to the largest exponent, keep 24 + 2 bits, truncate
each addend toward zero, add exactly, truncate the sum toward zero to 24 bits."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from nmrgnn_b200.params import GNNParams, baseline_path  # noqa: E402


def trunc_to(x, q):
    return np.trunc(x / q) * q


def exp_of(x):
    with np.errstate(divide="ignore"):
        e = np.floor(np.log2(np.abs(x)))
    return np.where(x == 0, -1000.0, e)


def rz_chain(prods, guard=2, nseg=1):
    """prods: [n_instr, rows, 16, cols] exact products -> (final accumulator, partial sums P_k [n_instr, rows, cols])"""
    n = prods.shape[0]
    acc = np.zeros(prods.shape[1:2] + prods.shape[3:])
    total = np.zeros_like(acc)
    seg = n // nseg
    for i in range(n):
        if i % seg == 0 and i:
            total = (total + acc).astype(np.float32).astype(np.float64)
            acc = np.zeros_like(acc)
        p = prods[i]
        emax = np.maximum(exp_of(acc), exp_of(p).max(axis=1))
        q = np.exp2(emax - 23 - guard)
        s = trunc_to(acc, q) + trunc_to(p, q[:, None, :]).sum(axis=1)
        q2 = np.exp2(exp_of(s) - 23)
        acc = np.where(s == 0, 0.0, trunc_to(s, q2))
    return (total + acc).astype(np.float32).astype(np.float64) if nseg > 1 else acc


def f16(x):
    return x.astype(np.float16).astype(np.float64)


def main():
    p = GNNParams.load(baseline_path())
    g = load_golden(sys.argv[1] if len(sys.argv) > 1 else "prot300")
    rows = slice(0, int(sys.argv[2]) if len(sys.argv) > 2 else 128)
    e3 = g["edge_features"].astype(np.float64)
    nl = g["nlist"]
    inv = g["inv_degree"].astype(np.float64)[rows]
    for l in range(4):
        h = (g["embed"] if l == 0 else g[f"mp_nodes_{l - 1}"]).astype(np.float64)
        T = np.einsum("ijn,ijl->iln", e3[rows], h[nl[rows]]).astype(np.float32).astype(np.float64)   # [rows, 256, 3]
        w = p.mp_w[l].astype(np.float64)                                                            # [l_in, m, n]
        # instruction k = (ps * 3 + n) * 2 + ks covers features 32 ps + 16 ks .. + 15 of channel n
        A = np.empty((48, T.shape[0], 16))
        B = np.empty((48, 16, 256))
        for ps in range(8):
            for n in range(3):
                for ks in range(2):
                    k = (ps * 3 + n) * 2 + ks
                    f0 = 32 * ps + 16 * ks
                    A[k] = T[:, f0:f0 + 16, n]
                    B[k] = w[f0:f0 + 16, :, n]
        sc = np.exp2(-np.ceil(np.log2(np.abs(T).max(axis=(1, 2)) / 1024.0)))       # row scaling to the fp16 range
        Ah = f16(A * sc[None, :, None])
        Bh = f16(B)
        prods = Ah[:, :, :, None] * Bh[:, None, :, :]                                # [48, rows, 16, 256]
        P = prods.sum(axis=2)                                                        # per-instruction sums, exact
        exact = P.sum(axis=0)
        r = exact / sc[:, None] * inv[:, None]
        slope = 1.0 / (1.0 + np.exp(-r))                                             # softplus'
        wgt = (slope * inv[:, None] / sc[:, None])                                   # d(output) / d(accumulator)
        out_scale = np.abs(np.log1p(np.exp(r)) + h[rows]).max()

        def report(name, d):
            e = (d - exact) * wgt / out_scale
            print(f"  {name:28s} rms {np.sqrt((e ** 2).mean()):.2e}  max {np.abs(e).max():.2e}")

        print(f"layer {l}: rows {T.shape[0]}, frac D>0 {np.mean(exact > 0):.2f}")
        # fp32 round-to-nearest sequential accumulation of the instruction sums, as a yardstick
        acc = np.zeros_like(exact)
        for k in range(48):
            for j in range(16):
                acc = (acc + prods[k, :, j, :]).astype(np.float32).astype(np.float64)
        report("fp32 RN sequential (FFMA-like)", acc)
        for nseg in (1, 2, 4):
            d = rz_chain(prods, nseg=nseg)
            e = d - exact
            w2 = wgt * wgt
            c = -(w2 * e * exact).sum() / (w2 * exact * exact).sum()
            report(f"RZ nseg={nseg} raw", d)
            report(f"RZ nseg={nseg} const c={c / 2 ** -24:.1f}", d * (1 + c))
            if nseg == 1:
                nk = np.arange(48, 0, -1.0)                                            # n - k + 1
                X1 = (nk[:, None, None] * P).sum(axis=0)
                X0 = exact
                # two-parameter fit: e ~ -(c' X1 + b X0)
                Amat = np.stack([(wgt * X1).ravel(), (wgt * X0).ravel()], axis=1)
                sol, *_ = np.linalg.lstsq(Amat, -(wgt * e).ravel(), rcond=None)
                report(f"RZ pos c'={sol[0] / 2 ** -24:.2f} b={sol[1] / 2 ** -24:.2f}", d + sol[0] * X1 + sol[1] * X0)
                report("RZ pos theory c'=0.72 b=0", d + 0.7213 * 2 ** -24 * X1)
                Afull = (wgt[None] * P).reshape(48, -1).T
                solf, *_ = np.linalg.lstsq(Afull, -(wgt * e).ravel(), rcond=None)
                report("RZ fit48", d + np.tensordot(solf, P, axes=1))
                print("    gamma_k (x 2^-24):", np.round(solf / 2 ** -24, 1)[::6])


if __name__ == "__main__":
    main()
