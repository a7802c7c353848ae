"""GPU diagnostic: stage times (CUDA events inside nmrgnn_forward) and the error against the committed fp64 golden peaks
of BASELINE config 2 (or 3) with runtime options given as name=value arguments; `--` separates option sets.
  python tools/diag_stage_times.py [config=3] mp_l1_prefetch=1 -- mp_single_acc=1"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nmrgnn_b200  # noqa: E402
from nmrgnn_b200 import workloads, _capi  # noqa: E402


def main():
    args = sys.argv[1:]
    cfg = 2
    if args and args[0].startswith("config="):
        cfg = int(args.pop(0).split("=")[1])
    sets = [[]]
    for a in args:
        if a == "--":
            sets.append([])
        else:
            sets[-1].append(a)
    sets = [[]] + [s for s in sets if s]
    m = nmrgnn_b200.load_model()
    h = m.handle
    h.set_option("tc_min_atoms", 0)
    if cfg == 2:
        atoms, nlist, edges, inv, offs = workloads.protein_batch(64, first_seed=0)
    else:
        atoms, nlist, edges, inv, offs = workloads.small_molecule_batch(1024, first_seed=0)
    ref = np.load(os.path.join(ROOT, "tests", "golden", f"full_config{cfg}.npz"))["peaks_f64"]
    g = (atoms, nlist, edges, inv)
    dev = torch.device("cuda", 0)
    n, k = atoms.shape[0], nlist.shape[1]
    d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in g]
    out = torch.empty(n, dtype=torch.float32, device=dev)
    s = int(torch.cuda.current_stream().cuda_stream) or 1
    for opts in sets:
        for o in opts:
            name, _, v = o.partition("=")
            h.set_option(name, int(v))
        y = m(g).astype(np.float64)
        e = np.abs(y - ref) / (1e-4 * np.abs(ref) + 1e-4)
        h.set_option("profile", 1)
        ts = []
        for it in range(10):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, k, out, _capi.MEM_DEVICE, s)
            h.synchronize(s)
            st = h.stage_times()
            ts.append([st["edge"], st["embed"]] + list(st["mp_layers"]) + [st["fc_readout"]])
        h.set_option("profile", 0)
        t = np.median(np.array(ts[3:]), axis=0)
        print(f"{' '.join(opts) or 'defaults':40s}: stage ms {np.round(t, 4).tolist()} total {t.sum():.4f} | tol_ratio max "
              f"{e.max():.3f} p99.99 {np.quantile(e, 0.9999):.3f} rms {np.sqrt(np.mean(e * e)):.4f}", flush=True)
        for o in opts:
            name = o.partition("=")[0]
            h.set_option(name, {"fc_pipe": 1, "edge_table": 1, "mp_chain_segments": 1}.get(name, 0))


if __name__ == "__main__":
    main()
