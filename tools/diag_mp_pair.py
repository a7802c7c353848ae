"""GPU diagnostic: CTA-pair (cta_group::2) form of the MP-layer kernel against the one-CTA form: bit-equality of one
layer and of the whole forward, launch times of both (CUDA events, same stream)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nmrgnn_b200
from nmrgnn_b200 import workloads, _capi

m = nmrgnn_b200.load_model()
h = m.handle
h.set_option("tc_min_atoms", 0)
dev = torch.device("cuda", 0)
s = int(torch.cuda.current_stream().cuda_stream) or 1
for n_graphs in (1, 3, 64):          # odd and even tile counts, then BASELINE configs[1]
    atoms, nlist, edges, inv, offs = workloads.protein_batch(n_graphs, first_seed=0)
    n = atoms.shape[0]
    d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)]
    g = torch.Generator(device="cpu").manual_seed(1)
    hh = (torch.randn(n, 256, generator=g) * 0.5 + 1.0).to(dev)
    ef = (torch.randn(n, 16, 3, generator=g) * 0.1).to(dev)
    res = {}
    for pair in (0, 1):
        h.set_option("mp_pair", pair)
        o = torch.empty_like(hh)
        h.mp_layer(1, hh, d_in[1], ef, d_in[3], n, 16, o, _capi.MEM_DEVICE, s)
        h.synchronize(s)
        out = torch.empty(n, dtype=torch.float32, device=dev)
        for _ in range(3):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
        h.synchronize(s)
        h.set_option("profile", 1)
        ts = []
        for _ in range(5):
            h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, 16, out, _capi.MEM_DEVICE, s)
            h.synchronize(s)
            st = h.stage_times()
            ts.append([st['edge'], st['embed']] + list(st['mp_layers']) + [st['fc_readout']])
        h.set_option("profile", 0)
        res[pair] = (o.cpu().numpy(), out.cpu().numpy(), np.median(np.array(ts), axis=0))
    same_layer = np.array_equal(res[0][0], res[1][0])
    same_fwd = np.array_equal(res[0][1], res[1][1])
    print(f"n_atoms {n} tiles {(n + 127) // 128}: layer bit-equal {same_layer} (max diff {np.abs(res[0][0] - res[1][0]).max():.3e}), "
          f"forward bit-equal {same_fwd}; stage ms one-CTA {np.round(res[0][2], 4).tolist()} pair {np.round(res[1][2], 4).tolist()}")
