"""GPU diagnostic: where the roles of the MP-layer kernel wait (cycle counters of the last MP launch of a forward)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nmrgnn_b200
from nmrgnn_b200 import workloads, _capi

SMALL = bool(os.environ.get("SMALL"))
b = workloads.small_molecule_batch(1024, first_seed=0) if SMALL else workloads.protein_batch(64, first_seed=0)
atoms, nlist, edges, inv, offs = b
K = nlist.shape[1]
n = atoms.shape[0]
m = nmrgnn_b200.load_model()
h = m.handle
if os.environ.get("NSPLIT"):
    h.set_option("mp_nsplit", int(os.environ["NSPLIT"]))
if os.environ.get("ONE"):
    h.set_option("mp_single_acc", 1)
if os.environ.get("NSEG"):
    h.set_option("mp_chain_segments", int(os.environ["NSEG"]))
dev = torch.device("cuda", 0)
d_in = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (atoms, nlist, edges, inv)]
out = torch.empty(n, dtype=torch.float32, device=dev)
s = int(torch.cuda.current_stream().cuda_stream) or 1
for it in range(3):
    h.forward(d_in[0], d_in[1], d_in[2], d_in[3], n, K, out, _capi.MEM_DEVICE, s)
h.synchronize(s)
h.set_option("mp_role_counters", 1)
# one MP layer only, so the counters belong to a single launch
hh = torch.randn(n, 256, device=dev) * 0.5 + 1.0
ef = torch.randn(n, K, 3, device=dev) * 0.1
o2 = torch.empty_like(hh)
h.mp_layer(1, hh, d_in[1], ef, d_in[3], n, K, o2, _capi.MEM_DEVICE, s)
h.synchronize(s)
h.set_option("mp_role_counters", 2)
