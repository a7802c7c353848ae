"""GPU diagnostic: where the roles of the edge-MLP kernel wait (cycle counters of one launch)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import nmrgnn_b200
from nmrgnn_b200 import workloads, _capi

b = workloads.protein_batch(64, first_seed=0)
atoms, nlist, edges, inv, offs = b
n = atoms.shape[0]
m = nmrgnn_b200.load_model()
h = m.handle
dev = torch.device("cuda", 0)
d_e = torch.from_numpy(np.ascontiguousarray(edges)).to(dev)
out = torch.empty((n, 16, 3), dtype=torch.float32, device=dev)
s = int(torch.cuda.current_stream().cuda_stream) or 1
for it in range(3):
    h.edge_features(d_e, n * 16, out, _capi.MEM_DEVICE, s)
h.synchronize(s)
h.set_option("mp_role_counters", 1)
h.edge_features(d_e, n * 16, out, _capi.MEM_DEVICE, s)
h.synchronize(s)
h.set_option("mp_role_counters", 2)
print("(edge kernel: 'mma total' = MMA thread total; 'epilogue' column = MMA waiting for X; 'producers' column = MMA waiting "
      "for W; \"W'\" column = epilogue warp 2 waiting for D)")
