#!/usr/bin/env python
"""Export the reference's pretrained inference tensors into this repo's own
weights file (nmrgnn_b200/models/baseline/baseline.npz).

Runs only in the build container (needs /root/reference).  The GPU box has no
/root/reference, so load_model() reads the exported file; the TensorBundle
reader stays available for user checkpoints (load_model(model_file=...)).

Source: nmrgnn/models/baseline/variables/variables.{index,data-00000-of-00001}
(MIT-licensed reference, TF 2.3.2 SavedModel); 23 float32 tensors, 1 070 477
parameters; Adam slots and metric scalars are dropped.  peak_std / peak_avg and
the RBF grid are cross-checked against the Const nodes baked in saved_model.pb.
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nmrgnn_b200.params import GNNParams, baseline_path, baseline_standards, rbf_centers  # noqa: E402
from oracle.savedmodel_interp import SavedModelInterpreter, reference_dir  # noqa: E402


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main():
    src = reference_dir()
    it = SavedModelInterpreter(src)
    fn = it.funcs[[k for k in it.funcs if k.startswith("__inference__wrapped_model")][0]]
    consts = {n.name: it._make_ndarray(n.attr["value"].tensor) for n in fn.node_def if n.op == "Const"}
    std, avg = consts["gnn-model/mul_3/y"], consts["gnn-model/mul_4/y"]
    bstd, bavg = baseline_standards(10)
    assert np.array_equal(std.view(np.uint32), bstd.view(np.uint32)), "peak_std constant drifted"
    assert np.array_equal(avg.view(np.uint32), bavg.view(np.uint32)), "peak_avg constant drifted"
    top = {n.name: it._make_ndarray(n.attr["value"].tensor) for n in it.mg.graph_def.node
           if n.op == "Const" and n.attr["dtype"].type == 1}
    c, gap = rbf_centers(0.005, 0.20, 128)
    assert np.array_equal(top["Const"].view(np.uint32), c.view(np.uint32)), "RBF centres differ"
    assert top["Const_1"] == gap, "RBF gap differs"

    p = GNNParams.from_tf_checkpoint(src, peak_std=std, peak_avg=avg)
    p.meta.update(
        source="ur-whitelab/nmrgnn nmrgnn/models/baseline (TF 2.3.2 SavedModel, MIT licence)",
        sha256_index=sha(os.path.join(src, "variables", "variables.index")),
        sha256_data=sha(os.path.join(src, "variables", "variables.data-00000-of-00001")),
        sha256_saved_model=sha(os.path.join(src, "saved_model.pb")),
        hypers=dict(atom_feature_size=256, edge_feature_size=3, edge_hidden_size=128, mp_layers=4,
                    fc_layers=4, edge_fc_layers=4, noise=0.025, dropout=True, neighbor_number=16),
    )
    out = baseline_path()
    os.makedirs(os.path.dirname(out), exist_ok=True)
    p.save(out)
    q = GNNParams.load(out)
    n = sum(W.size + b.size for W, b in q.edge_fc) + q.embed.size + sum(w.size for w in q.mp_w) \
        + sum(W.size + b.size for W, b in q.fc) + q.out[0].size + q.out[1].size
    print(f"wrote {out}: {n} parameters, {os.path.getsize(out)} bytes")
    assert n == 1070477


if __name__ == "__main__":
    main()
