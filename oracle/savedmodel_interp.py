"""TEST INFRASTRUCTURE — not product code.  Only tests/, tools/ fixture generators,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this package.

NumPy interpreter for the reference's *own traced inference graph*: it parses
``nmrgnn/models/baseline/saved_model.pb`` (TF SavedModel written by TF 2.3.2 /
Keras 2.4.0), follows the ``serving_default`` signature into
``__inference_signature_wrapper_*`` -> ``__inference__wrapped_model_4657940`` and
executes every node (189 of them, 25 op types) with NumPy kernels, reading the
variables from the TensorBundle next to it.  TensorFlow itself cannot be
installed in this image, so this is the closest available
"run the reference here": op order, einsum lowering
(``lmn,ijl->mnij`` / ``mnij,ijn->mi`` / ``mi,i->im``), activations and the baked
constants (RBF centres, gap, peak_std, peak_avg) all come from the reference's
artefact, not from our reading of nmrgnn/model.py.  Only the per-op kernels
(MatMul, Einsum, Softplus, Exp, ...) are restated, as plain IEEE arithmetic.

It needs /root/reference (or any copy of the SavedModel) and therefore runs only
in the build container: tools/make_golden.py uses it to write tests/golden/*.npz.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from nmrgnn_b200.tensorbundle import TensorBundle, _proto_fields, resolve_prefix

_SUFFIX = "/.ATTRIBUTES/VARIABLE_VALUE"


def tf_softplus(x: np.ndarray) -> np.ndarray:
    """tensorflow/core/kernels/softplus_op.h functor::Softplus: thresholded
    log(exp(x)+1) with threshold = log(eps)+2."""
    eps = np.finfo(x.dtype).eps
    thr = np.log(x.dtype.type(eps)) + x.dtype.type(2)
    with np.errstate(over="ignore"):
        ex = np.exp(x)
    mid = np.log(ex + x.dtype.type(1))
    return np.where(x > -thr, x, np.where(x < thr, ex, mid)).astype(x.dtype)


def _strided_slice(x, begin, end, strides, attr):
    def mask(name):
        return int(attr[name].i) if name in attr else 0
    bm, em, elm, nam, sam = (mask("begin_mask"), mask("end_mask"), mask("ellipsis_mask"),
                             mask("new_axis_mask"), mask("shrink_axis_mask"))
    idx: List[object] = []
    for i in range(len(begin)):
        bit = 1 << i
        if elm & bit:
            idx.append(Ellipsis)
        elif nam & bit:
            idx.append(np.newaxis)
        elif sam & bit:
            idx.append(int(begin[i]))
        else:
            b = None if bm & bit else int(begin[i])
            e = None if em & bit else int(end[i])
            idx.append(slice(b, e, int(strides[i])))
    return x[tuple(idx)]


class SavedModelInterpreter:
    def __init__(self, saved_model_dir: str, dtype=np.float32):
        from tensorboard.compat.proto import meta_graph_pb2
        from tensorboard.util import tensor_util

        self._make_ndarray = tensor_util.make_ndarray
        self.dtype = np.dtype(dtype)
        raw = open(os.path.join(saved_model_dir, "saved_model.pb"), "rb").read()
        mg = None
        for field, _, val in _proto_fields(raw):
            if field == 2:
                mg = meta_graph_pb2.MetaGraphDef()
                mg.ParseFromString(val)
        if mg is None:
            raise ValueError("no MetaGraphDef in saved_model.pb")
        self.mg = mg
        self.funcs = {f.signature.name: f for f in mg.graph_def.library.function}
        self.top = {n.name: n for n in mg.graph_def.node}
        # variables: VarHandleOp shared_name -> value, through the checkpoint's object graph
        tb = TensorBundle(resolve_prefix(saved_model_dir))
        by_name: Dict[str, np.ndarray] = {}
        names = tb.object_graph_names()
        seen: Dict[str, int] = {}
        for key in tb.keys():  # sorted order == creation order for variables/<i>
            pass
        # full_name is not unique for MPLayer/w (Keras reuses the name); the graph
        # uniquifies them as w, w_1, w_2, w_3 in creation order == variables/8..11.
        def order(k):
            short = k[:-len(_SUFFIX)] if k.endswith(_SUFFIX) else k
            parts = short.split("/")
            return (parts[0], int(parts[1]) if len(parts) > 1 and parts[1].isdigit() else -1, short)
        for key in sorted(names, key=order):
            if ".OPTIMIZER_SLOT" in key or key not in tb.entries:
                continue
            if tb.entries[key].dtype not in (1, 2, 3, 9):
                continue
            full = names[key]
            n = seen.get(full, 0)
            seen[full] = n + 1
            by_name[full if n == 0 else f"{full}_{n}"] = tb.read(key)
        self.variables = by_name
        self.trace: Dict[str, np.ndarray] = {}
        self.keep_trace = False

    # ------------------------------------------------------------------ public
    def __call__(self, atoms, nlist, edges, inv_degree) -> np.ndarray:
        sig = self.mg.signature_def["serving_default"]
        feeds = {
            sig.inputs["input_1"].name: np.asarray(atoms, self.dtype),
            sig.inputs["input_2"].name: np.asarray(nlist, np.int32),
            sig.inputs["input_3"].name: np.asarray(edges, self.dtype),
            sig.inputs["input_4"].name: np.asarray(inv_degree, self.dtype),
        }
        out_name = sig.outputs["output_1"].name
        cache: Dict[str, List[object]] = {}
        return self._eval_top(out_name, feeds, cache)

    # ---------------------------------------------------------------- plumbing
    def _eval_top(self, tensor_name: str, feeds, cache):
        if tensor_name in feeds:
            return feeds[tensor_name]
        node_name, _, idx = tensor_name.partition(":")
        idx = int(idx or 0)
        if node_name + ":0" in feeds and idx == 0:
            return feeds[node_name + ":0"]
        if node_name not in cache:
            node = self.top[node_name]
            ins = [self._eval_top(i, feeds, cache) for i in node.input if not i.startswith("^")]
            cache[node_name] = self._run_node(node, ins, prefix="")
        return cache[node_name][idx]

    def _run_function(self, fname: str, args: Sequence[object], prefix: str) -> List[object]:
        fn = self.funcs[fname]
        env: Dict[str, List[object]] = {}
        for a, v in zip(fn.signature.input_arg, args):
            env[a.name] = [v]
        nodes = {n.name: n for n in fn.node_def}

        def get(ref: str):
            parts = ref.split(":")
            name = parts[0]
            idx = int(parts[-1]) if len(parts) > 1 else 0
            if name not in env:
                node = nodes[name]
                ins = [get(i) for i in node.input if not i.startswith("^")]
                env[name] = self._run_node(node, ins, prefix=prefix + fname + "/")
            return env[name][idx]

        return [get(fn.ret[o.name]) for o in fn.signature.output_arg]

    # ------------------------------------------------------------------ kernels
    def _run_node(self, node, ins, prefix: str) -> List[object]:
        op = node.op
        a = node.attr
        dt = self.dtype

        def fl(x):  # cast float tensors to the working dtype
            x = np.asarray(x)
            return x.astype(dt) if x.dtype.kind == "f" else x

        if op == "Const":
            out = fl(self._make_ndarray(a["value"].tensor))
        elif op == "Placeholder":
            raise KeyError(f"unfed placeholder {node.name}")
        elif op == "VarHandleOp":
            out = ("var", a["shared_name"].s.decode())
        elif op == "ReadVariableOp":
            out = fl(self.variables[ins[0][1]])
        elif op in ("StatefulPartitionedCall", "PartitionedCall"):
            return self._run_function(a["f"].func.name, ins, prefix)
        elif op in ("Identity", "StopGradient"):
            out = ins[0]
        elif op == "NoOp":
            return []
        elif op == "Greater":
            out = ins[0] > ins[1]
        elif op == "Cast":
            out = fl(np.asarray(ins[0]).astype(np.float32)) if a["DstT"].type == 1 else np.asarray(ins[0])
        elif op == "StridedSlice":
            out = _strided_slice(ins[0], ins[1], ins[2], ins[3], a)
        elif op == "Sub":
            out = ins[0] - ins[1]
        elif op == "AddV2":
            out = ins[0] + ins[1]
        elif op == "Mul":
            out = ins[0] * ins[1]
        elif op == "RealDiv":
            out = ins[0] / ins[1]
        elif op == "Neg":
            out = -ins[0]
        elif op == "Pow":
            out = np.power(ins[0], ins[1])
        elif op == "Exp":
            out = np.exp(ins[0])
        elif op == "Softplus":
            out = tf_softplus(ins[0])
        elif op == "BiasAdd":
            out = ins[0] + ins[1]
        elif op == "MatMul":
            x, y = ins
            if a["transpose_a"].b:
                x = x.T
            if a["transpose_b"].b:
                y = y.T
            out = x @ y
        elif op == "Einsum":
            out = np.einsum(a["equation"].s.decode(), *ins, optimize=True)
        elif op == "GatherV2":
            out = np.take(ins[0], ins[1], axis=int(ins[2]))
        elif op == "Shape":
            out = np.asarray(np.shape(ins[0]), np.int32)
        elif op == "Prod":
            out = np.prod(ins[0], axis=tuple(np.atleast_1d(ins[1]).tolist()),
                          keepdims=bool(a["keep_dims"].b)).astype(np.asarray(ins[0]).dtype)
        elif op == "Sum":
            out = np.sum(ins[0], axis=tuple(np.atleast_1d(ins[1]).tolist()),
                         keepdims=bool(a["keep_dims"].b), dtype=np.asarray(ins[0]).dtype)
        elif op == "ConcatV2":
            out = np.concatenate([np.atleast_1d(t) for t in ins[:-1]], axis=int(ins[-1]))
        elif op == "Pack":
            out = np.stack(ins, axis=int(a["axis"].i))
        elif op == "Reshape":
            out = np.reshape(ins[0], tuple(int(s) for s in ins[1]))
        elif op == "Transpose":
            out = np.transpose(ins[0], tuple(int(s) for s in ins[1]))
        else:
            raise NotImplementedError(f"op {op} ({node.name}) not supported by the oracle interpreter")
        if self.keep_trace and isinstance(out, np.ndarray):
            self.trace[node.name] = out
        return [out]


def reference_dir(root: Optional[str] = None) -> str:
    root = root or os.environ.get("NMRGNN_REFERENCE", "/root/reference")
    return os.path.join(root, "nmrgnn", "models", "baseline")
