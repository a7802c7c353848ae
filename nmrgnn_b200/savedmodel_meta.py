"""What a reference-written TensorFlow SavedModel says about the model besides its weights, read without TensorFlow:
the hyper-parameters Keras recorded for ``GNNModel`` (activations, RBF range; nmrgnn/model.py:22-36, 206-220), the
``RBFExpansion`` layer config (nmrgnn/layers.py:131-135) and the per-element standards that ``GNNModel.call`` bakes
into the traced graph as constants (nmrgnn/model.py:222-228, 272-273).  ``nmrgnn.load_model`` gets all of this back
through ``tf.keras.models.load_model``; a loader that only read the checkpoint would silently predict wrong values for
a model trained with relu / tanh or with other standards.

Only the few protobuf fields involved are walked (field numbers of saved_model.proto, meta_graph.proto,
saved_object_graph.proto, graph.proto, function.proto, node_def.proto, attr_value.proto, tensor.proto)."""
from __future__ import annotations

import json
import os
from typing import Dict, Optional

import numpy as np

from .tensorbundle import _parse_shape, _proto_fields

_DT_FLOAT = 1


def _sub(buf: bytes, field: int):
    for f, wt, v in _proto_fields(buf):
        if f == field and wt == 2:
            yield v


def _tensor(buf: bytes) -> Optional[np.ndarray]:
    """TensorProto -> float32 array (dtype DT_FLOAT only)."""
    dtype, shape, content, vals = 0, (), None, []
    for f, wt, v in _proto_fields(buf):
        if f == 1 and wt == 0:
            dtype = v
        elif f == 2 and wt == 2:
            shape = _parse_shape(v)
        elif f == 4 and wt == 2:
            content = v
        elif f == 5 and wt == 2:                      # packed float_val
            vals.extend(np.frombuffer(v, "<f4").tolist())
        elif f == 5 and wt == 5:
            vals.append(float(np.frombuffer(v, "<f4")[0]))
    if dtype != _DT_FLOAT:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None:
        return np.frombuffer(content, "<f4").reshape(shape).copy()
    if len(vals) == 1 and n > 1:                       # splat encoding
        return np.full(shape, vals[0], np.float32)
    if len(vals) == n:
        return np.asarray(vals, np.float32).reshape(shape)
    return None


def read_savedmodel_meta(model_dir: str) -> Dict[str, object]:
    """{'hypers': {...} | None, 'rbf': {'low','high','count'} | None, 'constants': {node name: float32 array}} of the
    SavedModel in ``model_dir``; raises FileNotFoundError if there is no saved_model.pb."""
    pb = os.path.join(model_dir, "saved_model.pb")
    with open(pb, "rb") as f:
        buf = f.read()
    hypers, rbf, consts = None, None, {}
    for mg in _sub(buf, 2):                                   # SavedModel.meta_graphs
        for og in _sub(mg, 7):                                # MetaGraphDef.object_graph_def
            for node in _sub(og, 1):                          # SavedObjectGraph.nodes
                for uo in _sub(node, 4):                      # SavedObject.user_object
                    for md in _sub(uo, 3):                    # SavedUserObject.metadata (JSON written by Keras)
                        try:
                            meta = json.loads(md.decode("utf-8"))
                        except (UnicodeDecodeError, ValueError):
                            continue
                        cls, cfg = meta.get("class_name"), meta.get("config") or {}
                        if cls == "RBFExpansion" and {"low", "high", "count"} <= set(cfg):
                            rbf = {k: cfg[k] for k in ("low", "high", "count")}
                        elif cls in ("GNNModel", "EdgeFCBlock", "MPBlock", "FCBlock") and hypers is None:
                            # (the blocks carry the model's kerastuner.HyperParameters in their configs,
                            #  nmrgnn/model.py:140-144, 171-175, 198-202)
                            hp = cfg.get("hypers") or {}
                            vals = (hp.get("config") or {}).get("values") if isinstance(hp, dict) else None
                            if vals:
                                hypers = dict(vals)
        for gd in _sub(mg, 2):                                # MetaGraphDef.graph_def
            for lib in _sub(gd, 2):                           # GraphDef.library
                for fn in _sub(lib, 1):                       # FunctionDefLibrary.function
                    for nd in _sub(fn, 3):                    # FunctionDef.node_def
                        name, op, value = "", "", None
                        for f, wt, v in _proto_fields(nd):
                            if f == 1 and wt == 2:
                                name = v.decode()
                            elif f == 2 and wt == 2:
                                op = v.decode()
                            elif f == 5 and wt == 2:          # attr map entry
                                key, av = "", None
                                for f2, wt2, v2 in _proto_fields(v):
                                    if f2 == 1:
                                        key = v2.decode()
                                    elif f2 == 2:
                                        av = v2
                                if key == "value" and av is not None:
                                    for t in _sub(av, 8):     # AttrValue.tensor
                                        value = t
                        if op == "Const" and value is not None and name.endswith(("/mul_3/y", "/mul_4/y")) \
                                and name not in consts:
                            arr = _tensor(value)
                            if arr is not None and arr.ndim == 1:
                                consts[name] = arr
    return {"hypers": hypers, "rbf": rbf, "constants": consts}


def standards_from_constants(consts: Dict[str, np.ndarray], num_elem: int):
    """(peak_std, peak_avg) of ``GNNModel.call``'s readout: the second operands of its ``mul_3`` (x peak_std) and
    ``mul_4`` (x peak_avg) nodes; None if the graph does not have exactly one of each with ``num_elem`` entries."""
    std = [v for k, v in consts.items() if k.endswith("/mul_3/y") and v.shape == (num_elem,)]
    avg = [v for k, v in consts.items() if k.endswith("/mul_4/y") and v.shape == (num_elem,)]
    if len(std) != 1 or len(avg) != 1:
        return None
    return std[0], avg[0]
