"""nmrgnn_b200 — B200-native (sm_100a) forward path of the nmrgnn chemical-shift GNN.

Drop-in for the reference's inference surface (nmrgnn/__init__.py:26-31):
``load_model``, ``universe2graph``, ``check_peaks`` and the layer class names in
``custom_objects``.
"""
__version__ = "0.1.0"

from .params import GNNParams  # noqa: F401
from .model import (EdgeFCBlock, FCBlock, GNNModel, MPBlock, MPLayer, RBFExpansion,  # noqa: F401
                    build_GNNModel)
from .library import (check_peaks, load_baseline, load_embeddings, load_model, load_standards,  # noqa: F401
                      universe2graph)
from .graph import Universe, batch_graphs, build_graph, read_pdb  # noqa: F401
from .evalstruct import eval_struct  # noqa: F401
from .batchstream import BatchStream  # noqa: F401

custom_objects = {c.__name__: c for c in (MPLayer, RBFExpansion, EdgeFCBlock, MPBlock, FCBlock)}
