"""Graph-sharded evaluation over several GPUs (one process per GPU, torch.distributed).

The reference evaluates one graph per call on one device (nmrgnn/main.py:236-245); graphs never
interact (tf.gather indexes within one graph's node tensor, nmrgnn/layers.py:33), so a batch of
independent graphs shards by whole graphs with no data-path collective.  Every rank holds the full
batch description (or just its shard), runs the forward on the graphs it owns and ONE all-gather of
the float32 peaks (padded to the largest shard) reassembles the output in the original atom order.

    dist.init_process_group("nccl", ...)            # torchrun; gloo works for host tensors
    sm = ShardedModel(nmrgnn_b200.load_model(device=local_rank))
    peaks = sm(batch)                               # batch = (atoms, nlist, edges, inv_degree, graph_offsets)

`local_forward` is injectable so that the partition / reassembly logic is testable on CPU (gloo,
world_size 2) with the oracle standing in for the CUDA forward.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .workloads import shard_graphs, take_graphs

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


class ShardPlan:
    """Which graphs each rank owns and where their atoms sit in the full batch."""

    def __init__(self, graph_offsets: Sequence[int], world_size: int):
        self.offsets = np.asarray(graph_offsets, np.int64)
        if self.offsets.ndim != 1 or self.offsets.size < 1 or self.offsets[0] != 0 or np.any(np.diff(self.offsets) < 0):
            raise ValueError("graph_offsets must start at 0 and be non-decreasing")
        self.world_size = int(world_size)
        self.owned: List[np.ndarray] = shard_graphs(self.offsets, self.world_size)
        sizes = np.diff(self.offsets)
        self.counts = np.array([int(sizes[g].sum()) for g in self.owned], np.int64)   # atoms per rank
        self.max_count = int(self.counts.max()) if self.counts.size else 0

    @property
    def n_atoms(self) -> int:
        return int(self.offsets[-1])

    def atom_index(self, rank: int) -> np.ndarray:
        """Positions (in the full batch) of the atoms of `rank`'s shard, in shard order."""
        parts = [np.arange(self.offsets[g], self.offsets[g + 1]) for g in self.owned[rank]]
        return np.concatenate(parts) if parts else np.zeros(0, np.int64)

    def scatter_back(self, gathered: np.ndarray) -> np.ndarray:
        """[world, max_count] padded per-rank peaks -> [n_atoms] in the original order."""
        gathered = np.asarray(gathered).reshape(self.world_size, self.max_count)
        out = np.empty(self.n_atoms, gathered.dtype)
        for r in range(self.world_size):
            out[self.atom_index(r)] = gathered[r, :self.counts[r]]
        return out


class ShardedModel:
    def __init__(self, model=None, local_forward: Optional[Callable] = None, group=None):
        if model is None and local_forward is None:
            raise ValueError("need a model or a local_forward callable")
        self.model = model
        self._forward = local_forward if local_forward is not None else (lambda g: model(g))
        self.group = group

    def _world(self) -> Tuple[int, int]:
        if dist is not None and dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def __call__(self, batch) -> np.ndarray:
        atoms, nlist, edges, inv, offs = batch
        rank, world = self._world()
        plan = ShardPlan(offs, world)
        local = take_graphs((atoms, nlist, edges, inv, np.asarray(offs, np.int64)), plan.owned[rank])
        y = np.asarray(self._forward(local[:4]), np.float32).reshape(-1)
        if y.shape[0] != plan.counts[rank]:
            raise RuntimeError("local forward returned the wrong number of peaks")
        if world == 1:
            return plan.scatter_back(y[None, :])
        # the one collective: all-gather of the padded peaks
        backend = dist.get_backend(self.group)
        use_cuda = backend == "nccl"
        dev = torch.device("cuda", self.model.device) if use_cuda and self.model is not None else (
            torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu"))
        send = torch.zeros(plan.max_count, dtype=torch.float32, device=dev)
        send[:y.shape[0]] = torch.from_numpy(y).to(dev)
        recv = torch.empty(world * plan.max_count, dtype=torch.float32, device=dev)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        return plan.scatter_back(recv.cpu().numpy())
