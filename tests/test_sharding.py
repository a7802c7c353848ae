"""CPU tests of the multi-GPU host logic (graph partition, the one all-gather, reassembly) with the gloo
backend and world_size 2; the NumPy oracle stands in for the CUDA forward (it is the checker here, the
CUDA path is covered by the -m gpu tests)."""
import os
import socket

import numpy as np
import pytest

from conftest import load_golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_plan_balances_and_covers():
    from nmrgnn_b200.sharding import ShardPlan
    rng = np.random.default_rng(0)
    sizes = rng.integers(20, 3000, size=37)
    offs = np.concatenate([[0], np.cumsum(sizes)])
    for world in (1, 2, 3, 8):
        plan = ShardPlan(offs, world)
        owned = np.concatenate(plan.owned)
        assert sorted(owned.tolist()) == list(range(37))                 # every graph exactly once
        assert plan.counts.sum() == offs[-1]
        assert plan.counts.max() - plan.counts.min() <= sizes.max()      # greedy balance bound
        idx = np.concatenate([plan.atom_index(r) for r in range(world)])
        assert np.array_equal(np.sort(idx), np.arange(offs[-1]))
        # reassembly inverts the partition
        full = rng.normal(size=int(offs[-1])).astype(np.float32)
        padded = np.zeros((world, plan.max_count), np.float32)
        for r in range(world):
            padded[r, :plan.counts[r]] = full[plan.atom_index(r)]
        assert np.array_equal(plan.scatter_back(padded), full)


def test_shard_plan_edge_cases():
    from nmrgnn_b200.sharding import ShardPlan
    plan = ShardPlan(np.array([0]), 2)                                    # no graphs
    assert plan.n_atoms == 0 and plan.max_count == 0
    assert plan.scatter_back(np.zeros((2, 0), np.float32)).shape == (0,)
    plan = ShardPlan(np.array([0, 5]), 4)                                 # fewer graphs than ranks
    assert sorted(plan.counts.tolist()) == [0, 0, 0, 5]
    with pytest.raises(ValueError):
        ShardPlan(np.array([1, 5]), 2)


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nmrgnn_b200.params import GNNParams, baseline_path
        from nmrgnn_b200.sharding import ShardedModel
        from oracle import forward as orc
        params = GNNParams.load(baseline_path())
        g = load_golden("prot3_batch")
        batch = (g["atoms"], g["nlist"], g["edges"], g["inv_degree"], g["graph_offsets"])
        sm = ShardedModel(local_forward=lambda t: orc.forward(params, *t))
        y = sm(batch)
        if rank == 0:
            np.save(out_path, y)
    finally:
        dist.destroy_process_group()


def test_sharded_forward_gloo_world2(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "peaks.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    y = np.load(out)
    g = load_golden("prot3_batch")
    assert y.shape == g["peaks"].shape
    # graphs are independent: the sharded result equals the single-process oracle bit for bit
    from nmrgnn_b200.params import GNNParams, baseline_path
    from oracle import forward as orc
    params = GNNParams.load(baseline_path())
    ref = orc.forward_per_graph(params, g["atoms"], g["nlist"], g["edges"], g["inv_degree"], g["graph_offsets"],
                                reference_order=False)
    assert np.array_equal(y, ref)
    np.testing.assert_allclose(y, g["peaks"], rtol=1e-4, atol=1e-4)


def _gpu_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import nmrgnn_b200
        from nmrgnn_b200.sharding import ShardedModel
        g = load_golden("prot3_batch")
        batch = (g["atoms"], g["nlist"], g["edges"], g["inv_degree"], g["graph_offsets"])
        m = nmrgnn_b200.load_model(device=rank)
        m.handle.set_option("tc_min_atoms", 0)
        sm = ShardedModel(m)                              # peer-memory reassembly (nmrgnn_forward_sharded)
        assert sm.collective == "peer"
        y_host = sm(batch)
        y_host2 = sm(batch)                               # second epoch: the other parity of the gather buffers
        dev = torch.device("cuda", rank)
        tb = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in batch[:4]) + (batch[4],)
        y_dev = sm(tb)                                    # device-resident in and out
        assert y_dev.is_cuda
        y_nccl = ShardedModel(m, collective="torch")(batch)   # comparison arm: one NCCL all-gather
        single = m(batch[:4])
        ok = (np.array_equal(y_host, y_host2) and np.array_equal(y_dev.cpu().numpy(), y_host)
              and np.array_equal(y_nccl, y_host) and np.array_equal(single, y_host))
        flag = torch.tensor([int(ok)], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            np.savez(out_path, y=y_host, ok=int(flag.item()))
        sm.close()
        m.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_cuda_forward_two_gpus(tmp_path):
    """World size 2 on two GPUs, NCCL process group: the CUDA forward on each rank's graphs + the peer-memory
    reassembly equal the single-GPU forward bit for bit on both ranks (host and device inputs, two consecutive
    epochs, and the torch.distributed all-gather arm)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    import torch.multiprocessing as mp
    out = str(tmp_path / "peaks.npz")
    mp.spawn(_gpu_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    g = load_golden("prot3_batch")
    assert int(z["ok"]) == 1
    tol = 1e-4 * np.abs(g["peaks_f64"]) + 1e-4
    assert np.all(np.abs(z["y"] - g["peaks_f64"]) <= tol)
