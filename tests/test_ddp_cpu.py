"""Host-side multi-GPU logic on CPU: graph sharding and the flat-bucket gradient all-reduce over a
world_size-2 gloo group (the N > 1 path of bench.py / training)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_graphs_balanced_and_complete():
    from x2gnn_b200 import ddp, synth
    b = synth.qm9_batch(32, seed=2)
    tri = synth.triplets_host(b["edge_index"], len(b["x"]))[0]
    e_graph = np.repeat(np.arange(32), b["edge_num"])          # graph id of each bond
    costs = np.bincount(e_graph[tri[1]], minlength=32)          # triplets per molecule
    for world in (1, 2, 4, 8):
        parts = ddp.shard_graphs(costs.tolist(), world)
        flat = sorted(g for p in parts for g in p)
        assert flat == list(range(32))                          # every molecule exactly once
        loads = [int(costs[p].sum()) for p in parts]
        assert max(loads) - min(loads) <= int(costs.max())      # LPT bound
        assert ddp.shard_graphs(costs.tolist(), world) == parts  # deterministic
    with pytest.raises(ValueError):
        ddp.shard_graphs([1, 2], 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from x2gnn_b200 import ddp
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    for i, p in enumerate(net.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    list(net.parameters())[1].grad = None                       # an unused parameter
    bucket = ddp.FlatGradBucket(net.parameters())
    bucket.pack()
    bucket.allreduce(average=True)
    bucket.unpack()
    res = [p.grad.clone() for p in net.parameters()]
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_qm9_shard_partitions_the_global_batch():
    """bench.py's multi-GPU workload: every rank generates the same global list and keeps whole molecules;
    world == 1 is the plain seed-0 batch; the shares are disjoint, complete and balanced by triplets."""
    from x2gnn_b200 import synth
    one, ids = synth.qm9_shard(12, 1, 0, seed=3)
    ref = synth.qm9_batch(12, seed=3)
    assert ids == list(range(12))
    for k, v in ref.items():
        if hasattr(v, "shape"):
            assert np.array_equal(one[k], v), k
    world = 3
    seen, loads = [], []
    for r in range(world):
        b, mine = synth.qm9_shard(12, world, r, seed=3)
        seen += mine
        assert b["num_graphs"] == len(mine) and int(b["edge_num"].sum()) == b["edge_index"].shape[1]
        assert int(b["edge_index"].max()) < len(b["x"])          # node ids are local to the rank's batch
        loads.append(synth.triplets_host(b["edge_index"], len(b["x"]))[0].shape[1])
    assert sorted(seen) == list(range(12 * world))
    assert max(loads) - min(loads) < 0.15 * max(loads), loads


def test_flat_bucket_allreduce_gloo_world2(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out)
    mean_rank = (1 + 2) / 2
    for i, g in enumerate(res):
        want = 0.0 if i == 1 else mean_rank * (i + 1)
        assert torch.allclose(g, torch.full_like(g, want))
