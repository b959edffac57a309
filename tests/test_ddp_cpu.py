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


def test_qm9_shard_weak_scaling_batches():
    """bench.py's multi-GPU workload (weak scaling): rank 0 -- and world == 1 -- is the plain seed-0 batch; every
    other rank gets its own molecules with the SAME triplet count to within 0.5 %, so the per-rank work is
    the same at every N; node ids are local to the rank's batch; ids are disjoint global molecule numbers."""
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
        assert b["num_graphs"] == len(mine) == 12 and int(b["edge_num"].sum()) == b["edge_index"].shape[1]
        assert int(b["edge_index"].max()) < len(b["x"])          # node ids are local to the rank's batch
        loads.append(synth.triplets_host(b["edge_index"], len(b["x"]))[0].shape[1])
        if r == 0:
            assert np.array_equal(b["atom_pos"], ref["atom_pos"])
        again, _ = synth.qm9_shard(12, world, r, seed=3)
        assert np.array_equal(again["atom_pos"], b["atom_pos"])  # deterministic
    assert sorted(seen) == list(range(12 * world))
    assert max(abs(t - loads[0]) for t in loads) <= 0.005 * loads[0], loads


def _grad_worker(rank, world, port, out):
    """Ranks with DIFFERENT numbers of samples: per-rank mean loss x global_mean_scale, flat all-reduce."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from x2gnn_b200 import ddp
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 1))
    g = torch.Generator().manual_seed(1)
    xs, ys = torch.randn(11, 5, generator=g), torch.randn(11, generator=g)
    lo, hi = (0, 4) if rank == 0 else (4, 11)                  # 4 and 7 samples
    loss = torch.nn.functional.smooth_l1_loss(net(xs[lo:hi]).squeeze(1), ys[lo:hi])
    (loss * ddp.global_mean_scale(hi - lo, 11, world)).backward()
    bucket = ddp.FlatGradBucket(net.parameters())
    bucket.pack()
    bucket.allreduce(average=True)
    if rank == 0:
        ref = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.SiLU(), torch.nn.Linear(7, 1))
        ref.load_state_dict(net.state_dict())
        torch.nn.functional.smooth_l1_loss(ref(xs).squeeze(1), ys).backward()
        want = torch.cat([p.grad.reshape(-1) for p in ref.parameters()])
        torch.save({"got": bucket.flat.clone(), "want": want}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_equals_single_rank_on_the_same_global_batch(tmp_path):
    """SURVEY.md 4 item 5 on the host side: uneven shards + global_mean_scale + one flat all-reduce give the
    gradient of the single-process step on the whole batch (a mean of per-rank means would not)."""
    world, port, out = 2, _free_port(), str(tmp_path / "g.pt")
    mp.spawn(_grad_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out)
    assert torch.allclose(res["got"], res["want"], rtol=1e-5, atol=1e-7)


def test_flat_bucket_allreduce_gloo_world2(tmp_path):
    world, port, out = 2, _free_port(), str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    res = torch.load(out)
    mean_rank = (1 + 2) / 2
    for i, g in enumerate(res):
        want = 0.0 if i == 1 else mean_rank * (i + 1)
        assert torch.allclose(g, torch.full_like(g, want))
