"""Data-parallel host logic on CPU: world_size-2 gloo processes.

The model forward needs a GPU (no CPU fallback), so these tests drive the REAL ``GradBucket`` /
``shard_indices`` code with a small stand-in nn.Module: the averaged per-shard gradients after one
all-reduce must equal the full-batch gradient, replicas must stay bit-identical after an optimizer step,
and the epoch partition must be disjoint and complete."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _net():
    torch.manual_seed(7)
    return torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.LeakyReLU(0.2), torch.nn.Linear(5, 2), torch.nn.Softmax(dim=1))


def _data():
    g = torch.Generator().manual_seed(3)
    return torch.randn(8, 6, generator=g), torch.nn.functional.one_hot(torch.randint(0, 2, (8,), generator=g), 2).float()


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from multilevel_gnn_b200.train import GradBucket, shard_indices
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = _net()
    x, y = _data()
    params = list(net.parameters())
    bucket = GradBucket(params)
    opt = torch.optim.Adam(params, lr=1e-2)
    idx = shard_indices(8, rank, world, per_rank_batch=4, epoch_seed=5)[0]
    bucket.zero()
    loss = torch.nn.BCELoss()(net(x[idx]), y[idx])
    loss.backward()
    bucket.all_reduce(world)
    grads = bucket.flat.clone()
    opt.step()
    q.put((rank, idx.tolist(), grads.tolist(), torch.cat([p.detach().reshape(-1) for p in params]).tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_allreduce_matches_full_batch():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, i0, g0, w0), (_, i1, g1, w1) = res
    g0, g1, w0, w1 = (torch.tensor(v) for v in (g0, g1, w0, w1))
    assert sorted(i0 + i1) == list(range(8)) and not set(i0) & set(i1)
    assert torch.equal(g0, g1) and torch.equal(w0, w1)            # replicas stay identical
    net = _net()
    x, y = _data()
    # mean of the two shard means == full-batch mean (equal shard sizes)
    torch.nn.BCELoss()(net(x), y).backward()
    full = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert torch.allclose(g0, full, rtol=1e-5, atol=1e-7)


def test_shard_indices_partition():
    from multilevel_gnn_b200.train import shard_indices
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            batches = shard_indices(300, r, world, per_rank_batch=8, epoch_seed=11)
            assert len(batches) == 300 // (8 * world)
            assert all(len(b) == 8 for b in batches)
            seen += [int(i) for b in batches for i in b]
        assert len(seen) == len(set(seen))
    a = shard_indices(100, 0, 2, 4, epoch_seed=1)
    b = shard_indices(100, 0, 2, 4, epoch_seed=2)
    assert not all(torch.equal(u, v) for u, v in zip(a, b))       # reshuffled every epoch


def test_bucket_skips_unused_parameters():
    """lin_l.weight of every SAGEConv never gets a gradient: it must stay out of the bucket."""
    import multilevel_gnn_b200 as m
    from multilevel_gnn_b200.train import Trainer
    model = m.MultilevelGNN(m.configs.make_args("gbm", conv_channel_list=[8, 4], head_dim=16))
    m.synth.multilevel_params(model)
    tr = Trainer(model, model.args, None, world_size=1, fused_adam=False)
    names = {n for n, p in model.named_parameters() if any(p is q for q in tr.params)}
    assert not any(n.endswith("lin_l.weight") for n in names) and "info_mask" not in names
    assert tr.flat.numel() == sum(p.numel() for p in tr.params)
    assert all(p.grad.data_ptr() >= tr.flat.data_ptr() for p in tr.params)
