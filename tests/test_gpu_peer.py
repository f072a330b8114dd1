"""The fused data-parallel update (reduce-scatter -> Adam -> all-gather over peer memory, csrc/peer_adam.cu) against
torch.optim.Adam on the averaged gradient.  One GPU is enough to exercise the whole protocol: several arenas of ONE
process stand in for the ranks and their kernels run concurrently on separate streams, synchronising through the same
system-scope flags they use across NVLink.  (The real multi-process path is bench.py --gpus N.)"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


@pytest.fixture(scope="module")
def mlg():
    import multilevel_gnn_b200 as m
    m._cabi.lib()
    return m


def _virtual_ranks(mlg, shapes, world, seed, wd):
    from multilevel_gnn_b200.train import GradBucket, PeerAdam, PeerArena
    g = torch.Generator().manual_seed(seed)
    init = [torch.randn(s, generator=g) for s in shapes]
    n = sum(t.numel() for t in init)
    arenas = [PeerArena(n, world, r, DEV, exchange=False) for r in range(world)]
    for a in arenas:
        a.peer_base = [b.base for b in arenas]
    ranks = []
    for r in range(world):
        params = [torch.nn.Parameter(t.clone().to(DEV)) for t in init]
        bucket = GradBucket(params, flat=arenas[r].grad)
        opt = PeerAdam(bucket, arenas[r], lr=1e-2, betas=(0.9, 0.999), weight_decay=wd, timeout_s=2.0)
        ranks.append((params, bucket, opt, torch.cuda.Stream()))
    return init, arenas, ranks, g


@pytest.mark.parametrize("world,wd", [(1, 0.0), (1, 1e-2)])
def test_peer_adam_matches_torch_adam(mlg, world, wd):
    shapes = [(257, 33), (1000,), (64, 64), (3,), (50001,)]
    init, arenas, ranks, g = _virtual_ranks(mlg, shapes, world, 7 + world, wd)
    ref = [torch.nn.Parameter(t.clone().double()) for t in init]
    ref_opt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999), weight_decay=wd)
    for step in range(4):
        grads = [[torch.randn(s, generator=g) for s in shapes] for _ in range(world)]
        for p, gs in zip(ref, zip(*grads)):
            p.grad = torch.stack([x.double() for x in gs]).sum(0) / world
        ref_opt.step()
        for (params, bucket, opt, stream), gr in zip(ranks, grads):
            bucket.store([x.to(DEV) for x in gr])
        torch.cuda.synchronize()
        for params, bucket, opt, stream in ranks:      # all ranks' kernels in flight at once
            with torch.cuda.stream(stream):
                opt.step()
        torch.cuda.synchronize()
        assert all(a.status() == 0 for a in arenas), "a rank timed out waiting for its peers"
    for r, (params, bucket, opt, stream) in enumerate(ranks):
        for p, q, p0 in zip(params, ref, ranks[0][0]):
            assert torch.equal(p.data, p0.data), "replicas must stay bitwise identical"
            torch.testing.assert_close(p.data.cpu().double(), q.data, rtol=2e-5, atol=2e-6)
        assert float(opt.step_dev.item()) == 4.0


def test_peer_adam_in_cuda_graph(mlg):
    """The update kernel replays inside a CUDA graph (device-side step counter and epoch)."""
    shapes = [(4096,), (31, 7)]
    init, arenas, ranks, g = _virtual_ranks(mlg, shapes, 1, 3, 0.0)
    params, bucket, opt, stream = ranks[0]
    ref = [torch.nn.Parameter(t.clone().double()) for t in init]
    ref_opt = torch.optim.Adam(ref, lr=1e-2, betas=(0.9, 0.999))
    static = [torch.zeros(s, device=DEV) for s in shapes]
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        bucket.store(static)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        bucket.store(static)
        opt.step()
    # the capture itself does not run the kernel: three replays = three steps
    for step in range(3):
        gr = [torch.randn(s, generator=g) for s in shapes]
        for p, x, st in zip(ref, gr, static):
            p.grad = x.double()
            st.copy_(x)
        ref_opt.step()
        graph.replay()
    torch.cuda.synchronize()
    assert arenas[0].status() == 0
    for p, q in zip(params, ref):
        torch.testing.assert_close(p.data.cpu().double(), q.data, rtol=2e-5, atol=2e-6)
