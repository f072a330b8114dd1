"""Structural checks of the product path: WHICH kernels ran, not what they computed.  The file name sorts last on purpose:
under ``pytest -x`` a structural regression (a default flipped, a kernel renamed) must never stop the numerical parity tests
from running (round-1 lesson: one stale trace-shape assertion hid 21 parity tests).
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mlg():
    import multilevel_gnn_b200 as m
    m._cabi.lib()
    return m


def _gbm_model(mlg, batch_size=3):
    from multilevel_gnn_b200 import configs, synth
    args = configs.make_args("gbm")
    torch.manual_seed(3)
    model = mlg.MultilevelGNN(args)
    synth.multilevel_params(model)
    model.to(DEV).train()
    model.pathway_indexs = model.pathway_indexs.to(DEV)
    b = synth.multilevel_batch(batch_size=batch_size, seed=5).to(DEV)
    params = [p for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
    return model, args, b, params


def _trace(model, b, params):
    """One forward + backward: (repo kernel tags from _cabi.KernelTimer, aten op names from the CPU-side profiler)."""
    from multilevel_gnn_b200 import _cabi
    timer = _cabi.KernelTimer()
    _cabi.TIMER = timer
    try:
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU]) as prof:
            pred, feat = model(b)
            loss = pred.square().sum() + feat.square().mean()
            torch.autograd.grad(loss, params, allow_unused=True)
        torch.cuda.synchronize()
    finally:
        _cabi.TIMER = None
    return set(timer.summary()), [e.key for e in prof.key_averages()]


def test_default_path_runs_the_native_kernels(mlg):
    """The shipped defaults at the gbm shape: factored first layer, transform-first second layer, fused pool, native head
    GEMMs -- by the repo's own kernel tags."""
    model, args, b, params = _gbm_model(mlg)
    tags, ops = _trace(model, b, params)
    assert {"sage_rank1_fwd", "sage_rank1_bwd", "sage_aggr_fwd", "sage_aggr_bwd", "pool_fwd", "pool_bwd"} <= tags, tags


def test_no_library_activation_backward_on_the_sage_chain(mlg):
    """The LeakyReLU derivative of every SAGE layer is applied inside a repo kernel (consumer epilogue, GEMM epilogue or the
    layer's own backward kernel): no aten::leaky_relu_backward pass on the default path; the unfused chain has them."""
    model, args, b, params = _gbm_model(mlg)
    _, ops1 = _trace(model, b, params)
    type(model).FUSE_ACT_BACKWARD = False
    try:
        _, ops0 = _trace(model, b, params)
    finally:
        type(model).FUSE_ACT_BACKWARD = True
    n1 = sum("leaky_relu_backward" in k for k in ops1)
    n0 = sum("leaky_relu_backward" in k for k in ops0)
    assert n0 >= 1 and n1 == 0, (ops0, ops1)
