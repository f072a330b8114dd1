"""Algebra of the fused DiffPool kernel (csrc/diffpool_fused.cu) checked WITHOUT a GPU: the same source compiled as host
C++ (tests/host_emu/diffpool_emu.cpp, -DMLG_HOST_EMU: every 'parallel for + barrier' phase runs its items sequentially)
against the golden vectors produced by the reference's own DiffPool (models/diff_pooling.py:116-133) and against autograd
of the CPU oracle: pooled features, link / entropy terms, dL/dx and all 18 parameter gradients.  Test infrastructure only:
the product path is the CUDA build of that source (tests/test_gpu_parity.py::test_diffpool_golden)."""
import ctypes
import os
import subprocess

import pytest
import torch

from conftest import assert_close, load_golden
from oracle import restated as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ("gnn_pool.layers.0.lin_rel.weight", "gnn_pool.layers.0.lin_root.weight", "gnn_pool.layers.0.lin_root.bias",
        "gnn_embed.layers.0.lin_rel.weight", "gnn_embed.layers.0.lin_root.weight", "gnn_embed.layers.0.lin_root.bias")
AFTER = ("layers.0.lin_rel.weight", "layers.0.lin_root.weight", "layers.0.lin_root.bias")


def weight_names(layers):
    out = []
    for l in range(layers):
        out += ["diffpool_layers.%d.%s" % (l, k) for k in KEYS]
        out += ["after_pool_layers.%d.%s" % (l, k) for k in AFTER]
    return out


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libdiffpool_emu.so")
    src = os.path.join(ROOT, "tests", "host_emu", "diffpool_emu.cpp")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", so, src], check=True)
    lib = ctypes.CDLL(so)
    lib.emu_diffpool_smem_floats.restype = ctypes.c_long
    lib.emu_diffpool_state_floats.restype = ctypes.c_long
    return lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def run_emu(lib, sd, x, adj, dims, g_out=None, coef=None, state=None):
    """Forward: (out, stats, state) -- ``state`` is what the forward kernel stores for backward.  Backward (g_out given):
    with ``state`` the stored forward is read back, without it the forward pass is recomputed; both must agree."""
    layers = len(dims)
    names = weight_names(layers)
    ws = [sd[k].detach().float().contiguous() for k in names]
    warr = (ctypes.c_void_p * len(ws))(*[w.data_ptr() for w in ws])
    darr = (ctypes.c_int64 * (4 * layers))(*[v for d in dims for v in d])
    b = x.shape[0]
    x = x.detach().float().contiguous()
    adj = adj.float().contiguous()
    if g_out is None:
        out = torch.zeros(b, dims[-1][2], dims[-1][3])
        stats = torch.zeros(b, 2 * layers)
        st = torch.zeros(b * lib.emu_diffpool_state_floats(layers, darr))
        lib.emu_diffpool_fwd(_ptr(x), _ptr(adj), warr, layers, darr, b, _ptr(out), _ptr(stats), _ptr(st))
        return out, stats, st
    gx = torch.zeros_like(x)
    nfl = sum(w.numel() for w in ws)
    gw = torch.zeros(nfl)
    g_out, coef = g_out.float().contiguous(), coef.float().contiguous()
    n = lib.emu_diffpool_bwd(_ptr(g_out), _ptr(coef), _ptr(x), _ptr(adj), warr, layers, darr, b, _ptr(gx), _ptr(gw),
                             _ptr(state) if state is not None else None)
    assert n == nfl
    grads, off = {}, 0
    for k, w in zip(names, ws):
        grads[k] = gw[off:off + w.numel()].view_as(w)
        off += w.numel()
    return gx, grads


def dims_of(c):
    import math
    n, k1 = c["n"], math.ceil(0.25 * c["n"])
    k2 = math.ceil(0.25 * k1)
    return [(n, c["c"], k1, c["hid"]), (k1, c["hid"], k2, c["outd"])]


def losses_from_stats(stats, dims, b):
    """l = sum_l sqrt(sum_b F) / numel(adj_l), e = sum_l sum_b E / (b n_l)   (dense_diff_pool: adj_0 is shared [n, n],
    adj_1 is the batched pooled adjacency [b, k, k])."""
    l = e = 0.0
    for i, d in enumerate(dims):
        numel = d[0] * d[0] * (1 if i == 0 else b)
        l = l + stats[:, 2 * i].sum().sqrt() / numel
        e = e + stats[:, 2 * i + 1].sum() / (b * d[0])
    return l, e


@pytest.mark.parametrize("name", ["small", "ref_shape"])
def test_fused_diffpool_algebra_matches_reference_golden(emu, name):
    c = load_golden("diffpool")[name]
    dims = dims_of(c)
    b = c["x"].shape[0]
    assert emu.emu_diffpool_smem_floats(2, (ctypes.c_int64 * 8)(*[v for d in dims for v in d])) * 4 <= 227 * 1024
    out, stats, state = run_emu(emu, c["state_dict"], c["x"], c["adj"], dims)
    l, e = losses_from_stats(stats, dims, b)
    assert_close(out, c["out"], what=name + ".out")
    assert_close(l, c["link"], what=name + ".link")
    assert_close(e, c["ent"], what=name + ".ent")
    # the golden's loss: (out * R).sum() + 3 l + 0.5 e
    coef = []
    for i, d in enumerate(dims):
        numel = d[0] * d[0] * (1 if i == 0 else b)
        coef += [3.0 / (float(stats[:, 2 * i].sum().sqrt()) * numel), 0.5 / (b * d[0])]
    for st in (state, None):   # stored forward state / recomputed forward
        gx, grads = run_emu(emu, c["state_dict"], c["x"], c["adj"], dims, g_out=c["R"], coef=torch.tensor(coef), state=st)
        assert_close(gx, c["g_x"], rtol=2e-4, what=name + ".g_x")
        for k, g in grads.items():
            assert_close(g, c["g_params"][k], rtol=2e-4, what=name + ".g_" + k)


def test_fused_diffpool_algebra_single_layer_vs_oracle(emu):
    """One pooling layer (num_layers = 1 is a legal DiffPool configuration) and odd sizes, against autograd of the oracle."""
    g = torch.Generator().manual_seed(3)
    b, n, c, k, h = 3, 13, 5, 4, 7
    x = torch.randn(b, n, c, generator=g)
    a = torch.rand(n, n, generator=g)
    adj = a * (a > 0.4) + torch.eye(n)          # not symmetric, some row sums below the clamp are impossible here but deg varies
    sd = {}
    for kname, shape in zip(weight_names(1), [(k, c), (k, c), (k,), (h, c), (h, c), (h,), (h, h), (h, h), (h,)]):
        sd[kname] = (torch.randn(*shape, generator=g) * 0.5).requires_grad_()
    xr = x.clone().requires_grad_()
    out_r, l_r, e_r = R.diffpool_forward(sd, xr, adj, num_layers=1)
    dims = [(n, c, k, h)]
    out, stats, state = run_emu(emu, sd, x, adj, dims)
    l, e = losses_from_stats(stats, dims, b)
    assert_close(out, out_r, what="out")
    assert_close(l, l_r, what="link")
    assert_close(e, e_r, what="ent")
    Rw = torch.randn(out_r.shape, generator=g)
    names = weight_names(1)
    gr = torch.autograd.grad((out_r * Rw).sum() + 2.0 * l_r + 0.7 * e_r, [xr] + [sd[kk] for kk in names])
    coef = torch.tensor([2.0 / (float(stats[:, 0].sum().sqrt()) * n * n), 0.7 / (b * n)])
    for st in (state, None):
        gx, grads = run_emu(emu, sd, x, adj, dims, g_out=Rw, coef=coef, state=st)
        assert_close(gx, gr[0], rtol=2e-4, what="g_x")
        for kk, gg in zip(names, gr[1:]):
            assert_close(grads[kk], gg, rtol=2e-4, what="g_" + kk)
