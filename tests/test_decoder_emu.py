"""Algebra of the grouped decoder kernel (csrc/decoder_grouped.cu) checked WITHOUT a GPU: the same source compiled as host
C++ (tests/host_emu/decoder_emu.cpp, -DMLG_HOST_EMU) against autograd of the CPU oracle (oracle.restated.foreach_decoder,
models/vae.py:216-222), plus the state_dict mapping of models/decoder.py.  Test infrastructure only: the product path is the
CUDA build of that source (tests/test_gpu_vae.py)."""
import ctypes
import os
import subprocess

import pytest
import torch

import multilevel_gnn_b200 as mlg
from conftest import assert_close
from oracle import restated as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emu") / "libdecoder_emu.so")
    src = os.path.join(ROOT, "tests", "host_emu", "decoder_emu.cpp")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", "-o", so, src], check=True)
    return ctypes.CDLL(so)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


CASES = {
    # (B, F, hidden per pathway, outputs per pathway): aligned sizes, ragged sizes (scalar operand paths, K tails), one row
    "aligned": (8, 96, [64] * 5, [32, 8, 64, 4, 16]),
    "ragged": (7, 10, [16, 5, 32, 8], [3, 17, 1, 9]),
    "one_row": (1, 6, [4, 4], [5, 2]),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_grouped_decoder_algebra_matches_oracle(emu, name):
    B, F, hidden, outs = CASES[name]
    torch.manual_seed(1)
    dec = mlg.models.decoder.GroupedDecoder(F, hidden, outs)
    S = len(outs)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, S, F, generator=g)
    Rw = torch.randn(B, sum(outs), generator=g)
    # oracle: the reference's per-block keys come out of the state_dict hook
    sd = {"decoder." + k: v.clone().requires_grad_() for k, v in dec.state_dict().items()}
    assert set(sd) == {"decoder.%d.%s" % (i, k) for i in range(S) for k in mlg.models.decoder.KEYS}
    xr = x.clone().requires_grad_()
    y_r = R.foreach_decoder(sd, xr)
    names = sorted(sd)
    g_r = torch.autograd.grad((y_r * Rw).sum(), [xr] + [sd[k] for k in names])
    # host build of the kernel source
    packed, table = dec.packed.detach().contiguous(), dec.table.contiguous()
    out = torch.zeros(B, dec.total_out)
    h = torch.zeros(B, dec.total_hidden)
    emu.emu_decoder_fwd(_ptr(x), _ptr(packed), _ptr(table), B, S, F, dec.d_max, ctypes.c_longlong(dec.total_out),
                        ctypes.c_longlong(dec.total_hidden), _ptr(out), _ptr(h))
    assert_close(out, y_r, what=name + ".pred")
    gx = torch.zeros_like(x)
    gp = torch.zeros_like(packed)
    # two row chunks when possible: the second accumulates the parameter gradients
    cuts = [(0, B)] if B < 2 else [(0, B // 2), (B // 2, B)]
    for ci, (b0, b1) in enumerate(cuts):
        emu.emu_decoder_bwd(_ptr(Rw[b0:b1].contiguous()), _ptr(x[b0:b1].contiguous()), _ptr(h[b0:b1].contiguous()), _ptr(packed),
                            _ptr(table), b1 - b0, S, F, dec.d_max, ctypes.c_longlong(dec.total_out),
                            ctypes.c_longlong(dec.total_hidden), _ptr(gx[b0:b1]), _ptr(gp), 1 if ci else 0)
    assert_close(gx, g_r[0], rtol=2e-4, what=name + ".g_x")
    for i in range(S):
        for key, view in zip(mlg.models.decoder.KEYS, dec.block(i, gp)):
            k = "decoder.%d.%s" % (i, key)
            assert_close(view, g_r[1 + names.index(k)], rtol=2e-4, what=name + ".g_" + k)


def test_grouped_decoder_state_dict_round_trip():
    """The packed parameter loads from / saves to the reference's per-block keys (vae.py:54-74 ModuleList of Sequential)."""
    ref = torch.nn.ModuleList([torch.nn.Sequential(torch.nn.Linear(6, d), torch.nn.ReLU(), torch.nn.Linear(d, n))
                               for d, n in [(4, 3), (8, 5), (4, 1)]])
    dec = mlg.models.decoder.GroupedDecoder(6, [4, 8, 4], [3, 5, 1])
    assert [k for k, _ in dec.named_parameters()] == ["packed"]
    dec.load_state_dict(ref.state_dict())
    sd = dec.state_dict()
    assert list(sd) == list(ref.state_dict())
    for k, v in ref.state_dict().items():
        assert torch.equal(sd[k], v), k
    with pytest.raises(RuntimeError):
        bad = dict(ref.state_dict())
        bad.pop("1.2.bias")
        dec.load_state_dict(bad)
    with pytest.raises(Exception):
        dec(torch.zeros(2, 3, 6))      # CPU tensors: no CPU path
