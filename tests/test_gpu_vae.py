"""VAE.predict_head -- the reference's only DiffPool call site (models/vae.py:233-265) -- and the per-pathway decoders
(vae.py:54-74,216-222) on the GPU against golden vectors produced by the reference's own VAE class (oracle/make_golden.py):
prediction, link / entropy terms, gradients w.r.t. the input and every parameter on the path."""
import pytest
import torch

from conftest import assert_close, load_golden

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
VAEG = load_golden("vae")


@pytest.fixture(scope="module")
def mlg():
    import multilevel_gnn_b200 as m
    m._cabi.lib()
    return m


def _build(mlg, c):
    args = mlg.configs.make_args("lgg", **c["overrides"])
    torch.manual_seed(0)
    model = mlg.VAE(args, pathway_indexs=c["pathway_indexs"])
    model.reconstruct_head(args)
    model.set_pathway_similarity_matrix(c["sim"])
    missing, unexpected = model.load_state_dict(c["state_dict"], strict=False)
    assert not unexpected, unexpected
    assert all(not k.startswith(("diff_pooling.", "conv_model.", "head.")) for k in missing), missing
    return model.to(DEV).eval(), args


@pytest.mark.parametrize("name", sorted(VAEG))
def test_predict_head_golden(mlg, name):
    c = VAEG[name]
    model, args = _build(mlg, c)
    x = c["x"].to(DEV).requires_grad_()
    pred, feat, l, e = model.predict_head(x, c["age"].to(DEV))
    assert feat is x
    assert_close(pred, c["pred"], what=name + ".pred")
    assert_close(torch.as_tensor(l), c["link"], what=name + ".link")
    assert_close(torch.as_tensor(e), c["ent"], what=name + ".ent")
    names = [k for k, g in c["g_params"].items() if g is not None]
    params = dict(model.named_parameters())
    gs = torch.autograd.grad((pred * c["R"].to(DEV)).sum() + 3.0 * l + 0.5 * e, [x] + [params[k] for k in names])
    assert_close(gs[0], c["g_x"], rtol=2e-4, what=name + ".g_x")
    for k, g in zip(names, gs[1:]):
        assert_close(g, c["g_params"][k], rtol=2e-4, what=name + ".g_" + k)


def test_foreach_decoder_golden(mlg):
    c = VAEG["diffpool_pathway"]
    model, args = _build(mlg, c)
    h = c["h"].to(DEV).requires_grad_()
    dec = model.foreach_decoder(h)
    assert_close(dec, c["dec"], what="dec")
    dn = list(c["g_dec"])
    params = dict(model.named_parameters())
    gd = torch.autograd.grad((dec * c["Rd"].to(DEV)).sum(), [h] + [params[k] for k in dn])
    assert_close(gd[0], c["g_h"], rtol=2e-4, what="g_h")
    for k, g in zip(dn, gd[1:]):
        assert_close(g, c["g_dec"][k], rtol=2e-4, what="g_" + k)


def test_vae_encoder_and_forward_shapes(mlg):
    """encoder / forward / train_step of the drop-in run end to end on the kernels (shape + finiteness; the GNN stack and
    the pool underneath are parity-tested through MultilevelGNN)."""
    from multilevel_gnn_b200 import synth
    args = mlg.configs.make_args("lgg", decoder_type="foreach_diffhidden", reorder_type="diff_pooling",
                                 diff_pooling_location="pathway", reorder_pathway=False)
    torch.manual_seed(1)
    _, seg = synth.pool_layout(seed=0)
    model = mlg.VAE(args, pathway_indexs=seg)
    synth.multilevel_params(model)
    model.reconstruct_head(args)
    model.set_pathway_similarity_matrix(torch.rand(146, 146))
    model.to(DEV).train()
    b = synth.multilevel_batch(batch_size=2, seed=4).to(DEV)
    out = model(b)
    assert out["pred_x"].shape == (2, 25015) and out["embedding"].shape == (2, 438, 2 * 96)
    assert bool(torch.isfinite(out["pred_x"]).all())
    pred, feat, l, e, _ = model.train_step(b)
    assert pred.shape == (2, 2) and feat.shape == (2, 96, 146, 3)      # mu half of the latent: C*P channels (vae.py:97)
    (pred.sum() + l + e + out["pred_x"].square().mean()).backward()
    assert model.node_embedding.grad is not None and bool(torch.isfinite(model.node_embedding.grad).all())
