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
    # one packed parameter holds every block (models/decoder.py); its gradient is compared block by block with the
    # reference's per-Linear gradients
    assert [k for k, _ in model.named_parameters() if k.startswith("decoder.")] == ["decoder.packed"]
    g_h, g_packed = torch.autograd.grad((dec * c["Rd"].to(DEV)).sum(), [h, model.decoder.packed])
    assert_close(g_h, c["g_h"], rtol=2e-4, what="g_h")
    seen = 0
    for i in range(len(model.decoder)):
        for key, view in zip(mlg.models.decoder.KEYS, model.decoder.block(i, g_packed)):
            k = "decoder.%d.%s" % (i, key)
            assert_close(view, c["g_dec"][k], rtol=2e-4, what="g_" + k)
            seen += 1
    assert seen == len(c["g_dec"])
    # and the state_dict still carries the reference's keys
    sd = model.state_dict()
    for k in c["g_dec"]:
        assert torch.equal(sd[k].cpu(), c["state_dict"][k]), k


def test_grouped_decoder_row_chunks_and_full_size(mlg):
    """The lgg-sized decoder stack (438 pathways, F = 96, D = 128, 15 405 output genes) with a batch larger than one launch
    holds in shared memory (row chunks, accumulated parameter gradients) against the per-block library formulation in fp64."""
    torch.manual_seed(3)
    g = torch.Generator().manual_seed(4)
    S, F, D = 438, 96, 128
    outs = torch.randint(1, 70, (S,), generator=g).tolist()
    dec = mlg.models.decoder.GroupedDecoder(F, [D] * S, outs).to(DEV)
    B = int(mlg._cabi.lib().mlg_decoder_max_rows(F, D, 1)) + 37
    x = torch.randn(B, S, F, generator=g).to(DEV).requires_grad_()
    Rw = torch.randn(B, sum(outs), generator=g).to(DEV)
    y = dec(x)
    gx, gp = torch.autograd.grad((y * Rw).sum(), [x, dec.packed])
    xr = x.detach().double().requires_grad_()
    blocks = [[t.double().requires_grad_() for t in dec.block(i)] for i in range(S)]
    yr = torch.cat([torch.relu(xr[:, i] @ w1.t() + b1) @ w2.t() + b2 for i, (w1, b1, w2, b2) in enumerate(blocks)], dim=1)
    flat = [t for blk in blocks for t in blk]
    gr = torch.autograd.grad((yr * Rw.double()).sum(), [xr] + flat)
    assert_close(y, yr.float(), what="pred")
    assert_close(gx, gr[0].float(), rtol=2e-4, what="g_x")
    q = 1
    for i in range(S):
        for view in dec.block(i, gp):
            assert_close(view, gr[q].float(), rtol=2e-4, what="g_block%d[%d]" % (i, q))
            q += 1


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
