"""Container-only: freeze golden vectors from the REFERENCE's own module files.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Runs /root/reference's unmodified
``models/gcn_lib/*``, ``models/multilevel_gnn.py``, ``models/deepergcn.py`` and
``models/diff_pooling.py`` on CPU behind ``oracle/pyg_stub.py`` on seeded inputs and saves
inputs + state_dict + outputs + every gradient as small ``tests/golden/*.pt`` files.

    python -m oracle.make_golden            # regenerate everything

The reference hard-codes 5135 genes / 25015 gene slots inside ``MultilevelGNN.__init__``; to keep
the fixture small the INSTANCE is re-sized after construction (node_num, node_embedding,
learnable_pca_params), its ``forward`` code is untouched.  A full-size (gbm.yaml shape) check of
the restatement against the live reference lives in tests/test_oracle_pin.py (container-only).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_import  # noqa: E402
from oracle.pyg_stub import Data  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def random_graph(n, e, g, with_isolated=True):
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, n - (2 if with_isolated else 0), (e,), generator=g)   # last 2 nodes: no in-edges
    return torch.stack([src, dst])


def grads_of(loss, tensors):
    gs = torch.autograd.grad(loss, tensors, allow_unused=True)
    return [None if g is None else g.detach().clone() for g in gs]


GEN_CASES = [
    # name, kwargs, H
    ("softmax_learn_t_msgnorm", dict(aggr="softmax", t=0.7, learn_t=True, msg_norm=True, learn_msg_scale=True, norm="layer"), 32),
    ("softmax_fixed_t", dict(aggr="softmax", t=1.3, learn_t=False, msg_norm=False, norm="layer"), 64),
    ("softmax_sg", dict(aggr="softmax_sg", t=2.0, learn_t=True, msg_norm=True, learn_msg_scale=False, norm="batch"), 16),
    ("softmax_sum", dict(aggr="softmax_sum", t=1.0, learn_t=True, y=0.3, learn_y=True, msg_norm=True, norm="layer", mlp_layers=1), 128),
    ("power_p2", dict(aggr="power", p=2.0, learn_p=True, msg_norm=False, norm="layer"), 32),
    ("power_p1", dict(aggr="power", p=1.0, learn_p=False, msg_norm=True, norm="layer"), 18),
    ("power_sum", dict(aggr="power_sum", p=3.0, learn_p=True, y=-0.2, learn_y=True, msg_norm=False, norm="layer"), 20),
    ("add", dict(aggr="add", msg_norm=True, norm="layer"), 32),
    ("mean", dict(aggr="mean", msg_norm=False, norm="layer", mlp_layers=1), 136),
    ("softmax_wide", dict(aggr="softmax", t=1.0, learn_t=True, msg_norm=True, norm="layer", mlp_layers=1), 160),
    ("max", dict(aggr="max", msg_norm=True, norm="layer"), 40),
]


def make_genconv(ns):
    out = {}
    for i, (name, kw, H) in enumerate(GEN_CASES):
        g = gen(100 + i)
        torch.manual_seed(100 + i)
        n, e = (37, 260) if H < 128 else (14, 70)
        conv = ns.torch_vertex.GENConv(H, H, encode_edge=True, edge_feat_dim=H, **kw)
        conv.train()
        x = (torch.randn(n, H, generator=g) * 1.5).requires_grad_()
        ea = torch.randn(e, H, generator=g).requires_grad_()
        ei = random_graph(n, e, g)
        R = torch.randn(n, H, generator=g)
        y = conv(x, ei, ea)
        params = [p for p in conv.parameters() if p.requires_grad]
        names = [k for k, p in conv.named_parameters() if p.requires_grad]
        gs = grads_of((y * R).sum(), [x, ea] + params)
        # aggregate-only view (GenMessagePassing.aggregate on explicit messages)
        msg = torch.rand(e, H, generator=g) * 3 + 0.01
        msg_in = msg.clone().requires_grad_()
        agg = conv.aggregate(msg_in * 1.0, ei[1], dim_size=n)
        g_msg = grads_of((agg * R).sum(), [msg_in])[0]
        out[name] = dict(kw=kw, H=H, x=x.detach(), edge_attr=ea.detach(), edge_index=ei, R=R,
                         state_dict={k: v.detach().clone() for k, v in conv.state_dict().items()},
                         y=y.detach(), g_x=gs[0], g_edge_attr=gs[1],
                         g_params={k: v for k, v in zip(names, gs[2:])},
                         msg=msg, agg=agg.detach(), g_msg=g_msg)
    torch.save(out, os.path.join(OUT, "genconv.pt"))


PATHWAY_CASES = [
    ("softmax_learn_t", dict(aggr="softmax", t=0.8, learn_t=True, norm="layer"), 32, True),
    ("softmax_sum_mask", dict(aggr="softmax_sum", t=1.0, learn_t=True, y=0.2, learn_y=True, norm="batch"), 24, True),
    ("mean", dict(aggr="mean", norm="layer", mlp_layers=1), 20, False),
    ("max", dict(aggr="max", norm="layer"), 16, False),
]


def make_pathwayconv(ns):
    """PathwayConv (torch_vertex.py:107-178) from the reference's own class: outer-product message through msg_encoder,
    PathwayMessagePassing aggregation, residual, MLP, ReLU, optional mask."""
    out = {}
    for i, (name, kw, H, use_mask) in enumerate(PATHWAY_CASES):
        g = gen(700 + i)
        torch.manual_seed(700 + i)
        n, e = 41, 230
        conv = ns.torch_vertex.PathwayConv(H, H, **kw)
        conv.train()
        x = torch.randn(n, H, generator=g).requires_grad_()
        ea = torch.randn(e, 2, generator=g).requires_grad_()
        ei = random_graph(n, e, g)
        mask = (torch.rand(n, 1, generator=g) > 0.3).float() if use_mask else None
        R = torch.randn(n, H, generator=g)
        y = conv(x, ei, ea, mask)
        names = [k for k, p in conv.named_parameters() if p.requires_grad]
        gs = grads_of((y * R).sum(), [x, ea] + [p for p in conv.parameters() if p.requires_grad])
        out[name] = dict(kw=kw, H=H, x=x.detach(), edge_attr=ea.detach(), edge_index=ei, mask=mask, R=R,
                         state_dict={k: v.detach().clone() for k, v in conv.state_dict().items()},
                         y=y.detach(), g_x=gs[0], g_edge_attr=gs[1], g_params={k: v for k, v in zip(names, gs[2:])})
    torch.save(out, os.path.join(OUT, "pathwayconv.pt"))


def make_sage(ns):
    out = {}
    for i, (conv_name, cin, cout) in enumerate([("sage", 64, 64), ("sage", 64, 32), ("rsage", 32, 32), ("sage", 10, 6)]):
        g = gen(200 + i)
        torch.manual_seed(200 + i)
        n, e = 45, 300
        conv = ns.torch_vertex.GraphConv(cin, cout, conv=conv_name, act="leakyrelu", norm=None, mlp_norm="none")
        x = torch.randn(n, cin, generator=g).requires_grad_()
        ei = random_graph(n, e, g)
        ei[:, :7] = ei[0, :7]                     # a few pre-existing self loops (dropped + re-added with w=1)
        ea = torch.rand(e, 1, generator=g) * 2 - 0.5
        R = torch.randn(n, cout, generator=g)
        y = conv(x, ei, ea)
        params = [p for p in conv.parameters()]
        names = [k for k, _ in conv.named_parameters()]
        gs = grads_of((y * R).sum(), [x] + params)
        out["%s_%d_%d" % (conv_name, cin, cout)] = dict(
            conv=conv_name, cin=cin, cout=cout, x=x.detach(), edge_index=ei, edge_attr=ea, R=R,
            state_dict={k: v.detach().clone() for k, v in conv.state_dict().items()},
            y=y.detach(), g_x=gs[0], g_params={k: v for k, v in zip(names, gs[1:])})
    torch.save(out, os.path.join(OUT, "sage.pt"))


def make_knn(ns):
    out = {}
    g = gen(300)
    for name, (b, n, d, k, dil) in {"b2_n70_d8_k5": (2, 70, 8, 5, 1), "b1_n130_d33_k9_d2": (1, 130, 33, 9, 2),
                                    "b3_n64_d3_k16": (3, 64, 3, 16, 1)}.items():
        x = torch.randn(b * n, d, generator=g)
        batch = torch.arange(b).repeat_interleave(n)
        ei = ns.torch_edge.knn_graph_matrix(x, k * dil, batch)
        dg = ns.torch_edge.DilatedKnnGraph(k, dil)
        dg.eval()
        ei_d = dg(x, batch)
        xd = x.view(b, n, d).transpose(1, 2).unsqueeze(-1).contiguous()
        dense = ns.dense_edge.dense_knn_matrix(xd, k * dil)
        dense_d = ns.dense_edge.DenseDilatedKnnGraph(k, dil)(xd)
        out[name] = dict(b=b, n=n, d=d, k=k, dil=dil, x=x, batch=batch, edge_index_full=ei, edge_index_dilated=ei_d,
                         dense_full=dense, dense_dilated=dense_d)
    # exact ties: integer grid points
    xg = torch.stack(torch.meshgrid(torch.arange(6.), torch.arange(6.), indexing="ij"), -1).reshape(-1, 2)
    out["grid_ties"] = dict(b=1, n=36, d=2, k=5, dil=1, x=xg, batch=torch.zeros(36, dtype=torch.long),
                            edge_index_full=ns.torch_edge.knn_graph_matrix(xg, 5, None))
    torch.save(out, os.path.join(OUT, "knn.pt"))


def small_multilevel_args(config, **kw):
    over = dict(conv_channel_list=[8, 4], head_dim=16, hidden_channels=16, final_channels=8,
                node_embedding_dim=12 if config == "gbm" else 8)
    over.update(kw)
    return over


def resize_multilevel(model, genes, slots, g):
    """shrink the reference INSTANCE (its forward code is untouched)"""
    model.node_num = genes
    emb_dim = model.node_embedding.shape[1]
    model.node_embedding = torch.nn.Parameter(torch.randn(3 * genes, emb_dim, generator=g) * 0.3)
    model.learnable_pca_params = torch.nn.Parameter(torch.randn(slots, model.pca_dim, generator=g) * 0.2)
    model.set_info_mask((torch.rand(slots, 1, generator=g) < 0.6).float())


def small_multilevel_batch(bsz, genes, slots, g):
    n = 3 * genes
    e_intra = genes * 6
    src = torch.randint(0, genes, (e_intra,), generator=g)
    dst = torch.randint(0, genes, (e_intra,), generator=g)
    gi = torch.arange(genes)
    ei = torch.cat([torch.stack([3 * src, 3 * dst]), torch.stack([3 * gi + 1, 3 * gi]), torch.stack([3 * gi + 2, 3 * gi])], 1)
    ea = torch.cat([torch.rand(e_intra, generator=g), torch.ones(genes), -torch.ones(genes)]).unsqueeze(1)
    perm = torch.randperm(ei.shape[1], generator=g)
    ei, ea = ei[:, perm], ea[perm]
    off = (torch.arange(bsz) * n).view(-1, 1, 1)
    match = torch.randint(0, n, (slots,), generator=g)
    match[torch.rand(slots, generator=g) < 0.05] = -1
    seg = torch.sort(torch.randint(0, 438, (slots,), generator=g)).values
    lab = (torch.rand(bsz, generator=g) < 0.5).long()
    return dict(x=torch.randn(bsz * n, 1, generator=g),
                edge_index=(ei.unsqueeze(0) + off).permute(1, 0, 2).reshape(2, -1).contiguous(),
                edge_attr=ea.repeat(bsz, 1), gene_pca_match=match.unsqueeze(0).repeat(bsz, 1),
                raw_indice=seg.unsqueeze(0).repeat(bsz, 1), age=torch.rand(bsz, generator=g),
                y=torch.nn.functional.one_hot(lab, 2).float().reshape(-1))


def make_multilevel(ns):
    out = {}
    cases = {"gbm": {}, "kirc": {}, "lgg_rsage_res": dict(gnn_name="rsage", resgnn=True, final_channels=16, node_embedding_dim=16),
             "gbm_repeatmask": dict(repeat_mask=True, repeat_cyclic=1, repeat_norm=True, num_layers=3)}
    for i, (name, extra) in enumerate(cases.items()):
        config = name.split("_")[0]
        g = gen(400 + i)
        torch.manual_seed(400 + i)
        over = small_multilevel_args(config, **extra)
        args = ref_import.default_args(config + ".yaml", **over)
        model = ns.multilevel_gnn.MultilevelGNN(args)
        genes, slots, bsz = 30, 500, 3
        resize_multilevel(model, genes, slots, g)
        fields = small_multilevel_batch(bsz, genes, slots, g)
        model.set_pathway_indexs(fields["raw_indice"][0].clone())
        model.eval()                       # dropout = identity
        acts = {}
        hooks = [layer.register_forward_hook(lambda m, a, o, j=j: acts.__setitem__("gnn%d" % j, o.detach().clone()))
                 for j, layer in enumerate(model.gnn_model)]
        pred, feat = model(Data(**fields))
        for h in hooks:
            h.remove()
        floss = model.get_feature_loss(feat)
        weight = torch.tensor([[0.7, 1.4]]).repeat(bsz, 1)
        loss = torch.nn.BCELoss(weight=weight)(pred.float(), fields["y"].reshape(-1, 2)) + floss
        params = {k: p for k, p in model.named_parameters() if p.requires_grad}
        gs = grads_of(loss, list(params.values()))
        out[name] = dict(config=config, overrides=over, genes=genes, slots=slots, batch=fields, weight=weight,
                         state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                         pred=pred.detach(), pca_feature=feat.detach(), acts=acts, feature_loss=torch.as_tensor(floss).detach(),
                         loss=loss.detach(), grads={k: g_ for k, g_ in zip(params.keys(), gs)})
    torch.save(out, os.path.join(OUT, "multilevel.pt"))


def make_diffpool(ns):
    out = {}
    for i, (name, (b, n, c, hid, outd)) in enumerate({"small": (5, 22, 6, 8, 10), "ref_shape": (4, 146, 32, 32, 64)}.items()):
        g = gen(500 + i)
        torch.manual_seed(500 + i)
        args = ref_import.default_args("lgg.yaml")
        dp = ns.diff_pooling.DiffPool(c, 2, n, 2, hid, outd, args)
        dp.train()
        x = torch.randn(b, n, c, generator=g).requires_grad_()
        a = torch.rand(n, n, generator=g)
        adj = (a + a.t()) * 0.5 + torch.eye(n)
        xo, l, e = dp(x, adj)
        R = torch.randn(xo.shape, generator=g)
        params = {k: p for k, p in dp.named_parameters()}
        gs = grads_of((xo * R).sum() + 3.0 * l + 0.5 * e, [x] + list(params.values()))
        out[name] = dict(b=b, n=n, c=c, hid=hid, outd=outd, x=x.detach(), adj=adj, R=R,
                         state_dict={k: v.detach().clone() for k, v in dp.state_dict().items()},
                         out=xo.detach(), link=l.detach(), ent=e.detach(), g_x=gs[0],
                         g_params={k: g_ for k, g_ in zip(params.keys(), gs[1:])})
    torch.save(out, os.path.join(OUT, "diffpool.pt"))


def make_deepergcn(ns):
    out = {}
    for i, (name, extra) in enumerate({"resplus_softmax": dict(block="res+", gcn_aggr="softmax", learn_t=True, msg_norm=True, learn_msg_scale=True),
                                       "res_power": dict(block="res", gcn_aggr="power", p=2.0, learn_p=True),
                                       "plain_max": dict(block="plain", gcn_aggr="max", msg_norm=True)}.items()):
        g = gen(600 + i)
        torch.manual_seed(600 + i)
        P = 8
        base = dict(hidden_channels=16, num_layers=3, conv="gen", norm="layer", mlp_layers=2, conv_encode_edge=True,
                    use_edge_attr=True, global_edge=None, use_column="stringdb::score", pathway_global_node=True,
                    pathway_num=P, pathway_readout="maxpool", pre_readout_drop=True, num_layer_head=2, use_age=True,
                    pre_concat_age=True, dropout=0.0, node_embedding=False, feature_drop=False)
        base.update(extra)
        args = ref_import.default_args(None, **base)
        model = ns.deepergcn.DeeperGCN(args)
        model.train()
        sizes = torch.tensor([40 + P, 33 + P])
        n = int(sizes.sum())
        e = 400
        ei = torch.cat([random_graph(40 + P, 220, g, False), random_graph(33 + P, 180, g, False) + (40 + P)], 1)
        fields = dict(x=torch.randn(n, 3, generator=g), edge_index=ei, edge_attr=torch.rand(e, 1, generator=g),
                      batch=torch.arange(2).repeat_interleave(sizes), age=torch.rand(2, generator=g),
                      pathway_node_attr=torch.randn(2 * P, 6, generator=g), node_size=sizes)
        pred = model(Data(**fields))
        R = torch.randn(pred.shape, generator=g)
        params = {k: p for k, p in model.named_parameters() if p.requires_grad}
        gs = grads_of((pred * R).sum(), list(params.values()))
        out[name] = dict(overrides=base, batch=fields, R=R,
                         state_dict={k: v.detach().clone() for k, v in model.state_dict().items()},
                         pred=pred.detach(), grads={k: g_ for k, g_ in zip(params.keys(), gs)})
    torch.save(out, os.path.join(OUT, "deepergcn.pt"))


def make_vae(ns):
    """VAE.predict_head (the reference's only DiffPool call site, models/vae.py:233-265) for both diff_pooling locations and
    the plain pooled head, plus the per-pathway decoders (vae.py:216-222), from the reference's own VAE class."""
    out = {}
    cases = {"diffpool_pathway": dict(reorder_type="diff_pooling", diff_pooling_location="pathway"),
             "diffpool_head": dict(reorder_type="diff_pooling", diff_pooling_location="head"),
             "plain_pool": dict(reorder_type="pca")}
    for i, (name, extra) in enumerate(cases.items()):
        g = gen(700 + i)
        torch.manual_seed(700 + i)
        over = dict(decoder_type="foreach_diffhidden", decoder_dim=64, head_dim=32 if i == 0 else 8, **extra)
        args = ref_import.default_args("lgg.yaml", **over)
        pidx = torch.repeat_interleave(torch.arange(16), torch.randint(1, 40, (16,), generator=g))   # 16 decoder blocks
        model = ns.vae.VAE(args, pathway_indexs=pidx)
        model.reconstruct_head(args)
        a = torch.rand(146, 146, generator=g)
        sim = (a + a.t()) * 0.5
        model.set_pathway_similarity_matrix(sim.numpy())
        model.eval()
        bsz = 2
        x = torch.randn(bsz, 32, 146, 9, generator=g).requires_grad_()
        age = torch.rand(bsz, generator=g)
        pred, feat, l, e = model.predict_head(x, age)
        R = torch.randn(pred.shape, generator=g)
        keep = ("diff_pooling.", "conv_model.", "head.", "decoder.")
        params = {k: p for k, p in model.named_parameters() if k.startswith(keep[:3])}
        gs = grads_of((pred * R).sum() + 3.0 * l + 0.5 * e, [x] + list(params.values()))
        h = torch.randn(bsz, 16, 96, generator=g).requires_grad_()
        dec = model.foreach_decoder(h)
        Rd = torch.randn(dec.shape, generator=g)
        dparams = {k: p for k, p in model.named_parameters() if k.startswith("decoder.")}
        gd = grads_of((dec * Rd).sum(), [h] + list(dparams.values()))
        out[name] = dict(overrides=over, pathway_indexs=pidx, sim=sim, x=x.detach(), age=age, R=R, h=h.detach(), Rd=Rd,
                         state_dict={k: v.detach().clone() for k, v in model.state_dict().items() if k.startswith(keep)},
                         pred=pred.detach(), link=torch.as_tensor(l).detach(), ent=torch.as_tensor(e).detach(), g_x=gs[0],
                         g_params={k: g_ for k, g_ in zip(params.keys(), gs[1:])},
                         dec=dec.detach(), g_h=gd[0], g_dec={k: g_ for k, g_ in zip(dparams.keys(), gd[1:])})
        if i > 0:      # the decoder blocks are exercised once
            for k in ("h", "Rd", "dec", "g_h", "g_dec"):
                out[name].pop(k)
            out[name]["state_dict"] = {k: v for k, v in out[name]["state_dict"].items() if not k.startswith("decoder.")}
    torch.save(out, os.path.join(OUT, "vae.pt"))


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_import.load()
    only = set(sys.argv[1:])
    for fn in (make_genconv, make_pathwayconv, make_sage, make_knn, make_multilevel, make_diffpool, make_deepergcn, make_vae):
        if only and fn.__name__ not in only:
            continue
        fn(ns)
        print("wrote", fn.__name__)
    for f in sorted(os.listdir(OUT)):
        print("%-20s %8.1f KB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024))


if __name__ == "__main__":
    main()
