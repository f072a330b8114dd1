#!/usr/bin/env python
"""Diagnostic: at B = 40 graphs, gradient of a loss that touches ONE replica only (feat[b]^2), default (factored +
transform-first) path and buffered path, both against autograd of the CPU oracle."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multilevel_gnn_b200 as m
from multilevel_gnn_b200 import configs, functional as Fn, synth
from oracle import restated as R

DEV = "cuda:0"
args = configs.make_args("gbm")
torch.manual_seed(5)
model = m.MultilevelGNN(args)
synth.multilevel_params(model)
model.eval()
batch = synth.multilevel_batch(batch_size=40, seed=9)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
names = [n for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")
         and not n.startswith(("head", "conv_model"))]
model.to(DEV)
model.pathway_indexs = model.pathway_indexs.to(DEV)
gb = batch.to(DEV)
params = dict(model.named_parameters())

def cuda_grads(bsel, factored, tfirst):
    Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST = factored, tfirst
    try:
        pred, feat = model(gb)
        g = torch.autograd.grad(feat[bsel].square().sum(), [params[k] for k in names], allow_unused=True)
    finally:
        Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST = True, True
    return [x.detach().cpu() for x in g]

for bsel in (0, 31, 32, 35, 39):
    leaf = {k: (v.clone().requires_grad_() if k in names else v) for k, v in sd.items()}
    pred_r, feat_r = R.multilevel_forward(leaf, batch, args)
    g_r = torch.autograd.grad(feat_r[bsel].square().sum(), [leaf[k] for k in names], allow_unused=True)
    for label, (fa, tf) in (("default", (True, True)), ("buffered", (False, False)), ("factored+buffered2", (True, False)),
                            ("buffered1+tfirst", (False, True))):
        g = cuda_grads(bsel, fa, tf)
        out = []
        for n, a, c in zip(names, g, g_r):
            sc = float(c.abs().max().clamp_min(1e-30))
            out.append("%s %.1e" % (n.split(".")[-2] + "." + n.split(".")[-1] if "." in n else n, float((a - c).abs().max()) / sc))
        print("replica", bsel, label, "| ".join(out))

# ---- train mode: where do the two paths start to disagree? ----
print("---- train mode (dropout on) ----")
model.train()
all_names = [n for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
all_params = [params[k] for k in all_names]

def train_run(factored, tfirst):
    Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST = factored, tfirst
    torch.manual_seed(11)
    grabbed = {}
    try:
        pred, feat = model(gb)
        feat.register_hook(lambda g: grabbed.__setitem__("g_feat", g.detach().clone()))
        loss = (pred * torch.arange(pred.numel(), device=DEV).reshape(pred.shape)).sum() + feat.square().mean()
        g = torch.autograd.grad(loss, all_params, allow_unused=True)
        torch.cuda.synchronize()
    finally:
        Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST = True, True
    return pred.detach(), feat.detach(), [x.detach() for x in g], grabbed["g_feat"]

def rel(a, c):
    return float((a - c).abs().max()) / float(c.abs().max().clamp_min(1e-30))

rA = train_run(True, True)
rA2 = train_run(True, True)
rB = train_run(False, False)
rB2 = train_run(False, False)
print("default twice: feat", rel(rA[1], rA2[1]), "g_feat", rel(rA[3], rA2[3]), "grads", max(rel(a, c) for a, c in zip(rA[2], rA2[2])))
print("buffered twice: feat", rel(rB[1], rB2[1]), "g_feat", rel(rB[3], rB2[3]), "grads", max(rel(a, c) for a, c in zip(rB[2], rB2[2])))
print("default vs buffered: pred", rel(rA[0], rB[0]), "feat", rel(rA[1], rB[1]), "g_feat", rel(rA[3], rB[3]))
for n, a, c in zip(all_names, rA[2], rB[2]):
    print("   ", n, "%.2e" % rel(a, c))
gf = (rA[3] - rB[3]).abs()
print("g_feat mismatch per replica:", [round(float(gf[b].max() / rB[3].abs().max()), 5) for b in range(gf.shape[0])])
