#!/usr/bin/env python
"""Diagnostic for tests/test_gpu_wide_batch.py::test_reformulated_layers_with_more_than_32_graphs: factored + transform-first
path vs the buffered path at B = 40 in eval mode, train mode without dropout and train mode with dropout; also checks that
both runs draw the same dropout words."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multilevel_gnn_b200 as m
from multilevel_gnn_b200 import configs, functional as Fn, synth

DEV = "cuda:0"
args = configs.make_args("gbm")
torch.manual_seed(5)
model = m.MultilevelGNN(args)
synth.multilevel_params(model)
model.to(DEV)
model.pathway_indexs = model.pathway_indexs.to(DEV)
b = synth.multilevel_batch(batch_size=40, seed=9).to(DEV)
names = [n for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
params = [p for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
drawn = []
orig = Fn._drop_bits
def rec(n, device):
    t = orig(n, device)
    drawn.append(t.clone())
    return t
Fn._drop_bits = rec

def run(factored, tfirst, self_mask=True, feat_term=True):
    Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST, Fn.RANK1_SELF_MASK = factored, tfirst, self_mask
    torch.manual_seed(11)
    drawn.clear()
    pred, feat = model(b)
    loss = (pred * torch.arange(pred.numel(), device=DEV).reshape(pred.shape)).sum()
    if feat_term:
        loss = loss + feat.square().mean()
    g = torch.autograd.grad(loss, params, allow_unused=True)
    torch.cuda.synchronize()
    Fn.FACTORED_RANK1, Fn.TRANSFORM_FIRST, Fn.RANK1_SELF_MASK = True, True, True
    return pred.detach(), feat.detach(), g, [d.clone() for d in drawn]

for mode in ("eval", "train_p0", "train"):
    model.train(mode != "eval")
    model.drop1.p = 0.0 if mode == "train_p0" else 0.25
    model.head[2].p = 0.0 if mode == "train_p0" else 0.5
    for label, kw in (("default", {}), ("leaky-pass", dict(self_mask=False)), ("no feat term", dict(feat_term=False))):
        p0, f0, g0, d0 = run(False, False, **{k: v for k, v in kw.items() if k == "feat_term"})
        p1, f1, g1, d1 = run(True, True, **kw)
        same_bits = len(d0) == len(d1) and all(torch.equal(x, y) for x, y in zip(d0, d1))
        worst = []
        for n, a, c in zip(names, g1, g0):
            if a is None:
                continue
            sc = float(c.abs().max().clamp_min(1e-30))
            worst.append((float((a - c).abs().max()) / sc, n))
        worst.sort(reverse=True)
        print(mode, label, "bits equal:", same_bits, len(d0), "pred diff %.2e" % float((p1 - p0).abs().max()),
              "worst grads:", [("%.2e" % e, n) for e, n in worst[:3]])
