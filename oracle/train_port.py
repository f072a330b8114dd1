"""CPU port of one reference training step -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

Follows train.py:38-68 (forward, get_feature_loss, zero_grad, BCELoss(weight) + feature loss,
backward, Adam.step) with the restated forward of oracle/restated.py, i.e. the reference's own op
sequence (per-edge lin_r matmul, scatter-mean, materialised [B,C,G,P] pool + scatter).  Used by
tests (one-step parity of the Trainer) and by bench.py's cpu_baseline / --impl reference legs.
"""
import torch

from . import restated as R


class CpuTrainer:
    def __init__(self, state_dict, args, criterion_weight, pathway_indexs):
        self.args = args
        self.sd = {k: v.detach().clone().float() for k, v in state_dict.items()}
        self.train_keys = [k for k, v in self.sd.items()
                           if v.is_floating_point() and k != "info_mask" and not k.endswith("lin_l.weight")]
        for k in self.train_keys:
            self.sd[k].requires_grad_()
        self.opt = torch.optim.Adam([self.sd[k] for k in self.train_keys], lr=args.lr,
                                    betas=(args.beta1, args.beta2), weight_decay=args.wd)
        self.weight = criterion_weight
        self.pathway_indexs = pathway_indexs

    def loss(self, batch):
        pred, feat = R.multilevel_forward(self.sd, batch, self.args)
        fl = R.feature_loss(feat, self.sd["learnable_pca_params"], self.sd["info_mask"], self.pathway_indexs,
                            pca_loss=self.args.pca_loss, pca_indep_loss=self.args.pca_indep_loss)
        w = self.weight if self.args.weight_balance else None
        return R.bce_loss(pred, batch.y.reshape(-1, 2), w) + fl

    def step(self, batch):
        self.opt.zero_grad()
        loss = self.loss(batch)
        loss.backward()
        self.opt.step()
        return loss.detach()
