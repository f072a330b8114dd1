"""One training step with the REFERENCE's own modules on the CPU -- TEST INFRASTRUCTURE / CPU BASELINE ONLY.

``RefTrainer`` instantiates the reference's unmodified ``models/multilevel_gnn.py::MultilevelGNN`` (from /root/reference in
the build container, from the staged copy oracle/_ref/ on the GPU box; both behind oracle/pyg_stub.py) and steps it exactly
as train.py:38-68 / 112-125 do: model(batch), get_feature_loss, BCELoss(weight) (+ feature loss), zero_grad, backward,
torch.optim.Adam.step.  Used by bench.py's ``--impl reference`` / ``cpu_baseline`` legs (kind "reference") and by the GPU
parity test that steps the CUDA trainer and this one side by side.
"""
import torch

from . import ref_import
from .pyg_stub import Data


def available():
    return ref_import.available()


class RefTrainer:
    def __init__(self, config, state_dict=None, criterion_weight=None, pathway_indexs=None, info_mask=None, **overrides):
        self.ns = ref_import.load()
        self.args = ref_import.default_args(config + ".yaml", **overrides)
        self.model = self.ns.multilevel_gnn.MultilevelGNN(self.args)
        if info_mask is not None:
            self.model.set_info_mask(info_mask.clone())
        if pathway_indexs is not None:
            self.model.set_pathway_indexs(pathway_indexs.clone())
        if state_dict is not None:
            self.model.load_state_dict({k: v.detach().cpu().clone() for k, v in state_dict.items()}, strict=True)
        a = self.args
        self.opt = torch.optim.Adam(self.model.parameters(), lr=a.lr, betas=(a.beta1, a.beta2), weight_decay=a.wd)
        self.crit = torch.nn.BCELoss(weight=criterion_weight) if a.weight_balance else torch.nn.BCELoss()

    def step(self, batch, train_mode=True):
        self.model.train(train_mode)
        fields = {k: v for k, v in vars(batch).items() if torch.is_tensor(v)}
        pred, feat = self.model(Data(**fields))
        loss_feature = self.model.get_feature_loss(feat)
        self.opt.zero_grad()
        loss = self.crit(pred.to(torch.float32), fields["y"].reshape(-1, 2).to(torch.float32))
        loss = loss + loss_feature
        loss.backward()
        self.opt.step()
        return loss.detach()

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.model.state_dict().items()}
