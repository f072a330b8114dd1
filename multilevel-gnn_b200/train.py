"""The caller of the hot path: one training step as the reference's ``train()`` runs it
(train.py:38-68; optimizer / criterion set-up train.py:112-125), plus the data-parallel wrapper the
reference lacks: patient graphs are independent, so every rank steps on its own
batch and ONE NCCL all-reduce (op=AVG, the 1/world scale folded in) over a flat fp32 gradient bucket
synchronises the replicas.  Parameters that never receive a gradient (``lin_l.weight`` of every
SAGEConv, ``info_mask``) are left out of the bucket and of the optimizer state.
"""
import torch
import torch.distributed as dist


class Trainer:
    def __init__(self, model, args, criterion_weight=None, world_size=1, fused_adam=True):
        self.model, self.args, self.world = model, args, world_size
        self.params = [p for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
        dev = self.params[0].device
        # flat gradient bucket: every p.grad is a view into it, so the all-reduce needs no packing
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        kw = dict(lr=args.lr, betas=(args.beta1, args.beta2), weight_decay=args.wd)
        if fused_adam and dev.type == "cuda":
            kw["fused"] = True
        self.opt = torch.optim.Adam(self.params, **kw)
        self.weight = criterion_weight
        self.crit = torch.nn.BCELoss(weight=criterion_weight) if args.weight_balance else torch.nn.BCELoss()

    def loss(self, batch):
        pred, feat = self.model(batch)
        loss = self.crit(pred.to(torch.float32), batch.y.reshape(-1, 2).to(torch.float32))
        return loss + self.model.get_feature_loss(feat)

    def step(self, batch):
        """forward, loss, backward, (all-reduce), Adam.  Returns the loss as a device scalar (the
        reference's ``loss.item()`` host sync is the caller's choice)."""
        self.model.train()
        self.flat.zero_()                        # == optimizer.zero_grad() with grads kept as bucket views
        loss = self.loss(batch)
        loss.backward()
        if self.world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        if self.args.clip_grad:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=20, norm_type=2)
        self.opt.step()
        return loss.detach()
