"""The caller of the hot path: one training step as the reference's ``train()`` runs it
(train.py:38-68; optimizer / criterion set-up train.py:112-125), plus the data-parallel wrapper the
reference lacks: patient graphs are independent, so every rank steps on its own
batch and ONE NCCL all-reduce (op=AVG, the 1/world scale folded in) over a flat fp32 gradient bucket
synchronises the replicas.  Parameters that never receive a gradient (``lin_l.weight`` of every
SAGEConv, ``info_mask``) are left out of the bucket and of the optimizer state.
"""
import torch
import torch.distributed as dist


def shard_indices(n_items, rank, world, per_rank_batch, epoch_seed=0, drop_last=True):
    """Per-epoch partition of the patient list: one shuffled permutation (same seed on every rank), rank r takes
    batches r, r+world, ... of size ``per_rank_batch`` -- disjoint across ranks (train.py:303-323 uses a single
    DataLoader with shuffle + drop_last; this is its data-parallel split).  Returns a list of index tensors."""
    g = torch.Generator()
    g.manual_seed(epoch_seed)
    perm = torch.randperm(n_items, generator=g)
    step = per_rank_batch * world
    n_steps = n_items // step if drop_last else -(-n_items // step)
    out = []
    for s in range(n_steps):
        lo = s * step + rank * per_rank_batch
        out.append(perm[lo:lo + per_rank_batch])
    return out


class GradBucket:
    """Flat fp32 gradient bucket: every ``p.grad`` is a view into one contiguous buffer, so the data-parallel
    synchronisation is ONE all-reduce with no packing/unpacking (NCCL: op=AVG folds the 1/world scale in;
    gloo has no AVG, so SUM + scale).  Parameters that never get a gradient are not in the bucket."""

    def __init__(self, params):
        self.params = list(params)
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce(self, world):
        if world <= 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)


class Trainer:
    def __init__(self, model, args, criterion_weight=None, world_size=1, fused_adam=True):
        self.model, self.args, self.world = model, args, world_size
        self.params = [p for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
        dev = self.params[0].device
        self.bucket = GradBucket(self.params)
        self.flat = self.bucket.flat
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
        self.bucket.zero()                       # == optimizer.zero_grad() with grads kept as bucket views
        loss = self.loss(batch)
        loss.backward()
        self.bucket.all_reduce(self.world)
        if self.args.clip_grad:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=20, norm_type=2)
        self.opt.step()
        return loss.detach()
