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
            # backward kernels that produce a whole parameter gradient may write it here directly (functional.grad_slot)
            p._mlg_grad_slot = {"view": p.grad, "claimed": False}
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def store(self, grads):
        """Write a list of gradients (one per bucket parameter, None = no gradient) into the bucket with one
        multi-tensor copy.  Replaces zero() + AccumulateGrad's per-parameter in-place add (one tiny kernel per
        parameter: ~20 launches per step on the gbm model)."""
        if any(g is None for g in grads):
            if not hasattr(self, "_has_none"):
                self._has_none = True
            for p, g in zip(self.params, grads):
                if g is None:
                    p.grad.zero_()
        dst, src = [], []
        for p, g in zip(self.params, grads):
            p._mlg_grad_slot["claimed"] = False
            if g is not None and g.data_ptr() != p.grad.data_ptr():    # same pointer: written in place by its kernel
                dst.append(p.grad)
                src.append(g)
        if dst:
            torch._foreach_copy_(dst, src)

    def all_reduce(self, world):
        if world <= 1:
            return
        if dist.get_backend() == "nccl":
            dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)


class FlatAdam:
    """torch.optim.Adam (amsgrad off) as ONE kernel over flat buffers (mlg_adam_step): parameters are re-homed as views
    of one contiguous fp32 buffer laid out like the gradient bucket; the step counter is a device scalar so the
    update replays inside a CUDA graph.  lr / betas / eps / weight_decay as in train.py:112."""

    def __init__(self, bucket, lr, betas, eps=1e-8, weight_decay=0.0):
        self.bucket, self.lr, self.betas, self.eps, self.wd = bucket, lr, betas, eps, weight_decay
        flat_g = bucket.flat
        self.flat_p = torch.empty_like(flat_g)
        off = 0
        for p in bucket.params:
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p)
            off += n
        self.exp_avg = torch.zeros_like(flat_g)
        self.exp_avg_sq = torch.zeros_like(flat_g)
        self.step_dev = torch.zeros(1, dtype=torch.float32, device=flat_g.device)

    def step(self):
        from . import _cabi
        L = _cabi.lib()
        with torch.cuda.device(self.flat_p.device):
            _cabi.check(L.mlg_adam_step(_cabi.fptr(self.flat_p), _cabi.fptr(self.bucket.flat), _cabi.fptr(self.exp_avg),
                                        _cabi.fptr(self.exp_avg_sq), _cabi.fptr(self.step_dev), self.flat_p.numel(),
                                        float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                        float(self.wd), _cabi.stream_ptr()), "mlg_adam_step")


class Trainer:
    """One training step of train.py:38-68 on the B200 kernels; ``capture()`` turns the whole step
    (forward, loss, backward, NCCL all-reduce, fused Adam) into ONE CUDA graph replayed per step."""

    def __init__(self, model, args, criterion_weight=None, world_size=1, fused_adam=True):
        self.model, self.args, self.world = model, args, world_size
        self.params = [p for n, p in model.named_parameters() if p.requires_grad and not n.endswith("lin_l.weight")]
        dev = self.params[0].device
        self.bucket = GradBucket(self.params)
        self.flat = self.bucket.flat
        if fused_adam and dev.type == "cuda":
            self.opt = FlatAdam(self.bucket, lr=args.lr, betas=(args.beta1, args.beta2), weight_decay=args.wd)
        else:
            self.opt = torch.optim.Adam(self.params, lr=args.lr, betas=(args.beta1, args.beta2), weight_decay=args.wd)
        self.weight = criterion_weight
        self.crit = torch.nn.BCELoss(weight=criterion_weight) if args.weight_balance else torch.nn.BCELoss()
        self.graph = None
        self.graph_update = None
        self.static_batch = None
        self.static_loss = None
        self._copy_stream = None

    def loss(self, batch):
        pred, feat = self.model(batch)
        loss = self.crit(pred.to(torch.float32), batch.y.reshape(-1, 2).to(torch.float32))
        return loss + self.model.get_feature_loss(feat)

    def _fwd_bwd(self, batch):
        self.model.train()
        loss = self.loss(batch)
        # == optimizer.zero_grad(); loss.backward() with the gradients landing in the bucket views (p.grad)
        from . import functional as Fn
        Fn.GRAD_SLOTS_ENABLED = loss.is_cuda
        try:
            grads = torch.autograd.grad(loss, self.params, allow_unused=True)
        finally:
            Fn.GRAD_SLOTS_ENABLED = False
        self.bucket.store(grads)
        return loss.detach()

    def _update(self):
        if self.args.clip_grad:
            torch.nn.utils.clip_grad_norm_(self.params, max_norm=20, norm_type=2)
        self.opt.step()

    def _step_eager(self, batch):
        loss = self._fwd_bwd(batch)
        self.bucket.all_reduce(self.world)
        self._update()
        return loss

    def capture(self, batch, warmup=3):
        """Capture the step as CUDA graphs over ``batch``'s device tensors (they become the static input
        buffers).  The CSR / pool layouts are built during the warm-up steps (their one-time host syncs are not
        capturable).  Single GPU: ONE graph (fwd, loss, bwd, Adam).  Data parallel: graph A (fwd, loss, bwd) ->
        the NCCL all-reduce issued eagerly on the same stream -> graph B (Adam); the collective stays outside
        the capture so that NCCL's own stream/event management never interferes with it."""
        self.static_batch = batch
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager(batch)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if self.world > 1:
            self.graph_update = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_loss = self._fwd_bwd(batch)
            with torch.cuda.graph(self.graph_update, pool=self.graph.pool()):
                self._update()
        else:
            self.graph_update = None
            with torch.cuda.graph(self.graph):
                self.static_loss = self._step_eager(batch)
        return self

    def _replay(self):
        self.graph.replay()
        if self.graph_update is not None:
            self.bucket.all_reduce(self.world)
            self.graph_update.replay()
        return self.static_loss

    # fields that are dataset constants when a batch carries a ``topology_key``: the key promises that the edge list,
    # edge weights, node->graph vector and pooling layout are identical for every batch with that key (one gene network
    # shared by all patients, dataloader/multiloader.py:687-698) -- exactly what MultilevelGNN.forward already assumes
    # when it caches the CSR / pool layout under that key.  They are uploaded once, not per step.
    STATIC_TOPOLOGY_FIELDS = ("edge_index", "edge_attr", "batch", "gene_pca_match", "raw_indice")

    def _step_fields(self, host_batch):
        key = getattr(self.static_batch, "topology_key", None)
        same = key is not None and getattr(host_batch, "topology_key", None) == key
        return [k for k, v in vars(self.static_batch).items()
                if torch.is_tensor(v) and not (same and k in self.STATIC_TOPOLOGY_FIELDS)]

    def h2d_bytes(self, host_batch):
        """Bytes ``load_batch`` / ``prefetch`` copy host -> device for this batch."""
        return sum(getattr(host_batch, k).numel() * getattr(host_batch, k).element_size()
                   for k in self._step_fields(host_batch))

    def load_batch(self, host_batch):
        """Copy a (pinned) host batch into the captured graph's static device buffers (async H2D)."""
        for k in self._step_fields(host_batch):
            getattr(self.static_batch, k).copy_(getattr(host_batch, k), non_blocking=True)

    def prefetch(self, host_batch):
        """Start the H2D copy of the NEXT batch on a side stream into staging buffers while the current step's
        graph is still running; ``step_prefetched`` then moves it into the static buffers (device-to-device)
        and replays.  This is the double buffering the reference's synchronous ``batch.to(device)``
        (train.py:42) lacks."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
            self._staging = {k: torch.empty_like(getattr(self.static_batch, k)) for k in self._step_fields(host_batch)}
            self._staged = torch.cuda.Event()
            self._consumed = torch.cuda.Event()
            self._consumed.record()
        self._copy_stream.wait_event(self._consumed)        # staging may not be overwritten before it was consumed
        with torch.cuda.stream(self._copy_stream):
            if set(self._staging) != set(self._step_fields(host_batch)):
                raise ValueError("prefetch: this batch's topology_key differs from the one the staging buffers were "
                                 "sized for; use step(batch) for a batch with a different topology")
            for k, v in self._staging.items():
                v.copy_(getattr(host_batch, k), non_blocking=True)
            self._staged.record()

    def step_prefetched(self):
        cur = torch.cuda.current_stream()
        cur.wait_event(self._staged)
        for k, v in self._staging.items():
            getattr(self.static_batch, k).copy_(v, non_blocking=True)
        self._consumed.record()
        return self._replay()

    def step(self, batch=None):
        """forward, loss, backward, (all-reduce), Adam.  Returns the loss as a device scalar (the
        reference's ``loss.item()`` host sync is the caller's choice).  With a captured graph, ``batch`` must
        be the static batch (or None); use ``load_batch`` to feed new data."""
        if self.graph is not None:
            if batch is not None and batch is not self.static_batch:
                self.load_batch(batch)
            return self._replay()
        return self._step_eager(batch)
