"""DiffPool drop-in (models/diff_pooling.py:11-133 of the reference; PyG 2.2.0 ``DenseSAGEConv`` and
``dense_diff_pool`` semantics restated from SURVEY.md Appendix A).

Same constructors, ``forward(x, adj, mask=None) -> (x, l_total, e_total)`` and state_dict keys
(``diffpool_layers.{i}.gnn_pool.layers.0.lin_rel.weight`` ...).  Two paths behind ``DiffPool.forward``:
* the reference's size (one shared 146-node adjacency, 146 -> 37 -> 10 clusters, hundreds of samples): the whole module
  -- both DenseSAGE stacks, softmax, S^T.X, S^T.A.S, link and entropy terms -- is ONE kernel per direction
  (``csrc/diffpool_fused.cu``, one persistent CTA per sample, shared-memory resident), see ``DiffPool._fused_plan``;
* everything else goes layer by layer through ``dense_ops``: the contractions (A.X, S^T.X, S^T.A.S, S.S^T) and the
  DenseSAGE projections on the hand-written tcgen05/TMA bf16 GEMM (``mlg_gemm_bf16``) once the operands are large
  enough for the tensor pipe (``dense_ops.TENSOR_CORE_MIN``), e.g. the synthetic N=10k / K=2.5k / C=1024 shape of
  SURVEY section 8(d); the library fp32 product only for small operands that fit neither path.
"""
from math import ceil

import torch
from torch import nn
from torch.nn import functional as F

from .. import dense_ops


class DenseSAGEConv(nn.Module):
    """out = lin_rel((A.X) / clamp(rowsum(A), 1)) + lin_root(X), optional L2 normalisation and mask."""

    def __init__(self, in_channels, out_channels, normalize=False, bias=True):
        super().__init__()
        self.in_channels, self.out_channels, self.normalize = in_channels, out_channels, normalize
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=False)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=bias)

    @staticmethod
    def aggregate(x, adj):
        """(A.X) / clamp(rowsum(A), 1): depends on the input only, so the assignment conv and the embedding conv of one
        DiffPool layer (same x, same adj) share it."""
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        return dense_ops.matmul(adj, x) / adj.sum(dim=-1, keepdim=True).clamp(min=1)

    def forward(self, x, adj, mask=None, agg=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        out = self.aggregate(x, adj) if agg is None else agg
        out = dense_ops.linear(out, self.lin_rel) + dense_ops.linear(x, self.lin_root)
        if self.normalize:
            out = F.normalize(out, p=2.0, dim=-1)
        if mask is not None:
            out = out * mask.view(x.size(0), x.size(1), 1).to(x.dtype)
        return out


class DenseGraphConv(nn.Module):
    def __init__(self, in_channels, out_channels, aggr="add", bias=True):
        super().__init__()
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=bias)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, adj, mask=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        out = self.lin_rel(dense_ops.matmul(adj, x)) + self.lin_root(x)
        if mask is not None:
            out = out * mask.view(x.size(0), x.size(1), 1).to(x.dtype)
        return out


def dense_diff_pool(x, adj, s, mask=None, normalize=True):
    """softmax(S); S^T X; S^T A S; link loss ||A - S S^T||_F / numel(A); entropy mean(sum(-S log(S + 1e-15)))."""
    x = x.unsqueeze(0) if x.dim() == 2 else x
    adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
    s = s.unsqueeze(0) if s.dim() == 2 else s
    s = torch.softmax(s, dim=-1)
    if mask is not None:
        m = mask.view(x.size(0), x.size(1), 1).to(x.dtype)
        x, s = x * m, s * m
    st = s.transpose(1, 2)
    out = dense_ops.matmul(st, x)
    out_adj = dense_ops.matmul(dense_ops.matmul(st, adj), s)
    link = torch.norm(adj - dense_ops.matmul(s, st), p=2)
    if normalize:
        link = link / adj.numel()
    ent = (-s * torch.log(s + 1e-15)).sum(dim=-1).mean()
    return out, out_adj, link, ent


class SAGEConvolutions(nn.Module):
    def __init__(self, num_layers, in_channels, out_channels, residual=True):
        super().__init__()
        self.num_layers, self.residual = num_layers, residual
        self.layers, self.bns = nn.ModuleList(), nn.ModuleList()
        for i in range(num_layers - 1):
            self.layers.append(DenseSAGEConv(in_channels if i == 0 else out_channels, out_channels, normalize=True))
            self.bns.append(nn.BatchNorm1d(out_channels))
        self.layers.append(DenseSAGEConv(in_channels if num_layers == 1 else out_channels, out_channels, normalize=True))

    def forward(self, x, adj, mask=None, agg=None):
        """``agg``: the first layer's neighbourhood mean when the caller already has it (DiffPoolLayer)."""
        for i in range(self.num_layers - 1):
            x_new = F.relu(self.layers[i](x, adj, mask, agg if i == 0 else None))
            b, n, c = x_new.size()
            x_new = self.bns[i](x_new.view(-1, c)).view(b, n, c)
            x = x + x_new if (self.residual and x.shape == x_new.shape) else x_new
        return self.layers[self.num_layers - 1](x, adj, mask, agg if self.num_layers == 1 else None)


class DiffPoolLayer(nn.Module):
    def __init__(self, dim_input, dim_embedding, current_num_clusters, no_new_clusters):
        super().__init__()
        self.gnn_pool = SAGEConvolutions(1, dim_input, no_new_clusters)
        self.gnn_embed = SAGEConvolutions(1, dim_input, dim_embedding)

    def forward(self, x, adj, mask=None):
        agg = DenseSAGEConv.aggregate(x, adj)       # one A.X for both convs (the reference computes it twice)
        s = self.gnn_pool(x, adj, mask, agg)
        x = self.gnn_embed(x, adj, mask, agg)
        return dense_diff_pool(x, adj, s, mask)


class DiffPool(nn.Module):
    def __init__(self, num_features, num_classes, max_num_nodes, num_layers, gnn_hidden_dim, gnn_output_dim, args,
                 encode_edge=False, pre_sum_aggr=False):
        super().__init__()
        self.args = args
        self.encode_edge = encode_edge
        self.max_num_nodes = max_num_nodes
        self.pooling_type = args.pooling_type
        self.num_pooling_layers = num_layers
        coarse = 0.1 if num_layers == 1 else 0.25        # the DiffPool paper's coarsening factors
        if pre_sum_aggr:
            self.initial_embed = DenseGraphConv(num_features, gnn_output_dim)
        else:
            self.initial_embed = SAGEConvolutions(1, num_features, gnn_output_dim)
        clusters = ceil(coarse * max_num_nodes)
        current = max_num_nodes
        pools, afters = [], []
        for i in range(num_layers):
            d_in = num_features if i == 0 else gnn_hidden_dim
            d_out = gnn_output_dim if i == num_layers - 1 else gnn_hidden_dim
            pools.append(DiffPoolLayer(d_in, d_out, current, clusters))
            current, clusters = clusters, ceil(clusters * coarse)
            afters.append(SAGEConvolutions(args.after_pooling_layer, d_out, d_out))
        self.diffpool_layers = nn.ModuleList(pools)
        self.after_pool_layers = nn.ModuleList(afters)

    FUSED = True      # reference-sized problems through the one-kernel path (csrc/diffpool_fused.cu)

    def _fused_plan(self, x, adj, mask):
        """(dims, weights) when the fused kernel applies: CUDA fp32, no mask, ONE shared 2-D adjacency that needs no
        gradient, single-conv SAGEConvolutions everywhere, at most two pooling layers, and the per-sample working set
        fits one SM's shared memory; else None (module path: library GEMMs / the tensor-core path for GEMM-sized inputs)."""
        from .. import functional as Fn
        if not (self.FUSED and x.is_cuda and x.dim() == 3 and adj.dim() == 2 and mask is None and not adj.requires_grad
                and x.dtype == torch.float32 and 1 <= self.num_pooling_layers <= 2):
            return None
        dims, weights = [], []
        n, c = x.shape[1], x.shape[2]
        for lay, aft in zip(self.diffpool_layers, self.after_pool_layers):
            convs = (lay.gnn_pool, lay.gnn_embed, aft)
            if any(sc.num_layers != 1 or not sc.layers[0].normalize for sc in convs):
                return None
            k, h = lay.gnn_pool.layers[0].out_channels, lay.gnn_embed.layers[0].out_channels
            if (lay.gnn_pool.layers[0].in_channels != c or lay.gnn_embed.layers[0].in_channels != c
                    or aft.layers[0].in_channels != h or aft.layers[0].out_channels != h):
                return None
            dims.append((n, c, k, h))
            for sc in convs:
                conv = sc.layers[0]
                if conv.lin_root.bias is None or conv._forward_hooks:
                    return None
                weights += [conv.lin_rel.weight, conv.lin_root.weight, conv.lin_root.bias]
            n, c = k, h
        dims = tuple(dims)
        return (dims, weights) if Fn.DiffPoolFused.supported(dims) else None

    def forward(self, x, adj, mask=None):
        from .. import _cabi
        _cabi.require_cuda(x)            # no CPU path: the product computes on the B200 kernels only
        plan = self._fused_plan(x, adj, mask)
        if plan is not None:
            from .. import functional as Fn
            return Fn.DiffPoolFused.apply(x, adj, plan[0], *plan[1])
        l_total, e_total = 0, 0
        with dense_ops.operand_cache():      # one bf16 copy per operand and orientation for the whole forward
            for i in range(self.num_pooling_layers):
                x, adj, l, e = self.diffpool_layers[i](x, adj, mask if i == 0 else None)
                x = self.after_pool_layers[i](x, adj)
                l_total = l_total + l
                e_total = e_total + e
        return x, l_total, e_total
