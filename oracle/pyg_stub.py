"""Container-only stand-in for the third-party packages the reference imports.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned versions being imitated
(/root/reference/requirements.txt:147-151): torch-geometric 2.2.0,
torch-scatter 2.1.0, torch-cluster 1.6.0, torch-sparse 0.6.16.  Their sources
are not in the container; the semantics below restate their published
behaviour.  ``install()`` registers the fakes in
``sys.modules`` so that ``/root/reference``'s files import unmodified.
"""
import inspect
import sys
import types

import torch
import torch.nn.functional as F
from torch import nn


# ----------------------------------------------------------------------------
# torch_scatter
# ----------------------------------------------------------------------------
def _expand_index(index, src, dim):
    if dim < 0:
        dim += src.dim()
    if index.dim() == 1:
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.view(shape)
    return index.expand_as(src), dim


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    index, dim = _expand_index(index, src, dim)
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ("sum", "add"):
        return src.new_zeros(shape).scatter_add_(dim, index, src)
    if reduce == "mean":
        tot = src.new_zeros(shape).scatter_add_(dim, index, src)
        cnt = src.new_zeros(shape).scatter_add_(dim, index, torch.ones_like(src))
        return tot / cnt.clamp(min=1)
    if reduce in ("max", "min"):
        amode = "amax" if reduce == "max" else "amin"
        res = src.new_zeros(shape).scatter_reduce(dim, index, src, amode, include_self=False)
        return res
    raise ValueError(reduce)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "mean")


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "max"), None


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "min"), None


def scatter_softmax(src, index, dim=-1, eps=1e-12, dim_size=None):
    # torch_scatter 2.1.0 composite: max per index, exp(src - max), sum, divide
    # (no epsilon in the denominator).
    index, dim = _expand_index(index, src, dim)
    n = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = n
    mx = src.new_zeros(shape).scatter_reduce(dim, index, src, "amax", include_self=False)
    rec = (src - mx.gather(dim, index)).exp()
    den = src.new_zeros(shape).scatter_add_(dim, index, rec)
    return rec / den.gather(dim, index)


# ----------------------------------------------------------------------------
# torch_geometric.utils
# ----------------------------------------------------------------------------
def degree(index, num_nodes=None, dtype=None):
    if num_nodes is None:
        num_nodes = int(index.max()) + 1
    out = torch.zeros(num_nodes, dtype=dtype or torch.get_default_dtype(), device=index.device)
    return out.scatter_add_(0, index, torch.ones(index.numel(), dtype=out.dtype, device=index.device))


def remove_self_loops(edge_index, edge_attr=None):
    keep = edge_index[0] != edge_index[1]
    return edge_index[:, keep], (None if edge_attr is None else edge_attr[keep])


def add_self_loops(edge_index, edge_attr=None, fill_value=1.0, num_nodes=None):
    if num_nodes is None:
        num_nodes = int(edge_index.max()) + 1
    loop = torch.arange(num_nodes, dtype=edge_index.dtype, device=edge_index.device)
    edge_index = torch.cat([edge_index, loop.unsqueeze(0).repeat(2, 1)], dim=1)
    if edge_attr is not None:
        fill = edge_attr.new_full((num_nodes,) + tuple(edge_attr.shape[1:]), fill_value)
        edge_attr = torch.cat([edge_attr, fill], dim=0)
    return edge_index, edge_attr


def to_dense_batch(*a, **k):  # imported by the reference, never called on the path
    raise NotImplementedError


def to_dense_adj(*a, **k):
    raise NotImplementedError


# ----------------------------------------------------------------------------
# torch_geometric.nn
# ----------------------------------------------------------------------------
class MessagePassing(nn.Module):
    """flow=source_to_target, node_dim=-2 message passing skeleton."""

    def __init__(self, aggr="add", flow="source_to_target", node_dim=-2, **kwargs):
        super().__init__()
        self.aggr = aggr
        self.node_dim = node_dim

    @staticmethod
    def _params(fn, skip=0):
        return list(inspect.signature(fn).parameters.keys())[skip:]

    def propagate(self, edge_index, size=None, **kwargs):
        n = None
        for v in kwargs.values():
            if torch.is_tensor(v) and v.dim() >= 2:
                n = v.size(self.node_dim)
                break
        if size is not None and size[1] is not None:
            n = size[1]
        margs = {}
        for name in self._params(self.message):
            if name.endswith("_j") or name.endswith("_i"):
                base = kwargs.get(name[:-2])
                sel = edge_index[0] if name.endswith("_j") else edge_index[1]
                margs[name] = None if base is None else base.index_select(self.node_dim, sel)
            elif name in kwargs:
                margs[name] = kwargs[name]
        msg = self.message(**margs)
        out = self.aggregate(msg, edge_index[1], ptr=None, dim_size=n)
        uargs = {k: kwargs[k] for k in self._params(self.update, skip=1) if k in kwargs}
        return self.update(out, **uargs)

    def message(self, x_j):
        return x_j

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        red = {"add": "sum", "sum": "sum", "mean": "mean", "max": "max", "min": "min"}[self.aggr]
        return scatter(inputs, index, dim=self.node_dim, dim_size=dim_size, reduce=red)

    def update(self, inputs):
        return inputs


class SAGEConv(MessagePassing):
    """Parameter shell of PyG 2.2.0 SAGEConv (the reference overrides forward/message/update)."""

    def __init__(self, in_channels, out_channels, aggr="mean", normalize=False,
                 root_weight=True, project=False, bias=True, **kwargs):
        super().__init__(aggr=aggr, **kwargs)
        self.in_channels, self.out_channels = in_channels, out_channels
        self.normalize, self.root_weight = normalize, root_weight
        self.lin_l = nn.Linear(in_channels, out_channels, bias=bias)
        if root_weight:
            self.lin_r = nn.Linear(in_channels, out_channels, bias=False)


class DenseSAGEConv(nn.Module):
    def __init__(self, in_channels, out_channels, normalize=False, bias=True):
        super().__init__()
        self.normalize = normalize
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=False)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=bias)

    def forward(self, x, adj, mask=None):
        x = x.unsqueeze(0) if x.dim() == 2 else x
        adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
        out = torch.matmul(adj, x)
        out = out / adj.sum(dim=-1, keepdim=True).clamp(min=1)
        out = self.lin_rel(out) + self.lin_root(x)
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
        return self.lin_rel(torch.matmul(adj, x)) + self.lin_root(x)


def dense_diff_pool(x, adj, s, mask=None, normalize=True):
    x = x.unsqueeze(0) if x.dim() == 2 else x
    adj = adj.unsqueeze(0) if adj.dim() == 2 else adj
    s = s.unsqueeze(0) if s.dim() == 2 else s
    s = torch.softmax(s, dim=-1)
    if mask is not None:
        m = mask.view(x.size(0), x.size(1), 1).to(x.dtype)
        x, s = x * m, s * m
    out = torch.matmul(s.transpose(1, 2), x)
    out_adj = torch.matmul(torch.matmul(s.transpose(1, 2), adj), s)
    link_loss = adj - torch.matmul(s, s.transpose(1, 2))
    link_loss = torch.norm(link_loss, p=2)
    if normalize:
        link_loss = link_loss / adj.numel()
    ent_loss = (-s * torch.log(s + 1e-15)).sum(dim=-1).mean()
    return out, out_adj, link_loss, ent_loss


def _global_pool(reduce):
    def pool(x, batch, size=None):
        size = int(batch.max()) + 1 if size is None else size
        return scatter(x, batch, dim=0, dim_size=size, reduce=reduce)
    return pool


global_add_pool = _global_pool("sum")
global_mean_pool = _global_pool("mean")
global_max_pool = _global_pool("max")


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("not on the hot path; not stubbed")


# ----------------------------------------------------------------------------
# torch_geometric.data
# ----------------------------------------------------------------------------
class Data:
    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device):
        for k, v in list(vars(self).items()):
            if torch.is_tensor(v):
                setattr(self, k, v.to(device))
        return self

    def keys(self):
        return list(vars(self).keys())


def knn_graph(x, k, batch=None, loop=False, flow="source_to_target"):
    raise NotImplementedError("torch_cluster.knn_graph: alternative kNN path, dependency absent")


def install():
    """Register the fake packages; idempotent."""
    if "torch_geometric" in sys.modules and getattr(sys.modules["torch_geometric"], "_mlg_stub", False):
        return
    this = sys.modules[__name__]

    ts = types.ModuleType("torch_scatter")
    for n in ("scatter", "scatter_add", "scatter_mean", "scatter_max", "scatter_min", "scatter_softmax"):
        setattr(ts, n, getattr(this, n))

    tg = types.ModuleType("torch_geometric")
    tg._mlg_stub = True
    tg_nn = types.ModuleType("torch_geometric.nn")
    for n in ("MessagePassing", "SAGEConv", "DenseSAGEConv", "DenseGraphConv", "dense_diff_pool",
              "global_add_pool", "global_mean_pool", "global_max_pool"):
        setattr(tg_nn, n, getattr(this, n))
    for n in ("EdgeConv", "GATConv", "GCNConv", "GINConv", "TopKPooling"):
        setattr(tg_nn, n, type(n, (_Unavailable,), {}))
    tg_utils = types.ModuleType("torch_geometric.utils")
    for n in ("degree", "remove_self_loops", "add_self_loops", "to_dense_batch", "to_dense_adj"):
        setattr(tg_utils, n, getattr(this, n))
    tg_data = types.ModuleType("torch_geometric.data")
    tg_data.Data = Data
    tg_data.InMemoryDataset = object
    tg_data.DataLoader = object
    tg_data.extract_zip = lambda *a, **k: None
    tg.nn, tg.utils, tg.data = tg_nn, tg_utils, tg_data

    tc = types.ModuleType("torch_cluster")
    tc.knn_graph = knn_graph
    h5 = types.ModuleType("h5py")

    sys.modules.update({
        "torch_scatter": ts, "torch_geometric": tg, "torch_geometric.nn": tg_nn,
        "torch_geometric.utils": tg_utils, "torch_geometric.data": tg_data,
        "torch_cluster": tc, "h5py": h5,
    })
