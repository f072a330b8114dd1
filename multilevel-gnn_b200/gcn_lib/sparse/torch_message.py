"""GenMessagePassing / MsgNorm drop-ins (models/gcn_lib/sparse/torch_message.py:8-85,168-179).

``aggregate(inputs, index, ptr, dim_size)`` keeps the reference signature and runs the CSR
segment-softmax / power-mean kernel on messages given directly; ``GENConv`` (torch_vertex.py) uses the
fully fused path instead (gather + message + aggregate + MsgNorm + residual in one kernel)."""
import torch
from torch import nn

from ... import functional as Fn
from ... import graph

_SOFTMAX = ('softmax_sg', 'softmax', 'softmax_sum')
_POWER = ('power', 'power_sum')


class GenMessagePassing(nn.Module):
    def __init__(self, aggr='softmax', t=1.0, learn_t=False, p=1.0, learn_p=False, y=0.0, learn_y=False):
        super().__init__()
        self.aggr = aggr
        self.node_dim = -2
        self.learn_t = False
        if aggr in _SOFTMAX:
            if learn_t and aggr in ('softmax', 'softmax_sum'):
                self.learn_t = True
                self.t = nn.Parameter(torch.Tensor([t]), requires_grad=True)
            else:
                self.t = t
            if aggr == 'softmax_sum':
                self.y = nn.Parameter(torch.Tensor([y]), requires_grad=learn_y)
        elif aggr in _POWER:
            self.p = nn.Parameter(torch.Tensor([p]), requires_grad=True) if learn_p else p
            if aggr == 'power_sum':
                self.y = nn.Parameter(torch.Tensor([y]), requires_grad=learn_y)
        elif aggr not in ('add', 'mean', 'max', None):
            raise NotImplementedError('To be implemented')

    # --- kernel argument helpers ----------------------------------------------------------------
    def _kernel_args(self):
        aggr = self.aggr if self.aggr is not None else 'add'
        t = getattr(self, 't', 1.0)
        p = getattr(self, 'p', 1.0)
        y = getattr(self, 'y', None) if aggr in ('softmax_sum', 'power_sum') else None
        learn = self.learn_t if aggr in _SOFTMAX else (torch.is_tensor(p) and aggr in _POWER)
        return aggr, t, p, y, learn

    def aggregate(self, inputs, index, ptr=None, dim_size=None):
        """inputs [E,H] messages, index [E] target of each message -> [dim_size, H]."""
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() else 0
        topo = graph.topology(torch.stack([index, index]), dim_size)
        aggr, t, p, y, learn = self._kernel_args()
        out = Fn.GenAggregate.apply(None, inputs, t, p, y, None, topo, aggr, 0.0, Fn.EPI_NONE, learn)
        if aggr == 'softmax_sum' or aggr == 'power_sum':
            self.sigmoid_y = torch.sigmoid(self.y)
        return out

    def propagate(self, edge_index, size=None, x=None, edge_attr=None, **kwargs):
        """source->target flow with GENConv's message (relu(x_j + e) + eps); unfused epilogue."""
        topo = graph.topology(edge_index, x.shape[0])
        aggr, t, p, y, learn = self._kernel_args()
        return Fn.GenAggregate.apply(x, edge_attr, t, p, y, None, topo, aggr, getattr(self, 'eps', 1e-7),
                                     Fn.EPI_NONE, learn)


class PathwayMessagePassing(GenMessagePassing):
    """torch_message.py:88-165: the same aggregation arithmetic as GenMessagePassing under another name (used by
    PathwayConv).  The reference's constructor cannot build the power modes (it calls ``super(GenMessagePassing, self)``
    on a class that is not a GenMessagePassing, :111): mirrored as the same TypeError."""

    def __init__(self, aggr='softmax', t=1.0, learn_t=False, p=1.0, learn_p=False, y=0.0, learn_y=False):
        if aggr in _POWER:
            raise TypeError("super(type, obj): obj must be an instance or subtype of type "
                            "(PathwayMessagePassing with a power aggregation fails the same way in the reference)")
        super().__init__(aggr=aggr, t=t, learn_t=learn_t, p=p, learn_p=learn_p, y=y, learn_y=learn_y)


class MsgNorm(nn.Module):
    """msg / max(||msg||_2, 1e-12) * ||x||_2 * msg_scale (torch_message.py:175-179).

    Inside ``GENConv`` this is the fused epilogue of the aggregation kernel (MLG_EPI_MSGNORM) and this
    ``forward`` is never called; it only keeps the module callable on its own, as plain tensor algebra
    on whatever device its inputs live on."""

    def __init__(self, learn_msg_scale=False):
        super().__init__()
        self.msg_scale = nn.Parameter(torch.Tensor([1.0]), requires_grad=learn_msg_scale)

    def forward(self, x, msg, p=2):
        nrm = torch.linalg.vector_norm(msg, ord=p, dim=1, keepdim=True).clamp_min(1e-12)
        return msg / nrm * torch.linalg.vector_norm(x, ord=p, dim=1, keepdim=True) * self.msg_scale
