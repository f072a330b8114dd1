"""Dilated kNN graph construction, sparse layout (models/gcn_lib/sparse/torch_edge.py:6-104).

The distance matrix is never materialised: ``mlg_knn_graph`` fuses the tiled fp32 distance with a
per-row top-k and the dilation stride; graphs of >= 4096 points search their candidates on the tensor cores (3xTF32) and
certify them in fp32 (csrc/knn.cu) -- same output.  Tie contract: ascending distance, lowest index first inside
exact fp32 ties (``torch.topk`` leaves tie order unspecified)."""
import torch
from torch import nn

from ... import _cabi


def _knn_call(xb, k, dilation, add_offset):
    """xb [B,N,D] fp32 CUDA -> (nbr, ctr) int64 [B*N*k]; k = neighbours kept after dilation."""
    L = _cabi.lib()
    _cabi.require_cuda(xb)
    xb = xb.detach().contiguous().float()
    B, N, D = xb.shape
    nbr = torch.empty(B * N * k, dtype=torch.int64, device=xb.device)
    ctr = torch.empty_like(nbr)
    ws = torch.empty((int(L.mlg_knn_workspace_bytes(B, N, D, k, dilation)) + 3) // 4, dtype=torch.float32, device=xb.device)
    with torch.cuda.device(xb.device):
        _cabi.check(L.mlg_knn_graph(_cabi.fptr(xb), B, N, D, k, dilation, int(add_offset), _cabi.lptr(nbr),
                                    _cabi.lptr(ctr), None, _cabi.fptr(ws), ws.numel() * 4, _cabi.stream_ptr()),
                    "mlg_knn_graph")
    return nbr, ctr


def pairwise_distance(x):
    """[B,N,D] -> [B,N,N] squared distances.  Kept for API parity only (plain tensor algebra); the kNN
    entry points below never materialise this matrix."""
    inner = -2 * torch.matmul(x, x.transpose(2, 1))
    sq = torch.sum(x * x, dim=-1, keepdim=True)
    return sq + inner + sq.transpose(2, 1)


def _batch_size(batch):
    return 1 if batch is None else int(batch[-1]) + 1


def knn_matrix(x, k=16, batch=None, dilation=1):
    """x [B*N, D] (equal-sized graphs, as the reference's ``view`` assumes) -> (nn_idx, center_idx),
    each [1, B*N*k], ids carrying the per-graph node offset.  With ``dilation`` d the k*d nearest are
    ranked and ranks 0, d, 2d, ... are returned (Dilated fused in)."""
    bsz = _batch_size(batch)
    xb = x.view(bsz, -1, x.shape[-1])
    nbr, ctr = _knn_call(xb, k, dilation, True)
    return nbr.view(1, -1), ctr.view(1, -1)


def knn_graph_matrix(x, k=16, batch=None):
    """edge_index [2, B*N*k]: row 0 = neighbour, row 1 = centre."""
    nn_idx, center_idx = knn_matrix(x, k, batch)
    return torch.cat((nn_idx, center_idx), dim=0)


class Dilated(nn.Module):
    """Every ``dilation``-th neighbour of a centre-major list with k*dilation entries per centre; the
    stochastic branch (random k of k*d while training, with probability epsilon) is kept."""

    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index, batch=None):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            num = self.k * self.dilation
            pick = torch.randperm(num)[:self.k]
            return edge_index.view(2, -1, num)[:, :, pick].reshape(2, -1)
        return edge_index[:, ::self.dilation]


class DilatedKnnGraph(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0, knn='matrix'):
        super().__init__()
        if knn != 'matrix':
            raise NotImplementedError("knn='%s' goes through torch_cluster in the reference (dependency absent)" % knn)
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = Dilated(k, dilation, stochastic, epsilon)
        self.knn = knn_graph_matrix

    def forward(self, x, batch):
        if self.stochastic and self.training:
            # random-k-of-k*d needs the full ranked list
            return self._dilated(self.knn(x, self.k * self.dilation, batch), batch)
        nn_idx, center_idx = knn_matrix(x, self.k, batch, dilation=self.dilation)
        return torch.cat((nn_idx, center_idx), dim=0)
