"""Dilated kNN, dense layout (models/gcn_lib/dense/torch_edge.py:6-76): x [B,C,N,1] -> [2,B,N,k]."""
import torch
from torch import nn

from ..sparse.torch_edge import _knn_call, pairwise_distance  # noqa: F401


def dense_knn_matrix(x, k=16, dilation=1):
    """Row 0 = neighbour ids (no batch offset), row 1 = centre ids."""
    xb = x.transpose(2, 1).squeeze(-1)
    B, N, _ = xb.shape
    nbr, ctr = _knn_call(xb, k, dilation, False)
    return torch.stack((nbr.view(B, N, k), ctr.view(B, N, k)), dim=0)


class DenseDilated(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k

    def forward(self, edge_index):
        if self.stochastic and torch.rand(1) < self.epsilon and self.training:
            pick = torch.randperm(self.k * self.dilation)[:self.k]
            return edge_index[:, :, :, pick]
        return edge_index[:, :, :, ::self.dilation]


class DenseDilatedKnnGraph(nn.Module):
    def __init__(self, k=9, dilation=1, stochastic=False, epsilon=0.0):
        super().__init__()
        self.dilation, self.stochastic, self.epsilon, self.k = dilation, stochastic, epsilon, k
        self._dilated = DenseDilated(k, dilation, stochastic, epsilon)
        self.knn = dense_knn_matrix

    def forward(self, x):
        if self.stochastic and self.training:
            return self._dilated(self.knn(x, self.k * self.dilation))
        return dense_knn_matrix(x, self.k, dilation=self.dilation)
