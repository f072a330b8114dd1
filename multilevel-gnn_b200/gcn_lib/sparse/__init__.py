from .torch_nn import MLP, act_layer, norm_layer  # noqa: F401
from .torch_edge import (Dilated, DilatedKnnGraph, knn_graph_matrix, knn_matrix,  # noqa: F401
                         pairwise_distance)
from .torch_message import GenMessagePassing, MsgNorm  # noqa: F401
from .torch_vertex import (GENConv, PathwayConv, SAGEConv, RSAGEConv, GraphConv, DynConv, PlainDynBlock,  # noqa: F401
                           ResDynBlock, DenseDynBlock, ResGraphBlock, DenseGraphBlock)
