"""DeeperGCN drop-in (models/deepergcn.py:17-358 of the reference): L x GENConv in res+ / res /
plain blocks over one shared edge embedding, pathway "global node" rows overwritten before the
stack and read out after it.  Every GENConv layer runs the fused sm_100a aggregation kernel; the
CSR / CSC of the edge list is built once per forward and shared by all layers.

Differences that are not numerical: the per-graph Python loops with ``.cpu().numpy()`` syncs
(deepergcn.py:217-223,283-317) are replaced by one index tensor built from ``node_size``.
"""
import logging

import torch
import torch.nn as nn
import torch.nn.functional as F

from ..functional import AffineEdge
from ..gcn_lib.sparse.torch_nn import norm_layer
from ..gcn_lib.sparse.torch_vertex import GENConv


class DeeperGCN(nn.Module):
    # carry a scalar edge attribute's embedding in factored form (functional.AffineEdge) instead of an [E, H] tensor;
    # False = the literal per-layer edge GEMM path (kept for the equivalence test)
    AFFINE_EDGE = True

    def __init__(self, args):
        super().__init__()
        self.num_layers = args.num_layers
        self.dropout = args.dropout
        self.block = args.block
        self.mul_attr = args.mul_attr
        hidden = args.hidden_channels
        self.hidden_channels = hidden
        self.learn_t, self.learn_p, self.msg_norm = args.learn_t, args.learn_p, args.msg_norm
        if self.block not in ('res+', 'res', 'plain'):
            raise NotImplementedError('To be implemented' if self.block == 'dense' else 'Unknown block Type')
        if args.conv != 'gen':
            raise Exception('Unknown Conv Type')
        if args.gnn_encoder != 'linear':
            raise NotImplementedError("gnn_encoder='conv1x1' is not selected by any shipped config")
        self.pca_only = args.pca_only
        self.gnn_encoder = args.gnn_encoder
        self.no_inter_drop = args.no_inter_drop
        self.no_inter_norm = args.no_inter_norm
        self.feature_drop_flag = args.feature_drop

        self.gcns = nn.ModuleList()
        self.norms = nn.ModuleList()
        for _ in range(self.num_layers):
            self.gcns.append(GENConv(hidden, hidden, aggr=args.gcn_aggr, t=args.t, learn_t=self.learn_t,
                                     p=args.p, learn_p=self.learn_p, gnn_encoder=self.gnn_encoder,
                                     msg_norm=self.msg_norm, learn_msg_scale=args.learn_msg_scale,
                                     encode_edge=args.conv_encode_edge, edge_feat_dim=hidden, norm=args.norm,
                                     mlp_layers=args.mlp_layers, pca_only=self.pca_only))
            self.norms.append(norm_layer(args.norm, hidden))

        self.node_embedding = args.node_embedding
        if self.node_embedding:
            self.node_embedding_encoder = nn.Embedding(args.node_num, args.node_embedding_dim)
        in_dim = 3 + (args.node_embedding_dim if self.node_embedding else 0) + (2 if self.mul_attr else 0)
        self.node_features_encoder = nn.Linear(in_dim, hidden)
        self.edge_encoder = nn.Linear(7 if args.use_column is None else 1, hidden)
        self.global_edge = args.global_edge
        if args.global_edge == "onehot":
            self.edge_encoder = nn.Embedding(args.pathway_edge_num, hidden)
        self.use_edge_attr = args.use_edge_attr
        self.pathway_global_node = args.pathway_global_node
        if self.pathway_global_node:
            self.pathway_num = args.pathway_num
            self.pathway_features_encoder = nn.Linear(6, hidden)
        self.num_layer_head = args.num_layer_head
        self.pathway_readout = args.pathway_readout
        self.pre_concat_age = args.pre_concat_age
        self.feature_drop = nn.Dropout(0.25)
        if self.pathway_readout == 'MSA':
            raise NotImplementedError("pathway_readout='MSA' is not selected by any shipped config")
        if self.pathway_readout == 'maxpool':
            readout_in = (args.pathway_num // 4) * hidden + (1 if args.pre_concat_age else 0)
            if args.pre_readout_drop:
                self.readout_func = nn.Sequential(nn.Linear(readout_in, hidden), nn.ReLU())
            else:
                self.readout_func = nn.Sequential(nn.Linear(readout_in, hidden), nn.ReLU(), nn.Dropout(0.5))
        if args.graph_pooling not in ("sum", "mean", "max"):
            raise Exception('Unknown Pool Type')
        self.graph_pooling = args.graph_pooling
        self.use_age = args.use_age
        width = hidden + 1 if (args.use_age and not args.pre_concat_age) else hidden
        self.graph_pred_linear = nn.Sequential()
        for i in range(args.num_layer_head - 1):
            self.graph_pred_linear.add_module(str(2 * i), nn.Linear(width, width))
            self.graph_pred_linear.add_module(str(2 * i + 1), nn.ReLU())
            if args.head_dropout:
                self.graph_pred_linear.add_module("drop{}".format(i), nn.Dropout(self.dropout))
        self.graph_pred_linear.add_module(str(2 * args.num_layer_head), nn.Linear(width, args.num_tasks))
        if args.all_init:
            self.init_weight()
        elif args.head_init:
            for m in self.graph_pred_linear.modules():
                if isinstance(m, nn.Linear):
                    nn.init.xavier_uniform_(m.weight.data)
                    nn.init.constant_(m.bias.data, 0.0)

    # ------------------------------------------------------------------------------------------
    def _pathway_rows(self, node_size, device):
        ends = torch.cumsum(node_size, dim=0).tolist()
        P = self.pathway_num
        return torch.cat([torch.arange(e - P, e, device=device) for e in ends])

    def _pool(self, h, batch):
        nb = int(batch[-1]) + 1
        if self.graph_pooling == "max":
            idx = batch.view(-1, 1).expand_as(h)
            return h.new_zeros(nb, h.shape[1]).scatter_reduce(0, idx, h, "amax", include_self=False)
        out = h.new_zeros(nb, h.shape[1]).index_add_(0, batch, h)
        if self.graph_pooling == "mean":
            cnt = torch.bincount(batch, minlength=nb).clamp(min=1).to(h.dtype)
            out = out / cnt[:, None]
        return out

    def forward(self, input_batch):
        x, edge_index = input_batch.x, input_batch.edge_index
        edge_attr = input_batch.edge_attr.to(torch.long) if self.global_edge == "onehot" else input_batch.edge_attr
        batch, age = input_batch.batch, input_batch.age
        h = edge_emb = None
        if not self.pca_only:
            if self.node_embedding:
                x = torch.cat([x[:, :-1], self.node_embedding_encoder(x[:, -1].to(torch.long))], dim=-1)
            h = self.node_features_encoder(x)
            if self.use_edge_attr:
                if (self.AFFINE_EDGE and isinstance(self.edge_encoder, nn.Linear) and self.edge_encoder.in_features == 1
                        and edge_attr.is_cuda and edge_attr.numel() == edge_index.shape[1]
                        and self.edge_encoder.out_features % 4 == 0 and not self.gcns[0].bond_encoder):
                    # scalar edge attribute: Linear(1 -> H) is a_e * w + b; carried in factored form through every GENConv
                    edge_emb = AffineEdge(edge_attr.reshape(-1), self.edge_encoder.weight[:, 0], self.edge_encoder.bias)
                else:
                    edge_emb = self.edge_encoder(edge_attr)
        rows = None
        if self.pathway_global_node:
            pemb = self.pathway_features_encoder(input_batch.pathway_node_attr)
            if not self.pca_only:
                rows = self._pathway_rows(input_batch.node_size, h.device)
                h = h.index_copy(0, rows, pemb)
        if self.pca_only:
            raise NotImplementedError("pca_only (GNN bypass) is not selected by any shipped config")

        L, drop = self.num_layers, (lambda v: v if self.no_inter_drop else F.dropout(v, p=self.dropout, training=self.training))
        if self.block == 'res+':
            h = self.gcns[0](h, edge_index, edge_emb)
            for l in range(1, L):
                norm = self.norms[l - 1]
                if self.no_inter_norm:
                    a = F.relu(h)
                elif hasattr(norm, "forward_relu"):
                    a = norm.forward_relu(h)          # LayerNorm + ReLU in one pass each way
                else:
                    a = F.relu(norm(h))
                h = self.gcns[l](drop(a), edge_index, edge_emb) + h
            h = drop(self.norms[L - 1](h))
        elif self.block == 'res':
            h = F.dropout(F.relu(self.norms[0](self.gcns[0](h, edge_index, edge_emb))), p=self.dropout, training=self.training)
            for l in range(1, L):
                h = F.relu(self.norms[l](self.gcns[l](h, edge_index, edge_emb))) + h
                h = F.dropout(h, p=self.dropout, training=self.training)
        else:  # plain
            h = F.dropout(F.relu(self.norms[0](self.gcns[0](h, edge_index, edge_emb))), p=self.dropout, training=self.training)
            for l in range(1, L):
                h2 = self.gcns[l](h, edge_index, edge_emb)
                if not self.no_inter_norm:
                    h2 = self.norms[l](h2)
                h = drop(F.relu(h2) if l != L - 1 else h2)

        if self.pathway_global_node:
            pr = h.index_select(0, rows).view(-1, self.pathway_num, h.shape[1])        # [B, P, H]
            if self.pathway_readout is None:
                h_graph = self._pool(pr.reshape(-1, h.shape[1]), batch.index_select(0, rows))
            else:
                if self.feature_drop_flag:
                    pr = self.feature_drop(pr)
                h_graph = torch.flatten(F.max_pool1d(pr.transpose(1, 2), 4), start_dim=1)
                if self.pre_concat_age:
                    h_graph = torch.cat([h_graph, age[:, None]], dim=-1)
                h_graph = self.readout_func(h_graph)
        else:
            h_graph = self._pool(h, batch)
        if self.use_age and not self.pre_concat_age:
            h_graph = torch.cat([h_graph, age[:, None]], dim=-1)
        return F.softmax(self.graph_pred_linear(h_graph), dim=-1)

    def print_params(self, epoch=None, final=False):
        for flag, name, get in ((self.learn_t, 't', lambda g: g.t.item()), (self.learn_p, 'p', lambda g: g.p.item()),
                                (self.msg_norm, 's', lambda g: g.msg_norm.msg_scale.item())):
            if flag:
                vals = [get(g) for g in self.gcns]
                if final:
                    print('Final {} {}'.format(name, vals))
                else:
                    logging.info('Epoch {}, {} {}'.format(epoch, name, vals))

    def init_weight(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight.data)
                nn.init.constant_(m.bias.data, 0.0)
