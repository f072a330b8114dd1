"""Seeded synthetic inputs with the exact batch-field layout the reference's loader emits.

There is no access to the private TCGA drive the reference trains on (README.md:2), so every
benchmark / parity input is generated here, mirroring ``dataloader/multiloader.py``:
gene g owns nodes 3g (mRNA), 3g+1 (CNV), 3g+2 (methylation) (:613,647); with ``mute_edge "12"``
only mRNA-mRNA PPI edges survive, plus one CNV->mRNA (attr +1) and one MT->mRNA (attr -1) edge per
gene (:660-671); ALL patients share one edge list (:687-698); PyG ``Batch`` collate concatenates
graphs with cumulative node offsets.  Shapes: SURVEY.md section 8(d).
"""
import torch

GENES = 5135          # models/multilevel_gnn.py:34 (node_num); N = 3 * GENES nodes per graph
SLOTS = 25015         # models/multilevel_gnn.py:74 (gene slots G)
SEGMENTS = 146 * 3    # pathways x omics


class GraphBatch:
    """Attribute bag standing in for torch_geometric.data.Batch (only ``.to`` is needed)."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def to(self, device, non_blocking=False):
        out = GraphBatch()
        for k, v in vars(self).items():
            setattr(out, k, v.to(device, non_blocking=non_blocking) if torch.is_tensor(v) else v)
        return out

    def pin_memory(self):
        out = GraphBatch()
        for k, v in vars(self).items():
            setattr(out, k, v.pin_memory() if torch.is_tensor(v) else v)
        return out

    def nbytes(self):
        return sum(v.numel() * v.element_size() for v in vars(self).values() if torch.is_tensor(v))


def _gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def omics_topology(genes=GENES, intra_edges=82160, seed=0):
    """One patient graph: edge_index [2,E] int64 (row0 = source, row1 = target), edge_attr [E,1]."""
    g = _gen(seed)
    # power-law-ish target popularity so in-degrees are ragged like a PPI / GRN graph
    pop = torch.rand(genes, generator=g).pow(3.0) + 0.02
    dst = torch.multinomial(pop, intra_edges, replacement=True, generator=g)
    src = torch.randint(0, genes, (intra_edges,), generator=g)
    src = torch.where(src == dst, (src + 1) % genes, src)
    w = torch.rand(intra_edges, generator=g) * (1 - 1e-5) + 1e-5
    gi = torch.arange(genes)
    ei = torch.cat([torch.stack([3 * src, 3 * dst]),
                    torch.stack([3 * gi + 1, 3 * gi]),
                    torch.stack([3 * gi + 2, 3 * gi])], dim=1)
    ea = torch.cat([w, torch.ones(genes), -torch.ones(genes)]).unsqueeze(1)
    perm = torch.randperm(ei.shape[1], generator=g)   # loader order is not sorted by target
    return ei[:, perm].contiguous(), ea[perm].contiguous()


def pool_layout(genes=GENES, slots=SLOTS, segments=SEGMENTS, seed=0, missing=0.02):
    """gene_pca_match [slots] (node id or -1) and raw_indice [slots] (sorted segment ids)."""
    g = _gen(seed + 1)
    match = torch.randint(0, 3 * genes, (slots,), generator=g)
    match[torch.rand(slots, generator=g) < missing] = -1
    sizes = torch.multinomial(torch.ones(segments), slots - segments, replacement=True, generator=g)
    sizes = torch.bincount(sizes, minlength=segments) + 1          # every segment non-empty, mean ~57
    seg = torch.repeat_interleave(torch.arange(segments), sizes)
    return match, seg


def multilevel_batch(batch_size=32, genes=GENES, slots=SLOTS, intra_edges=82160, seed=0,
                     with_labels=True):
    """A gbm/kirc/lgg-shaped batch: fields x, edge_index, edge_attr,
    gene_pca_match, raw_indice, age, y, batch."""
    n = 3 * genes
    ei, ea = omics_topology(genes, intra_edges, seed)
    match, seg = pool_layout(genes, slots, SEGMENTS, seed)
    g = _gen(seed + 2)
    off = (torch.arange(batch_size) * n).view(-1, 1, 1)
    b = GraphBatch(
        x=torch.randn(batch_size * n, 1, generator=g),
        edge_index=(ei.unsqueeze(0) + off).permute(1, 0, 2).reshape(2, -1).contiguous(),
        edge_attr=ea.repeat(batch_size, 1),
        gene_pca_match=match.unsqueeze(0).repeat(batch_size, 1),
        raw_indice=seg.unsqueeze(0).repeat(batch_size, 1),
        age=torch.rand(batch_size, generator=g),
        batch=torch.arange(batch_size).repeat_interleave(n),
    )
    if with_labels:
        lab = (torch.rand(batch_size, generator=g) < 0.5).long()
        b.y = torch.nn.functional.one_hot(lab, 2).float().reshape(-1)   # PyG collate: [2] -> [2B]
    return b


def multilevel_params(model, slots=SLOTS, seed=0):
    """Install synthetic info_mask / projection weights / pathway_indexs (train.py:292-298 would
    derive them from sklearn MI + per-pathway PCA on the private data)."""
    g = _gen(seed + 3)
    _, seg = pool_layout(slots=slots, seed=seed)
    model.set_pathway_indexs(seg.clone())
    model.set_info_mask((torch.rand(slots, 1, generator=g) < 0.5).float())
    with torch.no_grad():
        model.learnable_pca_params.copy_(torch.randn(model.learnable_pca_params.shape, generator=g) * 0.05)
    return model


def knn_points(n=100000, dim=64, seed=0):
    return torch.randn(n, dim, generator=_gen(seed + 4))


def deepergcn_batch(edge_index, n_nodes, pathway_num=146, seed=0):
    """cfg4: one graph of n_nodes (+pathway rows) with a prebuilt edge_index over the first
    n_nodes rows; x [N+P,3], edge_attr [E,1] U(0,1), pathway_node_attr [P,6], node_size [1]."""
    g = _gen(seed + 5)
    n = n_nodes + pathway_num
    e = edge_index.shape[1]
    return GraphBatch(
        x=torch.randn(n, 3, generator=g),
        edge_index=edge_index,
        edge_attr=torch.rand(e, 1, generator=g),
        batch=torch.zeros(n, dtype=torch.long),
        age=torch.rand(1, generator=g),
        pathway_node_attr=torch.randn(pathway_num, 6, generator=g),
        node_size=torch.tensor([n]),
        y=torch.tensor([1.0, 0.0]),
    )


def diffpool_inputs(batch=576, nodes=146, channels=32, seed=0):
    """cfg3: x [b,146,C] N(0,1); adj = symmetric U(0,1) + I (vae.py:301-306)."""
    g = _gen(seed + 6)
    x = torch.randn(batch, nodes, channels, generator=g)
    a = torch.rand(nodes, nodes, generator=g)
    adj = (a + a.t()) * 0.5 + torch.eye(nodes)
    return x, adj
