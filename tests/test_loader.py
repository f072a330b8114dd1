"""The collate / loader replacement (multilevel-gnn_b200/data.py) against the layout PyG's Batch collate produces for the
reference's loader (train.py:316-327): same fields, same node offsets, plus the topology key."""
import types

import torch

from multilevel_gnn_b200 import data, synth


def _fold(genes=40, slots=600, edges=500, seed=0):
    ei, ea = synth.omics_topology(genes, edges, seed)
    match, seg = synth.pool_layout(genes, slots, synth.SEGMENTS, seed)
    return data.FoldTopology(ei, ea, match, seg, 3 * genes), (ei, ea, match, seg)


def _patients(n, nodes, seed=1):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        lab = int(torch.rand((), generator=g) < 0.5)
        out.append(types.SimpleNamespace(x=torch.randn(nodes, 1, generator=g), age=float(torch.rand((), generator=g)),
                                         y=torch.nn.functional.one_hot(torch.tensor(lab), 2).float()))
    return out


def test_collate_matches_pyg_batch_layout():
    topo, (ei, ea, match, seg) = _fold()
    pats = _patients(5, topo.n_nodes)
    b = data.collate(pats[:3], topo, pin=False)
    n = topo.n_nodes
    # PyG Batch: x concatenated on dim 0, edge_index concatenated on dim -1 with cumulative node offsets, python floats ->
    # [B], y [2] -> [2B], 2-D per-sample tensors [1, G] -> [B, G]
    assert torch.equal(b.x, torch.cat([p.x for p in pats[:3]], 0))
    assert torch.equal(b.edge_index, torch.cat([ei + i * n for i in range(3)], dim=1))
    assert torch.equal(b.edge_attr, torch.cat([ea] * 3, 0))
    assert torch.equal(b.gene_pca_match, match.unsqueeze(0).repeat(3, 1))
    assert torch.equal(b.raw_indice, seg.unsqueeze(0).repeat(3, 1))
    assert torch.equal(b.batch, torch.arange(3).repeat_interleave(n))
    assert torch.allclose(b.age, torch.tensor([p.age for p in pats[:3]]))
    assert torch.equal(b.y, torch.cat([p.y for p in pats[:3]]))
    assert b.topology_key.startswith("topo-") and b.topology_key.endswith("-b3")
    # the replicated topology is shared between batches of one size (no per-step allocation), and keyed by content
    b2 = data.collate(pats[2:5], topo, pin=False)
    assert b2.edge_index.data_ptr() == b.edge_index.data_ptr() and b2.topology_key == b.topology_key
    other, _ = _fold(seed=3)
    assert other.key != topo.key
    assert data.collate(pats[:2], topo, pin=False).topology_key != b.topology_key       # another batch size: another layout


def test_loader_epochs_and_ranks():
    topo, _ = _fold()
    pats = _patients(23, topo.n_nodes)
    seen = []
    for rank in range(2):
        ld = data.TopologyLoader(pats, topo, batch_size=4, rank=rank, world=2, seed=7, pin=False)
        assert len(ld) == 23 // 8
        for b in ld:
            assert b.x.shape == (4 * topo.n_nodes, 1)
            seen.append(b.x.view(4, -1)[:, 0])
    firsts = torch.cat(seen)
    assert firsts.unique().numel() == firsts.numel()            # ranks see disjoint patients
    ld = data.TopologyLoader(pats, topo, batch_size=4, seed=7, pin=False)
    e0 = [b.x.clone() for b in ld]
    ld.set_epoch(1)
    e1 = [b.x.clone() for b in ld]
    assert len(e0) == len(e1) == 5 and not all(torch.equal(a, c) for a, c in zip(e0, e1))
    ordered = data.TopologyLoader(pats, topo, batch_size=5, shuffle=False, drop_last=False, pin=False)
    assert [b.age.shape[0] for b in ordered] == [5, 5, 5, 5, 3]
