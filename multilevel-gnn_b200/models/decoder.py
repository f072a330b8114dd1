"""Per-pathway decoders of the VAE ("unpool") as one grouped kernel.

Reference: models/vae.py:54-74 builds ``nn.ModuleList([nn.Sequential(Linear(F, D_i), ReLU(), Linear(D_i, n_i)) ...])`` --
one block per pathway, D_i = decoder_dim ('foreach') or next_pow2(sqrt(n_i * final_channels)) ('foreach_diffhidden') -- and
``foreach_decoder`` (:216-222) runs them one by one on ``h[:, i, :]`` and concatenates: ~1750 launches forward.

``GroupedDecoder`` keeps every block's parameters in ONE packed ``nn.Parameter`` (W1_i, b1_i, W2_i, b2_i back to back, each
array starting on a 16-byte boundary) and runs all blocks in one launch per direction (csrc/decoder_grouped.cu: a CTA per
pathway).  The state_dict keeps the reference's keys -- ``{i}.0.weight``, ``{i}.0.bias``, ``{i}.2.weight``, ``{i}.2.bias``
under the module's prefix (``decoder.`` in VAE) -- through a pair of state_dict hooks, so checkpoints are interchangeable;
``named_parameters()`` shows the single packed tensor (what the optimizer and the gradient bucket should see).
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fn

KEYS = ("0.weight", "0.bias", "2.weight", "2.bias")


def _align4(v):
    return (v + 3) & ~3


class GroupedDecoder(nn.Module):
    def __init__(self, feat, hidden, outs):
        """feat: F = final_channels * pca_dim; hidden / outs: per-pathway D_i and n_i (same length)."""
        super().__init__()
        assert len(hidden) == len(outs) and len(outs) > 0
        self.feat, self.hidden, self.outs = int(feat), [int(d) for d in hidden], [int(n) for n in outs]
        rows, off, out_off, h_off = [], 0, 0, 0
        for d, n in zip(self.hidden, self.outs):
            w1 = off
            b1 = w1 + _align4(d * self.feat)
            w2 = b1 + _align4(d)
            b2 = w2 + _align4(n * d)
            off = b2 + _align4(n)
            rows.append([w1, b1, w2, b2, d, n, out_off, h_off])
            out_off += n
            h_off += d
        self.total_out, self.total_hidden, self.d_max = out_off, h_off, max(self.hidden)
        self.register_buffer("table", torch.tensor(rows, dtype=torch.int64), persistent=False)
        self._rows = rows
        self.packed = nn.Parameter(torch.zeros(off))
        self.reset_parameters()
        self.register_state_dict_post_hook(GroupedDecoder._unpack_into_state_dict)
        self.register_load_state_dict_pre_hook(GroupedDecoder._pack_from_state_dict)

    def __len__(self):
        return len(self.outs)

    def shapes(self, i):
        d, n = self.hidden[i], self.outs[i]
        return ((d, self.feat), (d,), (n, d), (n,))

    def block(self, i, flat=None):
        """(W1, b1, W2, b2) of pathway i as views of ``flat`` (default: the packed parameter's data)."""
        flat = self.packed.data if flat is None else flat
        return tuple(flat[o:o + math.prod(s)].view(s) for o, s in zip(self._rows[i][:4], self.shapes(i)))

    def reset_parameters(self):
        """xavier_uniform_ weights (MultilevelGNN.init_weight, multilevel_gnn.py:294-299) and nn.Linear's default bias."""
        with torch.no_grad():
            for i in range(len(self)):
                w1, b1, w2, b2 = self.block(i)
                nn.init.xavier_uniform_(w1)
                nn.init.xavier_uniform_(w2)
                b1.uniform_(-1 / math.sqrt(self.feat), 1 / math.sqrt(self.feat))
                b2.uniform_(-1 / math.sqrt(self.hidden[i]), 1 / math.sqrt(self.hidden[i]))

    # --- state_dict: the reference's per-block keys -------------------------------------------------------------------
    @staticmethod
    def _unpack_into_state_dict(module, state_dict, prefix, local_metadata):
        flat = state_dict.pop(prefix + "packed")
        for i in range(len(module)):
            for key, view in zip(KEYS, module.block(i, flat)):
                state_dict["%s%d.%s" % (prefix, i, key)] = view

    @staticmethod
    def _pack_from_state_dict(module, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        if prefix + "packed" in state_dict or (prefix + "0.0.weight") not in state_dict:
            return
        flat = module.packed.detach().clone()
        for i in range(len(module)):
            for key, view in zip(KEYS, module.block(i, flat)):
                name = "%s%d.%s" % (prefix, i, key)
                if name not in state_dict:
                    missing_keys.append(name)
                    continue
                src = state_dict.pop(name)
                if tuple(src.shape) != tuple(view.shape):
                    error_msgs.append("size mismatch for %s: %s vs %s" % (name, tuple(src.shape), tuple(view.shape)))
                    continue
                view.copy_(src)
        state_dict[prefix + "packed"] = flat

    def forward(self, h):
        """h [B, S, F] -> pred [B, sum_i n_i] (vae.py:216-222)."""
        if h.shape[1] != len(self) or h.shape[2] != self.feat:
            raise ValueError("GroupedDecoder: expected [B, %d, %d], got %s" % (len(self), self.feat, tuple(h.shape)))
        return Fn.DecoderGrouped.apply(h, self.packed, self.table, self.d_max, self.total_out, self.total_hidden)
