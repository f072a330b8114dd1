"""VAE drop-in (models/vae.py:48-330 of the reference): the pre-training variant of MultilevelGNN whose ``predict_head`` is
the ONLY call site of DiffPool in the reference (vae.py:233-265).

Same constructor (``args``, ``pca_params``, ``pathway_indexs``), same state_dict keys (``decoder.{i}.{0,2}.*`` or
``decoder.{0,2,4}.*``, ``diff_pooling.*``, ``enc_mu.*``, ``enc_log_sigma.*`` on top of MultilevelGNN's), same methods:
``encoder``, ``forward`` (dict), ``train_step`` / ``eval_step``, ``predict_head``, ``reconstruct_head``,
``flatten_decoder`` / ``foreach_decoder``, ``set_pathway_similarity_matrix`` / ``get_pathway_adj``.

What runs where: the gene-level GNN stack and the gene -> pathway pool are MultilevelGNN's kernels (``pool_genes``);
``predict_head`` feeds DiffPool (fused small-graph kernel at the reference's 146-pathway size, tensor-core GEMMs at
GEMM sizes) and the fused Linear/softmax head kernel; the per-pathway decoders ("unpool", vae.py:54-74,216-222; SURVEY
section 8 row f4) are one grouped kernel per direction over a packed parameter (models/decoder.py).
"""
import math

import torch
import torch.nn as nn

from .. import functional as Fn
from .decoder import GroupedDecoder
from .diff_pooling import DiffPool
from .multilevel_gnn import MultilevelGNN


def next_pow2(n):
    """Smallest power of two >= n (vae.py:14-29)."""
    return 1 if n <= 1 else 1 << (int(n) - 1).bit_length()


class VAE(MultilevelGNN):
    def __init__(self, args, pca_params=None, pathway_indexs=None):
        super().__init__(args, pca_params, pathway_indexs)
        self.node_num = self.GENES
        self.decoder_dim = args.decoder_dim
        self.decoder_type = args.decoder_type
        feat = args.final_channels * args.pca_dim
        if self.decoder_type == "flatten":
            self.decoder = nn.ModuleList([nn.Linear(args.final_channels * 146 * self.pca_dim * 3, self.decoder_dim), nn.ReLU(),
                                          nn.Linear(self.decoder_dim, self.decoder_dim), nn.ReLU(),
                                          nn.Linear(self.decoder_dim, self.node_num * 3)])
        elif self.decoder_type in ("foreach", "foreach_diffhidden"):
            counts = torch.bincount(pathway_indexs.reshape(-1).long(), minlength=int(pathway_indexs.max()) + 1).tolist()
            hidden = [self.decoder_dim if self.decoder_type == "foreach" else next_pow2(int(math.sqrt(n_out * args.final_channels)))
                      for n_out in counts]
            # one packed parameter + one kernel per direction for all blocks; state_dict keys stay decoder.{i}.{0,2}.*
            self.decoder = GroupedDecoder(feat, hidden, counts)
        if args.reorder_type == "diff_pooling":
            cin = {"pathway": args.final_channels, "head": args.conv_channel_list[-1]}.get(args.diff_pooling_location)
            if cin is not None:
                self.diff_pooling = DiffPool(cin, 2, args.pathway_num, args.diff_pooling_layer, args.diff_pooling_hidden_dim,
                                             args.diff_pooling_output_dim, args)
        self.enc_mu = nn.Linear(feat, feat)
        self.enc_log_sigma = nn.Linear(feat, feat)
        self.init_weight()

    # ------------------------------------------------------------------------------------------ encoder / decoders
    def encoder(self, input_batch):
        """vae.py:128-208: GNN stack + pool (no value mask), then the Gaussian heads on the per-pathway features."""
        x = self.pool_genes(input_batch, value_mask=False)                 # [B, C, 438, P]
        if self.decoder_type == "flatten":
            x = x.reshape(x.shape[0], x.shape[1], 146, self.pca_dim * 3)
        x = x.permute(0, 2, 1, 3).flatten(2)                               # [B, 438, C*P]  (flatten: [B, 146, C*3P])
        mu = self.enc_mu(x)
        sigma = torch.exp(self.enc_log_sigma(x))
        loss_std = -mu.flatten(1).permute(1, 0).std(1).mean()
        # mean |corr| between the feature columns of every pathway over the batch, off-diagonal (vae.py:206-207), batched
        m = mu.permute(1, 2, 0)                                            # [S, F, B]
        mc = m - m.mean(dim=2, keepdim=True)
        cov = mc @ mc.transpose(1, 2)
        d = cov.diagonal(dim1=1, dim2=2).clamp_min(0).sqrt()
        corr = (cov / (d[:, :, None] * d[:, None, :])).clamp(-1, 1)
        eye = torch.eye(corr.shape[-1], device=corr.device, dtype=corr.dtype)
        loss_corr = (corr * (1 - eye)).abs().mean()
        q_z = torch.distributions.Normal(loc=mu, scale=sigma + 1e-7)
        return q_z, torch.cat([mu, sigma], dim=-1), [loss_std, 0, loss_corr], None

    def flatten_decoder(self, h):
        x = h.flatten(1)
        for layer in self.decoder:
            x = layer(x)
        return x

    def foreach_decoder(self, h):
        """pred[:, genes of pathway i] = decoder[i](h[:, i, :]) for all pathways (vae.py:216-222), concatenated: one grouped
        kernel (models/decoder.py)."""
        return self.decoder(h)

    def forward(self, input_batch, x=None, gene_pca_match=None, raw_indice=None, age=None):
        q_z, h, loss, _ = self.encoder(input_batch)
        z = q_z.rsample()
        output = self.flatten_decoder(z) if self.decoder_type == "flatten" else self.foreach_decoder(z)
        return {"pred_x": output, "embedding": h, "q_z": q_z, "z": z, "loss": loss}

    # ------------------------------------------------------------------------------------------ classification head
    def _latent_to_image(self, h):
        b, _, c = h.shape
        if getattr(self.args, "channel_one", False):
            return h[:, :, :c // 2].reshape(b, 1, 146, -1)
        return h[:, :, :c // 2].permute(0, 2, 1).reshape(b, c // 2, 146, 3)

    def train_step(self, input_batch, require_grad=True):
        with torch.enable_grad() if require_grad else torch.no_grad():
            q_z, h, loss, gene_feature = self.encoder(input_batch)
            if getattr(self.args, "vae_generate_train_sample", False):
                h = q_z.rsample()
            h = self._latent_to_image(h)
        if self.args.reorder_pathway and self.reorder_idxs is not None:
            h = h[:, :, self.reorder_idxs, :]
        pred, pca_feature, l, e = self.predict_head(h, input_batch.age)
        return pred, pca_feature, l, e, gene_feature

    def eval_step(self, input_batch, require_grad=True):
        with torch.enable_grad() if require_grad else torch.no_grad():
            _, h, _, _ = self.encoder(input_batch)
            h = self._latent_to_image(h)
        if self.args.reorder_pathway and self.reorder_idxs is not None:
            h = h[:, :, self.reorder_idxs, :]
        return self.predict_head(h, input_batch.age)

    def predict_head(self, x, age):
        """vae.py:233-265.  x [B, C, 146, d] -> (pred [B, 2], pca_feature, link loss, entropy loss).  With
        reorder_type == 'diff_pooling' every (sample, feature column) pair is one 146-node graph over the shared pathway
        similarity matrix: [B, C, 146, d] -> [B*d, 146, C] -> DiffPool -> [B*d, 10, 64] -> [B, -1]."""
        args = self.args
        l = e = 0
        pca_feature = x
        if self.pca_prelinear:
            x = self.pre_linear(x)
        diff = args.reorder_type == "diff_pooling"
        if diff and args.diff_pooling_location == "pathway":
            b = x.shape[0]
            x = x.permute(0, 3, 2, 1).reshape(-1, args.pathway_num, args.final_channels)
            x, l, e = self.diff_pooling(x, self.get_pathway_adj().to(x.device))
            x = self.drop1(x.reshape(b, -1))
        else:
            for layer in self.conv_model:
                x = self._conv(layer, x) if x.is_cuda else layer(x)
            if diff and args.diff_pooling_location == "head":
                b = x.shape[0]
                x = x.permute(0, 3, 2, 1).reshape(-1, args.pathway_num, args.conv_channel_list[-1])
                x, l, e = self.diff_pooling(x, self.get_pathway_adj().to(x.device))
                x = self.drop1(x.reshape(b, -1))
            else:
                if args.reorder_type != "no_pooling":
                    x = self.pooling(x.contiguous())
                x = torch.flatten(self.drop1(x), start_dim=1)
        if args.use_age:
            x = torch.cat([x, age[:, None]], dim=-1)
        hd = self.head
        if Fn.HeadMLP.supported(x, hd[0], hd[3]) and not any(m._forward_hooks for m in hd):
            pred, _ = Fn.HeadMLP.apply(x, hd[0].weight, hd[0].bias, hd[3].weight, hd[3].bias, hd[2].p, self.training, None, None)
        else:
            pred = hd(x)
        return pred, pca_feature, l, e

    def reconstruct_head(self, args):
        """Rebuild the classifier for the chosen reorder_type (vae.py:267-300)."""
        age = 1 if self.args.use_age else 0
        if args.reorder_type == "no_pooling":
            d_in = args.conv_channel_list[-1] * 146 * (3 * self.pca_dim) + age
        elif args.reorder_type == "diff_pooling":
            n = self.args.pathway_num
            for _ in range(self.args.diff_pooling_layer):
                n = math.ceil(n * 0.25)
            d_in = self.args.diff_pooling_output_dim * n * (3 * self.pca_dim) + age
        else:
            d_in = args.conv_channel_list[-1] * (146 // self.pathway_pool_dim) * ((3 * self.pca_dim) // self.pca_pool_dim) + age
        self.head = nn.Sequential(nn.Linear(d_in, self.head_dim), nn.ReLU(), nn.Dropout(0.5), nn.Linear(self.head_dim, 2),
                                  nn.Softmax(dim=1))
        for m in self.head.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                nn.init.xavier_uniform_(m.weight.data)

    def get_pathway_adj(self):
        if self.args.pathway_similarity == "correlation":
            return self.pathway_similarity_matrix
        raise NotImplementedError("pathway_similarity=%r" % self.args.pathway_similarity)

    def set_pathway_similarity_matrix(self, pathway_similarity_matrix):
        self.pathway_similarity_matrix = (torch.as_tensor(pathway_similarity_matrix)
                                          + torch.eye(self.args.pathway_num)).to(torch.float32)
