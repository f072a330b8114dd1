"""Basic layers with the reference's names and Sequential indices
(models/gcn_lib/sparse/torch_nn.py:9-75), so state_dict keys such as ``gconv.nn.0.weight`` and
``feature_encoder.{0,1,3}.*`` are unchanged.  These stay torch modules (cuBLAS / ATen): SURVEY.md
section 2 row 4 keeps them as "next" epilogue-fusion candidates."""
import torch
from torch import nn

_ACTS = {
    "relu": lambda neg, n: nn.ReLU(False),
    "leakyrelu": lambda neg, n: nn.LeakyReLU(neg, False),
    "prelu": lambda neg, n: nn.PReLU(num_parameters=n, init=neg),
    "elu": lambda neg, n: nn.ELU(),
    "tanh": lambda neg, n: nn.Tanh(),
}


def act_layer(act_type, inplace=False, neg_slope=0.2, n_prelu=1):
    key = act_type.lower()
    if key not in _ACTS:
        raise NotImplementedError('activation layer [%s] is not found' % key)
    return _ACTS[key](neg_slope, n_prelu)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm (same parameters / state_dict keys); tall fp32 CUDA inputs go through mlg_layernorm_fwd / _bwd
    (functional.LayerNormFn), anything else through the library."""

    def forward(self, x):
        from ... import functional as Fn
        if Fn.LayerNormFn.supported(x, self.normalized_shape):
            return Fn.LayerNormFn.apply(x, self.weight, self.bias, self.eps)
        return super().forward(x)

    def forward_relu(self, x):
        """relu(self(x)) -- one pass each way on the kernel path (the ReLU that follows the norm in the res+ block and in
        the MLP stages)."""
        from ... import functional as Fn
        if Fn.LayerNormFn.supported(x, self.normalized_shape) and not self._forward_hooks:
            return Fn.LayerNormFn.apply(x, self.weight, self.bias, self.eps, True)
        return torch.relu(self(x))


def norm_layer(norm_type, nc):
    key = norm_type.lower()
    if key == 'batch':
        return nn.BatchNorm1d(nc, affine=True)
    if key == 'layer':
        return LayerNorm(nc, elementwise_affine=True)
    if key == 'instance':
        return nn.InstanceNorm1d(nc, affine=False)
    raise NotImplementedError('normalization layer [%s] is not found' % key)


class MLP(nn.Sequential):
    """Lin -> [norm] -> [act] -> [Dropout2d] per stage; ``last_lin`` leaves the final Linear bare."""

    def __init__(self, channels, act='relu', norm=None, bias=True, drop=0., last_lin=False):
        layers = []
        n_stage = len(channels) - 1
        for s in range(n_stage):
            layers.append(nn.Linear(channels[s], channels[s + 1], bias))
            if last_lin and s == n_stage - 1:
                continue
            if isinstance(norm, str) and norm.lower() != 'none':
                layers.append(norm_layer(norm, channels[s + 1]))
            if act is not None and act.lower() != 'none':
                layers.append(act_layer(act))
            if drop > 0:
                layers.append(nn.Dropout2d(drop))
        self.m = layers
        super().__init__(*layers)

    def forward(self, x):
        """Same chain as nn.Sequential; the Linears of a TALL CUDA input (node / edge rows) go through
        functional.tall_linear: 3xTF32 tensor-core forward and dX where the shape allows, tensor-core / fp32-FMA weight and
        bias gradient instead of the library's split-K SIMT GEMM + separate bias reductions."""
        from ... import functional as Fn
        mods = list(self)
        i = 0
        while i < len(mods):
            mod = mods[i]
            if isinstance(mod, nn.Linear) and torch.is_tensor(x) and x.is_cuda and x.dim() == 2:
                x = Fn.tall_linear(x, mod)
            elif (isinstance(mod, LayerNorm) and i + 1 < len(mods) and type(mods[i + 1]) is nn.ReLU
                  and not mods[i + 1]._forward_hooks and torch.is_tensor(x) and x.is_cuda):
                x = mod.forward_relu(x)      # norm + act in one pass
                i += 1
            else:
                x = mod(x)
            i += 1
        return x
