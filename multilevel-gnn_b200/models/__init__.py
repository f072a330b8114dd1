"""Model registry with the reference's names (models/__init__.py:11-23) for the in-scope families."""
from .multilevel_gnn import MultilevelGNN
from .deepergcn import DeeperGCN
from .diff_pooling import DiffPool, DiffPoolLayer, SAGEConvolutions, DenseSAGEConv, dense_diff_pool
from . import decoder
from .decoder import GroupedDecoder
from .vae import VAE

MODELS = {'deepergcn': DeeperGCN, 'multilevel_gnn': MultilevelGNN, 'vae': VAE}


def get_model(model_name):
    return MODELS[model_name]
