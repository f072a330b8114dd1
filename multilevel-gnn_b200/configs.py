"""Argument namespaces for the drop-in model constructors.

The reference's models read a flat ``args`` namespace built by ``opt.py`` (argparse defaults,
then YAML keys ``setattr``'d over them, /root/reference/opt.py:437-444).  Only the keys the
hot-path constructors/forwards read are kept here (``models/multilevel_gnn.py``,
``models/deepergcn.py``, ``models/diff_pooling.py``, ``train.py:112-125``); values are the
reference's defaults, and ``OVERLAYS`` holds what ``config/{gbm,kirc,lgg}.yaml`` change among
those keys.  ``tests/test_configs.py`` checks both against the reference tree when it is present.
"""
import argparse
import copy

DEFAULTS = dict(
    # --- GNN stack / MultilevelGNN (models/multilevel_gnn.py:16-130) ---
    gnn_name="gat", gnn_act="leakyrelu", gnn_mlp_norm="none", gnn_dropout=0.0, gnn_last_norm=False,
    num_layers=3, hidden_channels=128, final_channels=1, final_head=1,
    node_embedding=False, node_embedding_dim=32, embedding_init_type="xavier", emb_val=0.01,
    freeze_node_embedding=False, freeze_pca_weight=False, input_drop=None, input_emb_drop=None,
    dense_gnn=False, resgnn=False, repeat_mask=False, repeat_cyclic=2, repeat_norm=False,
    value_att_mask=False, merge_mode="mult", add_coef1=0.5, add_coef2=0.5,
    weighted_edge=False, edge_type="grnboost2", device_num=1, device=0,
    pca_compare=False, pca_prelinear=False, learnable_pca=False, pca_loss=False,
    pca_indep_loss=False, pca_loss_coef=1, pca_dim=2, pca_init_type=None, pca_match_mask=False,
    pathway_pool_dim=4, pca_pool_dim=2, mutual_info_mask=False, mutual_info_threshold=None,
    node_select_threshold=1, mutual_neighbors=3, head_dim=64, used_omics="012", use_age=False,
    feature_drop=False, conv_channel_list=[32, 64], conv_kernel_list=[1, 1],
    reduction_method="linear_projection", pca_lowrank_niter=2, reorder_pathway=False,
    pathway_num=146, freeze_mutual_select_init=False, random_state=1, remain_all_tf=False,
    # --- DeeperGCN (models/deepergcn.py:18-181) ---
    block="res+", conv="gen", gcn_aggr="max", t=1.0, learn_t=False, p=1.0, learn_p=False,
    msg_norm=False, learn_msg_scale=False, conv_encode_edge=False, norm="layer", mlp_layers=2,
    graph_pooling="mean", num_tasks=2, dropout=0.5, mul_attr=False, node_num=5606,
    gnn_encoder="linear", pca_only=False, no_inter_drop=False, no_inter_norm=False,
    head_init=False, all_init=True, init_emb=False, global_edge="onehot", use_column=None,
    use_edge_attr=False, pathway_global_node=False, num_layer_head=1, pathway_readout="maxpool",
    pre_concat_age=False, pre_readout_drop=False, head_dropout=False,
    # --- DiffPool (models/diff_pooling.py:70-114, opt.py:415-420) ---
    after_pooling_layer=1, pooling_type="correlation", diff_pooling_location="pathway",
    diff_pooling_layer=2, diff_pooling_hidden_dim=32, diff_pooling_output_dim=64,
    # --- VAE.predict_head / decoders (models/vae.py:48-87,233-300; opt.py:277-278,371-372,382-383) ---
    reorder_type="pca", pathway_similarity="correlation", decoder_dim=4096, decoder_type="flatten",
    channel_one=False, vae_generate_train_sample=False,
    # --- optimiser / loss (train.py:112-125) ---
    lr=1e-4, beta1=0.9, beta2=0.999, wd=0.0, weight_balance=False, weighted_loss=False,
    batch_weighted_loss=False, clip_grad=False, batch_size=4,
)

_COMMON = dict(
    dropout=0.25, conv_encode_edge=True, feature_drop=True, final_channels=32, final_head=4,
    freeze_mutual_select_init=True, gnn_name="sage", hidden_channels=64, init_emb=True,
    learnable_pca=True, mutual_info_mask=True, node_embedding=True, num_layer_head=2, num_layers=2,
    pathway_global_node=True, pca_indep_loss=True, pca_match_mask=True, pre_readout_drop=True,
    random_state=12345, use_column="stringdb::score", use_edge_attr=True, value_att_mask=True,
    weight_balance=True, weighted_edge=True,
)

OVERLAYS = {
    # config/gbm.yaml
    "gbm": dict(_COMMON, batch_size=32, head_dim=256, mutual_neighbors=7, node_embedding_dim=64,
                pre_concat_age=True, use_age=True),
    # config/kirc.yaml
    "kirc": dict(_COMMON, batch_size=64, head_dim=512, mutual_neighbors=15, pca_dim=3, lr=5e-5,
                 reorder_pathway=True, pathway_pool_dim=1, pca_pool_dim=1),
    # config/lgg.yaml
    "lgg": dict(_COMMON, batch_size=64, head_dim=512, mutual_neighbors=15, pca_dim=3, lr=5e-5,
                reorder_pathway=True, pre_concat_age=True, use_age=True, device=1),
}


def make_args(config=None, **overrides):
    """Namespace with the hot-path keys; ``config`` in {None, 'gbm', 'kirc', 'lgg'}."""
    d = copy.deepcopy(DEFAULTS)
    if config is not None:
        d.update(copy.deepcopy(OVERLAYS[config]))
    d.update(overrides)
    return argparse.Namespace(**d)


def deepergcn_args(hidden=128, layers=28, **overrides):
    """28-layer GENConv softmax DeeperGCN, res+, LayerNorm, msg_norm, learn_t,
    per-layer edge encoder, continuous edge_attr [E,1] (global_edge=None)."""
    base = dict(hidden_channels=hidden, num_layers=layers, gcn_aggr="softmax", block="res+",
                conv="gen", norm="layer", mlp_layers=2, msg_norm=True, learn_msg_scale=True,
                learn_t=True, conv_encode_edge=True, use_edge_attr=True, global_edge=None,
                use_column="stringdb::score", pathway_global_node=True, pathway_readout="maxpool",
                pre_readout_drop=True, num_layer_head=2, use_age=True, pre_concat_age=True,
                dropout=0.0, node_embedding=False)
    base.update(overrides)
    return make_args(None, **base)
