"""Backbone constructor kwargs of the reference's shipped configs (as constants).

Sources (reference tree, projects/configs/):
  ToC3D/ToC3D_fast.py:41-69, ToC3D/ToC3D_faster.py:64,
  ToC3D_1600_resolution/ToC3D_{fast,faster}_1600.py (img_size stays 320),
  StreamPETR/stream_petr_eva_vit_l{,_1600}.py:41-57.
"""

PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]

_VIT_L = dict(
    img_size=320, patch_size=16, window_size=16, global_window_size=20, in_chans=3,
    embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4 * 2 / 3,
    global_attn_indexes=(2, 5, 8, 11, 14, 17, 20, 23), qkv_bias=True, drop_path_rate=0.3,
    use_act_checkpoint=True, xattn=False, use_checkpoint=False,
)

_TOC3D = dict(
    rope=True, rope_acc=True, pc_range=PC_RANGE, pruning_num_queries=64,
    pruning_loc=[6, 12, 18], accelerate_global=True, token_selection_loss=None,
)

CONFIGS = {
    # name: (backbone type, ctor kwargs, (H_img, W_img))
    "toc3d_fast": ("ToC3DEVAViT", dict(_VIT_L, **_TOC3D, token_ratio=[0.7, 0.5, 0.5]), (320, 800)),
    "toc3d_faster": ("ToC3DEVAViT", dict(_VIT_L, **_TOC3D, token_ratio=[0.5, 0.4, 0.3]), (320, 800)),
    "toc3d_fast_1600": ("ToC3DEVAViT", dict(_VIT_L, **_TOC3D, token_ratio=[0.7, 0.5, 0.5]), (800, 1600)),
    "toc3d_faster_1600": ("ToC3DEVAViT", dict(_VIT_L, **_TOC3D, token_ratio=[0.5, 0.4, 0.3]), (800, 1600)),
    "eva_vit_l": ("EVA_ViT", dict(_VIT_L), (320, 800)),
    "eva_vit_l_1600": ("EVA_ViT", dict(_VIT_L), (800, 1600)),
}

# A structurally identical miniature used by CPU tests and the committed golden
# vectors (same block kinds: dense ws/global, accelerated ws/global, 3 stages).
TINY = dict(
    img_size=320, patch_size=16, window_size=16, global_window_size=20, in_chans=3,
    embed_dim=128, depth=8, num_heads=2, mlp_ratio=4 * 2 / 3,
    global_attn_indexes=(1, 3, 5, 7), qkv_bias=True, drop_path_rate=0.0,
    use_act_checkpoint=False, xattn=False, use_checkpoint=False,
    rope=True, rope_acc=True, pc_range=PC_RANGE, pruning_num_queries=64,
    pruning_loc=[2, 4, 6], accelerate_global=True, token_selection_loss=None,
    token_ratio=[0.7, 0.5, 0.3],
)
