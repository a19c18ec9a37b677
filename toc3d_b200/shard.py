"""Multi-GPU plumbing of the backbone path: one process per GPU, images sharded, one all-gather.

Every op of the path is independent per image except the 64 history queries shared by the views
of one frame (repeat_interleave, toc3d_utils.py:240), so the `(frames x views)` image list is cut
into contiguous per-rank chunks (the reference shards whole samples per rank,
datasets/samplers/distributed_sampler.py:41-44, and never gathers features; north_star asks for the
feature list to be all-gathered for the detection head).  The only collective is one
`all_gather_into_tensor` of `last_feat` (NHWC storage) plus the tiny per-stage masks / index lists.

torch.distributed is plumbing here: NCCL on the GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def partition(n_frames, views, world):
    """Contiguous image chunks [(start, stop)] * world over the n_frames*views images.

    A chunk must not straddle a frame boundary partially, because a launch needs
    `V_local % Bf_local == 0` with every local frame contributing the same number of views
    (toc3d_utils.py:240): either world divides the image count into whole frames, or each
    frame is cut into equal parts (e.g. 6 views over 2 GPUs -> 3 + 3).
    """
    total = n_frames * views
    if world <= 0 or total % world != 0:
        raise ValueError("cannot shard %d images (%d frames x %d views) over %d ranks evenly; use a batch "
                         "with frames*views %% ranks == 0" % (total, n_frames, views, world))
    per = total // world
    if not (per % views == 0 or views % per == 0):
        raise ValueError("a per-rank chunk of %d images would straddle frames of %d views" % (per, views))
    return [(r * per, (r + 1) * per) for r in range(world)]


_PER_FRAME = ("temp_queries", "temp_ref_points", "temp_vel", "temp_timestamp", "temp_ego_pose", "ego_pose_inv")


def local_inputs(inputs, views, rank, world):
    """Slice the keyword inputs of `forward` (x over images, temp_* over frames) for `rank`."""
    x = inputs["x"]
    n_frames = x.shape[0] // views
    a, b = partition(n_frames, views, world)[rank]
    fa, fb = a // views, (b - 1) // views + 1
    out = dict(inputs)
    out["x"] = x[a:b]
    for k in _PER_FRAME:
        if inputs.get(k) is not None:
            out[k] = inputs[k][fa:fb]
    for k in ("gumbel_noise", "teacher_scores"):          # per-image test / parity hooks of the plugin's forward
        if inputs.get(k) is not None:
            out[k] = [t[a:b] for t in inputs[k]]
    return out


def all_gather_rows(t, world, group=None):
    """Concatenate equally-shaped per-rank tensors along dim 0 (rank order)."""
    if world == 1:
        return t
    t = t.contiguous()
    out = t.new_empty((world * t.shape[0],) + tuple(t.shape[1:]))
    dist.all_gather_into_tensor(out, t, group=group)
    return out


def all_gather_last_feat(last_feat, world, group=None, out=None):
    """last_feat: (V_local, C, H, W) permuted view of NHWC storage (toc3d_eva_vit.py:294).
    Returns the (V, C, H, W) view of the gathered NHWC buffer, images in global order."""
    if world == 1:
        return last_feat
    nhwc = last_feat.permute(0, 2, 3, 1)
    if not nhwc.is_contiguous():
        nhwc = nhwc.contiguous()
    if out is None:
        out = nhwc.new_empty((world * nhwc.shape[0],) + tuple(nhwc.shape[1:]))
    dist.all_gather_into_tensor(out, nhwc, group=group)
    return out.permute(0, 3, 1, 2)


def all_gather_feature_list(feats, world, group=None):
    """Multi-scale feature list of the neck (tuple of (V_local, C, H_l, W_l) maps, possibly strided views):
    one all-gather per level, images in global order.  At 256 channels this moves 4x fewer bytes than
    gathering `last_feat`."""
    if world == 1:
        return tuple(feats)
    return tuple(all_gather_last_feat(f, world, group) for f in feats)


class ShardedBackbone:
    """Runs `backbone` on this rank's image chunk and returns the result for ALL images.

    forward takes the same keyword arguments as the backbone with the full batch (every rank is
    handed the same frame batch, as a data-parallel detector replica would be) and returns the same
    type with `last_feat`, token masks and index lists gathered in global image order.
    """

    def __init__(self, backbone, views=6, group=None):
        self.backbone, self.views, self.group = backbone, views, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def __call__(self, **inputs):
        out = self.backbone(**local_inputs(inputs, self.views, self.rank, self.world))
        if self.world == 1:
            return out
        g = lambda t: all_gather_rows(t, self.world, self.group)
        if isinstance(out, dict):
            return {k: all_gather_last_feat(v, self.world, self.group) for k, v in out.items()}
        feats = {k: all_gather_last_feat(v, self.world, self.group) for k, v in out.img_feats.items()}
        lst = lambda l: None if l is None else [g(t) for t in l]
        return type(out)(feats, lst(out.token_masks), out.attn_scores, keep_idx=lst(out.keep_idx),
                         drop_idx=lst(out.drop_idx), aux_outputs=out.aux_outputs)
