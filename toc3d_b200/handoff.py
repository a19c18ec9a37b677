"""Query hand-off from the detection head's memory bank to the backbone.

Host-side mirror of projects/mmdet3d_plugin/models/detectors/petr3d.py:116-143: before every backbone call
`Petr3D.extract_img_feat` takes the first `pruning_num_queries` entries of the StreamPETR head's memory
(`memory_embedding / reference_point / timestamp / egopose / velo`, kept in top-k-by-score order by
streampetr_head.py:315-377), detached, or zeros on the first frame of a scene.  Pure tensor slicing - no kernel.

The one behavioural difference is opt-in: the reference reads `prev_exists.bool().flatten()[0].item()` every
frame (petr3d.py:122), a device->host sync in front of the backbone.  `prev_exists` may be given here as a host
bool (the dataset already knows it), which keeps back-to-back frames free of syncs so the CUDA-graph replay of
the backbone is never stalled; a tensor is still accepted and read the reference's way.
"""
import torch

_FIELDS = (("temp_queries", "memory_embedding", None), ("temp_ref_points", "memory_reference_point", (3,)),
           ("temp_timestamp", "memory_timestamp", (1,)), ("temp_ego_pose", "memory_egopose", (4, 4)),
           ("temp_vel", "memory_velo", (2,)))


def memory_queries(head, prev_exists, batch, num_proposals, device, query_dim=None):
    """-> dict of the backbone's temporal kwargs (`temp_queries`, `temp_ref_points`, `temp_timestamp`,
    `temp_ego_pose`, `temp_vel`, `prev_exists`), exactly what petr3d.py:145-157 passes.

    head: object with the StreamPETR memory attributes (and `embed_dims` unless query_dim is given).
    prev_exists: host bool, or the reference's tensor (first element decides, petr3d.py:122)."""
    if prev_exists is None:
        raise AssertionError("prev_exists is required when query_backbone_selection is on (petr3d.py:121)")
    mid_frame = bool(prev_exists.bool().flatten()[0].item()) if torch.is_tensor(prev_exists) else bool(prev_exists)
    dim = query_dim if query_dim is not None else head.embed_dims
    out = {}
    if not mid_frame or getattr(head, "memory_embedding", None) is None:        # petr3d.py:125-130
        for name, _, tail in _FIELDS:
            out[name] = torch.zeros((batch, num_proposals) + (tail if tail is not None else (dim,)), device=device)
    else:                                                                        # petr3d.py:132-136
        for name, attr, _ in _FIELDS:
            out[name] = getattr(head, attr)[:, :num_proposals].detach()
    out["prev_exists"] = mid_frame
    return out
