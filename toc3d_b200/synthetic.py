"""Seeded synthetic inputs and weights shared by tests, bench and goldens.

Weights are a deterministic function of (seed, state-dict key), so the same
tensors can be loaded into the reference (in the build container), the CPU
oracle and the CUDA backbone without depending on module construction order.
"""
import zlib

import torch

from .configs import PC_RANGE

_LN_WEIGHT_SUFFIX = ("norm1.weight", "norm2.weight", "ffn_ln.weight", "in_conv.0.weight",
                     "time_embedding.1.weight")
_KEEP = ("freqs_cos", "freqs_sin", "pc_range")


def _gen(seed, key):
    g = torch.Generator()
    g.manual_seed((seed * 1000003 + zlib.crc32(key.encode())) & 0x7FFFFFFF)
    return g


def randomize_state_dict(template, seed=0, bias_std=0.0, weight_std=0.02):
    """template: {key: tensor}.  bias_std=0 reproduces the reference init pattern
    (LN 1/0, biases 0); bias_std>0 re-randomises LN/q/v/linear biases so that pad
    slots, RoPE rows and tie order become observable."""
    out = {}
    for k in sorted(template.keys()):
        t = template[k]
        if k.endswith(_KEEP) or not t.is_floating_point():
            out[k] = t.detach().clone()
            continue
        g = _gen(seed, k)
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        if k.endswith(_LN_WEIGHT_SUFFIX):
            v = 1.0 + bias_std * r
        elif k.endswith("gamma.bias"):
            v = 1.0 + bias_std * r
        elif k.endswith(".bias") or k.endswith("q_bias") or k.endswith("v_bias"):
            v = bias_std * r
        else:
            v = weight_std * r
        out[k] = v.to(t.dtype)
    return out


def _rigid(g, *lead):
    """Random rigid transforms (small yaw, metre-scale translation)."""
    yaw = (torch.rand(*lead, generator=g) - 0.5) * 0.6
    m = torch.eye(4).repeat(*lead, 1, 1)
    m[..., 0, 0] = yaw.cos(); m[..., 0, 1] = -yaw.sin()
    m[..., 1, 0] = yaw.sin(); m[..., 1, 1] = yaw.cos()
    m[..., :3, 3] = torch.randn(*lead, 3, generator=g) * torch.tensor([2.0, 2.0, 0.1])
    return m


def make_inputs(n_samples=1, views=6, hw=(320, 800), seed=0, num_queries=64, pose="identity"):
    """Synthetic frame batch: dict of CPU tensors with the reference forward's kwargs."""
    g = torch.Generator(); g.manual_seed(seed)
    Bf, V = n_samples, n_samples * views
    pc = torch.tensor(PC_RANGE)
    d = dict(
        x=torch.randn(V, 3, hw[0], hw[1], generator=g),
        temp_queries=torch.randn(Bf, num_queries, 256, generator=g),
        temp_ref_points=torch.rand(Bf, num_queries, 3, generator=g) * (pc[3:] - pc[:3]) + pc[:3],
        temp_vel=torch.randn(Bf, num_queries, 2, generator=g),
        temp_timestamp=torch.rand(Bf, num_queries, 1, generator=g, dtype=torch.float64),
        temp_ego_pose=torch.eye(4).expand(Bf, num_queries, 4, 4).contiguous(),
        ego_pose_inv=torch.eye(4).expand(Bf, 4, 4).contiguous(),
        prev_exists=True,
    )
    if pose == "random":
        d["temp_ego_pose"] = _rigid(g, Bf, num_queries)
        d["ego_pose_inv"] = _rigid(g, Bf)
    return d


def make_gumbel(V, N, stages=3, seed=1):
    """-log(-log(U)) noise, U~rand seed 1, (V,N,2) per stage (pin 2)."""
    g = torch.Generator(); g.manual_seed(seed)
    out = []
    for _ in range(stages):
        u = torch.rand(V, N, 2, generator=g).clamp_(1e-10, 1.0 - 1e-7)
        out.append(-torch.log(-torch.log(u)))
    return out
