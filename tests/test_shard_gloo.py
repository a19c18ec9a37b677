"""world_size-2 gloo tests (CPU) of the multi-GPU plumbing in toc3d_b200/shard.py.

The backbone stand-in is the CPU oracle (test infrastructure): running it on each rank's image
chunk and all-gathering must reproduce the oracle on the full batch (indices exactly, features to fp32 round-off), which checks the
partition rule, the per-frame slicing of the history-query inputs and the gather order.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from toc3d_b200 import TINY, ToC3DViTReturnType
from toc3d_b200 import shard as S
from toc3d_b200.synthetic import make_gumbel, make_inputs, randomize_state_dict


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleBackbone:
    """Adapter: the oracle behind the plugin's forward/return contract (tests only)."""

    def __init__(self, sd, cfg, gumbel, views_total):
        self.sd, self.cfg, self.gumbel, self.views_total = sd, cfg, gumbel, views_total
        self.offset = 0

    def __call__(self, x, temp_queries, temp_ref_points, temp_vel, temp_timestamp, temp_ego_pose, ego_pose_inv,
                 prev_exists):
        from oracle import toc3d_oracle as O
        gn = [g[self.offset:self.offset + x.shape[0]] for g in self.gumbel]
        with torch.no_grad():
            o = O.forward_toc3d(self.sd, self.cfg, x, temp_queries, temp_ref_points, temp_vel, temp_timestamp,
                                temp_ego_pose, ego_pose_inv, prev_exists, gn)
        return ToC3DViTReturnType({"last_feat": o["last_feat"]}, o["token_masks"], None, keep_idx=o["keep_idx"],
                                  drop_idx=o["drop_idx"])


def _setup(frames, views, hw):
    from toc3d_b200 import ToC3DEVAViT
    torch.manual_seed(0)
    sd = randomize_state_dict(ToC3DEVAViT(**TINY).state_dict(), seed=5, bias_std=0.1)
    inp = make_inputs(frames, views, hw, seed=5, pose="random")
    gn = make_gumbel(frames * views, (hw[0] // 16) * (hw[1] // 16), seed=6)
    return sd, inp, gn


def _worker(rank, world, port, frames, views, hw, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        sd, inp, gn = _setup(frames, views, hw)
        bb = _OracleBackbone(sd, TINY, gn, frames * views)
        bb.offset = S.partition(frames, views, world)[rank][0]
        out = S.ShardedBackbone(bb, views=views)(**inp)
        if rank == 0:
            q.put((out.img_feats["last_feat"], [t for t in out.keep_idx], [t for t in out.token_masks]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("frames,views", [(2, 2), (1, 2)])
def test_sharded_forward_equals_full_batch(frames, views):
    hw = (96, 160)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, frames, views, hw, q)) for r in range(world)]
    for p in procs:
        p.start()
    lf, keep, masks = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    sd, inp, gn = _setup(frames, views, hw)
    full = _OracleBackbone(sd, TINY, gn, frames * views)(**inp)
    assert lf.shape == full.img_feats["last_feat"].shape
    # fp32 BLAS blocking depends on the batch size: last-ulp differences, not bit equality
    assert (lf - full.img_feats["last_feat"]).abs().max().item() < 1e-3
    assert all(torch.equal(a, b) for a, b in zip(keep, full.keep_idx))
    assert all((a - b).abs().max().item() < 1e-4 for a, b in zip(masks, full.token_masks))


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(4 * 3 * 4 * 6, dtype=torch.float32).reshape(4, 4, 6, 3)          # NHWC storage
        loc = full[2 * rank:2 * rank + 2].permute(0, 3, 1, 2)                                  # (2, C, H, W) view
        feats = S.all_gather_feature_list((loc, loc[:, :, ::2, ::2]), world)
        if rank == 0:
            q.put([f.clone() for f in feats])
    finally:
        dist.destroy_process_group()


def test_feature_list_gather_order_and_strided_levels():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    f0, f1 = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    full = torch.arange(4 * 3 * 4 * 6, dtype=torch.float32).reshape(4, 4, 6, 3).permute(0, 3, 1, 2)
    assert torch.equal(f0, full) and torch.equal(f1, full[:, :, ::2, ::2])


def test_partition_rules():
    assert S.partition(4, 6, 8) == [(3 * r, 3 * r + 3) for r in range(8)]
    assert S.partition(4, 6, 4) == [(6 * r, 6 * r + 6) for r in range(4)]
    assert S.partition(1, 6, 2) == [(0, 3), (3, 6)]
    with pytest.raises(ValueError):
        S.partition(1, 6, 4)            # 6 images do not split over 4 ranks
    with pytest.raises(ValueError):
        S.partition(2, 6, 3)            # 4-image chunks would straddle 6-view frames
    inp = make_inputs(4, 6, (32, 32), seed=0)
    loc = S.local_inputs(inp, 6, rank=5, world=8)
    assert loc["x"].shape[0] == 3 and torch.equal(loc["x"], inp["x"][15:18])
    assert loc["temp_queries"].shape[0] == 1 and torch.equal(loc["temp_queries"], inp["temp_queries"][2:3])
    loc = S.local_inputs(inp, 6, rank=1, world=2)
    assert loc["x"].shape[0] == 12 and torch.equal(loc["ego_pose_inv"], inp["ego_pose_inv"][2:4])
