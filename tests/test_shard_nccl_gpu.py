"""The multi-GPU path on hardware: toc3d_b200.shard.ShardedBackbone over NCCL on 2 GPUs (skipped on a 1-GPU box).

Every rank is handed the full frame batch, runs the CUDA backbone on its contiguous image chunk and all-gathers
`last_feat`, the token masks and the keep / drop lists.  Views are independent (the kernels are bit-reproducible per
image), so the gathered result must EQUAL the single-GPU forward of the whole batch bit for bit."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, frames, views, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from toc3d_b200 import TINY, ToC3DEVAViT
        from toc3d_b200 import shard as S
        from toc3d_b200.synthetic import make_gumbel, make_inputs, randomize_state_dict
        torch.manual_seed(0)
        model = ToC3DEVAViT(**TINY).eval()
        model.load_state_dict(randomize_state_dict(model.state_dict(), seed=5, bias_std=0.1))
        model = model.cuda()
        hw = (160, 352)
        inp = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in make_inputs(frames, views, hw, seed=5, pose="random").items()}
        gn = [g.cuda() for g in make_gumbel(frames * views, (hw[0] // 16) * (hw[1] // 16), seed=6)]
        sb = S.ShardedBackbone(model, views=views)
        with torch.no_grad():
            out = sb(**inp, gumbel_noise=gn)
            full = model(**inp, gumbel_noise=gn) if rank == 0 else None
        torch.cuda.synchronize()
        ok = True
        if rank == 0:
            ok = torch.equal(out.img_feats["last_feat"], full.img_feats["last_feat"])
            ok = ok and out.img_feats["last_feat"].shape[0] == frames * views
            for a, b in zip(out.keep_idx + out.drop_idx + out.token_masks, full.keep_idx + full.drop_idx + full.token_masks):
                ok = ok and a.shape == b.shape and torch.equal(a, b)
        # graph-replay path (device-drawn noise): shapes, finiteness, every rank sees the same gathered result
        with torch.no_grad():
            o2 = sb(**inp)
        lf = o2.img_feats["last_feat"].contiguous()
        ref = lf.clone()
        dist.broadcast(ref, 0)
        ok = ok and bool(torch.isfinite(lf).all()) and torch.equal(lf, ref)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("frames,views", [(2, 2), (1, 6)])
def test_sharded_backbone_over_nccl_equals_single_gpu(frames, views):
    """(2, 2): one whole frame per rank; (1, 6): the six views of ONE frame cut 3 + 3 (both ranks score against the same
    history queries)."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, frames, views, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res
