"""GPU tests of OPT-IN kernels that were written without GPU time left in their round (default off in the plugin).
They run only with TOC3D_EXPERIMENTAL=1, so that an unverified kernel can never turn the regular `-m gpu` suite red:
    TOC3D_EXPERIMENTAL=1 python -m pytest tests/test_experimental_gpu.py -m gpu -x -q
Every option here must reproduce the default path bit for bit (same arithmetic, different launch structure)."""
import os

import pytest
import torch

from tests.helpers import build_model, case_setup, run_oracle, to_cuda
from toc3d_b200 import chain_plan

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("TOC3D_EXPERIMENTAL") != "1", reason="set TOC3D_EXPERIMENTAL=1")]
DEV = "cuda"


def _mlp_case(M, C, Hd, scatter, seed):
    g = torch.Generator().manual_seed(seed)
    Hp = (Hd + 31) // 32 * 32
    a = (torch.randn(M, C, generator=g)).to(DEV).bfloat16()
    # interleaved [32 x w1 | 32 x w2] rows, zero rows for the hidden padding (backbone.interleave_w12 layout)
    w12 = torch.zeros(2 * Hp, C)
    blocks = Hp // 32
    w1 = torch.randn(Hp, C, generator=g) * 0.05
    w2 = torch.randn(Hp, C, generator=g) * 0.05
    w1[Hd:] = 0; w2[Hd:] = 0
    w12.view(blocks, 2, 32, C)[:, 0] = w1.view(blocks, 32, C)
    w12.view(blocks, 2, 32, C)[:, 1] = w2.view(blocks, 32, C)
    b12 = torch.randn(2 * Hp, generator=g) * 0.1
    b12.view(blocks, 2, 32)[:, 0][torch.arange(Hp).view(blocks, 32) >= Hd] = 0
    b12.view(blocks, 2, 32)[:, 1][torch.arange(Hp).view(blocks, 32) >= Hd] = 0
    w3 = torch.randn(C, Hp, generator=g) * 0.05
    w3[:, Hd:] = 0
    u3 = torch.randn(C, generator=g)
    b3 = torch.randn(C, generator=g)
    R = M + 57
    x = torch.randn(R, C, generator=g)
    alt = torch.randn(M, C, generator=g)
    maps = {}
    if scatter:
        perm = torch.randperm(R, generator=g)[:M].int()
        omap = perm.clone(); omap[5::11] = -2; omap[3::17] = -1
        rmap = torch.arange(M, dtype=torch.int32)       # residual = alt rows (the compact side buffer)
        maps = dict(out_map=omap.to(DEV), resid_map=torch.full((M,), -2, dtype=torch.int32, device=DEV))
    return dict(a=a, w12=w12.to(DEV).bfloat16(), b12=b12.to(DEV), w3=w3.to(DEV).bfloat16(), u3=u3.to(DEV), b3=b3.to(DEV),
                x=x.to(DEV), alt=alt.to(DEV), maps=maps, Hp=Hp)


def _run_mlp(lib, c, M, C, Hd, chained, units=None):
    x, alt = c["x"].clone(), c["alt"].clone()
    hid = torch.zeros(M, c["Hp"], device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
    e0 = dict(bias=c["b12"], out=hid, row_stats=stats, tile_n=256)
    e1 = dict(bias=c["b3"], out=x, ldo=C, resid=x, out_alt=alt, ln_stats=stats, ln_u=c["u3"], ln_n=Hd, ln_eps=1e-6,
              tile_n=256, **c["maps"])
    if chained:
        plan = chain_plan.plan_mlp_chain(M, 2 * c["Hp"], C, C, units or lib.gemm_chain_units())
        sync = torch.zeros(2 * ((M + 255) // 256), device=DEV, dtype=torch.int32)
        sched = chain_plan.as_tensor(plan, DEV)
        for _ in range(2):                                # the second launch proves the counters were left at zero
            x.copy_(c["x"]); alt.copy_(c["alt"]); stats.zero_()
            lib.mlp_chain(c["a"], c["w12"], c["w3"], M, sched, sync, e0, e1)
        torch.cuda.synchronize()
        assert int(sync.abs().sum()) == 0, "chain counters not reset"
    else:
        lib.gemm(c["a"], c["w12"], lib.EPI_SWIGLU, M=M, **e0)
        lib.gemm(hid, c["w3"], lib.EPI_RESID, M=M, **e1)
    torch.cuda.synchronize()
    return x, alt, hid, stats


@pytest.mark.parametrize("M,C,Hd,scatter", [(300, 256, 341, False), (1000, 128, 200, True), (513, 1024, 2730, False),
                                            (4662, 1024, 2730, True), (8640, 1024, 2730, False)])
def test_mlp_chain_is_bit_identical_to_two_launches(lib, M, C, Hd, scatter):
    c = _mlp_case(M, C, Hd, scatter, seed=M)
    ref = _run_mlp(lib, c, M, C, Hd, chained=False)
    got = _run_mlp(lib, c, M, C, Hd, chained=True)
    for name, r, g in zip(("x", "alt", "hid", "stats"), ref, got):
        assert torch.equal(r, g), name


@pytest.mark.parametrize("units", [1, 3, 20])
def test_mlp_chain_with_few_pairs(lib, units):
    """Small grids make every pair walk through long mixed lists (many dependency waits per pair)."""
    M, C, Hd = 1500, 256, 341
    c = _mlp_case(M, C, Hd, True, seed=units)
    ref = _run_mlp(lib, c, M, C, Hd, chained=False)
    got = _run_mlp(lib, c, M, C, Hd, chained=True, units=units)
    for name, r, g in zip(("x", "alt", "hid", "stats"), ref, got):
        assert torch.equal(r, g), name


def _tail_case(M, C, Hd, seed):
    """proj (norm2 folded) + MLP: the operands of _mlp_case plus the proj weights and the norm2-fold vectors."""
    c = _mlp_case(M, C, Hd, False, seed)
    g = torch.Generator().manual_seed(seed + 1)
    c["ao"] = torch.randn(M, C, generator=g).to(DEV).bfloat16()
    c["wp"] = (torch.randn(C, C, generator=g) * 0.05).to(DEV).bfloat16()
    c["bp"] = (torch.randn(C, generator=g) * 0.1).to(DEV)
    c["u12"] = torch.randn(2 * c["Hp"], generator=g).to(DEV)
    return c


def _run_tail(lib, c, M, C, Hd, chained):
    x = c["x"][:M].clone()
    a = torch.zeros(M, C, device=DEV, dtype=torch.bfloat16)
    hid = torch.zeros(M, c["Hp"], device=DEV, dtype=torch.bfloat16)
    st, st2 = torch.ones(M, 2, device=DEV, dtype=torch.int64), torch.zeros(M, 2, device=DEV, dtype=torch.int64)
    e0 = dict(bias=c["bp"], out=x, ldo=C, resid=x, a_out=a, row_stats=st2, zero_stats=st, tile_n=256)
    e1 = dict(bias=c["b12"], out=hid, row_stats=st, ln_stats=st2, ln_u=c["u12"], ln_n=C, ln_eps=1e-6, tile_n=256)
    e2 = dict(bias=c["b3"], out=x, ldo=C, resid=x, ln_stats=st, ln_u=c["u3"], ln_n=Hd, ln_eps=1e-6, tile_n=256)
    if chained:
        probs = [(C, C, 256), (2 * c["Hp"], C, 256), (C, c["Hp"], 256)]
        plan = chain_plan.plan_chain(M, probs, lib.gemm_chain_units())
        sync = torch.zeros(4 * ((M + 255) // 256), device=DEV, dtype=torch.int32)
        sched = chain_plan.as_tensor(plan, DEV)
        for _ in range(2):
            x.copy_(c["x"][:M]); st.fill_(1); st2.zero_()
            lib.gemm_chain([(c["ao"], c["wp"], lib.EPI_RESID, e0), (a, c["w12"], lib.EPI_SWIGLU, e1),
                            (hid, c["w3"], lib.EPI_RESID, e2)], M, sched, sync)
        torch.cuda.synchronize()
        assert int(sync.abs().sum()) == 0, "chain counters not reset"
    else:
        lib.gemm(c["ao"], c["wp"], lib.EPI_RESID, M=M, **e0)
        lib.gemm(a, c["w12"], lib.EPI_SWIGLU, M=M, **e1)
        lib.gemm(hid, c["w3"], lib.EPI_RESID, M=M, **e2)
    torch.cuda.synchronize()
    return x, a, hid, st, st2


@pytest.mark.parametrize("M,C,Hd", [(300, 256, 341), (1000, 128, 200), (513, 1024, 2730), (4662, 1024, 2730), (6000, 1024, 2730)])
def test_block_tail_chain_is_bit_identical_to_three_launches(lib, M, C, Hd):
    c = _tail_case(M, C, Hd, seed=M + 7)
    ref = _run_tail(lib, c, M, C, Hd, chained=False)
    got = _run_tail(lib, c, M, C, Hd, chained=True)
    for name, r, g in zip(("x", "a", "hid", "stats", "stats2"), ref, got):
        assert torch.equal(r, g), name


def test_mlp_chain_rejects_oversubscribed_grid(lib):
    M, C, Hd = 300, 256, 341
    c = _mlp_case(M, C, Hd, False, seed=1)
    hid = torch.zeros(M, c["Hp"], device=DEV, dtype=torch.bfloat16)
    stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
    sched = torch.full((lib.gemm_chain_units() + 1, 4), -1, device=DEV, dtype=torch.int32)
    with pytest.raises(RuntimeError, match="co-resident"):
        lib.mlp_chain(c["a"], c["w12"], c["w3"], M, sched, torch.zeros(4, device=DEV, dtype=torch.int32),
                      dict(bias=c["b12"], out=hid, row_stats=stats), dict(bias=c["b3"], out=c["x"], ldo=C, resid=c["x"],
                                                                          ln_stats=stats, ln_u=c["u3"], ln_n=Hd, ln_eps=1e-6))


@pytest.mark.parametrize("case", ["tiny_prev_small", "tiny_dense"])
@pytest.mark.parametrize("option", ["fuse_mlp", "fuse_block_tail"])
def test_chained_forward_is_bit_identical(case, option):
    """fuse_mlp / fuse_block_tail through the whole plugin (eager and CUDA-graph replay): same features, masks and
    indices as the separate launches with the same arithmetic (fuse_block_tail <-> fold_norm2)."""
    fx, kind, cfg, model, sd, inp, gn = case_setup(case)
    ref = run_oracle(kind, cfg, sd, inp, gn)
    kw = dict(gumbel_noise=gn, teacher_scores=ref["scores"]) if kind != "dense" else {}
    outs = []
    for fuse in (False, True):
        m = build_model(kind, cfg)
        m.load_state_dict(sd)
        if fuse:
            setattr(m, option, True)
        elif option == "fuse_block_tail":
            m.fold_norm2 = True
        m = m.cuda()
        args = dict(x=inp["x"].cuda()) if kind == "dense" else to_cuda(inp)
        with torch.no_grad():
            o = m(**args, **kw)
            o2 = m(**args, **kw)                  # second call = graph replay where graphs are used
        feat = lambda r: (r["last_feat"] if isinstance(r, dict) else r.img_feats["last_feat"])
        assert torch.equal(feat(o), feat(o2))
        outs.append(o)
    a, b = outs
    feat = lambda r: (r["last_feat"] if isinstance(r, dict) else r.img_feats["last_feat"])
    assert torch.equal(feat(a), feat(b))
    if not isinstance(a, dict):
        assert all(torch.equal(p, q) for p, q in zip(a.keep_idx + a.drop_idx, b.keep_idx + b.drop_idx))
