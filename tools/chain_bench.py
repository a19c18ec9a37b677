"""Kernel-level timing of the SwiGLU MLP: two toc3d_gemm_bf16 launches vs the chained launch (toc3d_gemm_chain_bf16),
EVA-ViT-L shapes, CUDA events over trains of launches (warm L2, like tools/gemm_bench.py)."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from toc3d_b200 import lib, chain_plan

DEV = "cuda"
C, Hd, Hp = 1024, 2730, 2752


def main():
    lib.load()
    units = lib.gemm_chain_units()
    print("chain units", units)
    g = torch.Generator().manual_seed(0)
    w12 = (torch.randn(2 * Hp, C, generator=g) * 0.02).to(DEV).bfloat16()
    w3 = (torch.randn(C, Hp, generator=g) * 0.02).to(DEV).bfloat16()
    b12 = torch.zeros(2 * Hp, device=DEV); b3 = torch.zeros(C, device=DEV); u3 = torch.ones(C, device=DEV)
    for M in [int(a) for a in sys.argv[1:]] or [6000, 8640, 4662, 3400, 12000, 30000]:
        a = torch.randn(M, C, generator=g).to(DEV).bfloat16()
        x = torch.zeros(M, C, device=DEV)
        hid = torch.zeros(M, Hp, device=DEV, dtype=torch.bfloat16)
        stats = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
        e0 = dict(bias=b12, out=hid, row_stats=stats, tile_n=256)
        e1 = dict(bias=b3, out=x, ldo=C, resid=x, ln_stats=stats, ln_u=u3, ln_n=Hd, ln_eps=1e-6, tile_n=256)
        plan = chain_plan.plan_mlp_chain(M, 2 * Hp, C, C, units)
        sched = chain_plan.as_tensor(plan, DEV)
        sync = torch.zeros(2 * ((M + 255) // 256), device=DEV, dtype=torch.int32)

        def two():
            lib.gemm(a, w12, lib.EPI_SWIGLU, M=M, **e0)
            lib.gemm(hid, w3, lib.EPI_RESID, M=M, **e1)

        sh = plan.shape
        seq = [l for l in chain_plan._sequential(sh, min(units, sh.tiles0 + sh.tiles1)) if l]
        seq_plan = chain_plan.Plan(seq, len(seq), max(len(l) for l in seq) + 1, 0.0, "sequential", sh)
        sched_seq = chain_plan.as_tensor(seq_plan, DEV)

        def chain():
            lib.mlp_chain(a, w12, w3, M, sched, sync, e0, e1)

        def chain_seq():
            lib.mlp_chain(a, w12, w3, M, sched_seq, sync, e0, e1)

        res = {}
        for name, fn in (("two", two), ("chain", chain), ("seq", chain_seq)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            t0.record()
            for _ in range(n):
                fn()
            t1.record()
            torch.cuda.synchronize()
            res[name] = t0.elapsed_time(t1) / n * 1e3
        # block tail: proj (norm2 folded) + MLP as three launches vs one chain of three problems
        ao = torch.randn(M, C, generator=g).to(DEV).bfloat16()
        wp = (torch.randn(C, C, generator=g) * 0.02).to(DEV).bfloat16()
        bp = torch.zeros(C, device=DEV); u12 = torch.ones(2 * Hp, device=DEV)
        stats2 = torch.zeros(M, 2, device=DEV, dtype=torch.int64)
        t0p = dict(bias=bp, out=x, ldo=C, resid=x, a_out=a, row_stats=stats2, zero_stats=stats, tile_n=256)
        t1p = dict(e0, ln_stats=stats2, ln_u=u12, ln_n=C, ln_eps=1e-6)
        plan3 = chain_plan.plan_chain(M, [(C, C, 256)] + chain_plan.mlp_probs(2 * Hp, C, C), units)
        sched3 = chain_plan.as_tensor(plan3, DEV)
        sync3 = torch.zeros(4 * ((M + 255) // 256), device=DEV, dtype=torch.int32)

        def three():
            lib.gemm(ao, wp, lib.EPI_RESID, M=M, **t0p)
            lib.gemm(a, w12, lib.EPI_SWIGLU, M=M, **t1p)
            lib.gemm(hid, w3, lib.EPI_RESID, M=M, **e1)

        def chain3():
            lib.gemm_chain([(ao, wp, lib.EPI_RESID, t0p), (a, w12, lib.EPI_SWIGLU, t1p), (hid, w3, lib.EPI_RESID, e1)], M, sched3, sync3)

        for name, fn in (("three", three), ("chain3", chain3)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            for _ in range(20):
                fn()
            t1.record()
            torch.cuda.synchronize()
            res[name] = t0.elapsed_time(t1) / 20 * 1e3
        print("         proj+MLP: three launches %.1f us, chain of three %.1f us (%+.1f %%)  plan=%s" % (
            res["three"], res["chain3"], (res["chain3"] / res["three"] - 1) * 100, plan3.strategy))
        fl = 2.0 * M * C * (2 * Hp) + 2.0 * M * Hp * C
        print("M=%5d  two %.1f us (%.0f TF/s)  chain %.1f us (%.0f TF/s)  %+.1f %%  [sequential order %.1f us]  plan=%s model %.0f -> %.0f" % (
            M, res["two"], fl / res["two"] * 1e-6, res["chain"], fl / res["chain"] * 1e-6,
            (res["chain"] / res["two"] - 1) * 100, res["seq"], plan.strategy,
            chain_plan.separate_launch_makespan(M, chain_plan.mlp_probs(2 * Hp, C, C), units), plan.makespan))


if __name__ == "__main__":
    main()
