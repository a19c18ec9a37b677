"""Host-side planner of the chained MLP launch (toc3d_b200/chain_plan.py): coverage, deadlock freedom, balance.
CPU only - the kernel that consumes these schedules is covered by tests/test_experimental_gpu.py."""
import random

import pytest

from toc3d_b200 import chain_plan as cp

N0, K0, N1 = 5504, 1024, 1024          # EVA-ViT-L MLP: interleaved w1|w2 (2 x 2752), embed 1024
MLP = cp.mlp_probs(N0, K0, N1)
TAIL = [(1024, 1024, 256)] + MLP       # attention output projection (norm2 folded) in front of the MLP
# M of the shipped configs: dense rows (6 / 12 / 24 views at 800x320, 6 views at 1600x800) and compact rows of the stages
MS = [1, 100, 256, 257, 1000, 2600, 3400, 4662, 6000, 8640, 12000, 24000, 30000]


@pytest.mark.parametrize("M", MS)
@pytest.mark.parametrize("units", [74, 66, 8])
@pytest.mark.parametrize("probs", [MLP, TAIL], ids=["mlp", "proj+mlp"])
def test_plan_covers_every_tile_once_and_cannot_deadlock(M, units, probs):
    plan = cp.plan_chain(M, probs, units)
    sh = plan.shape
    assert plan.units <= units and plan.units == len(plan.lists)
    assert sorted(g for l in plan.lists for g in l) == list(range(sh.base[-1]))
    cp.verify(sh, plan.lists)
    # never worse than running the GEMMs' tiles one problem after the other, nor than separate launches (same model)
    seq = [l for l in cp._sequential(sh, plan.units) if l]
    assert plan.makespan <= cp.simulate(sh, seq) + 1e-6
    assert plan.makespan <= cp.separate_launch_makespan(M, probs, plan.units) + 1e-6
    t = cp.as_tensor(plan)
    assert tuple(t.shape) == (plan.units, plan.sched_len) and (t[:, -1] == -1).all()
    for p, l in enumerate(plan.lists):
        assert t[p, :len(l)].tolist() == l and (t[p, len(l):] == -1).all()


@pytest.mark.parametrize("M", [4662, 8640])
@pytest.mark.parametrize("probs", [MLP, TAIL], ids=["mlp", "proj+mlp"])
def test_schedule_completes_under_any_timing(M, probs):
    """Deadlock freedom is a property of the order, not of the cost model: random per-tile slowdowns still finish."""
    plan = cp.plan_chain(M, probs, 74)
    for seed in range(5):
        rnd = random.Random(seed)
        scale = {}
        t = cp.simulate(plan.shape, plan.lists, cost_scale=lambda g: scale.setdefault(g, rnd.choice([0.2, 1.0, 5.0])))
        assert t > 0


def test_verify_rejects_broken_schedules():
    sh = cp.chain_shape(512, [(512, 64, 256), (256, 256, 256)])   # 2 row blocks, 2 + 1 column blocks: 4 + 2 tiles
    assert (sh.num_m, sh.num_n, sh.tiles) == (2, (2, 1), (4, 2))
    a = lambda m, n: m * 2 + n
    b = lambda m: 4 + m
    good = [[a(0, 0), a(0, 1), b(0)], [a(1, 0), a(1, 1), b(1)]]
    cp.verify(sh, good)
    with pytest.raises(ValueError, match="exactly once"):
        cp.verify(sh, [[a(0, 0), a(0, 1), b(0)], [a(1, 0), b(1)]])
    with pytest.raises(ValueError, match="exactly once"):
        cp.verify(sh, [[a(0, 0), a(0, 0), a(0, 1), b(0)], [a(1, 0), a(1, 1), b(1)]])
    # pair 0 waits on row block 1 whose last producer sits behind pair 1's wait on row block 0, and vice versa
    cyc = [[a(0, 0), b(1), a(0, 1)], [a(1, 0), b(0), a(1, 1)]]
    with pytest.raises(ValueError, match="cyclic"):
        cp.verify(sh, cyc)
    with pytest.raises(ValueError, match="deadlock"):
        cp.simulate(sh, cyc)
    # a consumer in front of its own producer in the same list
    with pytest.raises(ValueError, match="cyclic"):
        cp.verify(sh, [[b(0), a(0, 0), a(0, 1)], [a(1, 0), a(1, 1), b(1)]])


def test_planner_beats_two_launches_on_the_shipped_shapes():
    """The modelled gain that motivates the kernel: ragged waves of either GEMM are filled with the other's tiles."""
    for M, least in ((6000, 0.10), (8640, 0.08), (4662, 0.15)):
        plan = cp.plan_mlp_chain(M, N0, K0, N1, 74)
        two = cp.separate_launch_makespan(M, MLP, 74)
        assert plan.makespan < (1.0 - least) * two, (M, plan.makespan, two)


def test_three_problem_dependencies_are_transitive():
    """proj -> SwiGLU -> w3: a w3 tile may never sit in front of a SwiGLU or proj tile of its own row block."""
    plan = cp.plan_chain(4662, TAIL, 74)
    sh = plan.shape
    when = {}
    for p, l in enumerate(plan.lists):
        for i, g in enumerate(l):
            when[g] = (p, i)
    for p, l in enumerate(plan.lists):
        for i, g in enumerate(l):
            q, m, _ = cp.tile_info(sh, g)
            for j in l[i + 1:]:
                qj, mj, _ = cp.tile_info(sh, j)
                assert not (mj == m and qj < q), "a producer sits behind its consumer in the same list"
