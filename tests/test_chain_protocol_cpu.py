"""The dependency-counter protocol of gemm_chain_kernel (gemm_tcgen05.cu), replayed on the CPU under random
interleavings: producers and epilogue warps of a tile wait for ready[q-1][m], epilogue warps publish ready[q][m] and
count done[q-1][m], the last consumer of a row block resets the boundary's counters.  Checks, for 2- and 3-problem
chains and both publishing variants, that every schedule the planner emits runs to completion whatever the
interleaving (no poll can come after the reset it depends on) and leaves the workspace zeroed."""
import random

import pytest

from toc3d_b200 import chain_plan as cp

EPI_WARPS = 8
MLP = cp.mlp_probs(512, 128, 256)                     # small shapes: 2 + 1 column blocks
TAIL = [(256, 128, 256)] + MLP


def _replay(plan, nprob, sig, rnd):
    sh = plan.shape
    target = [sh.num_n[q] * 2 * EPI_WARPS for q in range(nprob)]
    ready = [[0] * sh.num_m for _ in range(nprob - 1)]
    done = [[0] * sh.num_m for _ in range(nprob - 1)]

    def publish(q, m, n):
        if q < nprob - 1:
            ready[q][m] += n
        if q > 0:
            old = done[q - 1][m]
            done[q - 1][m] += n
            if old == target[q] - n:
                done[q - 1][m] = 0
                ready[q - 1][m] = 0

    # one generator per hardware role; `yield cond` blocks the role until cond() holds
    agents = []

    def add_pair(tiles):                                  # own scope per CTA pair (the generators run later)
        info = [cp.tile_info(sh, g) for g in tiles]
        loaded = [[False, False] for _ in tiles]          # producer of each CTA has passed its wait for tile i
        mma_done = [False] * len(tiles)
        epi_left = [2 * EPI_WARPS] * len(tiles)           # accumulator of tile i released by all warps of both CTAs
        arrived = [[0, 0] for _ in tiles]                 # SIG: per-CTA mbarrier arrivals of tile i

        def producer(cta):
            for i, (q, m, _) in enumerate(info):
                if q:
                    yield lambda q=q, m=m: ready[q - 1][m] >= target[q - 1]
                loaded[i][cta] = True
                yield None

        def mma():
            for i in range(len(info)):
                yield lambda i=i: all(loaded[i]) and (i < 2 or epi_left[i - 2] == 0)
                mma_done[i] = True

        def epilogue(cta):
            for i, (q, m, _) in enumerate(info):
                if q:
                    yield lambda q=q, m=m: ready[q - 1][m] >= target[q - 1]
                yield lambda i=i: mma_done[i]
                if sig:
                    arrived[i][cta] += 1
                epi_left[i] -= 1
                if not sig:
                    yield None                            # the fence takes time: other roles may run in between
                    publish(q, m, 1)

        def publisher(cta):
            for i, (q, m, _) in enumerate(info):
                yield lambda i=i: arrived[i][cta] == EPI_WARPS
                publish(q, m, EPI_WARPS)

        for cta in (0, 1):
            agents.append(producer(cta))
            agents.extend(epilogue(cta) for _ in range(EPI_WARPS))
            if sig:
                agents.append(publisher(cta))
        agents.append(mma())

    for tiles in plan.lists:
        add_pair(tiles)
    waiting = {}                                          # agent index -> condition it is blocked on
    live = set(range(len(agents)))
    while live:
        runnable = [a for a in live if a not in waiting or waiting[a]()]
        assert runnable, "deadlock: %d roles blocked" % len(live)
        a = rnd.choice(runnable)
        waiting.pop(a, None)
        try:
            cond = next(agents[a])
            if cond is not None and not cond():
                waiting[a] = cond
        except StopIteration:
            live.discard(a)
        for row in ready + done:
            assert all(0 <= v for v in row)
    assert all(v == 0 for row in ready + done for v in row), "counters not left at zero"


@pytest.mark.parametrize("sig", [False, True], ids=["warp-publish", "publisher-warp"])
@pytest.mark.parametrize("probs", [MLP, TAIL], ids=["mlp", "proj+mlp"])
@pytest.mark.parametrize("M,units", [(300, 2), (700, 3), (1500, 5), (1500, 1)])
def test_counter_protocol_completes_and_cleans_up(M, units, probs, sig):
    plan = cp.plan_chain(M, probs, units)
    for seed in range(3):
        _replay(plan, len(probs), sig, random.Random(seed))


def test_replay_detects_a_cyclic_schedule():
    sh = cp.chain_shape(512, MLP)
    a = lambda m, n: m * sh.num_n[0] + n
    b = lambda m: sh.base[1] + m
    bad = cp.Plan([[a(0, 0), b(1), a(0, 1)], [a(1, 0), b(0), a(1, 1)]], 2, 4, 0.0, "bad", sh)
    with pytest.raises(AssertionError, match="deadlock"):
        _replay(bad, 2, False, random.Random(0))
