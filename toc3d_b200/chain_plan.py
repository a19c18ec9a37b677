"""Host-side planner for toc3d_mlp_chain_bf16 (include/toc3d_b200.h): which CTA pair runs which tile, in what order.

The chain kernel runs the two GEMMs of the SwiGLU MLP (eva_vit.py:44-51) in one persistent launch.  Tiles of problem 1
(w3) wait for the problem-0 tiles (w1/w2 + SwiGLU) of their 256-row block.  The shapes of this path are static per
(token grid, stage, window size), so the work list of every CTA pair is planned here once and uploaded:

  * tile ids as in the header: g < tiles0 -> problem 0, (row block, column block) = divmod(g, num_n0);
    else problem 1 with divmod(g - tiles0, num_n1);
  * cost model in k-block units (one 256 x BN x 64 MMA step of a pair): a tile costs num_k + c_fix, its results are
    visible e_lat units after its last MMA (the epilogue runs under the next tile's mainloop);
  * candidates: "sequential" (all problem-0 tiles round-robin, then all problem-1 tiles - what two launches do, minus
    the launch boundary) and greedy list schedules that keep `reserve` problem-1 tiles back for full final waves and
    start the others as soon as their row block is complete; the candidate with the smallest simulated makespan wins.

Deadlock freedom does not depend on the cost model: verify() checks that every tile appears exactly once and that
list order + dependencies form a DAG, i.e. the lists can always be executed in order whatever the real timing is.
No GPU code here; tests/test_chain_plan.py runs on CPU.
"""
import heapq
from collections import namedtuple

BM_PAIR = 256          # rows of a pair tile (gemm_tcgen05.cu: 2 * BM)
BK = 64

Shape = namedtuple("Shape", "num_m num_n0 num_n1 k0 k1 tiles0 tiles1")
Plan = namedtuple("Plan", "lists units sched_len makespan strategy shape")


def chain_shape(M, N0, K0, N1, bn0=256, bn1=256):
    num_m = (M + BM_PAIR - 1) // BM_PAIR
    num_n0 = (N0 + bn0 - 1) // bn0
    num_n1 = (N1 + bn1 - 1) // bn1
    k0 = (K0 + BK - 1) // BK
    k1 = (N0 // 2 + BK - 1) // BK
    return Shape(num_m, num_n0, num_n1, k0, k1, num_m * num_n0, num_m * num_n1)


def tile_info(sh, g):
    """-> (problem, row block, column block)"""
    if g < sh.tiles0:
        return (0,) + divmod(g, sh.num_n0)
    return (1,) + divmod(g - sh.tiles0, sh.num_n1)


def verify(sh, lists):
    """Every tile exactly once; list order + row-block dependencies acyclic (executable in order).  Raises ValueError."""
    total = sh.tiles0 + sh.tiles1
    seen = [0] * total
    for l in lists:
        for g in l:
            if not 0 <= g < total:
                raise ValueError("tile id %d out of range" % g)
            seen[g] += 1
    if any(c != 1 for c in seen):
        raise ValueError("schedule does not cover every tile exactly once")
    # Kahn on: predecessor in the same list -> tile; every problem-0 tile of row block m -> every problem-1 tile of m.
    # Row-block nodes keep the edge count linear: A(m, *) -> R(m) -> B(m, *).
    indeg = [0] * (total + sh.num_m)
    nxt = [-1] * total
    for l in lists:
        for a, b in zip(l, l[1:]):
            nxt[a] = b
            indeg[b] += 1
    for g in range(sh.tiles0, total):
        indeg[g] += 1                                  # from its row-block node
    for m in range(sh.num_m):
        indeg[total + m] = sh.num_n0
    ready = [g for g in range(total) if indeg[g] == 0]
    done = 0
    while ready:
        g = ready.pop()
        done += 1
        succ = []
        if g < total:
            if nxt[g] >= 0:
                succ.append(nxt[g])
            if g < sh.tiles0:
                succ.append(total + g // sh.num_n0)
        else:
            m = g - total
            succ.extend(sh.tiles0 + m * sh.num_n1 + n for n in range(sh.num_n1))
        for s in succ:
            indeg[s] -= 1
            if indeg[s] == 0:
                ready.append(s)
    if done != total + sh.num_m:
        raise ValueError("schedule has a cyclic wait (would deadlock)")


def simulate(sh, lists, c_fix=1.5, e_lat=(12.0, 10.0), cost_scale=None):
    """Makespan of executing `lists` in order under the cost model (k-block units).  cost_scale: optional
    callable(g) -> factor, to test robustness against a wrong model."""
    total = sh.tiles0 + sh.tiles1
    row_ready = [0.0] * sh.num_m               # time the last problem-0 result of a row block is visible
    row_left = [sh.num_n0] * sh.num_m
    pos = [0] * len(lists)
    free = [0.0] * len(lists)
    end = 0.0
    heap = [(0.0, p) for p in range(len(lists)) if lists[p]]
    heapq.heapify(heap)
    blocked = {}                               # row block -> pairs waiting for it
    executed = 0
    while heap:
        t, p = heapq.heappop(heap)
        g = lists[p][pos[p]]
        q, m, _ = tile_info(sh, g)
        if q == 1 and row_left[m] > 0:
            blocked.setdefault(m, []).append(p)
            continue
        start = max(t, row_ready[m]) if q == 1 else t
        cost = (sh.k1 if q else sh.k0) + c_fix
        if cost_scale is not None:
            cost *= cost_scale(g)
        fin = start + cost
        executed += 1
        end = max(end, fin + e_lat[q])
        if q == 0:
            row_ready[m] = max(row_ready[m], fin + e_lat[0])
            row_left[m] -= 1
            if row_left[m] == 0:
                for w in blocked.pop(m, []):
                    heapq.heappush(heap, (free[w], w))
        pos[p] += 1
        free[p] = fin
        if pos[p] < len(lists[p]):
            heapq.heappush(heap, (fin, p))
    if executed != total:
        raise ValueError("schedule deadlocks in simulation")
    return end


def _sequential(sh, units):
    """Problem 0 column-major round-robin (the order of the stand-alone GEMM), then problem 1 row-major."""
    order = [m * sh.num_n0 + n for n in range(sh.num_n0) for m in range(sh.num_m)]
    order += [sh.tiles0 + j for j in range(sh.tiles1)]
    lists = [[] for _ in range(units)]
    for i, g in enumerate(order):
        lists[i % units].append(g)
    return lists


def _greedy(sh, units, reserve, c_fix, e_lat):
    """Event-driven list schedule: a free pair takes (1) an 'early' problem-1 tile whose row block is complete, else
    (2) the next problem-0 tile (row-major, so row blocks complete one after the other), else (3) the next remaining
    problem-1 tile.  The last `reserve` problem-1 tiles (by row) are never taken early."""
    n_early = max(0, sh.tiles1 - reserve)
    a_next, b_next = 0, 0                      # next problem-0 / problem-1 tile (both in row-major id order)
    row_ready = [0.0] * sh.num_m
    row_left = [sh.num_n0] * sh.num_m
    lists = [[] for _ in range(units)]
    heap = [(0.0, p) for p in range(units)]
    heapq.heapify(heap)
    cA, cB = sh.k0 + c_fix, sh.k1 + c_fix
    while heap and (a_next < sh.tiles0 or b_next < sh.tiles1):
        t, p = heapq.heappop(heap)
        take_b = False
        if b_next < sh.tiles1:
            mb = b_next // sh.num_n1
            complete = row_left[mb] == 0
            if a_next >= sh.tiles0:
                take_b = True                   # (3): only problem-1 tiles are left; all their producers are placed
            elif b_next < n_early and complete and row_ready[mb] <= t:
                take_b = True                   # (1)
        if take_b:
            mb = b_next // sh.num_n1
            start = max(t, row_ready[mb])
            lists[p].append(sh.tiles0 + b_next)
            b_next += 1
            heapq.heappush(heap, (start + cB, p))
        else:
            m = a_next // sh.num_n0
            lists[p].append(a_next)
            a_next += 1
            fin = t + cA
            row_left[m] -= 1
            row_ready[m] = max(row_ready[m], fin + e_lat[0])
            heapq.heappush(heap, (fin, p))
    return lists


def plan_mlp_chain(M, N0, K0, N1, max_units, bn0=256, bn1=256, c_fix=1.5, e_lat=(12.0, 10.0)):
    """-> Plan.  lists[p] = tile ids of CTA pair p in execution order; units = pairs used (<= max_units)."""
    sh = chain_shape(M, N0, K0, N1, bn0, bn1)
    units = max(1, min(max_units, sh.tiles0 + sh.tiles1))
    cands = [("sequential", _sequential(sh, units))]
    reserves = {0, sh.tiles1}
    w = 1
    while w * units <= sh.tiles1:
        reserves.add(w * units)
        w += 1
    reserves.add(sh.tiles1 % units)
    for r in sorted(reserves):
        cands.append(("greedy(reserve=%d)" % r, _greedy(sh, units, r, c_fix, e_lat)))
    best = None
    for name, lists in cands:
        lists = [l for l in lists if l]
        verify(sh, lists)
        t = simulate(sh, lists, c_fix, e_lat)
        if best is None or t < best[0] - 1e-9:
            best = (t, name, lists)
    t, name, lists = best
    return Plan(lists, len(lists), max(len(l) for l in lists) + 1, t, name, sh)


def two_launch_makespan(M, N0, K0, N1, units, bn0=256, bn1=256, c_fix=1.5, e_lat=(12.0, 10.0), launch=6.0):
    """The same cost model for the two separate launches the chain replaces (round-robin waves per GEMM, plus the
    exposed epilogue and prologue at the launch boundary) - for reporting the expected gain only."""
    sh = chain_shape(M, N0, K0, N1, bn0, bn1)
    w0 = -(-sh.tiles0 // units)
    w1 = -(-sh.tiles1 // units)
    return w0 * (sh.k0 + c_fix) + e_lat[0] + launch + w1 * (sh.k1 + c_fix) + e_lat[1]


def as_tensor(plan, device=None):
    """int32 [units, sched_len] tensor, -1 padded (the `sched` argument of toc3d_mlp_chain_bf16)."""
    import torch
    t = torch.full((plan.units, plan.sched_len), -1, dtype=torch.int32)
    for p, l in enumerate(plan.lists):
        t[p, :len(l)] = torch.tensor(l, dtype=torch.int32)
    return t.to(device) if device is not None else t
