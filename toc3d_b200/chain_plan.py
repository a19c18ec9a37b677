"""Host-side planner for toc3d_gemm_chain_bf16 (include/toc3d_b200.h): which CTA pair runs which tile, in what order.

The chain kernel runs consecutive GEMMs of a block in one persistent launch: the two GEMMs of the SwiGLU MLP
(eva_vit.py:44-51), optionally preceded by the attention output projection with norm2 folded (eva_vit.py:113,263).
A tile of problem q > 0 waits for all problem q-1 tiles of its 256-row block.  The shapes of this path are static per
(token grid, stage, window size), so the work list of every CTA pair is planned here once and uploaded:

  * tile ids as in the header: problem q owns ids [base[q], base[q+1]); (row block, column block) =
    divmod(id - base[q], num_n[q]);
  * cost model in k-block units (one 256 x BN x 64 MMA step of a pair): a tile costs num_k + c_fix, its results are
    visible e_lat units after its last MMA (the epilogue runs under the next tile's mainloop);
  * candidates: "sequential" (problem after problem, round-robin - what separate launches do, minus the launch
    boundaries) and greedy list schedules that keep `reserve` tiles of the last problem back for full final waves and
    start every other dependent tile as soon as its row block is complete; the smallest simulated makespan wins.

Deadlock freedom does not depend on the cost model: verify() checks that every tile appears exactly once and that
list order + dependencies form a DAG, i.e. the lists can always be executed in order whatever the real timing is.
No GPU code here; tests/test_chain_plan.py runs on CPU.
"""
import heapq
from collections import namedtuple

BM_PAIR = 256          # rows of a pair tile (gemm_tcgen05.cu: 2 * BM)
BK = 64
E_LAT = 12.0           # k-block units between a tile's last MMA and its results being visible to other pairs

Shape = namedtuple("Shape", "num_m num_n k tiles base")          # num_n / k / tiles: one entry per problem
Plan = namedtuple("Plan", "lists units sched_len makespan strategy shape")


def chain_shape(M, probs):
    """probs: [(N, K, tile_n), ...] in chain order."""
    num_m = (M + BM_PAIR - 1) // BM_PAIR
    num_n = tuple((N + bn - 1) // bn for N, _, bn in probs)
    k = tuple((K + BK - 1) // BK for _, K, _ in probs)
    tiles = tuple(num_m * n for n in num_n)
    base = [0]
    for t in tiles:
        base.append(base[-1] + t)
    return Shape(num_m, num_n, k, tiles, tuple(base))


def mlp_probs(N0, K0, N1, bn0=256, bn1=256):
    """The SwiGLU MLP: [M,K0] x [N0,K0]^T (interleaved w1|w2) -> hidden N0/2 -> [N1, N0/2]^T."""
    return [(N0, K0, bn0), (N1, N0 // 2, bn1)]


def tile_info(sh, g):
    """-> (problem, row block, column block)"""
    q = 0
    while g >= sh.base[q + 1]:
        q += 1
    return (q,) + divmod(g - sh.base[q], sh.num_n[q])


def verify(sh, lists):
    """Every tile exactly once; list order + row-block dependencies acyclic (executable in order).  Raises ValueError."""
    total = sh.base[-1]
    nq = len(sh.tiles)
    seen = [0] * total
    for l in lists:
        for g in l:
            if not 0 <= g < total:
                raise ValueError("tile id %d out of range" % g)
            seen[g] += 1
    if any(c != 1 for c in seen):
        raise ValueError("schedule does not cover every tile exactly once")
    # Kahn on: predecessor in the same list -> tile; every problem q-1 tile of row block m -> every problem q tile of m.
    # Row-block nodes R(q, m) keep the edge count linear: tiles(q-1, m, *) -> R(q, m) -> tiles(q, m, *).
    rnode = lambda q, m: total + (q - 1) * sh.num_m + m            # q in [1, nq)
    indeg = [0] * (total + (nq - 1) * sh.num_m)
    nxt = [-1] * total
    for l in lists:
        for a, b in zip(l, l[1:]):
            nxt[a] = b
            indeg[b] += 1
    for g in range(sh.base[1], total):
        indeg[g] += 1                                              # from its row-block node
    for q in range(1, nq):
        for m in range(sh.num_m):
            indeg[rnode(q, m)] = sh.num_n[q - 1]
    ready = [g for g in range(total) if indeg[g] == 0]
    done = 0
    while ready:
        g = ready.pop()
        done += 1
        succ = []
        if g < total:
            if nxt[g] >= 0:
                succ.append(nxt[g])
            q, m, _ = tile_info(sh, g)
            if q + 1 < nq:
                succ.append(rnode(q + 1, m))
        else:
            q, m = divmod(g - total, sh.num_m)
            q += 1
            succ.extend(sh.base[q] + m * sh.num_n[q] + n for n in range(sh.num_n[q]))
        for s in succ:
            indeg[s] -= 1
            if indeg[s] == 0:
                ready.append(s)
    if done != len(indeg):
        raise ValueError("schedule has a cyclic wait (would deadlock)")


def simulate(sh, lists, c_fix=1.5, e_lat=E_LAT, cost_scale=None):
    """Makespan of executing `lists` in order under the cost model (k-block units).  cost_scale: optional
    callable(g) -> factor, to test robustness against a wrong model."""
    nq = len(sh.tiles)
    row_ready = [[0.0] * sh.num_m for _ in range(nq)]      # time the last result of (problem, row block) is visible
    row_left = [[sh.num_n[q]] * sh.num_m for q in range(nq)]
    pos = [0] * len(lists)
    free = [0.0] * len(lists)
    end = 0.0
    heap = [(0.0, p) for p in range(len(lists)) if lists[p]]
    heapq.heapify(heap)
    blocked = {}                                           # (problem, row block) -> pairs waiting for it
    executed = 0
    while heap:
        t, p = heapq.heappop(heap)
        g = lists[p][pos[p]]
        q, m, _ = tile_info(sh, g)
        if q > 0 and row_left[q - 1][m] > 0:
            blocked.setdefault((q - 1, m), []).append(p)
            continue
        start = max(t, row_ready[q - 1][m]) if q > 0 else t
        cost = sh.k[q] + c_fix
        if cost_scale is not None:
            cost *= cost_scale(g)
        fin = start + cost
        executed += 1
        end = max(end, fin + e_lat)
        row_ready[q][m] = max(row_ready[q][m], fin + e_lat)
        row_left[q][m] -= 1
        if row_left[q][m] == 0:
            for w in blocked.pop((q, m), []):
                heapq.heappush(heap, (free[w], w))
        pos[p] += 1
        free[p] = fin
        if pos[p] < len(lists[p]):
            heapq.heappush(heap, (fin, p))
    if executed != sh.base[-1]:
        raise ValueError("schedule deadlocks in simulation")
    return end


def _sequential(sh, units):
    """Problem 0 column-major round-robin (the order of the stand-alone GEMM), then the other problems row-major."""
    order = [m * sh.num_n[0] + n for n in range(sh.num_n[0]) for m in range(sh.num_m)]
    order += list(range(sh.base[1], sh.base[-1]))
    lists = [[] for _ in range(units)]
    for i, g in enumerate(order):
        lists[i % units].append(g)
    return lists


def _greedy(sh, units, reserve, c_fix, e_lat):
    """Event-driven list schedule.  A free pair takes, deepest problem first, a dependent tile whose row block is
    complete and visible; else the next problem-0 tile (row-major, so row blocks complete one after the other); else,
    when problem 0 is used up, the next tile of the shallowest unfinished problem (all its producers are placed then)
    and waits for it.  The last `reserve` tiles of the last problem are only taken once every other problem is used
    up, so that they form whole waves at the end."""
    nq = len(sh.tiles)
    last = nq - 1
    n_early = max(0, sh.tiles[last] - reserve)
    nxt = [0] * nq                                         # next tile of each problem (row-major order)
    row_ready = [[0.0] * sh.num_m for _ in range(nq)]
    row_left = [[sh.num_n[q]] * sh.num_m for q in range(nq)]
    lists = [[] for _ in range(units)]
    heap = [(0.0, p) for p in range(units)]
    heapq.heapify(heap)
    remaining = sh.base[-1]
    while remaining:
        t, p = heapq.heappop(heap)
        pick = None
        for q in range(last, 0, -1):
            if nxt[q] >= sh.tiles[q]:
                continue
            m = nxt[q] // sh.num_n[q]
            if row_left[q - 1][m] or row_ready[q - 1][m] > t:
                continue
            if q == last and nxt[q] >= n_early and any(nxt[r] < sh.tiles[r] for r in range(last)):
                continue
            pick = q
            break
        if pick is None:
            pick = next(q for q in range(nq) if nxt[q] < sh.tiles[q])      # shallowest unfinished problem
        q = pick
        m = nxt[q] // sh.num_n[q]
        assert q == 0 or row_left[q - 1][m] == 0
        start = max(t, row_ready[q - 1][m]) if q else t
        lists[p].append(sh.base[q] + nxt[q])
        nxt[q] += 1
        remaining -= 1
        fin = start + sh.k[q] + c_fix
        row_left[q][m] -= 1
        row_ready[q][m] = max(row_ready[q][m], fin + e_lat)
        heapq.heappush(heap, (fin, p))
    return lists


def plan_chain(M, probs, max_units, c_fix=1.5, e_lat=E_LAT):
    """-> Plan.  lists[p] = tile ids of CTA pair p in execution order; units = pairs used (<= max_units)."""
    sh = chain_shape(M, probs)
    units = max(1, min(max_units, sh.base[-1]))
    cands = [("sequential", _sequential(sh, units))]
    tl = sh.tiles[-1]
    reserves = {0, tl, tl % units}
    w = 1
    while w * units <= tl:
        reserves.add(w * units)
        w += 1
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


def plan_mlp_chain(M, N0, K0, N1, max_units, bn0=256, bn1=256, **kw):
    return plan_chain(M, mlp_probs(N0, K0, N1, bn0, bn1), max_units, **kw)


def separate_launch_makespan(M, probs, units, c_fix=1.5, e_lat=E_LAT, launch=6.0):
    """The same cost model for the separate launches the chain replaces (round-robin waves per GEMM, plus the exposed
    epilogue and prologue at every launch boundary) - for reporting the expected gain only."""
    sh = chain_shape(M, probs)
    t = 0.0
    for q in range(len(sh.tiles)):
        t += -(-sh.tiles[q] // units) * (sh.k[q] + c_fix) + e_lat + (launch if q else 0.0)
    return t


def as_tensor(plan, device=None):
    """int32 [units, sched_len] tensor, -1 padded (the `sched` argument of toc3d_gemm_chain_bf16)."""
    import torch
    t = torch.full((plan.units, plan.sched_len), -1, dtype=torch.int32)
    for p, l in enumerate(plan.lists):
        t[p, :len(l)] = torch.tensor(l, dtype=torch.int32)
    return t.to(device) if device is not None else t
