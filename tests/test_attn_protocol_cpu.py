"""Barrier protocol of the persistent attention kernel (attention.cu: window_attention_pp_kernel), replayed on the CPU under
random interleavings of its five roles (TMA producer, MMA issuer, softmax warps of slot 0 / 1, epilogue warps).  The model
keeps what the kernel keeps - mbarriers with phase parity, the in-order tensor pipe with commits, a ring of item buffers,
two TMEM slots whose O columns alias the score columns when a unit has more than 192 key columns, the 4-deep row-sum
exchange - and checks at every step that nobody reads what has not been written yet or overwrites what is still needed.
It is a model of the kernel, not a test of it (the GPU tests are); what it buys is that a change of the wait / arrive
order can be checked for deadlocks and hazards here first.
"""
import random

import pytest


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.phase = count, count, 0

    def arrive(self, n=1):
        self.pending -= n
        assert self.pending >= 0, "more arrivals than the barrier expects in one phase"
        if self.pending == 0:
            self.pending = self.count
            self.phase += 1

    def done(self, parity):                 # try_wait.parity: the phase with this parity has completed
        return (self.phase & 1) != parity


class Sim:
    """units: list of (item, tile, spad, last_tile_of_item)"""

    def __init__(self, items, nbuf, rng):
        self.rng, self.nbuf = rng, nbuf
        self.items = items                                  # per item: (tiles, spad)
        self.units = [(i, t, sp, t == T - 1) for i, (T, sp) in enumerate(items) for t in range(T)]
        self.full = [MBar(1) for _ in range(nbuf)]
        self.empty = [MBar(1) for _ in range(nbuf)]
        self.bar_s, self.bar_o = [MBar(1), MBar(1)], [MBar(1), MBar(1)]
        self.bar_p, self.bar_ofree = [MBar(4), MBar(4)], [MBar(4), MBar(4)]      # 4 warps each (one arrival per warp here)
        self.pipe = []                                      # tensor pipe: in-order list of ("qk"|"pv"|"commit", ...)
        # state checked for hazards
        self.buf_item = [None] * nbuf                       # item whose data sits in ring buffer b (None = being refilled)
        self.s_unit = [None, None]                          # unit whose scores are valid in slot s
        self.p_unit = [None, None]                          # unit whose probabilities are valid in slot s
        self.o_unit = [None, None]                          # unit whose output is valid (unread) in slot s
        self.xsum = [None] * 4
        self.done_units = set()
        self.log = []

    # ---- tensor pipe: executes in issue order, one op per call
    def pipe_step(self):
        if not self.pipe:
            return False
        op = self.pipe.pop(0)
        kind = op[0]
        if kind == "qk":
            _, u, s, buf, spad = op
            item = self.units[u][0]
            assert self.buf_item[buf] == item, "S = Q K^T reads a ring buffer that does not hold its item"
            assert self.p_unit[s] is None, "S(u) overwrites probabilities that P V has not consumed"
            if spad > 192:
                assert self.o_unit[s] is None, "S(u) with > 192 key columns overwrites an unread O"
            self.s_unit[s] = u
        elif kind == "pv":
            _, u, s, buf = op
            assert self.buf_item[buf] == self.units[u][0], "P V reads a ring buffer that does not hold its item"
            assert self.p_unit[s] == u, "P V reads probabilities that are not there"
            assert self.o_unit[s] is None, "P V overwrites an unread O"
            self.p_unit[s] = None
            self.o_unit[s] = u
        else:
            op[1].arrive()
        return True

    def roles(self):
        return [self.producer(), self.mma(), self.softmax(0), self.softmax(1), self.epilogue(), self.tensor_pipe()]

    def tensor_pipe(self):
        while True:
            if not self.pipe_step():
                yield "idle"
            else:
                yield None

    def producer(self):
        for i in range(len(self.items)):
            buf = i % self.nbuf
            while not self.empty[buf].done((((i // self.nbuf) & 1) ^ 1)):
                yield "wait"
            self.buf_item[buf] = None                       # being overwritten
            yield None
            self.buf_item[buf] = i                          # TMA completes ...
            self.full[buf].arrive()                         # ... and its bytes complete the FULL phase
            yield None

    def mma(self):
        prev = None
        k_of = [0, 0]
        for u, (item, t, spad, last) in enumerate(self.units):
            s, buf = u & 1, item % self.nbuf
            k = k_of[s]
            k_of[s] += 1
            if t == 0:
                while not self.full[buf].done((item // self.nbuf) & 1):
                    yield "wait"
            if spad > 192 and k > 0:
                while not self.bar_ofree[s].done((k - 1) & 1):
                    yield "wait"
            self.pipe.append(("qk", u, s, buf, spad))
            self.pipe.append(("commit", self.bar_s[s]))
            yield None
            if prev is not None:
                yield from self.do_pv(*prev)
            prev = (u, s, k, buf, spad, last)
        if prev is not None:
            yield from self.do_pv(*prev)

    def do_pv(self, u, s, k, buf, spad, last):
        while not self.bar_p[s].done(k & 1):
            yield "wait"
        if spad <= 192 and k > 0:
            while not self.bar_ofree[s].done((k - 1) & 1):
                yield "wait"
        self.pipe.append(("pv", u, s, buf))
        self.pipe.append(("commit", self.bar_o[s]))
        if last:
            self.pipe.append(("commit", self.empty[buf]))
        yield None

    def softmax(self, slot):
        k = 0
        for u, (item, t, spad, last) in enumerate(self.units):
            if (u & 1) != slot:
                continue
            while not self.bar_s[slot].done(k & 1):
                yield "wait"
            assert self.s_unit[slot] == u, "softmax reads scores of another unit"
            yield None                                      # max pass, exp pass
            self.s_unit[slot] = None
            self.p_unit[slot] = u
            assert self.xsum[u & 3] is None, "row-sum slot overwritten before the epilogue read it"
            self.xsum[u & 3] = u
            self.bar_p[slot].arrive(4)
            k += 1
            yield None

    def epilogue(self):
        for u, (item, t, spad, last) in enumerate(self.units):
            s = u & 1
            while not self.bar_o[s].done((u >> 1) & 1):
                yield "wait"
            assert self.o_unit[s] == u, "epilogue reads the output of another unit"
            assert self.xsum[u & 3] == u, "epilogue reads a row sum that is not this unit's"
            self.xsum[u & 3] = None
            self.o_unit[s] = None
            self.bar_ofree[s].arrive(4)
            yield None                                      # scale, stage, store
            self.done_units.add(u)
            yield None

    def run(self, max_steps=200000, starve=None):
        """starve: index of a role that is scheduled only rarely (adversarial interleavings)"""
        gens = self.roles()
        alive = [True] * len(gens)
        idle_streak = 0
        for _ in range(max_steps):
            if len(self.done_units) == len(self.units) and not any(alive[:5]):
                return
            order = list(range(len(gens)))
            self.rng.shuffle(order)
            progressed = False
            for g in order[: self.rng.randint(1, len(gens))]:
                if not alive[g] or (g == starve and self.rng.random() > 0.03):
                    continue
                try:
                    r = next(gens[g])
                    if r is None:
                        progressed = True
                except StopIteration:
                    alive[g] = False
                    progressed = True
            idle_streak = 0 if progressed else idle_streak + 1
            assert idle_streak < 20000, "deadlock: no role can make progress (units done: %d of %d)" % (
                len(self.done_units), len(self.units))
        raise AssertionError("did not finish")


def _items(rng, n, spads, max_tiles=2):
    return [(rng.randint(1, max_tiles), rng.choice(spads)) for _ in range(n)]


@pytest.mark.parametrize("seed", range(40))
@pytest.mark.parametrize("nbuf", [2, 3, 4])
def test_pp_protocol_random_interleavings(seed, nbuf):
    rng = random.Random(1000 * nbuf + seed)
    # mixtures of the shipped shapes: <= 192 key columns only (ToC3D stages), > 192 only, ragged dense windows (both)
    spads = [(144,), (192,), (208, 256), (16, 32, 64, 256), (16, 144, 208)][seed % 5]
    sim = Sim(_items(rng, rng.randint(1, 9), spads), nbuf, rng)
    sim.run(starve=(None, 0, 1, 2, 4, 5)[seed % 6])            # fair, or one role (producer, MMA, a slot, epilogue, pipe) slow
    assert sim.done_units == set(range(len(sim.units)))
    assert all(x is None for x in sim.xsum) and sim.p_unit == [None, None] and sim.o_unit == [None, None]


def test_pp_protocol_detects_a_missing_ofree_wait():
    """The model is sensitive: drop the OFREE wait in front of P V and an unread O gets overwritten (or the row-sum slot
    is reused too early) in some interleaving."""
    class Broken(Sim):
        def do_pv(self, u, s, k, buf, spad, last):
            while not self.bar_p[s].done(k & 1):
                yield "wait"
            self.pipe.append(("pv", u, s, buf))
            self.pipe.append(("commit", self.bar_o[s]))
            if last:
                self.pipe.append(("commit", self.empty[buf]))
            yield None

    failures = 0
    for seed in range(60):
        rng = random.Random(seed)
        try:
            Broken(_items(rng, 8, (144,)), 3, rng).run(starve=4)        # a slow epilogue warp
        except AssertionError:
            failures += 1
    assert failures > 0
