"""Discrete-event model of the specialised kernel's shared-memory ring (csrc/spec_kernel.cu): one producer, TEAMS
consumer teams, two carvings of the same bytes (t stages / u-v stages), one full/empty mbarrier pair per (team, stage)
pair (index i mod lcm(teams, stages), phase i / lcm) and the t -> u/v hand-over without a drain.

The model copies the kernel's index arithmetic and checks, under random interleavings of the actors and random
completion order of the bulk copies, that
  * nobody deadlocks,
  * a consumer that passed its wait finds exactly its tile, completely landed,
  * the producer never issues a copy into bytes that still hold a tile somebody has not released.
It is the regression test for the bug the first two-team version had: with one barrier per STAGE a team skips every
other phase of that barrier, and mbarrier parity cannot tell two phases apart (`per_stage_barriers=True` reproduces it).
"""
import math
import random

import pytest


class MBar:
    """mbarrier: `phase` counts completed phases; wait(parity) succeeds iff the phase with that parity has completed"""

    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def passed(self, parity):
        return (self.phase & 1) != parity

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.count

    def arrive(self):
        assert self.pending > 0, "more arrivals than the barrier expects in this phase"
        self.pending -= 1
        self._maybe_complete()

    def arrive_expect_tx(self, nbytes):
        self.tx += nbytes
        self.arrive()

    def complete_tx(self, nbytes):
        self.tx -= nbytes
        self._maybe_complete()


def simulate(teams, nt, nus, bt, bu, ring0, ring1, ring2, warps=2, seed=0, per_stage_barriers=False):
    rnd = random.Random(seed)
    lt = nt if per_stage_barriers else math.lcm(teams, nt)
    lu = nus if per_stage_barriers else math.lcm(teams, nus)
    fullT = [MBar(1) for _ in range(lt)]
    emptyT = [MBar(warps) for _ in range(lt)]
    fullU = [MBar(1) for _ in range(lu)]
    emptyU = [MBar(warps) for _ in range(lu)]
    live = []            # [lo, hi, tile, landed, releases_left]
    inflight = []        # (barrier, live entry)

    def issue(lo, hi, tile, bar):
        for e in live:
            assert hi <= e[0] or lo >= e[1], "copy of %r issued into bytes still holding %r" % (tile, e[2])
        e = [lo, hi, tile, False, warps]
        live.append(e)
        bar.arrive_expect_tx(hi - lo)
        inflight.append((bar, e))

    def producer():
        st = bi = pb = puse = 0
        for i in range(ring0):
            if i >= nt:
                yield (emptyT[pb], puse & 1)
                pb += 1
                if pb == lt:
                    pb, puse = 0, puse + 1
            issue(st * bt, st * bt + bt, ("t", i), fullT[bi])
            st = (st + 1) % nt
            bi = (bi + 1) % lt
        st = bi = pb = puse = k = 0
        for ring_ph in (ring1, ring2):
            for _ in range(ring_ph):
                if k >= nus:
                    yield (emptyU[pb], puse & 1)
                    pb += 1
                    if pb == lu:
                        pb, puse = 0, puse + 1
                else:
                    lo, hi = (st * bu) // bt, ((st + 1) * bu - 1) // bt
                    for s in range(lo, min(hi, nt - 1) + 1):
                        if ring0 > s:
                            last = s + ((ring0 - 1 - s) // nt) * nt
                            yield (emptyT[last % lt], (last // lt) & 1)
                issue(st * bu, st * bu + bu, ("u", k), fullU[bi])
                st = (st + 1) % nus
                bi = (bi + 1) % lu
                k += 1

    def consumer(team):
        def take(tile, full, empty, parity):
            yield (full, parity)
            e = [x for x in live if x[2] == tile]
            assert len(e) == 1 and e[0][3], "waiting for %r passed but the stage holds %r" % (tile, [x[2:4] for x in live])
            yield None                                   # ... computing on the tile ...
            empty.arrive()
            e[0][4] -= 1
            if e[0][4] == 0:
                live.remove(e[0])

        s, bi, use = team % nt, team % lt, team // lt
        for i in range(team, ring0, teams):
            yield from take(("t", i), fullT[bi], emptyT[bi], use & 1)
            s = (s + teams) % nt
            bi += teams
            while bi >= lt:
                bi, use = bi - lt, use + 1
        kbase = 0
        for ring_ph in (ring1, ring2):
            i0 = ((team - kbase) % teams + teams) % teams
            bi, use = (kbase + i0) % lu, (kbase + i0) // lu
            for i in range(i0, ring_ph, teams):
                yield from take(("u", kbase + i), fullU[bi], emptyU[bi], use & 1)
                bi += teams
                while bi >= lu:
                    bi, use = bi - lu, use + 1
            kbase += ring_ph

    actors = [producer()] + [consumer(t) for t in range(teams) for _ in range(warps)]
    blocked = [None] * len(actors)
    alive = set(range(len(actors)))
    steps = 0
    while alive or inflight:
        steps += 1
        assert steps < 2_000_000
        runnable = [a for a in alive if blocked[a] is None or blocked[a][0].passed(blocked[a][1])]
        choices = [("actor", a) for a in runnable] + [("land", k) for k in range(len(inflight))]
        assert choices, "deadlock: every actor waits and no copy is in flight"
        kind, idx = rnd.choice(choices)
        if kind == "land":                               # bulk copies complete in any order
            bar, e = inflight.pop(idx)
            e[3] = True
            bar.complete_tx(e[1] - e[0])
            continue
        blocked[idx] = None
        try:
            blocked[idx] = next(actors[idx])
        except StopIteration:
            alive.discard(idx)
    assert not live
    return steps


CASES = [
    # teams, nt, nus, bt, bu   (bytes in KB)
    (1, 2, 4, 48, 24),      # one surface type, bulk sets
    (1, 3, 13, 36, 8),      # one surface type, RCO
    (2, 3, 5, 60, 40),      # two surface types, bulk (the shipped geometry)
    (2, 3, 5, 64, 40),      # ... MOM5 (16 slots)
    (4, 7, 10, 30, 20),     # the finer variant that was measured and dropped
    (2, 2, 3, 56, 36),
]


@pytest.mark.parametrize("teams,nt,nus,bt,bu", CASES)
def test_ring_protocol_is_safe_and_live(teams, nt, nus, bt, bu):
    rnd = random.Random(teams * 1000 + nt * 100 + nus)
    for trial in range(150):
        ring0, ring1, ring2 = (rnd.choice([0, 1, 2, 3, 4, 5, 7, 8, 13, 29, 40]) for _ in range(3))
        simulate(teams, nt, nus, bt, bu, ring0, ring1, ring2, seed=trial)


def test_one_barrier_per_stage_breaks_with_two_teams():
    """the first two-team version: a team that runs ahead passes a wait on a phase two back and reads a stale stage"""
    failures = 0
    for trial in range(200):
        try:
            simulate(2, 3, 5, 60, 40, 29, 13, 13, seed=trial, per_stage_barriers=True)
        except AssertionError:
            failures += 1
    assert failures > 0
