"""Model check of the mbarrier protocol of ntt_pass_pipe_kernel (plonky3_b200/csrc/ntt.cu).

The kernel cannot be exercised without a GPU, but its synchronisation protocol is plain logic: one producer lane, NGROUP
consumer groups, a ring of NSTAGE tile stages (full[s] / empty[s] barriers) and two twiddle buffers per CTA (twfull[b] /
twempty[b]).  This test restates that protocol over a small discrete-event model of mbarriers with the hardware's PARITY
semantics (a wait can only tell the current phase from the one before it) and runs it under random and adversarial
schedules: loads may complete late and out of order, agents are interleaved arbitrarily.

Checked: no deadlock, no arrival-count overflow, every group processes a tile only while the stage really holds that tile's data
and the twiddle buffer its unit's twiddles.  The first version of the kernel let a group wait for its own tiles only; with
tiles that are processed faster than load latency varies, a group could then mistake the stage's older phase for the one it
waited for (hangs and traps on three-pass plans).  `observe_all=False` reproduces that protocol and must fail here.
"""
import random

import pytest


class Overflow(Exception):
    pass


class MBarrier:
    def __init__(self, count):
        self.init, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_complete(self):
        if self.pending == 0 and self.tx == 0:
            self.phase += 1
            self.pending = self.init

    def arrive(self, n=1, tx=0):
        if n > self.pending:
            raise Overflow("more arrivals than the phase expects")
        self.tx += tx
        self.pending -= n
        self._maybe_complete()

    def complete_tx(self, tx):
        self.tx -= tx
        self._maybe_complete()

    def test_wait(self, parity):
        """mbarrier.try_wait.parity: true iff the phase with this parity is the immediately preceding (completed) one."""
        return (self.phase & 1) != parity


class Model:
    def __init__(self, n_units, tpi, last_tpi, ngroup, nstage, observe_all, rng, slow_tile=None):
        self.n_units, self.tpi, self.last_tpi = n_units, tpi, last_tpi
        self.NG, self.NS, self.observe_all, self.rng, self.slow_tile = ngroup, nstage, observe_all, rng, slow_tile
        per_tile = ngroup if observe_all else 1            # arrivals on empty / twempty per tile (group granularity)
        self.full = [MBarrier(1) for _ in range(nstage)]
        self.empty = [MBarrier(per_tile) for _ in range(nstage)]
        self.twfull = [MBarrier(1) for _ in range(2)]
        self.twempty = [MBarrier(tpi * per_tile) for _ in range(2)]
        self.per_tile = per_tile
        self.stage_tag = [None] * nstage                   # which tile's data the stage holds
        self.tw_tag = [None] * 2                           # which unit's twiddles the buffer holds
        self.inflight = []                                 # (due_tick, kind, index, payload)
        self.tick = 0
        self.processed = []

    def tiles_of(self, ui):
        return self.last_tpi if ui == self.n_units - 1 else self.tpi

    # ---- agents are generators; `yield cond` blocks until cond() is true, bare `yield None` is a scheduling point
    def producer(self):
        q = 0
        for ui in range(self.n_units):
            b, ph = ui & 1, (ui >> 1) & 1
            yield lambda: self.twempty[b].test_wait(ph ^ 1)
            n = self.tiles_of(ui)
            if n < self.tpi:
                self.twempty[b].arrive((self.tpi - n) * self.per_tile)
            self.twfull[b].arrive(1, tx=1)
            self.inflight.append((self.tick + self.rng.randint(1, 6), "tw", b, ui))
            for _ in range(n):
                s, k = q % self.NS, q // self.NS
                yield lambda s=s, k=k: self.empty[s].test_wait((k & 1) ^ 1)
                self.full[s].arrive(1, tx=1)
                delay = self.rng.randint(1, 8)
                if self.slow_tile is not None and q == self.slow_tile:
                    delay = 400                              # one load stuck behind a slow DRAM page / refresh
                self.inflight.append((self.tick + delay, "tile", s, q))
                q += 1
                yield None

    def consumer(self, gid):
        q = 0
        for ui in range(self.n_units):
            b, ph = ui & 1, (ui >> 1) & 1
            tw_ready = False
            for _ in range(self.tiles_of(ui)):
                s, k = q % self.NS, q // self.NS
                owner = q % self.NG == gid
                if owner or self.observe_all:
                    yield lambda s=s, k=k: self.full[s].test_wait(k & 1)
                    if not tw_ready:
                        yield lambda b=b, ph=ph: self.twfull[b].test_wait(ph)
                        tw_ready = True
                if not owner:
                    if self.observe_all:
                        self.empty[s].arrive()
                        self.twempty[b].arrive()
                    q += 1
                    continue
                assert self.stage_tag[s] == q, f"group {gid} started tile {q} but stage {s} holds {self.stage_tag[s]}"
                assert self.tw_tag[b] == ui, f"group {gid}: twiddle buffer {b} holds unit {self.tw_tag[b]}, wanted {ui}"
                for _ in range(self.rng.randint(1, 3)):      # "processing": other agents run meanwhile
                    yield None
                assert self.stage_tag[s] == q, f"stage {s} was overwritten while group {gid} processed tile {q}"
                assert self.tw_tag[b] == ui, f"twiddle buffer {b} was overwritten while group {gid} used it"
                self.processed.append(q)
                self.empty[s].arrive()
                self.twempty[b].arrive()
                q += 1

    def run(self):
        agents = [self.producer()] + [self.consumer(g) for g in range(self.NG)]
        blocked = [None] * len(agents)
        alive = [True] * len(agents)
        for _ in range(2_000_000):
            self.tick += 1
            for item in [x for x in self.inflight if x[0] <= self.tick]:
                self.inflight.remove(item)
                _, kind, idx, payload = item
                if kind == "tile":
                    self.stage_tag[idx] = payload
                    self.full[idx].complete_tx(1)
                else:
                    self.tw_tag[idx] = payload
                    self.twfull[idx].complete_tx(1)
            runnable = [i for i, a in enumerate(alive) if a and (blocked[i] is None or blocked[i]())]
            if not any(alive):
                return
            if not runnable:
                if self.inflight:
                    continue
                raise AssertionError("deadlock: every agent is blocked and nothing is in flight")
            i = self.rng.choice(runnable)
            try:
                blocked[i] = next(agents[i])
            except StopIteration:
                alive[i] = False
        raise AssertionError("model did not terminate")


SHAPES = [(5, 13, 13), (40, 1, 1), (9, 3, 1), (12, 8, 6), (6, 5, 5), (30, 2, 2)]   # (units per CTA, tiles per unit, tiles of the last unit)


@pytest.mark.parametrize("n_units,tpi,last_tpi", SHAPES)
def test_every_group_observes_every_phase(n_units, tpi, last_tpi):
    total = (n_units - 1) * tpi + last_tpi
    for seed in range(40):
        rng = random.Random(seed)
        slow = rng.randrange(total) if seed % 2 else None
        m = Model(n_units, tpi, last_tpi, ngroup=4, nstage=6, observe_all=True, rng=rng, slow_tile=slow)
        m.run()
        assert sorted(m.processed) == list(range(total))


def test_owner_only_waits_are_unsound():
    """The shipped-and-fixed bug: without observing the skipped phases a late load breaks the ring."""
    failures = 0
    for seed in range(60):
        rng = random.Random(seed)
        m = Model(40, 8, 8, ngroup=4, nstage=6, observe_all=False, rng=rng, slow_tile=rng.randrange(20, 200))
        try:
            m.run()
            if sorted(m.processed) != list(range(40 * 8)):
                failures += 1
        except (AssertionError, Overflow):
            failures += 1
    assert failures > 0
