"""Host-side model of the fused KPConv's tile ring (pcrcg_b200/csrc/kpconv_fused.cu): 13 producer warps that CLAIM points from a
per-CTA counter and fill 8-point tiles in a 3-slot shared-memory ring, one MMA thread, 4 epilogue warps, all synchronised by
phase-parity mbarriers.  No GPU: the model restates the kernel's waits, arrivals and parities one for one and runs them under
randomised and adversarial timings (a discrete-event simulation), checking that

  * no producer ever writes into a slot whose previous tile has not been consumed by the MMAs,
  * the MMA thread always reads exactly the 8 rows of the tile it expects,
  * every barrier phase receives exactly its arrival count, and nothing deadlocks.

It documents a defect found in round 2: WITHOUT the "tiles_issued" gate a warp that runs two uses of a slot ahead passes the
parity test of the slot's "empty" barrier on the completion of tile i - 6 (parity waits cannot tell phase c from phase c + 2),
overwrites tile i - 3 and adds arrivals to its barrier -> wrong rows or a hang (one 4-GPU bench run hung).  The test shows the
un-gated protocol failing in the model and the gated one (the kernel as shipped) surviving the same schedules."""
import heapq
import random

import pytest

SLOTS, TILE, ACCS = 3, 8, 2


class Violation(Exception):
    pass


class MBar:
    """mbarrier restricted to what the kernel uses: init(count), arrive, try_wait.parity"""

    def __init__(self, count):
        self.count, self.pending, self.completed = count, count, 0

    def arrive(self):
        self.pending -= 1
        if self.pending == 0:
            self.completed += 1
            self.pending = self.count

    def test(self, parity):
        # try_wait.parity(P): true once the phase of parity P is over = the phase in progress has the other parity
        return (self.completed & 1) != parity


class Sim:
    def __init__(self, ntiles, warps, claims_in_flight, gate, seed, point_time, mma_time=0.05, epi_time=0.05):
        self.ntiles, self.warps, self.cif, self.gate = ntiles, warps, claims_in_flight, gate
        self.rng = random.Random(seed)
        self.point_time, self.mma_time, self.epi_time = point_time, mma_time, epi_time
        self.full = [MBar(TILE) for _ in range(SLOTS)]
        self.empty = [MBar(1) for _ in range(SLOTS)]
        self.accfull = [MBar(1) for _ in range(ACCS)]
        self.accempty = [MBar(4) for _ in range(ACCS)]
        self.next_point = 0
        self.tiles_issued = 0
        self.slot_rows = [dict() for _ in range(SLOTS)]        # row -> tile that wrote it (since the slot's last consumption)
        self.consumed = [False] * ntiles                        # the MMAs of the tile have completed (its commit has arrived)
        self.epilogued = 0
        self.now = 0.0
        self.heap, self.seq, self.waiting = [], 0, []

    # ---- scheduler ---------------------------------------------------------------------------------
    def spawn(self, gen):
        self._advance(gen)

    def _advance(self, gen):
        try:
            req = next(gen)
        except StopIteration:
            return
        if req[0] == "sleep":
            self.seq += 1
            heapq.heappush(self.heap, (self.now + req[1], self.seq, gen))
        else:                                   # ("wait", predicate)
            if req[1]():
                self._advance(gen)
            else:
                self.waiting.append((req[1], gen))

    def at(self, dt, fn):
        def g():
            yield ("sleep", dt)
            fn()
        self.spawn(g())

    def run(self):
        while self.heap:
            self.now, _, gen = heapq.heappop(self.heap)
            self._advance(gen)
            progress = True
            while progress:                     # wake every waiter whose predicate became true
                progress = False
                for k, (pred, g) in enumerate(self.waiting):
                    if pred():
                        del self.waiting[k]
                        self._advance(g)
                        progress = True
                        break
        if self.waiting or self.epilogued != self.ntiles:
            raise Violation(f"deadlock: {len(self.waiting)} actors blocked, {self.epilogued}/{self.ntiles} tiles finished")

    # ---- actors (each line mirrors a line of the kernel) ------------------------------------------------
    def claim(self):
        v = self.next_point
        self.next_point += 1
        return v

    def producer(self, w):
        m_end = self.ntiles * TILE
        q = []
        for _ in range(self.cif - 1):                            # m, nm (, fm_next): claims taken before the loop; the warps start
            q.append(self.claim())                               # together, so these interleave (warp w gets w, w + 13, ...)
            yield ("sleep", 0.0)
        while q[0] < m_end:
            q.append(self.claim())                               # fm (or the claim one further ahead)
            m = q.pop(0)
            yield ("sleep", self.point_time(self.rng, w, m))     # gather + influence + mma.sync of point m
            i, row = divmod(m, TILE)
            slot, use = i % SLOTS, i // SLOTS
            if self.gate:                                        # while (tiles_issued < i - (SLOTS - 1)) ...
                yield ("wait", lambda i=i: self.tiles_issued >= i - (SLOTS - 1))
            yield ("wait", lambda slot=slot, use=use: self.empty[slot].test((use & 1) ^ 1))
            if i >= SLOTS and not self.consumed[i - SLOTS]:
                raise Violation(f"warp {w} writes row {row} of tile {i} into slot {slot} before tile {i - SLOTS} was consumed "
                                f"(barrier completions {self.empty[slot].completed}, expected {use})")
            self.slot_rows[slot][row] = i
            self.full[slot].arrive()

    def mma(self):
        for i in range(self.ntiles):
            slot, acc = i % SLOTS, i % ACCS
            yield ("wait", lambda acc=acc, i=i: self.accempty[acc].test((((i // ACCS) & 1)) ^ 1))
            yield ("wait", lambda slot=slot, i=i: self.full[slot].test((i // SLOTS) & 1))
            rows = self.slot_rows[slot]
            if sorted(rows) != list(range(TILE)) or any(t != i for t in rows.values()):
                raise Violation(f"MMA of tile {i} reads slot {slot} holding {rows}")

            def done(i=i, slot=slot, acc=acc):                   # tcgen05.commit x2: arrive when the MMAs have completed
                self.consumed[i] = True
                self.slot_rows[slot] = dict()
                self.empty[slot].arrive()
                self.accfull[acc].arrive()
            self.at(self.mma_time * (0.5 + self.rng.random()), done)
            self.tiles_issued = i + 1
            yield ("sleep", 0.01)

    def epilogue(self, q):
        for i in range(self.ntiles):
            acc = i % ACCS
            yield ("wait", lambda acc=acc, i=i: self.accfull[acc].test((i // ACCS) & 1))
            yield ("sleep", self.epi_time * (0.5 + self.rng.random()))
            self.accempty[acc].arrive()
            if q == 0:
                self.epilogued += 1

    def go(self):
        for w in range(self.warps):
            self.spawn(self.producer(w))
        self.spawn(self.mma())
        for q in range(4):
            self.spawn(self.epilogue(q))
        self.run()
        for b in self.full + self.empty + self.accfull + self.accempty:
            assert b.pending == b.count, "a barrier phase was left with stray arrivals"


def heavy_tail(rng, w, m):
    """1-4 k-steps per point (neighbour lists of 1..64 entries in 16-neighbour steps) and an occasional long memory stall"""
    t = rng.choice((1, 1, 2, 3, 4)) * (0.8 + 0.4 * rng.random())
    if rng.random() < 0.03:
        t *= rng.choice((5, 10, 20))
    return t


def bench_like(rng, w, m):
    """a fixed part per point (index / query loads, tile store) + its k-steps, +-10 %"""
    r = rng.random()
    k = 1 if r < 0.03 else (2 if r < 0.82 else 3)
    return (0.6 + k) * (0.9 + 0.2 * rng.random())


def straggler(rng, w, m):
    """one warp stalls for a long time on an early point, another is slow for a while and then fast"""
    if m == 5:
        return 400.0
    if w == 7 and m < 40:
        return 25.0
    return 1.0


def run(gate, cif, seed, point_time=heavy_tail, ntiles=48, warps=13):
    Sim(ntiles, warps, cif, gate, seed, point_time).go()


def test_parity_wait_cannot_tell_two_phases_apart():
    b = MBar(1)
    want_use = 2                                   # a producer about to fill the slot's THIRD tile waits for the second completion
    assert b.test((want_use & 1) ^ 1), "on a barrier with 0 completions the wait for completion 2 passes at once: the ambiguity"
    b.arrive()
    assert not b.test((want_use & 1) ^ 1), "one completion behind: the parity wait blocks as intended"
    b.arrive()
    assert b.test((want_use & 1) ^ 1)


@pytest.mark.parametrize("cif", [3, 4])
def test_ungated_ring_fails_under_drift(cif):
    """The protocol as it was before the gate (3 claims per warp in flight; 4 with the claim issued one iteration ahead)."""
    failures = 0
    for seed in range(300):
        try:
            run(False, cif, seed)
        except Violation:
            failures += 1
    assert failures > 0, "the model no longer reproduces the round-2 defect: is it still the kernel's protocol?"
    # the point times of the bench workload (level 0, limit 34: 3 % of the points have one 16-neighbour k-step, 79 % two, 18 % three,
    # measured with the oracle's neighbour counts) do not trigger it, which is why the GPU parity tests and a few hundred bench
    # steps had passed before a run hung
    for seed in range(50):
        run(False, cif, seed, point_time=bench_like)


@pytest.mark.parametrize("cif", [3, 4])
def test_gated_ring_is_safe(cif):
    """The kernel as shipped (cif = 3); the gate also makes the deeper claim pipeline (cif = 4) safe."""
    for seed in range(1500):
        run(True, cif, seed)
    run(True, cif, 0, point_time=straggler)
    for seed in range(200):                       # extremes: everything instantaneous but one warp / a single warp / tiny grids
        run(True, cif, seed, point_time=lambda rng, w, m: 0.0 if w else 50.0 * rng.random())
        run(True, cif, seed, ntiles=1 + seed % 7, warps=1 + seed % 13)


def test_gated_ring_tail_tiles():
    """my_tiles = 0, 1, 2 (fewer tiles than slots) and claims past the end"""
    for nt in (0, 1, 2, 3, 4):
        for seed in range(50):
            run(True, 3, seed, ntiles=nt)


# ---- exhaustive exploration of a small configuration ------------------------------------------------------------------------------
def explore(W, cif, ntiles, TILE, SLOTS, gate, limit=8_000_000):
    """Every interleaving of a small configuration of the same protocol (W producer warps, TILE points per tile, the MMA thread,
    the completion of its commits as a separate, arbitrarily delayed event, one epilogue actor): depth-first over all enabled
    actions with a visited set.  -> ("safe", states) | ("violation: ...", detail) | ("deadlock", state) | ("limit", states)"""
    ACCS = 2
    m_end = ntiles * TILE
    # state: (next_point, tiles_issued, mma_i, mma_stage, pending commits (tuple of tile ids), epi_i,
    #         full (completed, pending) per slot, empty completed per slot, accfull completed per acc, accempty completed per acc,
    #         rows per slot (tuple of (row,tile) sorted), consumed count (in-order commits => consumed = set of first k tiles),
    #         producers: tuple of (queue tuple, stage))   stage: 0 = need claim (init or loop), 1 = at gate, 2 = at parity, 9 = done
    def initial():
        prods = tuple(((), 0) for _ in range(W))
        full = tuple((0, TILE) for _ in range(SLOTS))
        return (0, 0, 0, 0, (), 0, full, (0,) * SLOTS, (0,) * ACCS, (0,) * ACCS, tuple(() for _ in range(SLOTS)), 0, prods)

    def test(completed, parity):
        return (completed & 1) != parity

    seen = set()
    stack = [initial()]
    n = 0
    while stack:
        s = stack.pop()
        if s in seen:
            continue
        seen.add(s)
        n += 1
        if n > limit:
            return "limit", n
        (nextp, issued, mi, mstage, pend, ei, full, empty, accfull, accempty, rows, consumed, prods) = s
        succ = []
        # ---- producers
        for w, (q, st) in enumerate(prods):
            if st == 9:
                continue
            if st == 0:
                if len(q) < cif - 1:                       # initial claims
                    nq = q + (nextp,)
                    np_ = prods[:w] + ((nq, 0),) + prods[w + 1:]
                    succ.append((nextp + 1, issued, mi, mstage, pend, ei, full, empty, accfull, accempty, rows, consumed, np_))
                elif q[0] >= m_end:
                    np_ = prods[:w] + ((q, 9),) + prods[w + 1:]
                    succ.append((nextp, issued, mi, mstage, pend, ei, full, empty, accfull, accempty, rows, consumed, np_))
                else:                                       # loop head: claim fm, compute m, arrive at the store
                    nq = q + (nextp,)
                    np_ = prods[:w] + ((nq, 1 if gate else 2),) + prods[w + 1:]
                    succ.append((nextp + 1, issued, mi, mstage, pend, ei, full, empty, accfull, accempty, rows, consumed, np_))
            elif st == 1:
                i = q[0] // TILE
                if issued >= i - (SLOTS - 1):
                    np_ = prods[:w] + ((q, 2),) + prods[w + 1:]
                    succ.append((nextp, issued, mi, mstage, pend, ei, full, empty, accfull, accempty, rows, consumed, np_))
            elif st == 2:
                m = q[0]
                i, row = divmod(m, TILE)
                slot, use = i % SLOTS, i // SLOTS
                if test(empty[slot], (use & 1) ^ 1):
                    if i >= SLOTS and consumed <= i - SLOTS:
                        return "violation: overwrite", (w, m, i, empty[slot], use)
                    nrows = rows[:slot] + (tuple(sorted(rows[slot] + ((row, i),))),) + rows[slot + 1:]
                    c, p = full[slot]
                    p -= 1
                    if p == 0:
                        c, p = c + 1, TILE
                    nfull = full[:slot] + ((c, p),) + full[slot + 1:]
                    np_ = prods[:w] + ((q[1:], 0),) + prods[w + 1:]
                    succ.append((nextp, issued, mi, mstage, pend, ei, nfull, empty, accfull, accempty, nrows, consumed, np_))
        # ---- MMA thread
        if mi < ntiles:
            slot, acc = mi % SLOTS, mi % ACCS
            if mstage == 0:
                if test(accempty[acc], ((mi // ACCS) & 1) ^ 1):
                    succ.append((nextp, issued, mi, 1, pend, ei, full, empty, accfull, accempty, rows, consumed, prods))
            else:
                if test(full[slot][0], (mi // SLOTS) & 1):
                    if rows[slot] != tuple((r, mi) for r in range(TILE)):
                        return "violation: mma reads", (mi, rows[slot])
                    succ.append((nextp, mi + 1, mi + 1, 0, pend + (mi,), ei, full, empty, accfull, accempty, rows, consumed, prods))
        # ---- completion of the oldest pending commit
        if pend:
            t = pend[0]
            slot, acc = t % SLOTS, t % ACCS
            nrows = rows[:slot] + ((),) + rows[slot + 1:]
            nempty = empty[:slot] + (empty[slot] + 1,) + empty[slot + 1:]
            naf = accfull[:acc] + (accfull[acc] + 1,) + accfull[acc + 1:]
            succ.append((nextp, issued, mi, mstage, pend[1:], ei, full, nempty, naf, accempty, nrows, consumed + 1, prods))
        # ---- epilogue
        if ei < ntiles:
            acc = ei % ACCS
            if test(accfull[acc], (ei // ACCS) & 1):
                nae = accempty[:acc] + (accempty[acc] + 1,) + accempty[acc + 1:]
                succ.append((nextp, issued, mi, mstage, pend, ei + 1, full, empty, accfull, nae, rows, consumed, prods))
        if not succ:
            done = all(st == 9 for _, st in prods) and mi == ntiles and ei == ntiles and not pend
            if not done:
                return "deadlock", s
        # the producer warps are interchangeable: states that differ only by a permutation of them are one state
        stack.extend(t[:12] + (tuple(sorted(t[12])),) for t in succ)
    return "safe", n


def test_ungated_ring_has_a_violating_schedule_in_the_smallest_configuration():
    """3 warps, 3 claims in flight, one point per tile, 7 tiles: the warp holding claims (1, 2, 6) finishes tiles 1 and 2 while tile 0
    is still being computed and passes the parity wait of slot 0 for tile 6 on a barrier that has not completed once"""
    kind, detail = explore(3, 3, 7, 1, SLOTS, gate=False)
    assert kind == "violation: overwrite", (kind, detail)
    assert explore(3, 3, 6, 1, SLOTS, gate=False)[0] == "safe"          # no third use of a slot: nothing to confuse


def test_gated_ring_is_safe_under_every_interleaving_of_the_smallest_configuration():
    """the same configuration with the gate: the complete reachable state space (2.36 M states, ~0.4 M up to the permutation of the
    warps) holds no overwrite, no mixed tile under the MMAs and no deadlock"""
    kind, states = explore(3, 3, 7, 1, SLOTS, gate=True)
    assert kind == "safe" and states > 300_000, (kind, states)


def test_model_constants_are_the_kernels():
    """the model's ring geometry and the gate it assumes are what csrc/kpconv_fused.cu compiles"""
    import os
    import re
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pcrcg_b200", "csrc", "kpconv_fused.cu")).read()
    const = {k: int(v) for k, v in re.findall(r"constexpr int (FZ_\w+) = (\d+);", src)}
    assert const["FZ_SLOTS"] == SLOTS and const["FZ_TILE"] == TILE and const["FZ_PW"] == 13
    assert "mbar_init(full_bar(s), FZ_TILE)" in src and "mbar_init(empty_bar(s), 1)" in src and "mbar_init(accempty_bar(a), 4)" in src
    # the gate: producers read the issued-tile counter before the parity wait, the MMA thread publishes it after its commits
    gate = src.index("lds_acquire(tiles_issued) < i - (FZ_SLOTS - 1)")
    wait = src.index("mbar_wait(empty_bar(slot), (uint32_t)(((i / FZ_SLOTS) & 1) ^ 1))")
    assert gate < wait
    assert src.index("umma_commit(empty_bar(slot));") < src.index("sts_release(tiles_issued, i + 1);")
    # three claims per warp in flight: two before the loop, one per iteration
    body = src[src.index("int m = claim();"):src.index("store_point(m, 1.0f")]
    assert body.count("claim()") == 3
