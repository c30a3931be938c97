"""Model check of the persistent tcgen05 GEMM's barrier protocol (csrc/gemm_tcgen05.cu, variant 1) on CPU.

The kernel's three roles — TMA producer, MMA issuer, four epilogue warps — synchronise only through mbarriers
(phase-parity waits, expect_tx / complete_tx, tcgen05.commit arrivals) with hand-computed stage / phase indices.
This reproduces exactly those index computations as coroutines over a small mbarrier model, with an asynchronous
"TMA engine" and an in-order asynchronous "tensor core", under random schedules, and checks that
  * a shared-memory stage is never overwritten before the MMAs reading it have executed,
  * an accumulator is never written while the epilogue still drains it, and every epilogue sees exactly the
    K-slabs of its own tile,
  * nothing deadlocks.
A deliberately wrong parity (the classic off-by-one) must be caught."""
import random
from collections import deque

import pytest

STAGES = 4


class MBar:
    def __init__(self, count):
        self.count, self.pending, self.tx, self.phase = count, count, 0, 0

    def _maybe_flip(self):
        if self.pending == 0 and self.tx == 0:
            self.phase ^= 1
            self.pending = self.count

    def arrive(self, expect_tx=0):
        self.tx += expect_tx
        self.pending -= 1
        assert self.pending >= 0, "more arrivals than the barrier was initialised for"
        self._maybe_flip()

    def complete_tx(self, n):
        self.tx -= n
        self._maybe_flip()

    def done(self, parity):                 # mbarrier.try_wait.parity: has the phase with this parity completed?
        return self.phase != parity


def simulate(tiles_per_cta, num_kb, seed, bug=None, max_steps=300_000):
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(STAGES)]
    empty = [MBar(1) for _ in range(STAGES)]
    acc_full = [MBar(1) for _ in range(2)]
    acc_empty = [MBar(4) for _ in range(2)]
    smem = [None] * STAGES                   # (tile, kb) currently held by a stage
    acc = [[], []]                           # K-slabs accumulated into each TMEM accumulator
    acc_readers = [0, 0]                     # epilogue warps currently reading an accumulator
    tma_q, tc_q = deque(), deque()           # asynchronous engines
    seen = []

    def producer():
        it = 0
        for tile in range(tiles_per_cta):
            for kb in range(num_kb):
                s, ph = it % STAGES, (it // STAGES) & 1
                want = ph ^ 1 if bug != "producer_parity" else ph
                while not empty[s].done(want):
                    yield
                full[s].arrive(expect_tx=2)
                tma_q.append((s, (tile, kb)))             # A slab
                tma_q.append((s, (tile, kb)))             # W slab
                it += 1
                yield

    def mma():
        it = 0
        for local in range(tiles_per_cta):
            a, aph = local & 1, (local >> 1) & 1
            want = aph ^ 1 if bug != "acc_parity" else aph
            while not acc_empty[a].done(want):
                yield
            for kb in range(num_kb):
                s, ph = it % STAGES, (it // STAGES) & 1
                while not full[s].done(ph):
                    yield
                tc_q.append(("mma", s, a, (local, kb), kb == 0))
                tc_q.append(("commit", empty[s]))
                it += 1
                yield
            tc_q.append(("commit", acc_full[a]))
            yield

    def epilogue(warp):
        for local in range(tiles_per_cta):
            a, aph = local & 1, (local >> 1) & 1
            while not acc_full[a].done(aph):
                yield
            acc_readers[a] += 1
            for _ in range(2):                            # column chunks
                got = list(acc[a])
                assert got == [(local, kb) for kb in range(num_kb)], f"warp {warp} tile {local}: accumulator holds {got}"
                yield
            acc_readers[a] -= 1
            seen.append((warp, local))
            acc_empty[a].arrive()
            yield

    def tma_engine():
        while True:
            if tma_q and rng.random() < 0.6:
                s, what = tma_q.popleft()
                smem[s] = what                            # a late overwrite is caught when the MMA executes
                full[s].complete_tx(1)
            yield

    def tensor_core():
        while True:
            if tc_q and rng.random() < 0.6:
                op = tc_q.popleft()
                if op[0] == "commit":
                    op[1].arrive()
                else:
                    _, s, a, what, first = op
                    assert smem[s] == what, f"MMA for {what} found {smem[s]} in stage {s}"
                    assert acc_readers[a] == 0, f"MMA for {what} writes accumulator {a} while it is being drained"
                    if first:
                        acc[a] = []
                    acc[a].append(what)
            yield

    roles = {"producer": producer(), "mma": mma(), **{f"epi{w}": epilogue(w) for w in range(4)}}
    engines = [tma_engine(), tensor_core()]
    steps = 0
    while roles:
        steps += 1
        if steps > max_steps:
            raise TimeoutError(f"deadlock: waiting roles {sorted(roles)}")
        for e in engines:
            next(e)
        name = rng.choice(sorted(roles))
        for _ in range(rng.randint(1, 6)):
            try:
                next(roles[name])
            except StopIteration:
                del roles[name]
                break
    assert sorted(seen) == sorted((w, t) for w in range(4) for t in range(tiles_per_cta))


@pytest.mark.parametrize("tiles,num_kb", [(1, 1), (1, 12), (2, 3), (5, 12), (7, 5), (6, 48)])
def test_persistent_pipeline_indices_are_consistent(tiles, num_kb):
    for seed in range(12):
        simulate(tiles, num_kb, seed)


@pytest.mark.parametrize("bug", ["producer_parity", "acc_parity"])
def test_model_catches_a_wrong_parity(bug):
    caught = 0
    for seed in range(10):
        try:
            simulate(5, 12, seed, bug=bug, max_steps=40_000)
        except (AssertionError, TimeoutError):
            caught += 1
    assert caught == 10


def simulate_pair(num_kb, seed, stages=6, bug=None, max_steps=200_000):
    """Variant 2 (CTA pair): both CTAs produce, only the leader issues MMAs; the leader's full[s] counts the TMA bytes
    of both CTAs, commits arrive on empty[s] / acc_full of BOTH CTAs at once (multicast)."""
    rng = random.Random(seed)
    full = [MBar(1) for _ in range(stages)]                       # leader's copy only
    empty = [[MBar(1) for _ in range(stages)] for _ in range(2)]
    acc_full = [MBar(1), MBar(1)]
    smem = [[None] * stages for _ in range(2)]                    # [cta][stage] -> kb held
    acc = []
    tma_q, tc_q = deque(), deque()
    done = []

    def producer(cta):
        for kb in range(num_kb):
            s, ph = kb % stages, (kb // stages) & 1
            while not empty[cta][s].done(ph ^ 1 if bug != "pair_parity" else ph):
                yield
            if cta == 0:
                full[s].arrive(expect_tx=4)                       # A and W slabs of both CTAs
            tma_q.append((cta, s, kb))
            tma_q.append((cta, s, kb))
            yield

    def mma():
        for kb in range(num_kb):
            s, ph = kb % stages, (kb // stages) & 1
            while not full[s].done(ph):
                yield
            tc_q.append(("mma", s, kb))
            tc_q.append(("commit", [empty[0][s], empty[1][s]]))
            yield
        tc_q.append(("commit", acc_full))
        yield

    def epilogue(cta):
        while not acc_full[cta].done(0):
            yield
        assert acc == list(range(num_kb)), acc
        done.append(cta)
        yield

    def tma_engine():
        while True:
            if tma_q and rng.random() < 0.6:
                i = rng.randrange(min(len(tma_q), 3))             # the two CTAs' loads complete in any order
                cta, s, kb = tma_q[i]
                del tma_q[i]
                smem[cta][s] = kb
                full[s].complete_tx(1)
            yield

    def tensor_core():
        while True:
            if tc_q and rng.random() < 0.6:
                op = tc_q.popleft()
                if op[0] == "commit":
                    for b in op[1]:
                        b.arrive()
                else:
                    _, s, kb = op
                    assert smem[0][s] == kb and smem[1][s] == kb, (kb, smem[0][s], smem[1][s])
                    acc.append(kb)
            yield

    roles = {"p0": producer(0), "p1": producer(1), "mma": mma(), "e0": epilogue(0), "e1": epilogue(1)}
    engines = [tma_engine(), tensor_core()]
    steps = 0
    while roles:
        steps += 1
        if steps > max_steps:
            raise TimeoutError(f"deadlock: {sorted(roles)}")
        for e in engines:
            next(e)
        name = rng.choice(sorted(roles))
        for _ in range(rng.randint(1, 5)):
            try:
                next(roles[name])
            except StopIteration:
                del roles[name]
                break
    assert sorted(done) == [0, 1]


@pytest.mark.parametrize("num_kb", [1, 5, 6, 12, 48])
def test_cta_pair_protocol_is_consistent(num_kb):
    for seed in range(15):
        simulate_pair(num_kb, seed)


def test_cta_pair_model_catches_a_wrong_parity():
    caught = 0
    for seed in range(10):
        try:
            simulate_pair(24, seed, bug="pair_parity", max_steps=30_000)
        except (AssertionError, TimeoutError):
            caught += 1
    assert caught == 10
