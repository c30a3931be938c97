"""Model check of the LL all-reduce's slot-reuse rule (csrc/kernels_ll.cuh) on CPU.

The kernel pushes {data, flag} lines into a private slot of every peer's LL buffer and polls its own buffer, with
no barrier; slots are double-buffered by the parity of an LL-only op counter. This test runs the same protocol as
interleaved coroutines under a random scheduler (ranks drift apart by whole ops) and checks that every rank always
reduces exactly the lines of the op it is in — and that the checker is sharp: with a single buffer the same
schedule corrupts data or deadlocks."""
import random

import pytest


def simulate(world: int, ops: int, lines: int, buffers: int, seed: int, max_steps: int = 400_000):
    rng = random.Random(seed)
    # slot[dst][buf][src][line] = (value, flag)
    slot = [[[[(0, 0)] * lines for _ in range(world)] for _ in range(buffers)] for _ in range(world)]
    value = lambda r, k, i: (r + 1) * 1000 + k * 10 + i        # noqa: E731  rank r's input line i in op k
    results = [[None] * ops for _ in range(world)]

    def rank_prog(me):
        for k in range(ops):
            flag = k + 1
            buf = flag % buffers
            for i in range(lines):                               # push my lines to every peer
                for p in range(world):
                    if p != me:
                        slot[p][buf][me][i] = (value(me, k, i), flag)
                        yield
            acc = []
            for i in range(lines):                               # gather in rank order
                s = 0
                for p in range(world):
                    if p == me:
                        s += value(me, k, i)
                        continue
                    while True:
                        v, f = slot[me][buf][p][i]
                        if f == flag:
                            break
                        if f > flag:                             # a later op overwrote a line I still need
                            raise AssertionError(f"rank {me} op {k}: slot of rank {p} already holds op {f - 1}")
                        yield
                    s += v
                    yield
                acc.append(s)
            results[me][k] = acc

    progs = [rank_prog(r) for r in range(world)]
    alive = list(range(world))
    steps = 0
    while alive:
        steps += 1
        if steps > max_steps:
            raise TimeoutError("deadlock / livelock in the protocol model")
        # bursty scheduler: a random rank runs for a random stretch, so ranks get whole ops ahead of each other
        r = rng.choice(alive)
        for _ in range(rng.randint(1, 4 * lines * world)):
            try:
                next(progs[r])
            except StopIteration:
                alive.remove(r)
                break
    for k in range(ops):
        want = [sum(value(r, k, i) for r in range(world)) for i in range(lines)]
        for r in range(world):
            assert results[r][k] == want, (r, k)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_double_buffered_slots_are_safe_under_random_schedules(world):
    for seed in range(25):
        simulate(world, ops=12, lines=3, buffers=2, seed=seed)


def test_single_buffer_is_caught_by_the_model():
    bad = 0
    for seed in range(40):
        try:
            simulate(3, ops=8, lines=3, buffers=1, seed=seed, max_steps=60_000)
        except (AssertionError, TimeoutError):
            bad += 1
    assert bad > 0, "the model failed to expose the single-buffer hazard"
