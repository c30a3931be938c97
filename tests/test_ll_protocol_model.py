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


def _ll_kernel_reference(inputs, elems_per_word):
    """Line / word / tail index arithmetic of allreduce_ll_kernel (csrc/kernels_ll.cuh) replayed on Python lists:
    every rank's message is cut into 32-bit words of E elements, two words per 16-byte line; a trailing odd element
    (E == 2) travels alone in a zero-padded word and is written back alone."""
    world, n, E = len(inputs), len(inputs[0]), elems_per_word
    nwords = (n + E - 1) // E
    nlines = (nwords + 1) // 2
    outs = [[None] * n for _ in range(world)]

    def words_of(rank, line):
        w = []
        for k in range(2):
            wi = 2 * line + k
            if (wi + 1) * E <= n:
                w.append(tuple(inputs[rank][wi * E:(wi + 1) * E]))
            elif wi * E < n:
                w.append((inputs[rank][wi * E],) + (0,) * (E - 1))
            else:
                w.append((0,) * E)
        return w

    for me in range(world):
        for line in range(nlines):
            acc = None
            for p in range(world):                           # rank order, own contribution in place
                w = words_of(p, line)
                acc = w if acc is None else [tuple(a + b for a, b in zip(x, y)) for x, y in zip(acc, w)]
            for k in range(2):
                wi = 2 * line + k
                if (wi + 1) * E <= n:
                    outs[me][wi * E:(wi + 1) * E] = list(acc[k])
                elif wi * E < n:
                    outs[me][wi * E] = acc[k][0]
    return outs


@pytest.mark.parametrize("elems_per_word", [1, 2])
def test_ll_line_and_tail_indexing_covers_every_element_once(elems_per_word):
    rng = random.Random(elems_per_word)
    for n in list(range(1, 20)) + [255, 256, 257, 4097]:
        for world in (2, 3, 8):
            inputs = [[rng.randint(-50, 50) for _ in range(n)] for _ in range(world)]
            want = [sum(col) for col in zip(*inputs)]
            for out in _ll_kernel_reference(inputs, elems_per_word):
                assert out == want, (n, world)
