"""Model check of the sharded-optimizer step ordering (adapcc_b200/parallel/engine.py::_zero1_update + hooks).

In ZeRO-1 mode peers write updated parameter slices straight into each other's parameter buffers. The step is only
correct if (a) nobody's parameters are overwritten while that rank's forward/backward of the current step still
reads them and (b) nobody starts the next forward before every slice of the new parameters has arrived. The
engine relies on: the per-bucket reduce-scatter kernels (all-rank barriers inside), the 4-byte grad-norm all-reduce
after the last bucket, and one device barrier after the update. This simulates ranks as coroutines with those
synchronisation points under random schedules and checks both properties; removing either synchronisation must be
caught."""
import random

import pytest


class Barrier:
    """All-rank rendezvous usable repeatedly (models a collective kernel's internal flag barrier)."""
    def __init__(self, world):
        self.world, self.count, self.gen = world, 0, 0

    def arrive(self):
        self.count += 1
        g = self.gen
        if self.count == self.world:
            self.count, self.gen = 0, self.gen + 1
        return g

    def passed(self, g):
        return self.gen > g


def simulate(world, steps, buckets, seed, norm_allreduce=True, final_barrier=True, max_iters=200_000):
    rng = random.Random(seed)
    # params[r][s] = version of slice s in rank r's buffer (slice s is owned / updated by rank s)
    params = [[0] * world for _ in range(world)]
    rs_bar = [Barrier(world) for _ in range(buckets)]
    norm_bar, end_bar = Barrier(world), Barrier(world)

    def wait(bar):
        g = bar.arrive()
        while not bar.passed(g):
            yield

    def rank(me):
        for step in range(steps):
            # forward + backward: many reads of my whole parameter buffer, all must be of version `step`
            for b in range(buckets):
                for _ in range(3):
                    assert all(v == step for v in params[me]), f"rank {me} step {step}: reads mixed versions {params[me]}"
                    yield
                # this bucket's last gradient exists -> its reduce-scatter (all-rank barrier inside the kernel)
                yield from wait(rs_bar[b])
            if norm_allreduce:
                yield from wait(norm_bar)            # 4-byte all-reduce of the squared norm
            for p in range(world):                   # AdamW on my slice + multimem.st into every replica
                params[p][me] = step + 1
                yield
            if final_barrier:
                yield from wait(end_bar)

    progs = {r: rank(r) for r in range(world)}
    iters = 0
    while progs:
        iters += 1
        if iters > max_iters:
            raise TimeoutError("deadlock")
        r = rng.choice(sorted(progs))
        for _ in range(rng.randint(1, 12)):
            try:
                next(progs[r])
            except StopIteration:
                del progs[r]
                break
    assert all(v == steps for row in params for v in row)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_engine_ordering_is_race_free(world):
    for seed in range(30):
        simulate(world, steps=4, buckets=3, seed=seed)


def test_without_the_final_barrier_the_next_forward_races():
    caught = sum(1 for seed in range(30) if _fails(lambda: simulate(4, 4, 3, seed, final_barrier=False)))
    assert caught > 0


def test_bucket_barriers_alone_already_order_the_update_after_every_backward():
    # the reduce-scatter of the LAST bucket is an all-rank rendezvous that every rank enters only after its backward
    # finished, so even without the norm all-reduce no parameter is overwritten early (the all-reduce stays: it is
    # needed for the clip coefficient anyway)
    for seed in range(30):
        simulate(4, steps=4, buckets=3, seed=seed, norm_allreduce=False)


def _fails(fn):
    try:
        fn()
        return False
    except (AssertionError, TimeoutError):
        return True
