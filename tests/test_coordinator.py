"""Coordinator: ski-rental relay decision, heartbeat fault timeout, gRPC wire compatibility."""
import threading
import time

import pytest

from adapcc_b200.coord import Controller, Coordinator, Hooker, make_server
from adapcc_b200.coord import messages as pb


def _run(coord, arrivals, via_grpc=False, port=None):
    res = {}

    def worker(rank, delay, step=0):
        time.sleep(delay)
        if via_grpc:
            h, c = Hooker("127.0.0.1", port), Controller("127.0.0.1", port)
        else:
            h, c = Hooker(None, None, local=coord), Controller(None, None, local=coord)
        res[("h", rank)] = h.send_ready_request(step, rank)
        res[("c", rank)] = c.send_relay_request(step, rank)

    ths = [threading.Thread(target=worker, args=(r, d)) for r, d in arrivals.items()]
    [t.start() for t in ths]
    [t.join(30) for t in ths]
    return res


def test_late_worker_becomes_relay():
    c = Coordinator(world_size=4, relay_threshold=0.05, fault_tolerant_time=3.0)
    res = _run(c, {0: 0.0, 1: 0.002, 2: 0.004, 3: 0.6})
    active = res[("h", 0)]
    assert sorted(active) == [0, 1, 2]
    assert res[("h", 3)] == active and 3 not in res[("h", 3)]          # late -> relay
    assert all(res[("c", r)] == (active, 1) for r in range(4))          # controllers all see the decision
    assert c.straggler_gap(0) > 0.5


def test_everyone_on_time_all_active():
    c = Coordinator(world_size=4, relay_threshold=0.5)
    res = _run(c, {r: 0.001 * r for r in range(4)})
    assert sorted(res[("h", 0)]) == [0, 1, 2, 3]


def test_rent_buy_rule_matches_reference_formula():
    c = Coordinator(world_size=4)
    n, S, B = 4, c.accumulated_size, c.accumulated_bandwidth
    assert c.rent0() == pytest.approx(2 * (n - 1) * S / B)
    m = 2
    assert c.buy_cost(m) == pytest.approx(c.rent0() * ((m - 1) / m) / ((n - 1) / n) + n * S / B)
    assert c.should_stop(0.0, 4) and not c.should_stop(0.0, 1)
    assert c.should_stop(c.relay_threshold + 1e-3, 2)


def test_heartbeat_fault_timeout_reports_survivors():
    c = Coordinator(world_size=3, fault_tolerant_time=0.3)
    out = {}
    ths = [threading.Thread(target=lambda r=r: out.__setitem__(r, c.controller(5, r))) for r in (0, 2)]
    [t.start() for t in ths]
    [t.join(5) for t in ths]
    assert out[0][1] == 0 and sorted(out[0][0]) == [0, 2]                # status 0 + who is alive


def test_grpc_roundtrip_and_wire_format():
    c = Coordinator("127.0.0.1", 0, world_size=2, relay_threshold=0.05)
    srv = make_server(c)
    srv.start()
    try:
        res = _run(c, {0: 0.0, 1: 0.001}, via_grpc=True, port=c.port)
        assert sorted(res[("h", 0)]) == [0, 1] and res[("c", 1)][1] == 1
    finally:
        srv.stop(0)
    # same field numbers/types as the reference's coordinator.proto
    raw = pb.cont_response(active_list=[3, 1], status=1).SerializeToString()
    assert raw == b"\n\x02\x03\x01\x10\x01"
    assert pb.hook_request.FromString(b"\x08\x07\x10\x02").world_rank == 2


def test_concurrent_ranks_always_agree_on_the_active_set():
    """8 rank threads x 200 steps with random lateness against one coordinator: every rank must see the SAME active set
    for a step (relays included), nothing deadlocks, and per-step state stays bounded."""
    import random
    import threading
    import time

    from adapcc_b200.coord.server import Coordinator

    world, steps = 8, 200
    c = Coordinator(world_size=world, relay_threshold=0.002, fault_tolerant_time=2.0)
    seen = [[None] * world for _ in range(steps)]
    errors = []

    def rank(r):
        rng = random.Random(r)
        try:
            for s in range(steps):
                if rng.random() < 0.1:
                    time.sleep(rng.uniform(0.004, 0.015))
                seen[s][r] = tuple(sorted(c.hook(s, r)))
                c.controller(s, r)
        except Exception as e:                                   # pragma: no cover
            errors.append((r, repr(e)))

    threads = [threading.Thread(target=rank, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in threads) and not errors, errors
    sizes = set()
    for s in range(steps):
        sets = {a for a in seen[s] if a is not None}
        assert len(sets) == 1, (s, seen[s])
        sizes.add(len(next(iter(sets))))
    assert min(sizes) < world                                    # lateness did produce relay steps
    assert len(c._steps) <= c.keep_steps + 1


def test_dead_rank_is_not_waited_for_again():
    """A rank that misses the heartbeat deadline once is declared dead: the survivors get status 0 with the survivor
    list for that step, and every LATER step is decided among the survivors without running the fault timeout again."""
    import threading
    import time

    from adapcc_b200.coord import Coordinator

    c = Coordinator(world_size=3, relay_threshold=0.05, fault_tolerant_time=0.4)
    out = {}

    def survivor(rank, step):
        out[(rank, step, "hook")] = c.hook(step, rank)
        out[(rank, step, "ctl")] = c.controller(step, rank)

    t0 = time.time()
    ts = [threading.Thread(target=survivor, args=(r, 5)) for r in (0, 1)]       # rank 2 never reports
    [t.start() for t in ts]
    [t.join() for t in ts]
    for r in (0, 1):      # BOTH survivors learn about the fault in the step it happened (not only the first to time out)
        assert out[(r, 5, "ctl")][1] == 0 and sorted(out[(r, 5, "ctl")][0]) == [0, 1], out
    assert c.dead == {2} and time.time() - t0 >= 0.35
    t1 = time.time()
    ts = [threading.Thread(target=survivor, args=(r, 6)) for r in (0, 1)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert time.time() - t1 < 0.3, "the next step waited for the dead rank again"
    assert out[(0, 6, "ctl")] == ([0, 1], 1) or sorted(out[(0, 6, "ctl")][0]) == [0, 1]
    assert sorted(out[(1, 6, "hook")]) == [0, 1]
