"""Coordinator: per-step active-set negotiation (relay control) and heartbeat fault detection.

Behavioural parity with /root/reference/proto/rpc_server.py:20-108:

* ``hook_fetch(step, rank)`` — called by a worker when its first gradient bucket of ``step`` is
  ready. The first arrival becomes the leader and runs the **ski-rental (rent/buy) loop**: every
  ``time_slot`` it compares the cost of waiting one more slot ("rent") with starting a partial
  collective now ("buy"); it stops when ``waited + rent0 >= buy``, when ``waited > relay_threshold``
  or when everybody arrived. Workers arriving before the decision join the active list and block;
  workers arriving after it get the list without themselves — they are relays for this step.
* ``controller_fetch(step, rank)`` — per-step heartbeat from every worker's controller thread.
  If not all ``world_size`` workers report within ``fault_tolerant_time`` the call returns
  ``status=0`` with the survivors; otherwise it blocks until the leader decided and returns the
  step's active list with ``status=1``.

Fixed relative to the reference: state is created lazily per step and
garbage-collected (the reference pre-allocates 1,000,000 dict entries), all shared state is guarded
by one lock + condition variables (the reference shares lists between gRPC threads without locks
and busy-polls), and every wait has a timeout.
"""
from __future__ import annotations

import threading
import time
from concurrent import futures
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from . import messages as pb


@dataclass
class _Step:
    ready: List[int] = field(default_factory=list)       # arrival order; becomes the active list
    decided: bool = False
    heartbeats: List[int] = field(default_factory=list)
    first_arrival: float = 0.0
    decided_at: float = 0.0
    served_controllers: int = 0
    fault: bool = False                                  # a heartbeat deadline was missed in THIS step


class Coordinator:
    """Per-step negotiation service hosted by rank 0: ``hook`` collects the ranks whose first bucket is ready and
    closes the set with the reference's rent-or-buy (ski-rental) rule or after ``relay_threshold``; ``controller``
    is the heartbeat — all alive ranks must report within ``fault_tolerant_time`` or the missing ones are declared
    dead for the rest of the job. One lock + condition variable, per-step state created lazily and pruned
    (/root/reference/proto/rpc_server.py:20-110, whose shared dictionaries are unsynchronised)."""

    def __init__(self, ip: str = "127.0.0.1", port: int = 50051, world_size: int = 1, *,
                 relay_threshold: float = 0.1, time_slot_duration: float = 0.005,
                 fault_tolerant_time: float = 10.0, accumulated_size: float = 100 * 8 / 1024,
                 accumulated_bandwidth: Optional[float] = None, keep_steps: int = 64):
        self.ip, self.port, self.world_size = ip, port, world_size
        self.relay_threshold = relay_threshold
        self.time_slot_duration = time_slot_duration
        self.fault_tolerant_time = fault_tolerant_time
        self.accumulated_size = accumulated_size
        self.accumulated_bandwidth = accumulated_bandwidth if accumulated_bandwidth is not None else 50 * world_size
        self.keep_steps = keep_steps
        self._lock = threading.Lock()
        self._cv = threading.Condition(self._lock)
        self._steps: Dict[int, _Step] = {}
        self.dead: set = set()       # ranks declared failed (missed a heartbeat deadline): never waited for again
        self.arrival_log: Dict[int, List[Tuple[int, float]]] = {}   # straggler-gap measurement hook

    # -- cost model -----------------------------------------------------------------------
    def set_traffic(self, size_gb: float, bandwidth_gbs: float) -> None:
        """Bucket volume / aggregate bandwidth used by the rent-vs-buy rule (the reference hard
        codes 100*8/1024 and 50*world, /root/reference/proto/rpc_server.py:41-42)."""
        with self._lock:
            self.accumulated_size, self.accumulated_bandwidth = size_gb, bandwidth_gbs

    def rent0(self) -> float:
        n = self.world_size
        return 2 * (n - 1) * self.accumulated_size / self.accumulated_bandwidth

    def buy_cost(self, num_ready: int) -> float:
        n = self.world_size
        if n <= 1 or num_ready <= 1:
            return float("inf")
        co_n = (n - 1) / n
        co_m = (num_ready - 1) / num_ready
        return self.rent0() * (co_m / co_n) + n * self.accumulated_size / self.accumulated_bandwidth

    def alive_count(self) -> int:
        return self.world_size - len(self.dead)

    def should_stop(self, waited: float, num_ready: int) -> bool:
        if num_ready >= self.alive_count():
            return True
        if num_ready <= 1:
            return False
        return (waited + self.rent0()) >= self.buy_cost(num_ready) or waited > self.relay_threshold

    # -- state ----------------------------------------------------------------------------
    def _step(self, step: int) -> _Step:
        st = self._steps.get(step)
        if st is None:
            st = self._steps[step] = _Step()
            for old in [s for s in self._steps if s < step - self.keep_steps]:
                del self._steps[old]
            for old in [s for s in self.arrival_log if s < step - self.keep_steps]:
                del self.arrival_log[old]                   # same lifetime as the step state (long runs)
        return st

    # -- service methods (plain Python signatures; gRPC adapters below) ---------------------
    def hook(self, step: int, world_rank: int) -> List[int]:
        with self._cv:
            st = self._step(step)
            self.arrival_log.setdefault(step, []).append((world_rank, time.time()))
            if st.decided:
                return list(st.ready)                       # late: relay for this step
            if st.ready:                                    # waiting active worker
                if world_rank not in st.ready:
                    st.ready.append(world_rank)
                self._cv.notify_all()
                deadline = time.time() + self.relay_threshold + self.fault_tolerant_time
                while not st.decided and time.time() < deadline:
                    self._cv.wait(timeout=0.05)
                return list(st.ready)
            # leader
            st.ready.append(world_rank)
            st.first_arrival = time.time()
            waited = 0.0
            while not self.should_stop(waited, len(st.ready)):
                if self.world_size == 1:
                    break
                self._cv.wait(timeout=self.time_slot_duration)
                waited = time.time() - st.first_arrival
                if waited > self.relay_threshold + self.fault_tolerant_time:
                    break                                    # nobody else is coming
            st.decided = True
            st.decided_at = time.time()
            self._cv.notify_all()
            return list(st.ready)

    def controller(self, step: int, world_rank: int) -> Tuple[List[int], int]:
        with self._cv:
            st = self._step(step)
            if world_rank not in st.heartbeats:
                st.heartbeats.append(world_rank)
            self._cv.notify_all()
            deadline = time.time() + self.fault_tolerant_time
            while len([h for h in st.heartbeats if h not in self.dead]) < self.alive_count():
                left = deadline - time.time()
                if left <= 0:
                    # fault: the ranks that did not report are dead from now on — later steps expect only the
                    # survivors (the reference re-runs the 10 s timeout every step, /root/reference/proto/rpc_server.py:48-59)
                    self.dead |= {r for r in range(self.world_size) if r not in st.heartbeats}
                    st.fault = True
                    self._cv.notify_all()
                    break
                self._cv.wait(timeout=min(left, 0.05))
            if st.fault:
                # EVERY survivor of this step gets status 0 + the survivor list, not only the one whose deadline expired
                # first (the others leave the loop because the dead no longer count)
                return [h for h in st.heartbeats if h not in self.dead], 0
            deadline = time.time() + self.fault_tolerant_time + self.relay_threshold
            while not st.decided:
                left = deadline - time.time()
                if left <= 0:
                    return list(st.heartbeats), 0
                self._cv.wait(timeout=min(left, 0.05))
            st.served_controllers += 1
            return list(st.ready), 1

    def straggler_gap(self, step: int) -> Optional[float]:
        """max - min first-bucket arrival of a step (what units-test/get_wait_time.py records)."""
        with self._lock:
            a = self.arrival_log.get(step)
            if not a or len(a) < 2:
                return None
            ts = [t for _, t in a]
            return max(ts) - min(ts)

    # -- gRPC adapters ----------------------------------------------------------------------
    def controller_fetch(self, request, context=None):
        active, status = self.controller(request.step, request.world_rank)
        return pb.cont_response(active_list=active, status=status)

    def hook_fetch(self, request, context=None):
        return pb.hook_response(active_list=self.hook(request.step, request.world_rank))


def make_server(coordinator: Coordinator, max_workers: Optional[int] = None):
    """gRPC server bound to ``coordinator.ip:port`` (not started). Every rank holds up to two blocking RPCs per
    step (``controller_fetch`` until the step is decided, ``hook_fetch`` as leader or waiter), so the pool is sized
    from the world size: with a fixed pool the controllers of a large job could fill it and starve the hooks that
    would decide the step (-> a false fault after ``fault_tolerant_time``)."""
    import grpc

    if max_workers is None:
        max_workers = max(32, 2 * coordinator.world_size + 8)

    handlers = {
        "controller_fetch": grpc.unary_unary_rpc_method_handler(
            coordinator.controller_fetch, request_deserializer=pb.cont_request.FromString,
            response_serializer=pb.cont_response.SerializeToString),
        "hook_fetch": grpc.unary_unary_rpc_method_handler(
            coordinator.hook_fetch, request_deserializer=pb.hook_request.FromString,
            response_serializer=pb.hook_response.SerializeToString),
    }
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=max_workers))
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(f"{pb.PACKAGE}.{pb.SERVICE}", handlers),))
    bound = server.add_insecure_port(f"{coordinator.ip}:{coordinator.port}")
    if bound == 0:
        raise RuntimeError(f"cannot bind coordinator to {coordinator.ip}:{coordinator.port}")
    coordinator.port = bound
    return server


if __name__ == "__main__":
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--ip", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=50051)
    ap.add_argument("--world_size", type=int, default=4)
    ap.add_argument("--relay_threshold", type=float, default=0.1)
    ap.add_argument("--time_slot_duration", type=float, default=0.005)
    ap.add_argument("--fault_tolerant_time", type=float, default=10.0)
    ap.add_argument("--accumulated_size", type=float, default=100 * 8 / 1024)
    ap.add_argument("--accumulated_bandwidth", type=float, default=0.0)
    ap.add_argument("--parent", type=int, default=0, help="exit when this process disappears (the rank that spawned us)")
    a = ap.parse_args()
    srv = make_server(Coordinator(a.ip, a.port, a.world_size, relay_threshold=a.relay_threshold,
                                  time_slot_duration=a.time_slot_duration, fault_tolerant_time=a.fault_tolerant_time,
                                  accumulated_size=a.accumulated_size,
                                  accumulated_bandwidth=a.accumulated_bandwidth or None))
    srv.start()
    print("coordinator ready", flush=True)
    if a.parent:
        import os

        while srv.wait_for_termination(timeout=1.0):        # True = timed out, still serving
            try:
                os.kill(a.parent, 0)
            except OSError:
                break                                       # orphaned: the training job is gone
        srv.stop(0)
    else:
        srv.wait_for_termination()
