"""Coordinator clients — parity with /root/reference/proto/rpc_client.py:11-35.

``Controller.send_relay_request(step, rank) -> (active_list, status)`` and
``Hooker.send_ready_request(step, rank) -> active_list``. Both accept either a network address
(gRPC) or an in-process :class:`Coordinator` (tests, single-process jobs, world_size == 1).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from . import messages as pb
from .server import Coordinator


class _Stub:
    def __init__(self, coordinator_ip, port, local: Optional[Coordinator] = None, timeout: Optional[float] = None):
        self.coordinator_ip, self.port, self.local, self.timeout = coordinator_ip, port, local, timeout
        self._channel = None
        if local is None:
            import grpc

            self._channel = grpc.insecure_channel(f"{coordinator_ip}:{port}")
            self._controller = self._channel.unary_unary(pb.METHOD_CONTROLLER,
                                                         request_serializer=pb.cont_request.SerializeToString,
                                                         response_deserializer=pb.cont_response.FromString)
            self._hook = self._channel.unary_unary(pb.METHOD_HOOK,
                                                   request_serializer=pb.hook_request.SerializeToString,
                                                   response_deserializer=pb.hook_response.FromString)

    def close(self):
        if self._channel is not None:
            self._channel.close()
            self._channel = None


class Controller(_Stub):
    """Controller-thread client: one heartbeat per step; the reply is the step's active list and a status (0 = a
    heartbeat deadline was missed, the list holds the survivors) — /root/reference/proto/rpc_client.py:11-22."""

    def send_relay_request(self, step: int, world_rank: int) -> Tuple[List[int], int]:
        if self.local is not None:
            return self.local.controller(step, world_rank)
        r = self._controller(pb.cont_request(step=step, world_rank=world_rank), timeout=self.timeout,
                             wait_for_ready=True)
        return list(r.active_list), r.status


class Hooker(_Stub):
    """Hook client: called with the first gradient bucket of a step; the reply is the step's active list (a late caller
    learns that it relays) — /root/reference/proto/rpc_client.py:24-35."""

    def send_ready_request(self, step: int, world_rank: int) -> List[int]:
        if self.local is not None:
            return self.local.hook(step, world_rank)
        r = self._hook(pb.hook_request(step=step, world_rank=world_rank), timeout=self.timeout,
                       wait_for_ready=True)
        return list(r.active_list)
