"""Wire messages of the coordinator service, built at import time from a descriptor.

Schema parity with /root/reference/proto/protobuf/coordinator.proto:20-43 (package ``coordinator``,
messages ``cont_request{step,world_rank}``, ``cont_response{active_list[],status}``,
``hook_request{step,world_rank}``, ``hook_response{active_list[]}``). The reference checks in
protoc-generated ``coordinator_pb2*.py``; ``grpc_tools`` is not available here, so the classes are
created through ``google.protobuf``'s descriptor pool — same wire format, no generated code.
"""
from __future__ import annotations

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

PACKAGE = "coordinator"
SERVICE = "Coordinator"
_T = descriptor_pb2.FieldDescriptorProto


def _build_pool():
    fdp = descriptor_pb2.FileDescriptorProto(name="adapcc_b200/coordinator.proto", package=PACKAGE, syntax="proto3")

    def msg(name, fields):
        m = fdp.message_type.add(name=name)
        for i, (fname, repeated) in enumerate(fields, start=1):
            m.field.add(name=fname, number=i, type=_T.TYPE_INT32,
                        label=_T.LABEL_REPEATED if repeated else _T.LABEL_OPTIONAL)

    msg("cont_request", [("step", False), ("world_rank", False)])
    msg("cont_response", [("active_list", True), ("status", False)])
    msg("hook_request", [("step", False), ("world_rank", False)])
    msg("hook_response", [("active_list", True)])
    svc = fdp.service.add(name=SERVICE)
    svc.method.add(name="controller_fetch", input_type=f".{PACKAGE}.cont_request", output_type=f".{PACKAGE}.cont_response")
    svc.method.add(name="hook_fetch", input_type=f".{PACKAGE}.hook_request", output_type=f".{PACKAGE}.hook_response")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fdp)
    return pool


_POOL = _build_pool()


def _cls(name):
    return message_factory.GetMessageClass(_POOL.FindMessageTypeByName(f"{PACKAGE}.{name}"))


cont_request = _cls("cont_request")
cont_response = _cls("cont_response")
hook_request = _cls("hook_request")
hook_response = _cls("hook_response")

METHOD_CONTROLLER = f"/{PACKAGE}.{SERVICE}/controller_fetch"
METHOD_HOOK = f"/{PACKAGE}.{SERVICE}/hook_fetch"
