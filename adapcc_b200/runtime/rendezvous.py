"""Job-unique names for native rendezvous (abstract Unix sockets are per network namespace,
so concurrent jobs on one box must not collide — the reference's fixed SysV keys 1000/2000
did: /root/reference/csrc/shm_ipc.cpp:54-80)."""
from __future__ import annotations

import itertools
import os
import uuid

_counter = itertools.count()


def unique_name(tag: str = "ctx") -> str:
    """Same string on every rank of the job, different for every call and every job."""
    k = next(_counter)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            obj = [uuid.uuid4().hex[:12] if dist.get_rank() == 0 else None]
            dist.broadcast_object_list(obj, src=0)
            return f"{tag}-{obj[0]}-{k}"
    except Exception:
        pass
    job = os.environ.get("ADAPCC_JOB_ID") or f"{os.environ.get('MASTER_PORT', '0')}-{os.getppid()}"
    return f"{tag}-{job}-{k}"
