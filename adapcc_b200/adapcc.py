"""High-level API: the ``AdapCC`` class-level singleton.

Same surface as /root/reference/adapcc.py:6-76 — ``init(args, local_rank, world_rank, world_size)``,
``setup(prim)``, ``allreduce/reduce/boardcast/alltoall``, ``reconstruct_topology(args, prim)``,
``set_profile_freq``, ``clear(prim)`` and the ``communicator`` attribute — so a training script
written for the reference (see /root/reference/train_ddp.py:30-58) only changes its import.

``args`` needs: ``port, strategy_file, logical_graph, entry_point, parallel_degree, profile_freq``
(optional extras: ``backend, algo, wire_dtype, reduce_op, relay_mode, relay_control, policy,
staging_mb, heap_mb, work_dir``). ``entry_point``: 6 = detect + profile + synthesise,
7 = profile + synthesise, -1 = use the given strategy file as is.
"""
from __future__ import annotations

from .commu import CudaCommu
from .constants import (ALLGATHER, ALLREDUCE, ALLTOALL, BOARDCAST, DETECT, PROFILE, REDUCE,  # noqa: F401
                        REDUCESCATTER)


class AdapCC:
    """Class-level singleton, the library's public face: ``init → setup(prim) → allreduce / reduce / boardcast /
    alltoall / reducescatter / allgather → clear``, ``reconstruct_topology`` for on-the-fly re-profiling; state
    lives in ``AdapCC.communicator`` (/root/reference/adapcc.py:6-76)."""

    # meta info since the first registered
    communicator_path = None            # resolved lazily: adapcc_b200/_C/libadapcc.so
    communicator: CudaCommu = None
    local_rank = None
    world_rank = None
    world_size = None
    profile_freq = None

    @classmethod
    def init(cls, args, local_rank, world_rank, world_size, _native=None):
        """Create the communicator and run the workflow stages ``args.entry_point`` asks for: 6 = DETECT
        (topology -> logical graph) then PROFILE (link microbench -> synthesised strategy XML), 7 = PROFILE only,
        -1 = keep ``args.strategy_file``. Collective over all ranks (reference: /root/reference/adapcc.py:16-42)."""
        prev = cls.communicator
        if prev is not None and _native is None and not getattr(prev, "_cleared", False):
            # init() on top of a live communicator (no clear() in between): tear the old one down first instead of
            # leaking its coordinator server, controller thread and symmetric buffers
            prev.clear()
        dylib = None
        if getattr(args, "backend", "nccl") != "gloo":
            try:
                import torch

                if torch.cuda.is_available():
                    from .runtime.native import lib_path, load_library

                    dylib = load_library()
                    cls.communicator_path = str(lib_path())
            except ImportError:
                dylib = None
        cls.communicator = CudaCommu(args, dylib, local_rank, world_rank, world_size)
        cls.communicator.adopt_native(_native)         # reconstruct_topology: keep the symmetric buffers alive
        cls.local_rank, cls.world_rank, cls.world_size = local_rank, world_rank, world_size
        cls.profile_freq = getattr(args, "profile_freq", None)

        entry = getattr(args, "entry_point", -1)
        if entry == DETECT:
            cls.communicator.init_threads(DETECT)
            cls.communicator.exit_threads(DETECT)
            cls.communicator.init_threads(PROFILE)
            cls.communicator.exit_threads(PROFILE)
        elif entry == PROFILE:
            cls.communicator.init_threads(PROFILE)
            cls.communicator.exit_threads(PROFILE)
        elif entry == -1 or entry is None:
            pass
        else:
            print("no supported entry point for init.")

    @classmethod
    def _comm(cls) -> CudaCommu:
        if cls.communicator is None:
            raise RuntimeError("AdapCC.init(args, local_rank, world_rank, world_size) has not been called")
        return cls.communicator

    @classmethod
    def setup(cls, prim):
        """Build the data-plane context for ``prim`` (ALLREDUCE / REDUCE / BOARDCAST / ALLGATHER / ALLTOALL / REDUCESCATTER): symmetric buffers,
        strategy tables, coordinator + controller when relay control is on (reference: /root/reference/adapcc.py:44-46)."""
        cls._comm().init_threads(prim)

    @classmethod
    def allreduce(cls, tensor, size=None, chunk_bytes=None, active_gpus=None):
        """In-place all-reduce (sum) of the first ``size`` elements of ``tensor`` over ``active_gpus`` (default: all),
        asynchronous on the current stream; returns the tensor (reference: /root/reference/adapcc.py:48-50)."""
        return cls._comm().all_reduce(tensor, size, chunk_bytes, active_gpus)

    @classmethod
    def reduce(cls, tensor, size=None, chunk_bytes=None, active_gpus=None):
        """In-place reduce: every strategy tree's root ends up with the sum of its slice (direct algorithms: rank 0
        holds the whole result) — reference semantics, /root/reference/adapcc.py:52-54."""
        return cls._comm().reduce(tensor, size, chunk_bytes, active_gpus)

    @classmethod
    def boardcast(cls, tensor, size=None, chunk_bytes=None):
        """Broadcast (the reference's spelling): every rank receives the roots' data (reference: /root/reference/adapcc.py:56-58)."""
        return cls._comm().boardcast(tensor, size, chunk_bytes)

    @classmethod
    def alltoall(cls, tensor, size=None, chunk_bytes=None):
        """The reference declares this primitive but forwards to a method that does not exist
        (/root/reference/adapcc.py:59-61, latent AttributeError). Here it is a real dense
        all-to-all of equal splits, carried by the expert-parallel dispatch kernels."""
        from .parallel.alltoall import all_to_all_single

        return all_to_all_single(cls._comm(), tensor, size)

    @classmethod
    def reducescatter(cls, tensor, size=None, chunk_bytes=None, op="sum"):
        """In-place reduce-scatter (primitive id 5; the reference only declares the id, /root/reference/commu.py:19-26):
        returns ``(lo, hi)``, the element range of ``tensor`` holding this rank's reduced shard. GPU: the direct reduce
        kernel with root = self over peer memory (half the bytes of an all-reduce)."""
        return cls._comm().reduce_scatter(tensor, size, op)

    @classmethod
    def allgather(cls, tensor, size=None, chunk_bytes=None):
        """In-place all-gather (primitive id 3), the inverse of :meth:`reducescatter`: every rank contributes its shard
        of ``tensor`` and ends up with all of them. GPU: one multicast-store broadcast per shard."""
        return cls._comm().all_gather(tensor, size)

    @classmethod
    def reconstruct_topology(cls, args, prim):
        """Re-run init (re-profile / re-synthesise per ``args.entry_point``) and rebuild the context for ``prim``; call
        it every ``profile_freq`` steps from the training loop (reference: /root/reference/adapcc.py:64-68)."""
        old = cls.communicator
        old.exit_threads(prim)
        native = old.clear(keep_native=True)           # DDP buckets / flat gradients may live in its heap
        cls.init(args, cls.local_rank, cls.world_rank, cls.world_size, _native=native)
        old._successor = cls.communicator              # hooks registered on the old object follow
        # the DDP buckets did not change: keep what the hook learned about them at step 1 (sizes, relay scratch), or a
        # relay of the new communicator would not know how many ops to mirror
        cls.communicator.bucket_info = list(old.bucket_info)
        cls.communicator.relay_buffer = list(old.relay_buffer)
        cls.setup(prim)

    @classmethod
    def set_profile_freq(cls, freq):
        """Steps between ``reconstruct_topology`` calls (read by the training loop)."""
        cls.profile_freq = freq

    @classmethod
    def clear(cls, prim):
        """Tear down the context for ``prim`` and the control plane (collective; reference: /root/reference/adapcc.py:74-76)."""
        cls._comm().exit_threads(prim)
        cls._comm().clear()


def _main():
    """Primitive benchmark template — the ``__main__`` block of /root/reference/adapcc.py:81-117 (same flags, same
    ``ones(16) * i`` all-reduce with 8-byte chunks, printed per rank), under torchrun instead of mpirun and with
    ``--backend gloo`` for boxes without a GPU.

        torchrun --nproc-per-node 8 -m adapcc_b200.adapcc --entry_point -1 --strategy_file strategy/8.xml
    """
    import argparse
    import os
    import sys

    import torch
    import torch.distributed as dist

    parser = argparse.ArgumentParser(description="AdapCC primitive benchmark template")
    parser.add_argument("--port", type=str, default="5000")
    parser.add_argument("--strategy_file", type=str, default="./strategy/strategy.xml")
    parser.add_argument("--logical_graph", type=str, default="./topology/logical_graph.xml")
    parser.add_argument("--entry_point", type=int, default=-1)
    parser.add_argument("--parallel_degree", type=int, default=4)
    parser.add_argument("--profile_freq", type=int, default=500)
    parser.add_argument("--backend", type=str, default="nccl", choices=["nccl", "gloo"])
    args = parser.parse_args()

    env = os.environ
    local = int(env.get("LOCAL_RANK", env.get("OMPI_COMM_WORLD_LOCAL_RANK", 0)))
    world = int(env.get("WORLD_SIZE", env.get("OMPI_COMM_WORLD_SIZE", 1)))
    rank = int(env.get("RANK", env.get("OMPI_COMM_WORLD_RANK", 0)))
    cuda = args.backend == "nccl" and torch.cuda.is_available()
    if not cuda:
        args.backend = "gloo"
    env.setdefault("MASTER_ADDR", "127.0.0.1")
    env.setdefault("MASTER_PORT", "29400")
    if cuda:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    AdapCC.init(args, local, rank, world)
    for prim, name in ((ALLREDUCE, "allreduce"), (REDUCE, "reduce"), (BOARDCAST, "boardcast")):
        AdapCC.setup(prim)
        for i in range(1, 3):
            tensor = torch.ones(16, dtype=torch.float32) * i
            if cuda:
                tensor = tensor.to(local)
            size, chunk_bytes, active = int(tensor.numel()), 8, list(range(world))
            if prim == ALLREDUCE:
                out = AdapCC.communicator.all_reduce(tensor, size, chunk_bytes, active)
            elif prim == REDUCE:
                out = AdapCC.communicator.reduce(tensor, size, chunk_bytes, active)
            else:
                out = AdapCC.communicator.boardcast(tensor, size, chunk_bytes)
            if cuda:
                AdapCC.communicator.synchronize()
            sys.stdout.write("rank %d %s: %s\n" % (rank, name, out.cpu().numpy().tolist()))   # one write: ranks share a pipe
            sys.stdout.flush()
        AdapCC.communicator.exit_threads(prim)
    AdapCC.communicator.clear()
    dist.destroy_process_group()


if __name__ == "__main__":
    _main()
