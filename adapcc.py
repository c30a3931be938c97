"""Drop-in import shim: scripts written for the reference do ``from adapcc import *`` (train_ddp.py:23 there) and get
``AdapCC`` plus the primitive ids. The implementation lives in the ``adapcc_b200`` package."""
from adapcc_b200.adapcc import AdapCC  # noqa: F401
from adapcc_b200.constants import (ALLGATHER, ALLREDUCE, ALLTOALL, BOARDCAST, DETECT, PROFILE, REDUCE,  # noqa: F401
                                   REDUCESCATTER)

__all__ = ["AdapCC", "ALLREDUCE", "REDUCE", "BOARDCAST", "ALLGATHER", "ALLTOALL", "REDUCESCATTER", "DETECT", "PROFILE"]

if __name__ == "__main__":          # ``python adapcc.py ...`` = the primitive benchmark template
    from adapcc_b200.adapcc import _main

    _main()
