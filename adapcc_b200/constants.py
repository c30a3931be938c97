"""Primitive / dtype / op / algorithm ids shared by Python and the native runtime.

Primitive numbering is the reference's (/root/reference/commu.py:28-35,
/root/reference/csrc/include/trans.h:27-36).
"""
ALLREDUCE = 0
REDUCE = 1
BOARDCAST = 2
ALLGATHER = 3
ALLTOALL = 4
REDUCESCATTER = 5
DETECT = 6
PROFILE = 7

PRIMITIVE_NAMES = {ALLREDUCE: "allreduce", REDUCE: "reduce", BOARDCAST: "boardcast", ALLGATHER: "allgather",
                   ALLTOALL: "alltoall", REDUCESCATTER: "reducescatter", DETECT: "detect", PROFILE: "profile"}

DTYPE_IDS = {"float32": 0, "bfloat16": 1, "float16": 2}
OP_IDS = {"sum": 0, "avg": 1, "max": 2}
ALGO_IDS = {"auto": 0, "one_shot": 1, "two_shot": 2, "nvls": 3, "tree": 4}
ALGO_NAMES = {v: k for k, v in ALGO_IDS.items()}

RELAY_FORWARD = 0   # reference semantics: inactive ranks on a path forward data
RELAY_BYPASS = 1    # NVSwitch-aware: inactive ranks are contracted out of the trees

# TreeRoleFlags (csrc/common.h)
TR_HAS_LOCAL, TR_IN_REDUCE, TR_IN_BCAST, TR_WANT_RESULT, TR_PUBLISH, TR_PARENT_IS_ROOT = 1, 2, 4, 8, 16, 32
