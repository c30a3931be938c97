"""Single-process (gloo, world 1) walk through API misuse: calls before init, init twice, setup twice, unknown primitive ids,
clear twice, use after clear. Prints one line per case (tests/test_workflow_cpu.py::test_api_misuse_is_handled)."""
import os, sys, tempfile
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from types import SimpleNamespace
from adapcc_b200 import ALLREDUCE, REDUCE, BOARDCAST
from adapcc_b200.adapcc import AdapCC
os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=sys.argv[1], RANK="0", WORLD_SIZE="1")
dist.init_process_group("gloo", rank=0, world_size=1)
tmp = tempfile.mkdtemp()
args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "s.xml"), logical_graph=os.path.join(tmp, "lg.xml"), entry_point=-1,
                       parallel_degree=2, profile_freq=0, backend="gloo", work_dir=tmp, coordinator_port=int(sys.argv[2]))
for label, fn in [
    ("use before init", lambda: AdapCC.allreduce(torch.ones(4))),
]:
    try: fn(); print(label, "-> no error")
    except Exception as e: print(label, "->", type(e).__name__, str(e)[:100])
AdapCC.init(args, 0, 0, 1)
for label, fn in [
    ("allreduce before setup", lambda: AdapCC.allreduce(torch.ones(4))),
    ("double init", lambda: AdapCC.init(args, 0, 0, 1)),
    ("setup twice", lambda: (AdapCC.setup(ALLREDUCE), AdapCC.setup(ALLREDUCE))),
    ("setup other prims", lambda: (AdapCC.setup(REDUCE), AdapCC.setup(BOARDCAST))),
    ("reduce", lambda: print("  reduce ->", AdapCC.communicator.reduce(torch.ones(4)).tolist())),
    ("unknown prim", lambda: AdapCC.setup(42)),
    ("clear unknown prim", lambda: AdapCC.communicator.exit_threads(42)),
    ("clear", lambda: AdapCC.clear(ALLREDUCE)),
    ("clear twice", lambda: AdapCC.clear(ALLREDUCE)),
    ("use after clear", lambda: AdapCC.allreduce(torch.ones(4))),
]:
    try: fn(); print(label, "-> ok")
    except Exception as e: print(label, "->", type(e).__name__, str(e)[:120])
os._exit(0)
