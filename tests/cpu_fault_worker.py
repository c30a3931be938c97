"""torchrun worker (gloo): DDP training with the AdapCC hook in which the last rank dies at step 2; the survivors must keep
training on the active subset (tests/test_workflow_cpu.py::test_training_survives_a_dead_worker_on_cpu)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.distributed as dist
from types import SimpleNamespace
from adapcc_b200 import ALLREDUCE
from adapcc_b200.adapcc import AdapCC
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
tmp = sys.argv[1]
args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "s.xml"), logical_graph=os.path.join(tmp, "lg.xml"), entry_point=-1,
                       parallel_degree=2, profile_freq=0, backend="gloo", work_dir=tmp, coordinator_port=int(sys.argv[2]),
                       relay_threshold=0.05, fault_tolerant_time=1.0, relay_control=True)
AdapCC.init(args, rank, rank, world); AdapCC.setup(ALLREDUCE)
comm = AdapCC.communicator
model = torch.nn.Linear(16, 4)
ddp = torch.nn.parallel.DistributedDataParallel(model)
ddp.register_comm_hook(None, comm.cuda_allreduce_hook)
opt = torch.optim.SGD(ddp.parameters(), lr=0.1)
for step in range(6):
    if rank == world - 1 and step == 2:
        print(f"[rank {rank}] dying at step {step}", flush=True)
        os._exit(0)
    comm.update_relay(step)
    t0 = time.time()
    loss = ddp(torch.randn(8, 16)).pow(2).mean()
    opt.zero_grad(); loss.backward(); opt.step()
    print(f"[rank {rank}] step {step} active {comm.active_gpus} {1e3*(time.time()-t0):.0f} ms", flush=True)
print(f"[rank {rank}] finished", flush=True)
os._exit(0)
