"""torchrun worker (gloo): the same 30 SGD steps of an MLP on synthetic clusters with (a) stock DDP gradient averaging and
(b) DDP + the AdapCC comm hook; prints the largest parameter difference and both accuracies
(tests/test_workflow_cpu.py::test_hook_training_matches_stock_ddp_on_cpu)."""
import os
import sys
from types import SimpleNamespace

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adapcc_b200 import ALLREDUCE  # noqa: E402
from adapcc_b200.adapcc import AdapCC  # noqa: E402


def data(rank, step, n=64):
    g = torch.Generator().manual_seed(1000 * rank + step)
    y = torch.randint(0, 4, (n,), generator=g)
    centers = torch.tensor([[2.0, 0.0], [-2.0, 0.0], [0.0, 2.0], [0.0, -2.0]])
    x = centers[y] + 0.5 * torch.randn(n, 2, generator=g)
    return torch.cat([x, x ** 2], 1), y


def train(use_hook, comm, rank):
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(4, 32), torch.nn.Tanh(), torch.nn.Linear(32, 4))
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    if use_hook:
        ddp.register_comm_hook(None, comm.cuda_allreduce_hook)
    opt = torch.optim.SGD(ddp.parameters(), lr=0.2)
    for step in range(30):
        if use_hook:
            comm.update_relay(step)
        x, y = data(rank, step)
        loss = torch.nn.functional.cross_entropy(ddp(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    x, y = data(99, 0, 512)
    acc = (model(x).argmax(1) == y).float().mean().item()
    return torch.cat([p.detach().flatten() for p in model.parameters()]), acc


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tmp = sys.argv[1]
    args = SimpleNamespace(port=5000, strategy_file=os.path.join(tmp, "s.xml"), logical_graph=os.path.join(tmp, "lg.xml"),
                           entry_point=-1, parallel_degree=2, profile_freq=0, backend="gloo", work_dir=tmp,
                           coordinator_port=int(sys.argv[2]), relay_control=False)
    AdapCC.init(args, rank, rank, world)
    AdapCC.setup(ALLREDUCE)
    w_ref, acc_ref = train(False, None, rank)
    w_hook, acc_hook = train(True, AdapCC.communicator, rank)
    diff = (w_ref - w_hook).abs().max().item()
    sys.stdout.write(f"[rank {rank}] max param diff {diff:.3e} acc stock {acc_ref:.3f} acc hook {acc_hook:.3f}\n")
    sys.stdout.flush()
    AdapCC.clear(ALLREDUCE)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
