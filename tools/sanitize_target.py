"""Small single-GPU workload for compute-sanitizer (memcheck / racecheck): every hand-written kernel
runs once at a small size. `compute-sanitizer --tool memcheck python tools/sanitize_target.py`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adapcc_b200.constants import ALLREDUCE, BOARDCAST, REDUCE  # noqa: E402
from adapcc_b200.models.moe import MoEMLP  # noqa: E402
from adapcc_b200.ops import fused_adamw_, fused_ce_, sumsq_  # noqa: E402
from adapcc_b200.ops.layers import FusedLayerNorm, FusedLinear  # noqa: E402
from adapcc_b200.parallel.expert_parallel import ExpertExchange  # noqa: E402
from adapcc_b200.runtime.native import NativeComm  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
comm = NativeComm("san-%d" % os.getpid(), 0, 1, 0, staging_bytes=4 << 20, heap_bytes=32 << 20)
comm.set_tunable("force_kernel", 1)
for dtype, wire in ((torch.float32, None), (torch.float32, "bfloat16"), (torch.bfloat16, None)):
    x = torch.randn(70_001, device=dev).to(dtype)
    for algo in ["one_shot", "two_shot"] + (["nvls"] if comm.multicast else []):
        comm.all_reduce(x.clone(), op="avg", algo=algo, wire=wire)
    comm.set_tunable("pipe_min_bytes", 1 << 14)
    comm.set_tunable("pipe_piece_bytes", 1 << 14)
    comm.all_reduce(x.clone(), algo="two_shot", wire=wire)
    comm.set_tunable("pipe_min_bytes", 32 << 20)
    comm.reduce(x.clone(), root=0, algo="two_shot", wire=wire)
    comm.broadcast(x.clone(), root=0)
t = comm.symm_empty(33_000, torch.bfloat16)
t.normal_()
comm.all_reduce(t, algo="two_shot")
comm.all_to_all(torch.randn(4096, device=dev))
comm.load_strategy("<trees><root id='0' ip='a'/></trees>")
for prim in (ALLREDUCE, REDUCE, BOARDCAST):
    comm.tree_collective(prim, torch.randn(50_003, device=dev), chunk_bytes=4096)
comm.check()
n = 100_003 // 8 * 8
p = torch.randn(n, device=dev).bfloat16()
g = torch.randn(n, device=dev).bfloat16()
master, m, v, ss = p.float(), torch.zeros(n, device=dev), torch.zeros(n, device=dev), torch.zeros(1, device=dev)
sumsq_(g, ss)
fused_adamw_(p, g, master, m, v, lr=1e-3, step=1, max_norm=1.0, sumsq=ss)
fused_ce_((torch.randn(64, 50304, device=dev)).bfloat16(), torch.randint(0, 50262, (64,), device=dev), 50262)
ln = FusedLayerNorm(768).to(dev).bfloat16()
lin = FusedLinear(768, 2304).to(dev).bfloat16()
xx = torch.randn(513, 768, device=dev).bfloat16().requires_grad_(True)
lin(ln(xx)).float().pow(2).mean().backward()
ex = ExpertExchange(comm, 4, 64, 64)
moe = MoEMLP(4, 64, 128, top_k=2, exchange=ex).to(dev).bfloat16()
y = moe(torch.randn(96, 64, device=dev).bfloat16().requires_grad_(True))
y.float().sum().backward()
comm.check()
torch.cuda.synchronize()
comm.close()
print("sanitize target done")
