"""Property tests over RANDOM spanning trees (not just the chain / binary / star shapes the synthesizers emit):
* the CPU executor (the oracle of the GPU tree kernel) against plain sums for all-reduce over random active subsets in
  both relay modes, reduce-to-roots and broadcast-from-roots;
* the Python and the native (csrc/schedule.cpp) tree-role derivations against each other."""
import os
import random
import sys
import traceback

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import free_port

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_tree(nodes, rng):
    from adapcc_b200.strategy.trees import Tree
    order = nodes[:]; rng.shuffle(order)
    t = Tree(root=order[0]); t.nodes = [order[0]]
    for x in order[1:]:
        p = rng.choice(t.nodes)
        t.parent[x] = p; t.children.setdefault(p, []).append(x); t.nodes.append(x)
    # DFS order for nodes
    out = []
    def dfs(a):
        out.append(a)
        for c in t.kids(a): dfs(c)
    dfs(t.root); t.nodes = out
    for x in out: t.ip[x] = "127.0.0.1"
    return t

def _prop_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from adapcc_b200.strategy.trees import Strategy
    from adapcc_b200.strategy.cpu_executor import tree_collective_cpu
    from adapcc_b200.strategy import slice_bounds
    from adapcc_b200.constants import ALLREDUCE, REDUCE, BOARDCAST
    rng = random.Random(1234)
    problems = []
    try:
        for case in range(60):
            ntrees = rng.randint(1, 3)
            s = Strategy([random_tree(list(range(world)), rng) for _ in range(ntrees)], {})
            s.validate(world)
            n = rng.choice([1, 5, 64, 1001])
            chunk = rng.choice([8, 64, 4096])
            op = rng.choice(["sum", "avg", "max"])
            mode = rng.choice([0, 1])
            g = torch.Generator().manual_seed(case)
            data = torch.randn(world, n, generator=g)
            # allreduce over a random active subset
            act = sorted(rng.sample(range(world), rng.randint(1, world)))
            t = data[rank].clone()
            tree_collective_cpu(ALLREDUCE, t, s, rank, world, active=act, op=op, chunk_bytes=chunk, relay_mode=mode)
            sub = data[act]
            want = {"sum": sub.sum(0), "avg": sub.mean(0), "max": sub.max(0).values}[op] if rank in act else data[rank]
            if not torch.allclose(t, want, atol=1e-5): problems.append(f"case {case} allreduce act={act} mode={mode} op={op} n={n}")
            # reduce, all active: slice k lands on tree k's root
            t = data[rank].clone()
            tree_collective_cpu(REDUCE, t, s, rank, world, op="sum", chunk_bytes=chunk)
            b = slice_bounds(n, ntrees, 4)
            want = data[rank].clone()
            for k, tr in enumerate(s.trees):
                if tr.root == rank: want[b[k]:b[k+1]] = data.sum(0)[b[k]:b[k+1]]
            if not torch.allclose(t, want, atol=1e-5): problems.append(f"case {case} reduce n={n}")
            # broadcast, all active: slice k comes from tree k's root
            t = data[rank].clone()
            tree_collective_cpu(BOARDCAST, t, s, rank, world, chunk_bytes=chunk)
            want = torch.empty(n)
            for k, tr in enumerate(s.trees): want[b[k]:b[k+1]] = data[tr.root][b[k]:b[k+1]]
            if not torch.equal(t, want): problems.append(f"case {case} boardcast n={n}")
    except Exception:
        problems.append(traceback.format_exc()[-700:])
    q.put((rank, problems[:5]))
    dist.destroy_process_group()



def test_cpu_executor_on_random_trees_world6():
    world = 6
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_prop_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=20)
    assert sorted(r for r, _ in results) == list(range(world))
    assert all(not problems for _, problems in results), results


def test_python_and_native_tree_roles_agree_on_random_trees():
    from adapcc_b200.constants import ALLREDUCE, BOARDCAST, REDUCE
    from adapcc_b200.runtime.native import native_tree_role
    from adapcc_b200.strategy.relay import tree_role
    from adapcc_b200.strategy.trees import Strategy

    rng = random.Random(7)
    for _ in range(150):
        world = rng.choice([2, 3, 5, 8, 13, 16])
        s = Strategy([random_tree(list(range(world)), rng) for _ in range(rng.randint(1, 4))], {})
        xml = s.to_xml(compact=rng.random() < 0.5)
        for _ in range(6):
            act = sorted(rng.sample(range(world), rng.randint(1, world)))
            rank, ti = rng.randrange(world), rng.randrange(len(s.trees))
            prim, mode = rng.choice([ALLREDUCE, REDUCE, BOARDCAST]), rng.choice([0, 1])
            py = tree_role(s.trees[ti], rank, act, prim, mode)
            nat = native_tree_role(xml, world, ti, rank, act, prim, mode)
            assert (py.parent, py.flags, sorted(py.children)) == (nat["parent"], nat["flags"], sorted(nat["children"])), \
                (world, act, rank, ti, prim, mode)
