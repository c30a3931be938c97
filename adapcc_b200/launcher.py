"""Launcher — same CLI as /root/reference/launcher.py:19-86, torchrun instead of mpirun.

The reference composes ``mpirun -np N -H ips -mca pml ucx -x UCX_NET_DEVICES=... python <exec>
--port ... --entry_point ... --strategy_file ... --logical_graph ... --parallel_degree ...
--profile_freq ...``, writes ``topology/ip_table.txt`` (one line per rank) and scp's it to every
node. Here one ``torch.distributed.run`` agent is started per host (the local one directly, remote
ones through ssh), ranks get RANK/LOCAL_RANK/WORLD_SIZE from torchrun, and the same six flags are
forwarded to the exec file. ``--mpi-path`` / ``--net-device`` are accepted for CLI compatibility and
ignored (no MPI, no UCX on the data path).

    python -m adapcc_b200.launcher --num-process 8 --ips 127.0.0.1:8 --master 127.0.0.1 \
        --exec-file train_ddp.py --entry_point 7 --strategy_file strategy/8.xml
"""
from __future__ import annotations

import argparse
import os
import shlex
import subprocess
import sys
from typing import List, Tuple

from .dispatcher import Dispatcher
from .topology import write_ip_table


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--num-process", type=int, default=4)
    p.add_argument("--ips", type=str, default="127.0.0.1:4", help="host:ranks[,host:ranks...]")
    p.add_argument("--master", type=str, default="127.0.0.1")
    p.add_argument("--master-port", type=int, default=1234)
    p.add_argument("--mpi-path", type=str, default="", help="ignored (reference compatibility)")
    p.add_argument("--net-device", type=str, default="", help="ignored (reference compatibility)")
    p.add_argument("--exec-file", type=str, default="train_ddp.py")
    p.add_argument("--socket_port", type=str, default="5000")
    p.add_argument("--entry_point", type=int, default=-1, help="6:detect, 7:profile, other None")
    p.add_argument("--strategy_file", type=str, default="./strategy/strategy_test.xml")
    p.add_argument("--logical_graph", type=str, default="./topology/logical_graph_test.xml")
    p.add_argument("--parallel_degree", type=int, default=4)
    p.add_argument("--profile_freq", type=int, default=500)
    p.add_argument("--dry-run", action="store_true")
    p.add_argument("exec_args", nargs=argparse.REMAINDER, help="extra args for the exec file (after --)")
    return p


def parse_hosts(ips: str) -> List[Tuple[str, int]]:
    out = []
    for item in ips.split(","):
        item = item.strip()
        if not item:
            continue
        host, _, n = item.partition(":")
        out.append((host, int(n or 1)))
    return out


def ip_table(hosts: List[Tuple[str, int]]) -> List[str]:
    return [h for h, n in hosts for _ in range(n)]


def exec_flags(a) -> List[str]:
    return [f"--port={a.socket_port}", f"--entry_point={a.entry_point}", f"--strategy_file={a.strategy_file}",
            f"--logical_graph={a.logical_graph}", f"--parallel_degree={a.parallel_degree}",
            f"--profile_freq={a.profile_freq}"]


def commands(a) -> List[Tuple[str, List[str]]]:
    hosts = parse_hosts(a.ips)
    total = sum(n for _, n in hosts)
    if total != a.num_process:
        raise SystemExit(f"--num-process {a.num_process} != ranks listed in --ips ({total})")
    extra = [x for x in a.exec_args if x != "--"]
    cmds = []
    for node_rank, (host, n) in enumerate(hosts):
        cmd = [sys.executable, "-m", "torch.distributed.run", f"--nnodes={len(hosts)}", f"--node-rank={node_rank}",
               f"--nproc-per-node={n}", "--master-addr", a.master, "--master-port", str(a.master_port),
               a.exec_file, *exec_flags(a), *extra]
        cmds.append((host, cmd))
    return cmds


def main(argv=None) -> int:
    a = build_parser().parse_args(argv)
    hosts = parse_hosts(a.ips)
    table = ip_table(hosts)
    work = os.getcwd()
    path = os.path.join(work, "topology", "ip_table.txt")
    write_ip_table(path, table)
    Dispatcher([h for h, _ in hosts], dry_run=a.dry_run).dispatch_ip_table(path, os.path.join(work, "topology"))
    procs = []
    for host, cmd in commands(a):
        local = host in ("127.0.0.1", "localhost", a.master) and len(hosts) == 1 or host in ("127.0.0.1", "localhost")
        full = cmd if local else ["ssh", host, f"cd {shlex.quote(work)} && " + " ".join(shlex.quote(c) for c in cmd)]
        print("[launcher]", " ".join(full), flush=True)
        if not a.dry_run:
            procs.append(subprocess.Popen(full))
    rc = 0
    for p in procs:
        rc = p.wait() or rc
    return rc


if __name__ == "__main__":
    sys.exit(main())
