"""File plane: moves the ip table, detected topologies, profiles and strategies between nodes.

The reference shells out to ``scp`` for every file and every host, including the local one
(/root/reference/dispatcher.py:7-54). On a single NVSwitch box (and on any shared filesystem) the
files are already where they need to be, so local addresses are skipped and a barrier is all that
is required; remote hosts still go through ``scp`` so multi-server deployments keep working.
"""
from __future__ import annotations

import glob
import os
import shutil
import socket
import subprocess
from typing import List, Sequence

_LOCAL = {"127.0.0.1", "localhost", "::1", ""}


def _local_names() -> set:
    names = set(_LOCAL)
    try:
        host = socket.gethostname()
        names.add(host)
        names.update(socket.gethostbyname_ex(host)[2])
    except OSError:
        pass
    return names


class Dispatcher:
    """File plane between the servers of a job: ships detect files, profile records, the logical graph and the strategy
    XML (local hosts: copy, remote hosts: ``scp``) — /root/reference/dispatcher.py:1-54."""

    def __init__(self, ip_table: Sequence[str], scp: str = "scp", dry_run: bool = False):
        # ADAPCC_SHARED_FS=1: every host sees the same directory (one box, NFS, emulated multi-server
        # runs) -> plain local copies. Also the fallback when no scp binary exists.
        env = os.environ.get("ADAPCC_SHARED_FS", "auto")
        self.shared_fs = env == "1" or (env == "auto" and shutil.which(scp) is None)
        self.ip_table = list(ip_table)
        self.ip_dict = {}
        self.scp = scp
        self.dry_run = dry_run
        self.log: List[str] = []
        self.init_ip_dict()

    def init_ip_dict(self):
        self.ip_dict = {ip: True for ip in self.ip_table}

    def renew_ip_table(self, ip_table):
        self.ip_table = list(ip_table)
        self.init_ip_dict()

    # -- transport ---------------------------------------------------------------------------
    def _send(self, src_pattern: str, ip: str, dst_path: str) -> None:
        files = glob.glob(src_pattern) or [src_pattern]
        if self.shared_fs or ip in _local_names():
            for f in files:
                if not os.path.exists(f):
                    continue
                os.makedirs(dst_path, exist_ok=True)
                dst = os.path.join(dst_path, os.path.basename(f))
                if os.path.abspath(f) != os.path.abspath(dst):
                    shutil.copy2(f, dst)
            self.log.append(f"local {src_pattern} -> {dst_path}")
            return
        cmd = [self.scp, *files, f"{ip}:{dst_path}"]
        self.log.append(" ".join(cmd))
        if not self.dry_run:
            subprocess.run(cmd, check=False)

    # -- reference API ---------------------------------------------------------------------
    def dispatch_ip_table(self, src_file, dst_path):
        """master node sends the ip table to every node"""
        for ip in self.ip_dict:
            self._send(src_file, ip, dst_path)

    def dispatch_detected_topo(self, src_file, dst_path):
        """each local rank 0 sends its detect XML to every node"""
        for ip in self.ip_dict:
            self._send(src_file, ip, dst_path)

    def send_profiled_topo(self, src_file, dst_path):
        """each local rank 0 sends its profile to the master (world rank 0's node)"""
        if self.ip_table:
            self._send(src_file, self.ip_table[0], dst_path)

    def dispatch_strategy(self, src_file, dst_path):
        """master sends the synthesised strategy to every node"""
        for ip in self.ip_dict:
            self._send(src_file, ip, dst_path)
