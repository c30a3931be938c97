"""Topology files: ip table, detect XML, logical graph XML, profile dumps.

* ``ip_table.txt`` — line *i* = host of world rank *i* (/root/reference/launcher.py:64-79).
* detect XML — per server, written by the native detector (csrc/detect.cpp); the reference schema
  ``<cpu><pcie>[<nic/>]<gpu id/>…`` is accepted too (/root/reference/csrc/detect.cu:366-424).
* logical graph — ``<graph><server id ip><nic id><gpu id/>…`` with gpu id = world rank
  (/root/reference/commu.py:207-244, /root/reference/csrc/profile.cu:56-90).
* profile dump — text lines ``src, dst, type, value``; type 1 = bandwidth GB/s, type 0 = latency us,
  every (i, j) present, zeros where unmeasured (/root/reference/csrc/profile.cu:336-357). We add
  type 2 = peer write bandwidth and type 3 = NVLS (multimem) bandwidth; readers that only know
  types 0/1 (the reference's parser treats every non-zero type as bandwidth) must skip them, so
  they are written to a side file ``topo_profile_<rank>.ext``.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence, Tuple

from .strategy import xmlio


# ---- ip table -------------------------------------------------------------------------------
def read_ip_table(path) -> List[str]:
    with open(path, "r") as f:
        return [ln.strip() for ln in f.read().splitlines() if ln.strip()]


def write_ip_table(path, ips: Sequence[str]) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        f.write("".join(f"{ip}\n" for ip in ips))


def local_rank0_list(ip_table: Sequence[str]) -> List[int]:
    out, seen = [], set()
    for r, ip in enumerate(ip_table):
        if ip not in seen:
            out.append(r)
            seen.add(ip)
    return out


def server_groups(ip_table: Sequence[str]) -> Dict[int, List[int]]:
    """local-rank-0 world rank -> world ranks of that server (contiguous, as the reference assumes)."""
    groups: Dict[int, List[int]] = {}
    for r0 in local_rank0_list(ip_table):
        g, i = [], r0
        while i < len(ip_table) and ip_table[i] == ip_table[r0]:
            g.append(i)
            i += 1
        groups[r0] = g
    return groups


# ---- detect XML -> logical graph --------------------------------------------------------------
def count_gpus_nics(detect_doc: xmlio.Node) -> Tuple[int, int]:
    gpus = sum(1 for n in detect_doc.iter() if n.tag == "gpu")
    nics = sum(1 for n in detect_doc.iter() if n.tag == "nic")
    return gpus, nics


def build_logical_graph(detect_files: Sequence[str], ips: Sequence[str], first_ranks: Sequence[int],
                        gpus_per_server: Optional[Sequence[int]] = None) -> xmlio.Node:
    """Gather the per-server detect XMLs into one ``<graph>`` (reference:
    ``_gather_detect_graph``, /root/reference/commu.py:207-244): GPUs are split evenly over the
    server's NICs; gpu ids are world ranks."""
    graph = xmlio.Node("graph", {"version": "adapcc-b200"})
    for sid, (path, ip, r0) in enumerate(zip(detect_files, ips, first_ranks)):
        doc = xmlio.parse_file(path)
        g, n = count_gpus_nics(doc)
        if gpus_per_server is not None:
            g = gpus_per_server[sid]            # ranks launched on this server (<= GPUs detected)
        n = max(1, n)
        n = min(n, max(1, g))
        per = max(1, g // n)
        server = xmlio.Node("server", {"id": str(sid), "ip": ip})
        extra = {k: v for k, v in doc.attrs.items() if k in ("nvml",)}
        server.attrs.update(extra)
        nvs = [x for x in doc.iter() if x.tag == "gpu"]
        if nvs:
            server.attrs["nvlinks"] = nvs[0].attrs.get("nvlinks", "0")
            server.attrs["nvswitch_links"] = nvs[0].attrs.get("nvswitch_links", "0")
            server.attrs["multicast"] = nvs[0].attrs.get("multicast", "0")
        for k in range(n):
            nic = xmlio.Node("nic", {"id": str(k)})
            hi = g if k == n - 1 else per * (k + 1)
            for local in range(per * k, hi):
                nic.children.append(xmlio.Node("gpu", {"id": str(r0 + local)}))
            server.children.append(nic)
        graph.children.append(server)
    return graph


def logical_graph_ranks(path) -> Dict[str, List[int]]:
    """server ip -> world ranks, from a logical graph file."""
    doc = xmlio.parse_file(path)
    out: Dict[str, List[int]] = {}
    for srv in doc.find_all("server"):
        out[srv.attrs.get("ip", "")] = [int(g.attrs["id"]) for g in srv.iter() if g.tag == "gpu"]
    return out


# ---- profile dumps --------------------------------------------------------------------------
def write_profile(path, src: int, world: int, lat_us: Sequence[float], bw_gbs: Sequence[float],
                  write_gbs: Optional[Sequence[float]] = None, nvls_gbs: float = 0.0) -> None:
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        for dst in range(world):
            f.write(f"{src}, {dst}, 0, {float(lat_us[dst]):.6f}\n")
            f.write(f"{src}, {dst}, 1, {float(bw_gbs[dst]):.6f}\n")
    with open(str(path) + ".ext", "w") as f:
        for dst in range(world):
            if write_gbs is not None:
                f.write(f"{src}, {dst}, 2, {float(write_gbs[dst]):.6f}\n")
        f.write(f"{src}, {src}, 3, {float(nvls_gbs):.6f}\n")


def read_profiles(paths: Sequence[str], world: int):
    """-> (latency_us[world][world], bandwidth_gbs[world][world], ext dict). Parser parity with
    ``_gather_topo_profile`` (/root/reference/commu.py:246-270)."""
    lat = [[0.0] * world for _ in range(world)]
    bw = [[0.0] * world for _ in range(world)]
    ext = {"write": [[0.0] * world for _ in range(world)], "nvls": [0.0] * world}
    for p in paths:
        if not os.path.exists(p):
            continue
        with open(p, "r") as f:
            for ln in f:
                e = [x.strip() for x in ln.strip().split(",")]
                if len(e) < 4:
                    continue
                s, d, ty, v = int(e[0]), int(e[1]), int(e[2]), float(e[3])
                if s >= world or d >= world:
                    continue
                if ty == 0:
                    lat[s][d] = v
                else:
                    bw[s][d] = v
        if os.path.exists(str(p) + ".ext"):
            with open(str(p) + ".ext", "r") as f:
                for ln in f:
                    e = [x.strip() for x in ln.strip().split(",")]
                    if len(e) < 4:
                        continue
                    s, d, ty, v = int(e[0]), int(e[1]), int(e[2]), float(e[3])
                    if s >= world or d >= world:
                        continue
                    if ty == 2:
                        ext["write"][s][d] = v
                    elif ty == 3:
                        ext["nvls"][s] = v
    return lat, bw, ext


def accumulated_bandwidth(bw) -> float:
    """Sum of pair bandwidths / 2 — what the reference feeds the coordinator's cost model
    (/root/reference/commu.py:266-269)."""
    return sum(sum(row) for row in bw) / 2.0
