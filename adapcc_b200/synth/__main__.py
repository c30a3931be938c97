"""Stand-alone strategy synthesis (the counterpart of running the reference's ``gurobi/`` package by hand):

    python -m adapcc_b200.synth --shape 4-4 --policy auto --degree 4 --size 100e6 --out strategy/my.xml
    python -m adapcc_b200.synth --profile-dir topology/ --policy milp --out strategy/measured.xml

``--shape`` describes servers by GPU count (``8``, ``4-4``, ``4-2`` ...) with nominal NVLink / network figures;
``--profile-dir`` uses measured ``topo_profile_<rank>`` files plus ``ip_table.txt`` instead. Prints the trees, the chunk
size and the cost model's estimate next to the direct algorithms.
"""
import argparse
import os

from . import LinkModel, Synthesizer, direct_times, strategy_time
from ..strategy import Strategy
from ..topology import read_ip_table, read_profiles


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m adapcc_b200.synth", description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--shape", default="8", help="GPUs per server, dash separated (default: one 8-GPU server)")
    ap.add_argument("--profile-dir", default="", help="directory with ip_table.txt and topo_profile_<rank> files")
    ap.add_argument("--policy", default="auto", choices=["par-trees", "milp", "gurobi", "auto"])
    ap.add_argument("--intra", default="chain", choices=["chain", "binary", "star"])
    ap.add_argument("--degree", type=int, default=4, help="parallel transmissions (trees)")
    ap.add_argument("--size", type=float, default=25e6, help="message size in fp32 elements")
    ap.add_argument("--intra-gbs", type=float, default=700.0)
    ap.add_argument("--inter-gbs", type=float, default=45.0)
    ap.add_argument("--out", default="strategy/synth.xml")
    a = ap.parse_args(argv)

    if a.profile_dir:
        import glob
        import re
        import statistics

        files = [f for f in glob.glob(os.path.join(a.profile_dir, "topo_profile_*")) if re.search(r"topo_profile_\d+$", f)]
        if not files:
            ap.error(f"no topo_profile_<rank> files in {a.profile_dir}")
        table = os.path.join(a.profile_dir, "ip_table.txt")
        if os.path.exists(table):
            ips = read_ip_table(table)
        else:                                   # a single box: the world is whatever the records mention
            world = 1 + max(int(x) for f in files for ln in open(f) for x in ln.split(",")[:2] if x.strip().isdigit())
            ips = ["127.0.0.1"] * world
        lat, bw, _ = read_profiles(files, len(ips))
        # ranks whose own file is missing: use the reverse direction, then the median of what was measured
        known_bw = [v for row in bw for v in row if v > 0]
        known_lat = [v for row in lat for v in row if v > 0]
        for i in range(len(ips)):
            for j in range(len(ips)):
                if i != j and bw[i][j] <= 0:
                    bw[i][j] = bw[j][i] if bw[j][i] > 0 else statistics.median(known_bw)
                    lat[i][j] = lat[j][i] if lat[j][i] > 0 else statistics.median(known_lat)
    else:
        shape = [int(x) for x in a.shape.split("-")]
        ips = [f"node{s + 1}" for s, n in enumerate(shape) for _ in range(n)]
        w = len(ips)
        bw = [[0.0 if i == j else (a.intra_gbs if ips[i] == ips[j] else a.inter_gbs) for j in range(w)] for i in range(w)]
        lat = [[0.0 if i == j else (2.0 if ips[i] == ips[j] else 12.0) for j in range(w)] for i in range(w)]
    syn = Synthesizer(a.out, ip_table=ips, parallel_degree=a.degree, size=int(a.size), bandwidth_graph=bw,
                      latency_graph=lat, policy=a.policy, intra_policy=a.intra)
    chunk = syn.generate_strategy("reduce")
    s = Strategy.from_file(a.out, len(ips))
    s.validate(len(ips))
    lm = LinkModel(lat, bw)
    nbytes = a.size * 4
    print(open(a.out).read())
    print(f"world {len(ips)}, {len(s.trees)} trees, chunk {chunk} bytes, report {syn.last_report}")
    if len(set(ips)) == 1:
        # the same per-message algorithm plan the workflow writes (commu.py::_synthesis_strategy): size bands per algorithm
        from .plan import build_plan

        plan = build_plan(lm, s, nvls=len(ips) > 2, ll=len(ips) > 1)
        s.attrs.update(plan.to_attrs())
        s.save(a.out, compact=True)
        print(f"algorithm plan (written into {a.out}): staged [{plan.to_attrs()['bands']}]  heap [{plan.to_attrs()['bands_zc']}]")
    print(f"cost model for {nbytes / 1e6:.1f} MB: trees {strategy_time(s, lm, nbytes, chunk) * 1e6:.1f} us; "
          + ", ".join(f"{k} {v * 1e6:.1f} us" for k, v in direct_times(lm, nbytes).items())
          + ("  (direct algorithms assume one NVSwitch domain)" if len(set(ips)) > 1 else ""))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
