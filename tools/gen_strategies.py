"""Regenerates the shipped sample strategies (strategy/*.xml) and logical graphs (topology/*.xml)
with this repo's synthesizer — same file-name convention as the reference's hand-written samples
(server shape "4", "2-2", "4-4_1", ... /root/reference/strategy), same XML schemas, synthetic
addresses. Variants _1/_2/_3 are par-trees with chain / binary / star inside a server."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adapcc_b200.strategy import make_strategy  # noqa: E402
from adapcc_b200.synth import ParTrees, Solver  # noqa: E402
from adapcc_b200.topology import local_rank0_list  # noqa: E402


def ips_for(shape):
    out = []
    for s, n in enumerate(shape):
        out += [f"node{s + 1}"] * n
    return out


def link_model(ips, intra=700.0, inter=45.0):
    w = len(ips)
    bw = [[0.0 if i == j else (intra if ips[i] == ips[j] else inter) for j in range(w)] for i in range(w)]
    lat = [[0.0 if i == j else (2.0 if ips[i] == ips[j] else 12.0) for j in range(w)] for i in range(w)]
    return bw, lat


def main():
    shapes = {"4": [4], "8": [8], "2-2": [2, 2], "4-2": [4, 2], "3-3-3": [3, 3, 3], "4-4": [4, 4], "4-4-4": [4, 4, 4],
              "4-4-4-4": [4] * 4}
    for name, shape in shapes.items():
        ips = ips_for(shape)
        bw, lat = link_model(ips)
        variants = [("chain", "_1"), ("binary", "_2"), ("star", "_3")] if len(shape) > 1 and name.startswith("4-4") else [("binary" if len(shape) == 1 else "chain", "")]
        for intra, suffix in variants:
            s = ParTrees(intra).build(ips, local_rank0_list(ips), 4, bw, lat)
            s.attrs["chunk"] = str(4 << 20)
            s.save(os.path.join(ROOT, "strategy", f"{name}{suffix}.xml"), compact=True)
        if len(ips) <= 8:
            try:
                m = Solver(time_limit_s=5).solve(min(4, len(ips)), 100e6, bw, lat, ips)
                m.attrs["chunk"] = str(1 << 20)
                m.save(os.path.join(ROOT, "strategy", f"{name}_milp.xml"), compact=True)
            except Exception as e:  # noqa: BLE001
                print("milp skipped for", name, e)
        # logical graph sample: one server per line (<graph><server><nic><gpu/>... schema)
        lines = ['<?xml version="1.0" encoding="utf-8"?>', f'<graph version="adapcc-b200" shape="{name}">']
        r = 0
        for sid, n in enumerate(shape):
            gpus = "".join(f'<gpu id="{r + i}"/>' for i in range(n))
            lines.append(f'  <!-- server {sid}: world ranks {r}..{r + n - 1}, one NVLink/NVSwitch domain -->')
            lines.append(f'  <server ip="node{sid + 1}" id="{sid}" gpus="{n}"><nic id="0">{gpus}</nic></server>')
            r += n
        lines.append("</graph>")
        with open(os.path.join(ROOT, "topology", f"logical_graph_{name}.xml"), "w") as f:
            f.write("\n".join(lines) + "\n")
    # the tree of the reference's golden logs (0 <- 1 <- {2, 3}) as the default test strategy
    make_strategy(4, 2, "binary", ips_for([4])).save(os.path.join(ROOT, "strategy", "strategy_test.xml"), compact=True)
    with open(os.path.join(ROOT, "topology", "ip_table_example.txt"), "w") as f:
        f.write("".join(ip + "\n" for ip in ips_for([4, 4])))


if __name__ == "__main__":
    main()
