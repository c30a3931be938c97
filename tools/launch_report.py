"""Summarise an ``ncu --metrics gpu__time_duration.sum --csv`` launch list into a markdown table of
per-kernel-family time shares for ONE training step (the last 1/N-th of the launches).
Usage: python tools/launch_report.py gpurun_out/launches.csv STEPS_IN_CAPTURE out.md "title"."""
import collections
import csv
import re
import sys


def main():
    path, steps, out, title = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    recs = []
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        v *= {"us": 1e3, "ms": 1e6, "s": 1e9}.get(row.get("Metric Unit", "ns"), 1)
        recs.append((row["Kernel Name"], v))
    per = len(recs) // steps
    last = recs[-per:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, v in last:
        k = re.sub(r"\s+", " ", name)
        k = re.sub(r"^void ", "", k)
        k = k.replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"<.*", "", k)[:80]
        if "adapcc" in name:
            k = "**" + k + "** (ours)"
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v for _, v in last)
    ours = sum(t for k, (c, t) in agg.items() if "(ours)" in k)
    rows = [f"# {title}", "",
            f"`ncu --metrics gpu__time_duration.sum --clock-control none` over the whole run ({len(recs)} launches, "
            f"{steps} steps); table = the last step ({per} launches, {tot / 1e6:.2f} ms of kernel time; ncu serialises "
            "kernels and runs them cache-cold, so compare SHARES, not absolutes).", "",
            f"Our kernels: {100 * ours / tot:.1f}% of the step's kernel time.", "",
            "| kernel family | launches | ms | share |", "|---|---|---|---|"]
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        rows.append(f"| {k} | {c} | {t / 1e6:.3f} | {100 * t / tot:.1f}% |")
    open(out, "w").write("\n".join(rows) + "\n")
    print("\n".join(rows[:16]))


if __name__ == "__main__":
    main()
