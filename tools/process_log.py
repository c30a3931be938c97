"""Turn a training log into a one-number-per-line series: the ``Acc@1`` column of the progress lines
(/root/reference/models/image-classification/process_log.py → ``accuracy_*.txt`` / ``resnet18_*.txt``) or the
``mean gns:`` probe lines (process_gns.py → ``gns-split-all.txt``).

    python tools/process_log.py --metric acc1 nohup.out accuracy.txt
    python tools/process_log.py --metric gns  run.out   gns.txt
"""
import argparse
import re
import sys

PATTERNS = {
    "acc1": re.compile(r"Acc@1\s+([-+0-9.eE]+)"),          # current value of the meter: "Acc@1  12.50 ( 10.94)"
    "acc5": re.compile(r"Acc@5\s+([-+0-9.eE]+)"),
    "loss": re.compile(r"Loss\s+([-+0-9.eE]+)"),
    "gns": re.compile(r"mean gns:\s*([-+0-9.eE]+|nan|inf)"),
}


def extract(lines, metric):
    pat = PATTERNS[metric]
    out = []
    for line in lines:
        if metric != "gns" and line.lstrip().startswith("*"):   # validation summaries are not part of the training curve
            continue
        m = pat.search(line)
        if m:
            out.append(float(m.group(1)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("log")
    ap.add_argument("out", nargs="?", default="-")
    ap.add_argument("--metric", default="acc1", choices=sorted(PATTERNS))
    a = ap.parse_args()
    with open(a.log, errors="replace") as f:
        vals = extract(f, a.metric)
    dst = sys.stdout if a.out == "-" else open(a.out, "w")
    for v in vals:
        dst.write("%f\n" % v)
    if dst is not sys.stdout:
        dst.close()
        print(f"{len(vals)} values -> {a.out}")


if __name__ == "__main__":
    main()
