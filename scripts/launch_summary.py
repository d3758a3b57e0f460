"""Markdown summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`):
    python scripts/launch_summary.py launches.csv [first_kernel_substring [period]]
With a kernel-name substring, only one period (from its `period`-th occurrence, default the first, to just before the
next) is summarised."""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr, out = None, []
    for r in rows:
        if r and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] != "gpu__time_duration.sum":
                continue
            n = re.sub(r"^void ", "", d["Kernel Name"])
            n = re.sub(r"\(.*", "", n).replace("mcedm::", "")
            v = float(d["Metric Value"].replace(",", ""))
            v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(d["Metric Unit"], v)
            out.append((n, v))
    return out


def main():
    data = load(sys.argv[1])
    if len(sys.argv) > 2:
        idx = [i for i, (n, _) in enumerate(data) if sys.argv[2] in n]
        k = int(sys.argv[3]) if len(sys.argv) > 3 else 0
        if len(idx) >= k + 2:
            data = data[idx[k]:idx[k + 1]]
    agg = collections.OrderedDict()
    for n, v in data:
        a = agg.setdefault(n[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{len(data)} launches, {tot:.1f} us in kernels\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {t:.1f} | {100 * t / tot:.1f} % |")


if __name__ == "__main__":
    main()
