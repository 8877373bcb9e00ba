"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of ONE steady-state step.

usage: python tools/summarize_launches.py gpurun_out/launches.csv <step_index> > profiles/rNN_launches.md
A step is delimited by the adamw_kernel launch that ends it (train workload).  ncu times are cold-cache and
serialised: compare SHARES, not absolutes.
"""
import collections
import csv
import re
import sys


def main(path, step_index=3):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    launches = []
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
        launches.append((r[ki], v))
    ends = [i for i, (n, _) in enumerate(launches) if "adamw" in n]
    a, b = ends[step_index] + 1, ends[step_index + 1] + 1
    step = launches[a:b]
    total = sum(v for _, v in step)

    def short(n):
        n = re.sub(r"^void ", "", n)
        n = n.replace("tapclip::<unnamed>::", "")
        n = n[: n.rindex(">(") + 1] if ">(" in n else re.sub(r"\(.*", "", n)
        return n[:90]

    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in step:
        agg[short(n)][0] += 1
        agg[short(n)][1] += v
    print(f"# ncu launch list — one steady-state step ({len(step)} launches, {total / 1e3:.2f} ms serialised)\n")
    print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {c} | {t / 1e3:.3f} | {100 * t / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
