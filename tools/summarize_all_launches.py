"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list divided by the number of steps run.
usage: python tools/summarize_all_launches.py launches.csv n_steps > profiles/rNN_launches_x.md"""
import collections, csv, re, sys

path, nsteps = sys.argv[1], int(sys.argv[2])
rows = list(csv.reader(open(path)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[h], rows[h + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "s": 1e3}.get(r[ui], 1e-6)
    name = re.sub(r"^void |tapclip::|\(anonymous namespace\)::|unnamed>::", "", r[ki]).split("(")[0][:90]
    tot[name] += v / nsteps
    cnt[name] += 1 / nsteps
total = sum(tot.values())
print(f"# ncu launch list — per step ({sum(cnt.values()):.0f} launches, {total:.2f} ms serialised; {nsteps} steps averaged)\n")
print("| kernel | launches | ms | share |\n|---|---:|---:|---:|")
for k, v in tot.most_common():
    print(f"| `{k}` | {cnt[k]:.0f} | {v:.3f} | {100 * v / total:.1f}% |")
