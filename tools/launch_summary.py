"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown share table.
    python tools/launch_summary.py gpurun_out/launches.csv "title" "command" > profiles/x.md"""
import csv
import sys
from collections import defaultdict


def main():
    path, title, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    agg = defaultdict(lambda: [0, 0.0])
    with open(path) as f:
        rows = [r for r in csv.reader(f) if len(r) > 14]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rows[1:]:
        name = r[ki].replace("vpk::<unnamed>::", "vpk::")
        name = name.split("(")[0]
        a = agg[name]
        a[0] += 1
        a[1] += float(r[vi].replace(",", "")) * 1e-6
    total = sum(a[1] for a in agg.values())
    print(f"# {title}\n")
    print(f"Command: `{cmd}`\nPer-launch times are cold-cache and serialised: compare SHARES.\n")
    print("| kernel | launches | total ms | share |\n|---|---:|---:|---:|")
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name[:110]}` | {n} | {ms:.2f} | {100 * ms / total:.1f}% |")
    print(f"\nTotal {total:.1f} ms over {sum(a[0] for a in agg.values())} launches.")


if __name__ == "__main__":
    main()
