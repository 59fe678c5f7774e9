"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown)."""
import collections
import csv
import re
import sys


def main(path, skip=0, steps=1.0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))[skip:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        m = re.search(r"gemm_tcgen05_kernel<\(int\)(\d), \(int\)(\d+), \(int\)(\d)>", r["Kernel Name"])
        if m:
            name = f"gsl::gemm_tcgen05_kernel<cg{m.group(1)}, bn{m.group(2)}, epi{m.group(3)}>"
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | launches/step | avg us | ms/step | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print(f"| `{k[:90]}` | {v[0] / steps:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / 1e6 / steps:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values()) / steps:.0f} | | {tot / 1e6 / steps:.3f} | 100% |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
