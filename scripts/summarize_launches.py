"""Turn an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table (markdown).

    python scripts/summarize_launches.py launches.csv [first_row] [n_rows] [steps]
Rows [first_row, first_row + n_rows) are aggregated; `steps` = how many steps that range covers."""
import collections
import csv
import re
import sys

EPI = {0: "F16", 1: "F32", 2: "GELU", 3: "GELU_BWD", 4: "RES_F32", 5: "PERIODIC_F32", 6: "F16_ROWDOT"}


def short(name):
    m = re.search(r"gemm_tcgen05_kernel<(?:\(int\))?(\d), (?:\(int\))?(\d+), (?:\(int\))?(\d)(?:, (?:\((?:bool|int)\))?(\w+))?(?:, (?:\(int\))?(\d))?>", name)
    if m:
        split = {"1": ", SPLIT", "true": ", SPLIT", "2": ", SPLIT8"}.get(m.group(4), "")
        eg = ", 1 epilogue group" if m.group(5) == "1" else ""
        return f"gsl::gemm_tcgen05_kernel<cg{m.group(1)}, bn{m.group(2)}, {EPI.get(int(m.group(3)), m.group(3))}{split}{eg}>"
    return re.sub(r"\(.*", "", name).replace("void ", "")


def main(path, first=0, count=None, steps=1.0):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = rows[first:first + count] if count else rows[first:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = short(r["Kernel Name"])
        agg[name][0] += 1
        agg[name][1] += float(r["Metric Value"].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches/step | avg us | ms/step | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        print(f"| `{k[:90]}` | {v[0] / steps:.1f} | {v[1] / v[0] / 1e3:.1f} | {v[1] / 1e6 / steps:.3f} | {100 * v[1] / tot:.1f}% |")
    print(f"| **total** | {sum(v[0] for v in agg.values()) / steps:.0f} | | {tot / 1e6 / steps:.3f} | 100% |")


if __name__ == "__main__":
    a = sys.argv
    main(a[1], int(a[2]) if len(a) > 2 else 0, int(a[3]) if len(a) > 3 and int(a[3]) > 0 else None, float(a[4]) if len(a) > 4 else 1.0)
