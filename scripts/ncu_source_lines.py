"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line: samples, executed instructions, top stall reasons.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name K --launch-count 1 > k.csv; python scripts/ncu_source_lines.py k.csv [N]"""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path, newline="")))
    cur, hdr, agg = None, None, {}
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) != len(hdr) or not r[0].strip().isdigit() or r[2] != "-":
            continue
        idx = {n: i for i, n in enumerate(hdr)}
        stalls = {n: int(r[i] or 0) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n}
        agg[(cur, int(r[0]))] = (int(r[idx["# Samples"]] or 0), int(r[idx["Instructions Executed"]] or 0), r[1].strip()[:100],
                                 sorted(stalls.items(), key=lambda x: -x[1])[:3])
    tot = sum(v[0] for v in agg.values())
    print("total samples", tot)
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
        print(f"{k[0]}:{k[1]:<5d} {v[0]:6d} {100 * v[0] / max(tot, 1):5.1f}%  inst {v[1]:9d}  {v[3]}  | {v[2]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
