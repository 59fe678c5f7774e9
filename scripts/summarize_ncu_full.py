"""Reduce an `ncu --set full` report to the per-launch columns DESIGN.md / bench.py cite (run where the .ncu-rep lives):

    python scripts/summarize_ncu_full.py gpurun_out/r01t_full.ncu-rep > profiles/r01t_ncu_full_summary.csv
    python scripts/summarize_ncu_full.py gpurun_out/r01t_full.ncu-rep --traffic profiles/ncu_traffic.json

--traffic writes dram__bytes_read.sum + dram__bytes_write.sum of the fc1 (EPI_GELU, dropout launch) and fc2 (EPI_RES_F32, dropout launch)
GEMMs of scripts/dev_prof.py -- the `roofline.traffic` figure of bench.py."""
import csv
import io
import json
import subprocess
import sys

COLS = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__cycles_active.avg", "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def to_bytes(value, unit):
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]


def main():
    rep = sys.argv[1]
    hdr, units, rows = load(rep)
    idx = {n: i for i, n in enumerate(hdr)}
    cols = [c for c in COLS if c in idx]
    if "--traffic" in sys.argv:
        path = sys.argv[sys.argv.index("--traffic") + 1]
        r_i, w_i = idx["dram__bytes_read.sum"], idx["dram__bytes_write.sum"]

        def total(r):
            return to_bytes(r[r_i], units[r_i]) + to_bytes(r[w_i], units[w_i])
        gelu = [r for r in rows if "256, (int)2>" in r[idx["Kernel Name"]] or "256, 2>" in r[idx["Kernel Name"]]]
        res = [r for r in rows if "256, (int)4>" in r[idx["Kernel Name"]] or "256, 4>" in r[idx["Kernel Name"]]]
        # dev_prof.py launches each shape without and then with dropout 0.1 (timeit runs fn twice): the LAST launch of each kind has dropout on
        fc1, fc2 = total(gelu[-1]), total(res[-1])
        json.dump({"source": f"{rep.split('/')[-1]} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, per launch, M = 201728 rows, dropout 0.1)",
                   "gemm_EPI_GELU_fc1_bytes": int(fc1), "gemm_EPI_RES_F32_fc2_bytes": int(fc2), "ffn_pair_bytes": int(fc1 + fc2)},
                  open(path, "w"), indent=1)
        print(open(path).read())
        return
    w = csv.writer(sys.stdout)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in rows:
        w.writerow([r[idx[c]] for c in cols])


if __name__ == "__main__":
    main()
