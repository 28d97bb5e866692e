#!/usr/bin/env python
"""Summarise `ncu --set full` reports (gpurun_out/full_<TAG>_*.ncu-rep) into one CSV for profiles/: one row per captured
launch, units normalised (us, byte, %, GHz). usage: tools/ncu_summary.py TAG OUT.csv"""
import csv
import glob
import io
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__grid_size",
        "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"ns": ("us", 1e-3), "us": ("us", 1.0), "ms": ("us", 1e3), "s": ("us", 1e6), "byte": ("byte", 1.0), "Kbyte": ("byte", 1e3),
         "Mbyte": ("byte", 1e6), "Gbyte": ("byte", 1e9), "hz": ("Ghz", 1e-9), "Khz": ("Ghz", 1e-6), "Mhz": ("Ghz", 1e-3),
         "Ghz": ("Ghz", 1.0)}


def main():
    tag, out = sys.argv[1], sys.argv[2]
    rows_out, units_out = [], None
    for rep in sorted(glob.glob(f"gpurun_out/full_{tag}_*.ncu-rep")):
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        kn = hdr.index("Kernel Name")
        idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
        u_norm = []
        for r in rows[2:]:
            vals, u_norm = [], []
            for i in idx:
                if i < 0:
                    vals.append("")
                    u_norm.append("")
                    continue
                u = units[i]
                nu, sc = SCALE.get(u, (u, 1.0))
                try:
                    vals.append(f"{float(r[i].replace(',', '')) * sc:.6g}")
                except ValueError:
                    vals.append(r[i])
                u_norm.append(nu)
            rows_out.append([r[kn].split("(")[0].strip()] + vals)
        units_out = [""] + u_norm
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["Kernel Name"] + COLS)
        w.writerow(units_out or [""] * (len(COLS) + 1))
        w.writerows(rows_out)
    print(f"{len(rows_out)} launches -> {out}")


if __name__ == "__main__":
    main()
