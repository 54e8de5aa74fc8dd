#!/usr/bin/env python
"""Condenses an `ncu --set full` report into one CSV row per captured launch (the columns the roofline figures use).
usage: ncu_summary.py report.ncu-rep out.csv        (runs `ncu -i ... --page raw --csv` itself)"""
import csv
import subprocess
import sys

COLS = [("kernel", "Kernel Name"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("smem_dyn_kb", "launch__shared_mem_per_block_dynamic"),
        ("time_us", "gpu__time_duration.sum"), ("dram_read_gb", "dram__bytes_read.sum"), ("dram_write_gb", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("warp_inst", "smsp__inst_executed.sum"), ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("smem_bank_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")]


def to_unit(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
    return v * scale.get(unit, 1.0) if want else v


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(sys.argv[2], "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow([c for c, _ in COLS] + ["dram_total_gb", "achieved_dram_gbs"])
        for r in rows[2:]:
            out = []
            for name, col in COLS:
                if col not in idx:
                    out.append("")
                    continue
                v, u = r[idx[col]], units[idx[col]]
                if name == "kernel":
                    out.append(v.split("(")[0].replace("void ", ""))
                elif name in ("time_us", "dram_read_gb", "dram_write_gb"):
                    out.append(round(to_unit(v, u, True), 4))
                else:
                    out.append(v)
            tot = out[6] + out[7]
            out += [round(tot, 4), round(tot / (out[5] * 1e-6), 1) if out[5] else ""]
            wr.writerow(out)


if __name__ == "__main__":
    main()
