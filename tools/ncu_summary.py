#!/usr/bin/env python
"""Condenses an .ncu-rep (ncu --set full) into one line per launch: the numbers the roofline needs.

    python tools/ncu_summary.py report.ncu-rep > profiles/xxx.txt
"""
import csv, io, subprocess, sys

COLS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("sm__cycles_elapsed.avg.per_second", "sm_clk"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print("# " + path)
    print("# columns: " + ", ".join(f"{short}[{units[idx[m]]}]" if m in idx else f"{short}[n/a]" for m, short in COLS))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0][-44:]
        vals = []
        for m, short in COLS:
            if m not in idx:
                vals.append(f"{short}=n/a")
                continue
            v = r[idx[m]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            vals.append(f"{short}={v}{units[idx[m]] if short in ('time', 'dram_rd', 'dram_wr', 'l2->sm') else ''}")
        print(f"{r[idx['ID']]:>3} {name:44s} " + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
