"""Derive profiles/rNN_conv_traffic.json (the `roofline.traffic` of bench.py) from a condensed ncu step capture
(tools/ncu_summary.py output of `ncu --set full ... python tools/profile_step.py`): DRAM bytes read + written by the
convolution launches of one step.

    python tools/conv_traffic.py profiles/r02_ncu_step_cfg2.txt > profiles/r02_conv_traffic.json
"""
import json
import re
import sys

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TUNITS = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def field(line, name):
    m = re.search(name + r"=([0-9.e+-]+)([A-Za-z]*)", line)
    return (float(m.group(1)), m.group(2)) if m else (0.0, "")


def main():
    path = sys.argv[1]
    n, total, ms = 0, 0.0, 0.0
    for line in open(path):
        if not re.search(r"conv_(first|wide|stream|resident)|splitk_finish", line):
            continue
        rd, ru = field(line, "dram_rd")
        wr, wu = field(line, "dram_wr")
        t, tu = field(line, "time")
        total += rd * UNITS.get(ru, 1.0) + wr * UNITS.get(wu, 1.0)
        ms += t * TUNITS.get(tu, 1.0)
        n += 1
    print(json.dumps({
        "source": f"{path} (ncu --set full, one step of cfg2 = 256 x 100 x 100, computed at 104 x 104)",
        "conv_launches": n,
        "conv_dram_bytes_per_step": total,
        "conv_ms_under_ncu": ms,
        "note": "sum of dram__bytes_read.sum + dram__bytes_write.sum over the conv_first / conv_wide / conv_stream / "
                "conv_resident / splitk_finish launches of the step",
    }, indent=1))


if __name__ == "__main__":
    main()
