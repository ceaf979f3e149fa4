#!/usr/bin/env python
"""Turn one `ncu --set full` capture of a whole bench step into the committed evidence:

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > /tmp/raw.csv
    python profiles/extract_ncu.py /tmp/raw.csv profiles/ncu_rNN_x_summary.md [profiles/traffic.json]

The capture is expected to hold the launches of one step in launch order (warp, pyrdown l0.., collapse l_nb..l0),
which is the order bench.py's `roofline.launches_ms` uses; traffic.json maps those names to DRAM bytes per launch.
"""
import csv
import json
import sys

COLS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
        ("smsp__inst_executed.sum", "warp insts"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs")]


def fnum(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


def main():
    raw, out_md = sys.argv[1], sys.argv[2]
    out_json = sys.argv[3] if len(sys.argv) > 3 else None
    rows = list(csv.reader(open(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = []
    n_pyr = sum(1 for r in data if "pyrdown" in r[hdr.index("Kernel Name")])
    n_col = sum(1 for r in data if "collapse" in r[hdr.index("Kernel Name")])
    ip = ic = 0
    for r in data:
        k = r[hdr.index("Kernel Name")]
        if "warp" in k:
            names.append("warp")
        elif "pyrdown" in k:
            names.append(f"pyrdown_l{ip}")
            ip += 1
        elif "collapse" in k:
            names.append(f"collapse_l{n_col - 1 - ic}")
            ic += 1
        else:
            names.append(k)
    traffic = {}
    with open(out_md, "w") as f:
        f.write("| launch | kernel | " + " | ".join(c[1] for c in COLS) + " |\n|" + "---|" * (len(COLS) + 2) + "\n")
        for name, r in zip(names, data):
            k = r[hdr.index("Kernel Name")].split("::")[-1].split("(")[0].replace("unnamed>", "").strip()
            vals = []
            for key, _ in COLS:
                v = fnum(r[hdr.index(key)]) if key in hdr else None
                vals.append("" if v is None else f"{v:.4g}")
            f.write(f"| {name} | `{k}` | " + " | ".join(vals) + " |\n")
            rd, wr = fnum(r[hdr.index("dram__bytes_read.sum")]), fnum(r[hdr.index("dram__bytes_write.sum")])
            scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
            ru, wu = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
            traffic[name] = rd * scale.get(ru, 1.0) + wr * scale.get(wu, 1.0)
    if out_json:
        json.dump(traffic, open(out_json, "w"), indent=1)
    print(f"{len(data)} launches ({n_pyr} pyrdown, {n_col} collapse) -> {out_md}")


if __name__ == "__main__":
    main()
