"""Summarises an `ncu --set full` report of ONE hot-path step (scripts/probes/one_step.py) kernel by kernel, and records the DRAM
traffic of the conv launches in profiles/conv_traffic.json -- the number bench.py reports as roofline.traffic.

    ncu -i gpurun_out/X.ncu-rep --page raw --csv > raw.csv ;  python scripts/ncu_traffic.py raw.csv c3 32 profiles/r02_ncu_full_step_c3.txt
"""
import csv
import datetime
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CONV = ("conv_", "conv3x3", "stem_", "avgpool", "maxpool")


def num(v):
    try:
        return float(str(v).replace(",", ""))
    except ValueError:
        return 0.0


def main():
    raw, workload, B, out_txt = sys.argv[1], sys.argv[2], int(sys.argv[3]), Path(sys.argv[4])
    rows = list(csv.reader(open(raw, newline="")))
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hdr_i], rows[hdr_i + 1]
    col = {name: i for i, name in enumerate(hdr)}

    def get(r, name, want_unit=None):
        i = col.get(name)
        if i is None:
            return 0.0
        v, u = num(r[i]), units[i].lower()
        scale = {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6,
                 "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}
        return v * scale.get(u, 1.0)

    lines = ["kernel | time us | dram read MB | dram write MB | dram throughput % | sm throughput % | tensor pipe % | regs"]
    conv_bytes = conv_us = tot_us = 0.0
    for r in rows[hdr_i + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[col["Kernel Name"]]
        t = get(r, "gpu__time_duration.sum")
        rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
        dthr = num(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]) if "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed" in col else 0.0
        smthr = num(r[col["sm__throughput.avg.pct_of_peak_sustained_elapsed"]]) if "sm__throughput.avg.pct_of_peak_sustained_elapsed" in col else 0.0
        tp_name = next((c for c in col if c.startswith("sm__pipe_tensor") and "pct" in c), None)
        tens = num(r[col[tp_name]]) if tp_name else 0.0
        regs = num(r[col["launch__registers_per_thread"]]) if "launch__registers_per_thread" in col else 0
        lines.append(f"{name[:60]:60s} | {t:8.1f} | {rd / 1e6:8.1f} | {wr / 1e6:8.1f} | {dthr:5.1f} | {smthr:5.1f} | {tens:5.1f} | {int(regs)}")
        tot_us += t
        if any(k in name for k in CONV):
            conv_bytes += rd + wr
            conv_us += t
    lines.append(f"total {tot_us:.1f} us under ncu (serialised, cold cache); conv launches {conv_us:.1f} us, {conv_bytes / 1e6:.1f} MB of DRAM traffic")
    out_txt.write_text("\n".join(lines) + "\n")
    p = ROOT / "profiles" / "conv_traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d[f"{workload}_b{B}"] = {"dram_bytes": conv_bytes, "conv_us_under_ncu": conv_us, "source": f"profiles/{out_txt.name}",
                              "captured": datetime.date.today().isoformat()}
    p.write_text(json.dumps(d, indent=1) + "\n")
    print(lines[-1])


if __name__ == "__main__":
    main()
