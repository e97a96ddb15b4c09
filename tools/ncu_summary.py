"""Summarise an ncu report (`ncu -i X.ncu-rep --page raw --csv`) per kernel: duration, DRAM bytes, throughputs.

  python tools/ncu_summary.py raw.csv out_prefix     -> out_prefix.json, out_prefix.md
"""
import csv
import json
import sys
from collections import OrderedDict


def short(name):
    n = name.split("(")[0]
    n = n.replace("rlfc::<unnamed>::", "").replace("void ", "").replace("unnamed>::", "")
    return n


def main(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    want = OrderedDict([
        ("time_us", "gpu__time_duration.sum"), ("dram_read_B", "dram__bytes_read.sum"), ("dram_write_B", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("inst", "smsp__inst_executed.sum"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
    ])
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6}
    agg = OrderedDict()
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = short(r[idx["Kernel Name"]])
        rec = agg.setdefault(k, {"launches": 0, **{w: 0.0 for w in want}})
        rec["launches"] += 1
        for w, col in want.items():
            if col not in idx:
                continue
            try:
                v = float(r[idx[col]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[col]]
            if w in ("dram_read_B", "dram_write_B", "time_us"):
                v *= scale.get(u, 1)
            rec[w] += v
    out_rec = OrderedDict()
    for k, rec in agg.items():
        n = rec["launches"]
        out_rec[k] = {"launches_profiled": n, **{w: rec[w] / n for w in want}}
        out_rec[k]["dram_traffic_B_per_launch"] = out_rec[k]["dram_read_B"] + out_rec[k]["dram_write_B"]
        t = out_rec[k]["time_us"]
        out_rec[k]["dram_GBps"] = out_rec[k]["dram_traffic_B_per_launch"] / (t * 1e-6) / 1e9 if t else 0.0
    json.dump(out_rec, open(out + ".json", "w"), indent=1)
    with open(out + ".md", "w") as f:
        f.write("| kernel | launches | avg us | DRAM read MB | DRAM write MB | DRAM GB/s | dram % | sm % | warps active % | regs | L2 hit % |\n")
        f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
        for k, r in sorted(out_rec.items(), key=lambda kv: -kv[1]["time_us"]):
            f.write(f"| {k} | {r['launches_profiled']} | {r['time_us']:.1f} | {r['dram_read_B'] / 1e6:.1f} | {r['dram_write_B'] / 1e6:.1f} | "
                    f"{r['dram_GBps']:.0f} | {r['dram_pct']:.1f} | {r['sm_pct']:.1f} | {r['warps_active_pct']:.1f} | {r['regs']:.0f} | {r['l2_hit_pct']:.1f} |\n")
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
