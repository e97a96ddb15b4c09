"""Summarise `ncu --page source --csv` output: per-kernel stall mix and the most-stalled instructions."""
import csv
import sys


def main(path, ntop=14):
    rows = list(csv.reader(open(path)))
    sections, cur = [], None
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": rows[i + 1], "data": []}
            sections.append(cur)
            i += 2
            continue
        if cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
        i += 1
    for sec in sections:
        hdr, data = sec["hdr"], sec["data"]
        idx = {h: k for k, h in enumerate(hdr)}
        tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {s: sum(int(r[idx[s]] or 0) for r in data) for s in stalls}
        print(sec["name"][:70], "samples", tot, "instrs", len(data))
        print("  ", [(k, round(v / tot, 3)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]])
        for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:ntop]:
            st = {s: int(r[idx[s]] or 0) for s in stalls}
            main_s = max(st, key=st.get)
            print("    ", r[idx["# Samples"]], r[idx["Source"]].strip()[:64], main_s, st[main_s])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 14)
