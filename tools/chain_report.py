"""Summarise the RLFC_CHAIN_STATS printf lines of one launch (tools/build_variant.py stats -DRLFC_CHAIN_STATS=<serial>)."""
import re
import sys

for f in sys.argv[1:]:
    rows = []
    for ln in open(f):
        m = re.match(r'chain L(\d+) g(\d+) s(\d+) start (\d+) end (\d+) miss (\d+) spins (\d+) steps (\d+)', ln)
        if m:
            rows.append(tuple(int(x) for x in m.groups()))
    for L in sorted({x[0] for x in rows}):
        r = [x for x in rows if x[0] == L]
        t0 = min(x[3] for x in r)
        print(f, "level", L, "total us %.1f" % ((max(x[4] for x in r) - t0) / 1e3), "warps", len(r))
        for g in (1, 4):
            rr = sorted([x for x in r if x[1] == g], key=lambda x: x[2])
            for x in rr[:2] + rr[-2:]:
                print("  g%d s%2d end %7.1f miss %4d spins %5d  ns/step %.1f" % (g, x[2], (x[4] - t0) / 1e3, x[5], x[6], (x[4] - x[3]) / x[7]))
