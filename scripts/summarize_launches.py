#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total ms, share."""
import csv, gzip, sys, collections, re

def main(path, header):
    op = gzip.open if path.endswith(".gz") else open
    rows = []
    with op(path, "rt") as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    for r in rd:
        name = re.sub(r"\(.*$", "", r[ik]).strip()
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else v / 1e3 if r[iu] in ("us", "usecond") else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms; tot += ms; n += 1
    print(header)
    print("# %d launches, %.1f ms total device time (cold-cache, serialised: compare SHARES)" % (n, tot))
    print("%-92s %9s %12s %8s %12s" % ("kernel", "launches", "total_ms", "share", "avg_ms"))
    for name, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-92s %9d %12.3f %7.2f%% %12.4f" % (name[:92], c, ms, 100 * ms / tot, ms / c))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
