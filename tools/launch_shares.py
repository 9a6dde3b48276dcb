"""Per-kernel time shares from an ncu launch list (``ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...``):
    python tools/launch_shares.py X.csv "header line describing the command" > X_shares.txt
The per-launch times are cold-cache and serialised: compare a kernel's SHARE of the step, not the absolutes."""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= vi or not r[vi]:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        us = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        name = re.sub(r"\s+", " ", r[ki])
        t = tot.setdefault(name, [0.0, 0])
        t[0] += us
        t[1] += 1
    total = sum(t[0] for t in tot.values())
    if len(sys.argv) > 2:
        print(sys.argv[2])
    for name, (us, n) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:24]:
        print(f"{us:12.1f} us {100 * us / total:5.1f}%  x{n:<4d} {name[:110]}")
    print(f"all launches {total:.1f} us")


if __name__ == "__main__":
    main()
