"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`).
Run here, no GPU:  python tools/launch_summary.py gpurun_out/launches.csv [first_id last_id]"""
import csv
import sys
from collections import OrderedDict


def main(path, lo=None, hi=None):
    rows = [r for r in csv.reader(open(path, newline="")) if r]
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[h]
    iname, ival, iunit, iid = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit"), H.index("ID")
    tot = OrderedDict()
    n = 0
    for r in rows[h + 1:]:
        if len(r) <= ival or r[H.index("Metric Name")] != "gpu__time_duration.sum":
            continue
        k = int(r[iid])
        if (lo is not None and k < lo) or (hi is not None and k > hi):
            continue
        v = float(r[ival].replace(",", ""))
        v = v / 1e3 if r[iunit] in ("ns", "nsecond") else (v * 1e3 if r[iunit] in ("ms", "msecond") else v)
        name = r[iname][:60]
        a = tot.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        n += 1
    total = sum(a[1] for a in tot.values())
    print(f"launches {n} total us {total:.1f}")
    for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:60s} n={c:4d} {us:12.1f} us {100 * us / total:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], *(int(a) for a in sys.argv[2:4]))
