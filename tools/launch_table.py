"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list.
    python tools/launch_table.py launches.csv [repeats]
repeats: how many times the measured call ran (build timing: the LAST launch of each kernel is shown and one call's kernels are
summed); without dram metrics the table shows every kernel's launches, total time and SHARE of all GPU time instead."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr, agg, order = None, {}, []
for r in rows:
    if len(r) > 10 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        k = d["Kernel Name"].split("(")[0][:60]
        agg.setdefault(k, {}).setdefault(d["Metric Name"], []).append(float(d["Metric Value"].replace(",", "")))
        if k not in order:
            order.append(k)
have_dram = all("dram__bytes_read.sum" in agg[k] for k in order) and bool(order)
if have_dram:
    tot = 0.0
    print(f"{'kernel':46s} {'launches':>8s} {'last us':>9s} {'dram rd MB':>11s} {'dram wr MB':>11s}")
    for k in order:
        a = agg[k]
        t = a["gpu__time_duration.sum"]
        per_call = max(1, len(t) // rep)
        tot += t[-1] * per_call
        print(f"{k[:46]:46s} {len(t):8d} {t[-1] / 1e3:9.1f} {a['dram__bytes_read.sum'][-1] / 1e6:11.1f} {a['dram__bytes_write.sum'][-1] / 1e6:11.1f}")
    print(f"sum of one call's kernels (cold-cache, serialised by ncu): {tot / 1e3:.1f} us")
else:
    total = sum(sum(agg[k]["gpu__time_duration.sum"]) for k in order)
    print(f"{'kernel':60s} {'launches':>8s} {'total ms':>9s} {'share %':>8s} {'avg us':>9s}")
    for k in sorted(order, key=lambda k: -sum(agg[k]["gpu__time_duration.sum"])):
        t = agg[k]["gpu__time_duration.sum"]
        print(f"{k:60s} {len(t):8d} {sum(t) / 1e6:9.3f} {100 * sum(t) / total:8.2f} {sum(t) / len(t) / 1e3:9.1f}")
    print(f"all kernels: {total / 1e6:.3f} ms in {sum(len(agg[k]['gpu__time_duration.sum']) for k in order)} launches (per-launch times are cold-cache and serialised by ncu)")
