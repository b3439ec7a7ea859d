"""Per-kernel one-liners from an ncu metrics CSV (every kernel of the launch list, not just the two hot ones).

  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum \
      --clock-control none -c 700 --csv --log-file gpurun_out/kernels.csv python bench.py --steps 2 --warmup 3 --no-cpu
  python profiles/tools/kernel_table.py gpurun_out/kernels.csv profiles/r2/kernels_r2.txt

Times under ncu are cold-cache and serialised: the SHARE of a kernel is comparable with the live CUDA-event numbers, the absolute is not."""
import collections
import csv
import sys


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ki, mi, ui, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0, "inst": 1.0}
    per = collections.OrderedDict()
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")[:64]
        v = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
        d = per.setdefault((r[0], name), {})
        d[r[mi]] = v
    agg = collections.OrderedDict()
    for (_, name), d in per.items():
        a = agg.setdefault(name, collections.Counter())
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["dram"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
        a["issue"] += d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        a["inst"] += d.get("smsp__inst_executed.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    lines = [f"{'kernel':64s} {'launches':>8s} {'us/launch':>10s} {'share':>7s} {'DRAM MB/launch':>15s} {'DRAM GB/s':>10s} {'issue %':>8s} {'warp inst/launch':>17s}"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = a["n"]
        us = a["us"] / n
        lines.append(f"{name:64s} {n:8d} {us:10.1f} {100 * a['us'] / tot:6.1f}% {a['dram'] / n / 1e6:15.3f} {a['dram'] / max(a['us'], 1e-9) / 1e3:10.1f} {a['issue'] / n:8.1f} {a['inst'] / n:17.0f}")
    lines.append(f"total {tot:.1f} us over {sum(a['n'] for a in agg.values())} launches")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:45]))


if __name__ == "__main__":
    main()
