"""Turns the ncu captures brought back in gpurun_out/ into the text summaries committed here.
usage: python profiles/summarize.py <report.ncu-rep> <out.txt> [title] [kernel source file under loam_livox_b200/csrc/]
With the 4th argument the machine-readable companion (<out>.json: DRAM bytes per launch = bench.py's roofline.traffic) is stamped with the sha1 of that
source file; bench.py quotes the traffic only while the stamp matches the source it runs, so a stale capture is never reported."""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    lines = [title, "source: ncu --set full --clock-control none --import-source on (one GPU, under gpurun); per-launch values", ""]
    for r in body:
        lines.append("kernel: " + r[hdr.index("Kernel Name")][:100])
        for w in WANT:
            if w in hdr:
                lines.append(f"  {w:90s} {r[hdr.index(w)]:>18s} {units[hdr.index(w)]}")
        if "dram__bytes_read.sum" in hdr:
            def val(name):
                v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            lines.append(f"  traffic (dram read + write) = {(val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / 1e6:.3f} MB per launch")
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    # machine-readable companion (bench.py reads roofline.traffic from it): longest launch of the report
    if body and "dram__bytes_read.sum" in hdr:
        import json
        def _d(row):
            v, u = float(row[hdr.index('gpu__time_duration.sum')].replace(',', '')), units[hdr.index('gpu__time_duration.sum')]
            return v * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(u, 1.0)
        r = max(body, key=_d)   # the longest captured launch (for knn_blocks_kernel: the unseeded first ICP iteration the bench's roofline is quoted on)
        def val2(name):
            v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        def dur(name):
            v, u = float(r[hdr.index(name)].replace(",", "")), units[hdr.index(name)]
            return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
        sha = None
        if len(sys.argv) > 4:
            import hashlib, os
            with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "loam_livox_b200", "csrc", sys.argv[4]), "rb") as f:
                sha = hashlib.sha1(f.read()).hexdigest()
        # traffic = mean over the captured launches (bench.py's roofline is the mean over all launches of the timed region, seeded iterations included)
        tr = [float(b[hdr.index("dram__bytes_read.sum")].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index("dram__bytes_read.sum")], 1) +
              float(b[hdr.index("dram__bytes_write.sum")].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[hdr.index("dram__bytes_write.sum")], 1) for b in body]
        json.dump({"kernel": r[hdr.index("Kernel Name")][:100], "traffic_bytes_per_launch": sum(tr) / len(tr), "traffic_bytes_longest_launch": val2("dram__bytes_read.sum") + val2("dram__bytes_write.sum"),
                   "launches_captured": len(tr), "duration_us_under_ncu": dur("gpu__time_duration.sum"), "source": rep.split("/")[-1], "source_sha1": sha},
                  open(out.rsplit(".", 1)[0] + ".json", "w"))
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
