"""Markdown tables for DESIGN.md / BASELINE.md from the committed bench lines under profiles/r2/ (so the documents quote the files, not memory)."""
import glob
import json
import os
import sys

R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "r2")


def load(name):
    p = os.path.join(R, name)
    if not os.path.exists(p):
        return None
    txt = open(p).read().strip().splitlines()
    for ln in reversed(txt):
        if ln.startswith("{"):
            return json.loads(ln)
    return None


def row(label, name):
    d = load(name)
    if d is None:
        return f"| {label} | — | — | — | — | not measured | `{name}` |"
    e2e = d.get("e2e", {}).get("value")
    rf = d.get("roofline") or {}
    pv = d.get("pose_vs_oracle_max") or {}
    extra = []
    if pv:
        extra.append(f"pose vs oracle ≤ {pv.get('translation_m', 0):.1e} m / {pv.get('rotation_rad', 0):.1e} rad")
    if "icp_iterations_mean" in d:
        extra.append(f"{d['icp_iterations_mean']:.2f} ICP it.")
    if d.get("clocks"):
        extra.append(f"SM {d['clocks'].get('sm_mhz')} MHz {d['clocks'].get('reasons')}")
    return (f"| {label} | {d.get('n_gpus')} | **{d['value']:.1f}** | {e2e:.1f} | {d.get('ms_per_step', 0):.3f} | "
            f"{rf.get('kernel', '')[:18]} {100 * rf.get('frac', 0):.2f} % | `{name}`; " + "; ".join(extra) + " |")


def main():
    print("| Config | GPUs | scans/s (device-resident) | scans/s (e2e, host buffers) | ms / step | roofline (dominant kernel, frac of HBM peak) | source |")
    print("|---|---|---|---|---|---|---|")
    for label, name in (("C2 before the kernel work of round 2", "bench_n1_before_kernel_work.json"), ("C2", "bench_n1.json"), ("C2, TMA-staged kNN variant", "bench_n1_tma.json"),
                        ("C2, 2 contexts", "bench_n1_contexts2.json"), ("C2, 4 contexts", "bench_n1_contexts4.json"),
                        ("C2 replicas", "bench_replicas_w2.json"), ("C2 replicas", "bench_replicas_w4.json"), ("C2 replicas", "bench_replicas_w8.json"),
                        ("C4 sharded (20M map), 1 GPU", "bench_sharded_w1.json"), ("C4 sharded", "bench_sharded_w2.json"), ("C4 sharded", "bench_sharded_w4.json"), ("C4 sharded", "bench_sharded_w8.json"),
                        ("C3 stream, mode 0", "bench_c3_mode0.json"), ("C3 stream, mode 1", "bench_c3_mode1.json"), ("C5 triple-lidar", "bench_c5_w1.json"), ("C5 triple-lidar", "bench_c5_w8.json"),
                        ("CPU arm (oracle, best OpenMP team)", "bench_reference_arm.json")):
        print(row(label, name))
    d = load("bench_n1.json")
    if d:
        print("\nkernel ms per step:", d.get("kernel_ms_per_step"))
        for k in ("roofline", "roofline_other"):
            r = d.get(k) or {}
            print(f"{k}: {r.get('kernel')}: {r.get('algorithmic_bytes_per_launch', 0) / 1e6:.2f} MB / {r.get('kernel_ms', 0) * 1e3:.1f} us = {r.get('achieved', 0):.1f} GB/s = {100 * r.get('frac', 0):.2f} % of {r.get('peak')} ({r.get('peak_source')}); traffic {r.get('traffic')}; share of step {r.get('share_of_step')}")
        print("cpu_baseline:", d.get("cpu_baseline"))
    for n in (2, 4, 8):
        d = load(f"bench_sharded_w{n}.json")
        if d:
            print(f"sharded w{n}: shard", d.get("shard"), "exchange", d.get("exchange"))


if __name__ == "__main__":
    main()
