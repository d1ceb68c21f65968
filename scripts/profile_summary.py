#!/usr/bin/env python
"""Turn the raw ncu captures gpurun brings back (gpurun_out/, scratch) into the small text summaries that are
committed under profiles/.

    python scripts/profile_summary.py <tag> [--n-waters N --frames F]

reads   gpurun_out/launches_<tag>.csv      (ncu --metrics gpu__time_duration.sum launch list of bench.py)
        gpurun_out/prof_<tag>.ncu-rep      (ncu --set full capture of the hot kernels)
writes  profiles/<tag>_launches.txt        one bench step: kernel, grid, block, device time, share of the step
        profiles/<tag>_kernels.csv         per captured kernel: time, DRAM bytes, registers, occupancy, pipes
        profiles/<tag>_hotspots_<kernel>.txt  per-source-line instruction / stall shares + stall reasons
        profiles/dominant_kernel_traffic.json DRAM bytes per water-frame of the dominant kernel (bench.py reads it)
"""
import argparse
import csv
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SCRATCH = os.path.join(ROOT, "gpurun_out")

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_lim_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_lim_smem_blocks"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_throughput_pct"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_throughput_pct"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fp64_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "pipe_fma_fp32_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe_lsu_pct"),
    ("smsp__inst_executed.sum", "warp_insts"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_inst"),
    ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
]


def short(name):
    name = name.replace("void ", "").replace("wol::", "")
    return name.split("(")[0]


def launches(tag):
    path = os.path.join(SCRATCH, "launches_%s.csv" % tag)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols = rows[hdr]
    seq = []
    for r in rows[hdr + 1:]:
        d = dict(zip(cols, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            seq.append((short(d["Kernel Name"]), d["Grid Size"], d["Block Size"], float(d["Metric Value"]) / 1e3))
    # one step = from one cell_pass<..., 0> launch to the next
    starts = [i for i, s in enumerate(seq) if "cell_pass_kernel" in s[0] and s[0].rstrip(">").endswith("0")]
    if len(starts) < 3:
        return None
    a, b = starts[1], starts[2]
    step = seq[a:b]
    total = sum(s[3] for s in step)
    lines = ["# one bench step (launches %d..%d of %d captured); ncu per-launch times are cold-cache and serialised:" % (a, b - 1, len(seq)),
             "# compare SHARES, not absolutes", "%-58s %-16s %-14s %10s %7s" % ("kernel", "grid", "block", "time_us", "share")]
    for s in step:
        lines.append("%-58s %-16s %-14s %10.1f %6.1f%%" % (s[0][:58], s[1], s[2], s[3], 100.0 * s[3] / total))
    lines.append("%-58s %-16s %-14s %10.1f %6.1f%%" % ("TOTAL (our kernels, one step)", "", "", total, 100.0))
    open(os.path.join(OUT, "%s_launches.txt" % tag), "w").write("\n".join(lines) + "\n")
    return step


def ncu_csv(rep, page, extra=()):
    cmd = ["ncu", "-i", rep, "--page", page, "--csv"] + list(extra)
    return subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def kernels(tag):
    rep = os.path.join(SCRATCH, "prof_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return None
    rows = list(csv.reader(ncu_csv(rep, "raw").splitlines()))
    hdr, units = rows[0], rows[1]
    u = dict(zip(hdr, units))
    seen = OrderedDict()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        k = short(d["Kernel Name"])
        if k not in seen:
            seen[k] = d
    with open(os.path.join(OUT, "%s_kernels.csv" % tag), "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + ["%s[%s]" % (n, u.get(m, "")) for m, n in METRICS])
        for k, d in seen.items():
            w.writerow([k] + [d.get(m, "") for m, _ in METRICS])
    return seen, u


def hotspots(tag, kernel_regex, label, top=40):
    rep = os.path.join(SCRATCH, "prof_%s.ncu-rep" % tag)
    text = ncu_csv(rep, "source", ["--print-source", "cuda,sass", "--kernel-name", "regex:" + kernel_regex, "--launch-count", "1"])
    rows = list(csv.reader(text.splitlines()))
    agg = defaultdict(lambda: [0, 0, ""])
    stall = defaultdict(float)
    cur, hdr = "", None
    ti = ts = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
            continue
        d = dict(zip(hdr, r))
        try:
            inst, samp = int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0)
        except ValueError:
            continue
        a = agg[(cur, int(r[0]))]
        a[0] += inst
        a[1] += samp
        a[2] = r[1].strip()[:100]
        ti += inst
        ts += samp
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stall[k] += float(v or 0)
                except ValueError:
                    pass
    if ts == 0:
        return
    lines = ["# %s: warp-instructions %d, stall samples %d (ncu --set full, source page, aggregated per CUDA line)" % (label, ti, ts),
             "# stall reasons: " + ", ".join("%s %.1f%%" % (k[6:], 100 * v / sum(stall.values()))
                                            for k, v in sorted(stall.items(), key=lambda kv: -kv[1])[:8])]
    for key, (inst, samp, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        lines.append("%-20s %4d  inst %5.1f%%  stall %5.1f%%  %s" % (key[0], key[1], 100.0 * inst / max(ti, 1), 100.0 * samp / ts, src))
    open(os.path.join(OUT, "%s_hotspots_%s.txt" % (tag, label)), "w").write("\n".join(lines) + "\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--n-waters", type=int, default=1000000)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--dominant", default="q3b_tpc_kernel")
    ap.add_argument("--hot", nargs="*", default=["q3b_tpc_kernel", "q3b_tpc_widen_kernel"])
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    launches(a.tag)
    res = kernels(a.tag)
    if res:
        seen, u = res
        for k, d in seen.items():
            if a.dominant in k:
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
                rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
                wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
                json.dump({"kernel": k, "capture": "profiles/%s_kernels.csv" % a.tag, "n_waters": a.n_waters, "frames_per_launch": a.frames,
                           "dram_bytes_per_launch": rd + wr, "dram_bytes_per_water_frame": (rd + wr) / (a.n_waters * a.frames)},
                          open(os.path.join(OUT, "dominant_kernel_traffic.json"), "w"), indent=1)
                break
        for h in a.hot:
            hotspots(a.tag, h, h)
    print("profiles written for", a.tag)


if __name__ == "__main__":
    sys.exit(main())
