#!/usr/bin/env python
"""Per-kernel evidence from an ncu --set full capture: one row per kernel (last captured launch) with duration, DRAM traffic
and throughput against the measured HBM peak, FP64 / FP32 pipe utilisation, issue-slot utilisation, lanes per instruction --
and the SASS of each kernel from libwol.so.

    python scripts/kernel_table.py <tag>      reads gpurun_out/prof_<tag>.ncu-rep, writes profiles/<tag>_kernel_table.md,
                                              profiles/<tag>_sass_<kernel>.txt
"""
import csv
import json
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", "prof_%s.ncu-rep" % tag)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
text = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hdr, units = rows[0], rows[1]
u = dict(zip(hdr, units))
last = OrderedDict()
tscale0 = {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"]
    name = re.sub(r"^void ", "", name)
    name = name.split("(")[0].replace("wol::", "")
    def size(x):  # the largest launch of each kernel (the same kernel also runs on small inputs): grid first, then duration
        return (int(x["launch__grid_size"]), float(x["gpu__time_duration.sum"]) * tscale0[u["gpu__time_duration.sum"]])
    if name not in last or size(d) > size(last[name]):
        last[name] = d
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
tscale = {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}
lines = ["| kernel | grid x block | regs | time (ms) | DRAM read + write (MB) | DRAM GB/s | of measured HBM (%.1f GB/s) | FP64 pipe %% | FP32 FMA pipe %% | issue slots busy %% | lanes / instr | warps / SM |" % peak,
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for name, d in last.items():
    t = float(d["gpu__time_duration.sum"]) * tscale[u["gpu__time_duration.sum"]]
    rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
    wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
    gbs = (rd + wr) / (t * 1e-3) / 1e9
    warps = float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]) * 64 / 100.0
    lines.append("| `%s` | %s x %s | %s | %.3f | %.1f + %.1f | %.0f | %.1f %% | %.1f | %.1f | %.1f | %.1f | %.1f |" % (
        name, d["launch__grid_size"], d["launch__block_size"], d["launch__registers_per_thread"], t, rd / 1e6, wr / 1e6, gbs, 100 * gbs / peak,
        float(d["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]), float(d["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"]),
        float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]), warps))
out = os.path.join(ROOT, "profiles", "%s_kernel_table.md" % tag)
open(out, "w").write("# ncu --set full, the longest captured launch of each kernel (scripts/profile_all.py; 2 x 1M-water frames, 1M-water H-bond frame, "
                     "65536-water slab)\n\n" + "\n".join(lines) + "\n")
print(open(out).read())

# SASS of the same kernels from the shipped library
so = os.path.join(ROOT, "waterorderlib_b200", "libwol.so")
dump = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blocks = re.split(r"\n\s*Function : ", dump)
for name in last:
    short = name.split("<")[0]
    for b in blocks[1:]:
        mangled = b.split("\n", 1)[0].strip()
        demangled = subprocess.run(["c++filt", mangled], stdout=subprocess.PIPE, text=True).stdout.strip()
        def norm(x):
            x = x.replace("wol::", "").replace("void ", "").split("(")[0].replace(" ", "")
            return x.replace("false", "0").replace("true", "1")
        if norm(demangled) == norm(name):
            body = b.split("\n", 1)[1]
            ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", body)
            hist = {}
            for o in ops:
                k = o.split(".")[0]
                hist[k] = hist.get(k, 0) + 1
            top = ", ".join("%s %d" % kv for kv in sorted(hist.items(), key=lambda kv: -kv[1])[:24])
            fn = re.sub(r"[^A-Za-z0-9_]+", "_", name).strip("_")
            with open(os.path.join(ROOT, "profiles", "%s_sass_%s.txt" % (tag, fn)), "w") as f:
                f.write("// %s\n// %d instructions; most frequent: %s\n// async copy / barrier instructions: %s\n" % (
                    demangled, len(ops), top, ", ".join("%s %d" % (k, hist[k]) for k in ("UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "NANOSLEEP") if k in hist) or "none"))
                f.write(re.sub(r"\s*/\* 0x[0-9a-f]{16} \*/", "", body))  # without the hex encodings
            break
