#!/usr/bin/env python
"""Turn the raw evidence of tools/make_profiles.sh (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py [round-tag, default r02]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def rows(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def per_kernel(path):
    per, order = {}, []
    for r in rows(path):
        k = (r["ID"], r["Kernel Name"])
        if k not in per:
            per[k] = {}
            order.append(k)
        per[k][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    return [(k[1], per[k]) for k in order]


def short(name):
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name[:name.index("(")] if "(" in name else name


def launch_summary(src, dst, title, extra_cols=()):
    ks = per_kernel(src)
    agg = {}
    for name, m in ks:
        a = agg.setdefault(short(name), {"n": 0})
        a["n"] += 1
        for mk, v in m.items():
            a[mk] = a.get(mk, 0.0) + v
    tot = sum(a["gpu__time_duration.sum"] for a in agg.values())
    with open(dst, "w") as f:
        f.write(title + "\n")
        f.write("%-78s %5s %12s %7s" % ("kernel", "n", "time us", "share"))
        for c in extra_cols:
            f.write(" %16s" % c.split("__")[-1].replace(".sum", ""))
        f.write("\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
            f.write("%-78s %5d %12.1f %6.1f%%" % (name[:78], a["n"], a["gpu__time_duration.sum"] / 1e3,
                                                   100 * a["gpu__time_duration.sum"] / tot))
            for c in extra_cols:
                f.write(" %16.3e" % a.get(c, 0.0))
            f.write("\n")
        f.write("total %d launches, %.1f us (serialised under ncu)\n" % (len(ks), tot / 1e3))
    return agg, tot


os.makedirs(P, exist_ok=True)
for fn in ("bench.json", "bench_reference_arm.json"):
    src = os.path.join(G, "%s_%s" % (tag, fn))
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, "%s_%s" % (tag, fn)))
traffic = {}
tp = os.path.join(P, "traffic.json")
if os.path.exists(tp):
    traffic = json.load(open(tp))
src = os.path.join(G, tag + "_launches_office.csv")
if os.path.exists(src):
    launch_summary(src, os.path.join(P, tag + "_launches_office.txt"),
                   "ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --steps 4 --warmup 3 --no-esdf512 "
                   "(office replan, BASELINE config 2), launches 40..79 (resident replans): time per kernel")
src = os.path.join(G, tag + "_esdf512_launches.csv")
if os.path.exists(src):
    cols = ("dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum")
    agg, tot = launch_summary(src, os.path.join(P, tag + "_esdf512_launches.txt"),
                              "ONE whole 512^3 ESDF update (pillar V1, second update of the run = steady state), "
                              "ncu --cache-control none --clock-control none: every kernel", cols)
    rd = sum(a.get("dram__bytes_read.sum", 0) for a in agg.values())
    wr = sum(a.get("dram__bytes_write.sum", 0) for a in agg.values())
    with open(os.path.join(P, tag + "_esdf512_launches.txt"), "a") as f:
        f.write("DRAM read %.1f MB + write %.1f MB = %.1f MB per update; algorithmic 671.1 MB (5 B/voxel): ratio %.2f\n"
                % (rd / 1e6, wr / 1e6, (rd + wr) / 1e6, (rd + wr) / 671088640.0))
    traffic["esdf512_v1"] = {"bytes": rd + wr, "read": rd, "write": wr,
                             "source": "profiles/%s_esdf512_launches.txt (sum over the 33 kernels of one update)" % tag}
src = os.path.join(G, tag + "_frontier512_launches.csv")
if os.path.exists(src):
    cols = ("dram__bytes_read.sum", "dram__bytes_write.sum")
    agg, tot = launch_summary(src, os.path.join(P, tag + "_frontier512_launches.txt"),
                              "tools/frontier512.py (512^3 frontier search, large multi-kernel path), launches 239..357 = "
                              "the third search of the run; kernels of a split level that has nothing left to do return at once", cols)
    rd = sum(a.get("dram__bytes_read.sum", 0) for a in agg.values())
    wr = sum(a.get("dram__bytes_write.sum", 0) for a in agg.values())
    with open(os.path.join(P, tag + "_frontier512_launches.txt"), "a") as f:
        f.write("DRAM read %.1f MB + write %.1f MB over the window; algorithmic 268.4 MB (2 B/voxel)\n" % (rd / 1e6, wr / 1e6))
    traffic["frontier512"] = {"bytes": rd + wr, "read": rd, "write": wr,
                              "source": "profiles/%s_frontier512_launches.txt (sum over the launches of one search)" % tag}
rep = os.path.join(G, tag + "_esdf512_full.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "pipe_alu", "pipe_fma",
                          "l1tex__data_bank_conflicts", "smsp__inst_executed_op_shared"],
                         capture_output=True, text=True).stdout
    open(os.path.join(P, tag + "_esdf512_ncu_full.txt"), "w").write(
        "ncu --set full --cache-control none --clock-control none, tools/esdf512.py V1 2, kernels 33..35 "
        "(zpack, zy tile, x tile of the second update)\n" + out)
so = os.path.join(ROOT, "fuel_b200", "build", "esdf_tile.o")
if os.path.exists(so):
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines()
    keep, fn = [], None
    for l in sass:
        if "Function :" in l:
            fn = l.strip()
        if any(t in l for t in ("UBLKCP", "SYNCS", "CCTL.E.RML2", "UTMA", "UCGABAR")):
            if fn:
                keep.append(fn)
                fn = None
            keep.append(l.rstrip())
    open(os.path.join(P, tag + "_esdf_tile_sass_tma.txt"), "w").write(
        "cuobjdump -sass fuel_b200/build/esdf_tile.o | grep UBLKCP|SYNCS|CCTL.E.RML2|UCGABAR -- the bulk-async copies "
        "(cp.async.bulk -> UBLKCP.S.G), their mbarrier traffic (SYNCS.*), the L2 discards of the consumed partial "
        "(discard.global.L2 -> CCTL.E.RML2) and the cluster barriers of the 2-CTA long-line tiles (barrier.cluster -> "
        "UCGABAR_ARV / UCGABAR_WAIT; their DSMEM accesses are the LD.E / ST.E.U16 through mapa addresses) in the tile kernels\n" + "\n".join(keep) + "\n")
json.dump(traffic, open(tp, "w"), indent=1)
print("profiles/ updated:", sorted(os.listdir(P)))
