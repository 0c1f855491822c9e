#!/usr/bin/env python
"""Executed warp-instructions and stall samples per SOURCE LINE of one kernel, from an ncu report captured with
--import-source on (SASS view) joined with nvdisasm line info of the object file.
usage: python tools/ncu_lines.py report.ncu-rep build/obj.o kernel-substring [kernel-id] [top N] [divisor]"""
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

rep, obj, key = sys.argv[1], sys.argv[2], sys.argv[3]
kid = sys.argv[4] if len(sys.argv) > 4 else "1"
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
div = float(sys.argv[6]) if len(sys.argv) > 6 else 1.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cub = glob.glob(os.path.join(tmp, "*.cubin"))[0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
funcs, cur, line = {}, None, None
for l in dis.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        cur, line = m.group(1), None
        funcs[cur] = []
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,8}\*/", l):
        funcs[cur].append(line)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[2:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) > iI:
        body.append(r)
fn = [k for k in funcs if key in k]
fn = [k for k in fn if len(funcs[k]) == len(body)] or fn
lines = funcs[fn[0]]
print("kernel", rows[0][1][:90], "| sass rows", len(body), "| disasm rows", len(lines))
tot = sum(int(r[iI]) for r in body)
ts = sum(int(r[iS]) for r in body)
per = {}
for r, ln in zip(body, lines):
    d = per.setdefault(ln, [0, 0])
    d[0] += int(r[iI])
    d[1] += int(r[iS])
srcs = {}


def src(f, n):
    if f not in srcs:
        p = glob.glob(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f))
        srcs[f] = open(p[0]).read().split("\n") if p else []
    s = srcs[f]
    return s[n - 1].strip()[:78] if 0 < n <= len(s) else ""


print("total warp instr %d (%.1f per unit), samples %d" % (tot, tot / div, ts))
for ln, (c, sm) in sorted(per.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-26s %9.1f instr  %5.1f%% samples  %s" % ("%s:%d" % ln if ln else "?", c / div, 100 * sm / max(ts, 1),
                                                     src(*ln) if ln else ""))
