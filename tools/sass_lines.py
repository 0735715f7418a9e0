#!/usr/bin/env python
"""Per-source-line and per-opcode instruction counts of one ncu capture.

usage: tools/sass_lines.py <report.ncu-rep> <lib.so> [kernel-substring] [passes-per-launch]
Joins `ncu --page source --print-source sass` (executed counts per SASS instruction) with
`nvdisasm -g` line info of the cubin embedded in <lib.so> (the library must be the build that was profiled)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib = sys.argv[1:3]
want = sys.argv[3] if len(sys.argv) > 3 else "k_env_step32"
passes = float(sys.argv[4]) if len(sys.argv) > 4 else 40.0

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = rows[1]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
body = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":  # a second captured launch follows: keep the first
        break
    body.append(r)
sass = [(r[iS].strip(), int(r[iE] or 0), int(r[iSm] or 0)) for r in body if len(r) > iE]
warps = sass[0][1]  # the first instruction runs once per warp
print('stall samples total', sum(x for _, _, x in sass))

# line info from the cubin
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
lines = []
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    cur, infn, fn_lines = None, False, []
    for ln in out.splitlines():
        m = re.match(r"\s*\.text\.(\S+):", ln)
        if m:
            if infn and fn_lines:
                break
            infn = want in m.group(1) and "<false>" not in ln and ("ILi0ELi4ELb0" in m.group(1) or "ILb0EEE" in m.group(1))
            fn_lines = []
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inl = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur = (os.path.basename(m.group(1)), int(m.group(2)), tuple((os.path.basename(a), int(b)) for a, b in inl))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            fn_lines.append((m.group(2).strip(), cur))
    if infn and fn_lines:
        lines = fn_lines
        break
    lines = fn_lines if fn_lines else lines

print(f"sass rows {len(sass)}, disasm rows {len(lines)}, warps {warps}")
tot = sum(n for _, n, _ in sass)
print(f"warp instructions per launch {tot}  = {tot / warps / passes:.1f} per warp per pass")
ops = collections.Counter()
for s, n, _ in sass:
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", s)
    ops[m.group(2) if m else s] += n
print("opcode mix (per warp per pass):", ", ".join(f"{k} {v / warps / passes:.1f}" for k, v in ops.most_common(24)))
if len(lines) == len(sass):
    by = collections.Counter(); smp = collections.Counter()
    for (s, n, sm), (_, loc) in zip(sass, lines):
        key = (loc[0], loc[1]) if loc else ("?", 0)
        by[key] += n; smp[key] += sm
    src = {}
    print("\nper source line (innermost), instructions per warp per pass / stall samples:")
    for (f, l), n in by.most_common(70):
        if f not in src:
            p = os.path.join(os.path.dirname(os.path.abspath(lib)), "..", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src[f][l - 1].strip()[:110] if 0 < l <= len(src[f]) else ""
        print(f"{n / warps / passes:7.2f} {smp[(f, l)]:6d}  {f}:{l}  {text}")
else:
    print("line info does not match the capture (different build?)")
