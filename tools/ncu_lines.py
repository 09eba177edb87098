#!/usr/bin/env python
"""Per-source-line aggregation of an ncu report's SASS page, joined with nvdisasm line info of the built library.
usage: tools/ncu_lines.py report.ncu-rep [top_n] [--sass]   (run from the repo root)"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 40
show_sass = "--sass" in sys.argv
lib = "cerbos_b200/_lib/libcerbos_b200.so"

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
linemap = {}   # (func_key, addr) -> "file:line [inlined ...]"
func = None
cur = "?"
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        func = m.group(1)
        cur = "?"
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m and func:
        linemap[(func, int(m.group(1), 16))] = cur

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source=sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kernel = None
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        kernel = r[1]
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) >= 10:
        d = dict(zip(hdr, r))
        try:
            data.append((int(d["Address"], 16) if d["Address"].startswith("0x") else int(d["Address"]), d["Source"], int(d["Instructions Executed"] or 0), int(d["# Samples"] or 0),
                         int(d["Thread Instructions Executed"] or 0), d))
        except ValueError:
            pass
base = min(a for a, *_ in data)
# pick the function whose name fits the kernel (template arg)
m = re.search(r"\(bool\)(\d), \(int\)(\d)", kernel or "")
tag = f"ILb{m.group(1)}ELi{m.group(2)}E" if m else "check_kernel"
if "check_kernel_tiles" in (kernel or ""):
    tag = "check_kernel_tiles"
cands = [f for f in {k[0] for k in linemap} if "check_kernel" in f and "$" not in f and tag in f]
fn = cands[0]
# callee functions are laid out after the kernel in the same section group; ncu addresses are relative to kernel start.
tot = sum(x[2] for x in data)
ts = sum(x[3] for x in data)
agg = collections.defaultdict(lambda: [0, 0, 0])
unk = 0
for a, src, ie, sm, ti, d in data:
    key = linemap.get((fn, a - base))
    if key is None:
        key = "callee/unknown"
    agg[key][0] += ie
    agg[key][1] += sm
    agg[key][2] += ti
print(f"kernel {kernel[:80]}\ntotal warp-inst {tot}  samples {ts}")
for k, (ie, sm, ti) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k:>24}  inst {ie:>11} ({100*ie/max(tot,1):5.1f}%)  samp {sm:>7} ({100*sm/max(ts,1):5.1f}%)  thr/inst {ti/max(ie,1):5.1f}")
if show_sass:
    print("--- top SASS")
    for a, src, ie, sm, ti, d in sorted(data, key=lambda x: -x[3])[:top]:
        print(f"{a-base:6x} {linemap.get((fn, a-base),'?'):>22} inst {ie:>10} samp {sm:>6} thr {ti/max(ie,1):5.1f}  {src[:90]}")
if "--regions" in sys.argv:
    print("--- regions (cb_core.h line ranges)")
    import bisect
    regs = collections.defaultdict(lambda: [0, 0, 0])
    for k, (ie, sm, ti) in agg.items():
        m = re.match(r"(\S+):(\d+)", k)
        if not m:
            regs[k][0] += ie; regs[k][1] += sm; regs[k][2] += ti
            continue
        f, l = m.group(1), int(m.group(2))
        name = f if f != "cb_core.h" else f"cb_core.h:{l // 50 * 50}-{l // 50 * 50 + 49}"
        regs[name][0] += ie; regs[name][1] += sm; regs[name][2] += ti
    for k, (ie, sm, ti) in sorted(regs.items(), key=lambda kv: -kv[1][0]):
        print(f"{k:>28}  inst {ie:>11} ({100*ie/max(tot,1):5.1f}%)  samp {sm:>7} ({100*sm/max(ts,1):5.1f}%)  thr/inst {ti/max(ie,1):5.1f}")
