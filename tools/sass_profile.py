#!/usr/bin/env python3
"""Attributes an `ncu --page source --csv` (SASS view) capture to CUDA source lines without the GUI.

    python tools/sass_profile.py <object.o> <function substring> <ncu_source.csv> <kernel substring> [top N]

The CSV holds, per kernel launch, one row per SASS instruction (executed count, sampled stalls) but no source line; `nvdisasm -g`
of the same object holds the source line of every instruction. The two are joined by instruction offset. Output: warp
instructions executed, lanes per instruction and stall samples per source line (innermost inlined line), largest first.
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

csv.field_size_limit(10 ** 9)
obj, fsub, ncu_csv, ksub = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40

tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
line_of = {}          # offset -> "file:line"
cur = None; infn = False
for l in dis:
    if l.startswith(".text."):
        infn = fsub in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = f"{os.path.basename(m.group(1))}:{m.group(2)}"
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2).strip())

text = open(ncu_csv).read().split("\n")
starts = [i for i, l in enumerate(text) if l.startswith('"Kernel Name"')]
sec = [i for i in starts if ksub in text[i]]
if not sec:
    sys.exit(f"no kernel matching {ksub!r} in {ncu_csv}")
s0 = sec[0]; s1 = min([i for i in starts if i > s0] + [len(text)])
rows = list(csv.reader(io.StringIO("\n".join(text[s0 + 1:s1]))))
hdr = rows[0]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]; missing = 0
for r in rows[1:]:
    if len(r) <= it or not r[ii].isdigit():
        continue
    addr = int(r[ia], 16)
    if base is None:
        base = addr
    src = line_of.get(addr - base)
    if src is None:
        missing += 1; key = "?"
    else:
        key = src[0]
    for k, v in enumerate((int(r[ii]), int(r[it]), int(r[isamp]))):
        agg[key][k] += v; tot[k] += v
print(f"{text[s0][:120]}\nwarp instructions {tot[0]:,}  lanes/instruction {tot[1] / max(tot[0], 1):.1f}  samples {tot[2]:,}  (unmatched SASS rows: {missing})")
print(f"{'source line':34s} {'warp inst':>14s} {'share':>7s} {'lanes':>6s} {'samples':>9s} {'share':>7s}")
for key, (a, b, c) in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    print(f"{key:34s} {a:14,d} {100 * a / tot[0]:6.1f}% {b / max(a, 1):6.1f} {c:9,d} {100 * c / max(tot[2], 1):6.1f}%")
