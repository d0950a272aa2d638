"""Development aid: attribute ncu warp-stall samples of one kernel to CUDA source lines.
usage: ncu_lines.py <report.ncu-rep | exported source page .csv> <kernel-regex> <nth match (1-based)> <cubin-name-substring e.g. k_pitch> <mangled-substring> [top_n]"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kre, nth, cub, mangled = sys.argv[1:6]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "robust_speech_analysis_framework_b200", "libmshds_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith(cub + ".")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
func = None
line = file = None
off2line = {}
for l in sass.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        func = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        file = m.group(1).split("/")[-1]
        line = int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and func and mangled in func:
        off2line[int(m.group(1), 16)] = (file, line)
if rep.endswith(".csv"):
    out = open(rep).read()
else:
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{kre}:{nth}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.split("\n")))
print(rows[0][:2])
hdr = rows[1]
ai, ni, ii = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
base = int(rows[2][ai], 16)
agg, inst, tot = collections.Counter(), collections.Counter(), 0
seen = set()
for r in rows[2:]:
    try:
        a = int(r[ai], 16) - base
        if a in seen:           # ncu repeats the table when the kernel-id matches more than one view
            break
        seen.add(a)
        n = int(r[ni])
    except Exception:
        continue
    fl = off2line.get(a, ("?", 0))
    agg[fl] += n
    inst[fl] += int(r[ii])
    tot += n
src = {}
csrc = os.path.join(root, "robust_speech_analysis_framework_b200", "csrc")
for fn in os.listdir(csrc):
    if not os.path.isfile(os.path.join(csrc, fn)):
        continue
    src[fn] = open(os.path.join(csrc, fn)).read().split("\n")
print("total samples", tot, "total warp instructions", sum(inst.values()))
for (fn, ln), n in agg.most_common(int(sys.argv[6]) if len(sys.argv) > 6 else 28):
    text = src[fn][ln - 1].strip()[:110] if fn in src and 0 < ln <= len(src[fn]) else ""
    print(f"{100 * n / tot:5.1f}% inst={inst[(fn, ln)]:>11} {fn}:{ln}  {text}")
