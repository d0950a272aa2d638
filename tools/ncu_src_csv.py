#!/usr/bin/env python
"""Per-instruction / per-opcode stall breakdown of one kernel from an exported ncu source page
(`ncu -i rep --page source --csv --kernel-id ::regex:NAME:1 > file.csv`, see tools/gpu_measure_r02.sh).

    python tools/ncu_src_csv.py gpurun_out/r02_src_k_ac_frames_w.csv [top_n]
"""
import collections
import csv
import sys


def first_table(rows, hdr, c):
    """ncu repeats the table when the kernel-id matches more than one view: keep the first."""
    data = []
    for r in rows[2:]:
        if len(r) < len(hdr) - 2:
            if data:
                break
            continue
        if r[c["Address"]] == "Address":
            break
        if r[c["# Samples"]] != "":
            data.append(r)
    return data


def main():
    path = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    c = {h: i for i, h in enumerate(hdr)}
    data = first_table(rows, hdr, c)

    def iv(r, k):
        try:
            return int(r[c[k]] or 0)
        except (ValueError, IndexError):
            return 0

    tot = sum(iv(r, "# Samples") for r in data) or 1
    print(rows[0][1][:100])
    print("total samples", tot, "SASS instructions", len(data), "warp instructions executed", sum(iv(r, "Instructions Executed") for r in data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.Counter()
    for r in data:
        for s in stalls:
            agg[s] += iv(r, s)
    print("stall reasons:", ", ".join(f"{k[6:]} {100 * v / tot:.0f}%" for k, v in agg.most_common(8)))
    byop = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for r in data:
        src = r[c["Source"]].split()
        if not src:
            continue
        op = src[1] if src[0].startswith("@") and len(src) > 1 else src[0]
        op = ".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "ATOM", "RED")) else op.split(".")[0]
        b = byop[op]
        b[0] += iv(r, "# Samples")
        b[1] += iv(r, "Instructions Executed")
        for s in stalls:
            b[2][s] += iv(r, s)
    print(f"{'opcode':12s} {'samples':>8s} {'warp inst':>12s}  top stalls")
    for op, (n, inst, st) in sorted(byop.items(), key=lambda kv: -kv[1][0])[:top_n]:
        print(f"{op:12s} {100 * n / tot:7.1f}% {inst:12d}  " + ", ".join(f"{k[6:]} {100 * v / max(n, 1):.0f}%" for k, v in st.most_common(3)))
    loc = [r for r in data if r[c["Address Space"]] == "Local"]
    print("local-memory instructions:", len(loc), "executed", sum(iv(r, "Instructions Executed") for r in loc))
    print("hottest instructions:")
    for r in sorted(data, key=lambda r: -iv(r, "# Samples"))[:top_n]:
        st = sorted(((iv(r, s), s[6:]) for s in stalls), reverse=True)[:2]
        print(f"  {r[c['Address']][-6:]} {100 * iv(r, '# Samples') / tot:5.1f}%  {r[c['Source']][:64]:64s} " + ", ".join(f"{k} {v}" for v, k in st))


def by_line(path, cub, mangled, top_n=30):
    """Attribute the samples to CUDA source lines (needs the library built from the same sources: -lineinfo)."""
    import os
    import re
    import subprocess
    import tempfile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "robust_speech_analysis_framework_b200", "libmshds_b200.so")], cwd=tmp,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    off2line = {}
    for cubin in sorted(os.listdir(tmp)):
        if cub not in cubin:
            continue
        sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
        func = None
        cur = None
        for l in sass.split("\n"):
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                func = m.group(1)
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
            if m and func and mangled in func:
                off2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    c = {h: i for i, h in enumerate(hdr)}
    data = first_table(rows, hdr, c)
    base = int(data[0][c["Address"]], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    tot = 0
    seen = set()
    for r in data:
        a = int(r[c["Address"]], 16) - base
        if a in seen:
            continue
        seen.add(a)
        n = int(r[c["# Samples"]] or 0)
        fl = off2line.get(a) or ("?", 0)
        e = agg[fl]
        e[0] += n
        e[1] += int(r[c["Instructions Executed"]] or 0)
        for s_ in stalls:
            e[2][s_] += int(r[c[s_]] or 0)
        tot += n
    src = {}
    csrc = os.path.join(root, "robust_speech_analysis_framework_b200", "csrc")
    for fn in os.listdir(csrc):
        if os.path.isfile(os.path.join(csrc, fn)):
            src[fn] = open(os.path.join(csrc, fn)).read().split("\n")
    print("by source line (share of samples, warp instructions, top stalls):")
    for (fn, ln), (n, inst, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
        text = src[fn][ln - 1].strip()[:90] if fn in src and 0 < ln <= len(src[fn]) else ""
        print(f"{100 * n / max(tot, 1):5.1f}% inst={inst:>11} {fn}:{ln} [" + ", ".join(f"{k[6:]} {100 * v / max(n, 1):.0f}%" for k, v in st.most_common(2)) + f"]  {text}")


if __name__ == "__main__":
    if len(sys.argv) > 3 and not sys.argv[2].isdigit():
        by_line(sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 30)
    else:
        main()
