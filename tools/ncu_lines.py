#!/usr/bin/env python
"""Per-source-line view of one kernel of an ncu --set full report: joins the SASS page of the report (stall samples and executed
instructions per address) with `nvdisasm -gi` of the cubin built from the same sources (address -> innermost source line).
usage: ncu_lines.py <report.ncu-rep> <kernel-name> <engine.o> [top N]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_lines(obj, kernel):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], check=True, stdout=subprocess.PIPE, text=True).stdout
    out = {}
    inside = False
    frames = []
    fresh = True
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = (kernel in ln)
            continue
        if not inside:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            if fresh:
                frames = []
                fresh = False
            frames.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*)", ln)
        if m:
            out[int(m.group(1), 16)] = (list(frames), m.group(2))
            fresh = True
    return out


def main():
    rep, kernel, obj = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    hdr = rows[hi]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines = sass_lines(obj, kernel)
    base = None
    per_line = defaultdict(lambda: [0, 0, defaultdict(int)])
    per_outer = defaultdict(lambda: [0, 0])
    tot_s = tot_e = 0
    for r in rows[hi + 1:]:
        if len(r) <= iex or not r[ia].startswith("0x"):
            if len(r) > 0 and r[0] == "Kernel Name":
                break          # next launch of the same kernel
            continue
        addr = int(r[ia], 16)
        if base is None:
            base = addr
        off = addr - base
        frames, _ = lines.get(off, ([("?", 0)], ""))
        inner = frames[0] if frames else ("?", 0)
        s, e = int(r[isamp] or 0), int(r[iex] or 0)
        tot_s += s
        tot_e += e
        pl = per_line[inner]
        pl[0] += s
        pl[1] += e
        for i, h in stall_cols:
            v = int(r[i] or 0)
            if v:
                pl[2][h[6:]] += v
        # the outermost kernels_core.cuh frame (the kernel body line that called the helper)
        outer = [f for f in frames if f[0] == "kernels_core.cuh"]
        if outer:
            po = per_outer[outer[-1]]
            po[0] += s
            po[1] += e
    print("kernel %s: %d samples, %d warp instructions" % (kernel, tot_s, tot_e))
    print("== by innermost source line")
    for (f, l), (s, e, st) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        tops = ", ".join("%s %d" % (k, v) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print("%-18s %5d  samples %5.1f%%  instr %5.1f%%  [%s]" % (f, l, 100.0 * s / max(tot_s, 1), 100.0 * e / max(tot_e, 1), tops))
    print("== by kernel-body line (outermost kernels_core.cuh frame)")
    for (f, l), (s, e) in sorted(per_outer.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-18s %5d  samples %5.1f%%  instr %5.1f%%" % (f, l, 100.0 * s / max(tot_s, 1), 100.0 * e / max(tot_e, 1)))


if __name__ == "__main__":
    main()
