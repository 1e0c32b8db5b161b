#!/usr/bin/env python
"""Where does a kernel lose lanes?  Joins the per-instruction counters of an .ncu-rep (--import-source on) with the
inline call chains of the same build (nvdisasm -gi on the object) and prints, per source-level phase (the first
DEPTH call sites of the chain below the kernel), the share of executed warp instructions, the average number of
active lanes and the issue slots that full warps would have saved.
usage: python tools/ncu_lanes.py <rep> <kernel-regex> <object.o> <mangled-substring> [depth] [launch-index] [min-share]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def export(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                          "--print-source", "sass"], capture_output=True, text=True).stdout
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = [line]
        else:
            cur.append(line)
    if cur:
        blocks.append(cur)
    return [(b[0], list(csv.reader(io.StringIO("\n".join(b[1:]))))) for b in blocks]


def chains(obj, mangled):
    """offset -> list of 'file:line' frames, outermost call site first, for the function whose name contains `mangled`."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    where, frames, cur, active = {}, [], ["?"], False
    for line in dis.splitlines():
        if line.startswith(".text."):
            active = mangled in line
            frames = []
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            frames.append(f"{os.path.basename(m.group(1))}:{m.group(2)}")
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", line)
        if m:
            if frames:
                cur = list(reversed(frames))  # nvdisasm lists the innermost frame first
                frames = []
            where[int(m.group(1), 16)] = cur
    return where


def main():
    rep, kernel, obj, mangled = sys.argv[1:5]
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    launch = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    min_share = float(sys.argv[7]) if len(sys.argv) > 7 else 0.01
    name, rows = export(rep, kernel)[launch]
    h = rows[0]
    ia, iex, ith = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    data = [(int(r[ia], 16), int(r[iex]), int(r[ith])) for r in rows[1:] if len(r) > ith]
    base = data[0][0]
    where = chains(obj, mangled)
    agg = collections.OrderedDict()
    tot_w = sum(d[1] for d in data)
    tot_t = sum(d[2] for d in data)
    for a, w, t in data:
        fr = where.get(a - base, ["?"])
        # drop the kernel's own frame(s) at the front only when something is inlined below them
        key = " > ".join(fr[:depth])
        e = agg.setdefault(key, [0, 0])
        e[0] += w
        e[1] += t
    print(name.strip('",'))
    print(f"warp instructions {tot_w}, average active lanes {tot_t / max(tot_w, 1):.2f}; "
          f"issue slots full warps would save: {1 - tot_t / 32 / max(tot_w, 1):.3f}")
    print("inst share | lanes | slots lost (share of all) | phase (call sites, outermost first)")
    for key, (w, t) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if w / tot_w < min_share:
            continue
        print(f"   {w / tot_w:6.3f}   {t / max(w, 1):5.1f}   {(w - t / 32) / tot_w:6.3f}   {key}")


if __name__ == "__main__":
    main()
