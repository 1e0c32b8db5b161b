#!/usr/bin/env python
"""Inclusive cost of one kernel by INLINE CALL PATH (which phase of the point routine the instructions belong to),
from an .ncu-rep captured with --import-source on and the object file of the same build (-lineinfo).
usage: python tools/ncu_tree.py <rep> <kernel-regex> <object.o> <mangled-substring> [depth] [launch-block] [min-share]
The SASS of the kernel is mapped to its chain of call sites through `nvdisasm -gi`; every executed warp
instruction and stall sample is charged to each prefix of its chain (outermost call site first)."""
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
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    where, active, pending, chain_open = {}, False, [], False
    for line in dis.splitlines():
        if line.startswith("\t.section"):
            active = mangled in line
            continue
        m = re.match(r'\s*//## File "([^"]*)", line (\d+)(?: inlined at "([^"]*)", line (\d+))?', line)
        if m:
            if m.group(3) is None:
                # a chain of "inlined at" lines ends with its outermost call site on a plain line
                here = (os.path.basename(m.group(1)), int(m.group(2)))
                if not (chain_open and pending and pending[-1] == here):
                    pending = [here]
                chain_open = False
            else:
                if not chain_open:
                    pending = []
                chain_open = True
                # innermost first; each following line repeats the caller and adds its own caller
                if not pending or pending[-1] != (os.path.basename(m.group(1)), int(m.group(2))):
                    pending = [(os.path.basename(m.group(1)), int(m.group(2)))]
                pending.append((os.path.basename(m.group(3)), int(m.group(4))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S+)", line)
        if m:
            chain_open = False
        if m and active:
            where[int(m.group(1), 16)] = (tuple(reversed(pending)), m.group(2))
    return where


def main():
    rep, kernel, obj, mangled = sys.argv[1:5]
    depth = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    block = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    min_share = float(sys.argv[7]) if len(sys.argv) > 7 else 0.01
    name, rows = export(rep, kernel)[block]
    h = rows[0]
    ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    where = chains(obj, mangled)
    base = min(int(r[ai], 16) for r in rows[1:] if len(r) > ii)
    tree = collections.defaultdict(lambda: [0, 0, 0])  # path -> [samples, inst, fp64 inst]
    tot = [0, 0, 0]
    for r in rows[1:]:
        if len(r) <= ii:
            continue
        chain, op = where.get(int(r[ai], 16) - base, ((("?", 0),), "?"))
        n, s = int(r[ii] or 0), int(r[si] or 0)
        f = n if re.match(r"D(FMA|MUL|ADD|SETP)", op) else 0
        tot[0] += s
        tot[1] += n
        tot[2] += f
        for d in range(1, min(depth, len(chain)) + 1):
            t = tree[chain[:d]]
            t[0] += s
            t[1] += n
            t[2] += f
    print(name)
    print(f"total warp instructions {tot[1]}, fp64 share {tot[2] / max(tot[1], 1):.3f}, samples {tot[0]}")
    print("inst share | sample share | fp64 share of node | call path (outermost first)")
    for path in sorted(tree, key=lambda p: tuple((-tree[p[:d + 1]][1], p[d]) for d in range(len(p)))):
        s, n, f = tree[path]
        if n / max(tot[1], 1) < min_share:
            continue
        print(f"  {n / tot[1]:6.3f} {s / max(tot[0], 1):6.3f} {f / max(n, 1):6.3f}  " + "  " * (len(path) - 1) +
              f"{path[-1][0]}:{path[-1][1]}")


if __name__ == "__main__":
    main()
