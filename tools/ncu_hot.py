#!/usr/bin/env python
"""Source-level view of one kernel from an .ncu-rep (captured with --import-source on, built with -lineinfo):
dynamic opcode mix, stall reasons, and the hottest source lines.
usage: python tools/ncu_hot.py gpurun_out/full.ncu-rep k_kpass [top_n] [object.o]
With the object file (same build as the capture) the SASS is mapped back to file:line through nvdisasm -g."""
import collections
import csv
import io
import subprocess
import sys


def export(rep, kernel, view):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}",
                          "--print-source", view], capture_output=True, text=True).stdout
    blocks, cur = [], []
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            if cur:
                blocks.append(cur)
            cur = []
        else:
            cur.append(line)
    if cur:
        blocks.append(cur)
    return [list(csv.reader(io.StringIO("\n".join(b)))) for b in blocks]


def by_line(rows, h, obj, kernel, top):
    """Aggregate executed instructions / stall samples per source line using nvdisasm -g on the object."""
    import os
    import re
    import tempfile
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    where, cur, active = {}, ("?", 0), False
    for line in dis.splitlines():
        if line.startswith("\t.section") or ".text." in line and line.strip().startswith("//-----"):
            active = kernel in line
        m = re.match(r'\s*//## File "(.*)", line (\d+)', line)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
        if m and active:
            where[int(m.group(1), 16)] = cur
    ai, si, ii = h.index("Address"), h.index("# Samples"), h.index("Instructions Executed")
    base = min(int(r[ai], 16) for r in rows[1:] if len(r) > ii)
    ti = h.index("Thread Instructions Executed")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r in rows[1:]:
        if len(r) <= max(ii, ti):
            continue
        loc = where.get(int(r[ai], 16) - base, ("(library/unknown)", 0))
        agg[loc][0] += int(r[si] or 0)
        agg[loc][1] += int(r[ii] or 0)
        agg[loc][2] += int(r[ti] or 0)
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    files = collections.defaultdict(lambda: [0, 0])
    for (f, l), v in agg.items():
        files[f][0] += v[0]
        files[f][1] += v[1]
    print("per file (samples share, inst share):")
    for f, v in sorted(files.items(), key=lambda kv: -kv[1][0]):
        print(f"  {f:40s} {v[0] / tot_s:6.3f} {v[1] / tot_i:6.3f}")
    print("hottest lines (samples share, inst share, avg active lanes, file:line):")
    for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][0] - kv[1][1] * tot_s / tot_i)[:top]:
        print(f"  {v[0] / tot_s:6.3f} {v[1] / tot_i:6.3f} {v[2] / max(v[1], 1):5.1f}  {f}:{l}")


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    # optional: which matching launch (block index) and which mangled-name substring selects the SASS section
    block = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    mangled = sys.argv[6] if len(sys.argv) > 6 else kernel
    rows = export(rep, kernel, "sass")[block]
    h = rows[0]
    ci, si, ii = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    ops = collections.Counter()
    samples = collections.Counter()
    total_i = total_s = 0
    for r in rows[1:]:
        if len(r) <= ii:
            continue
        toks = r[ci].split()
        op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
        op = op.split(".")[0]
        n, s = int(r[ii] or 0), int(r[si] or 0)
        ops[op] += n
        samples[op] += s
        total_i += n
        total_s += s
    print(f"kernel {kernel}: {total_i} warp instructions, {total_s} stall samples")
    print("opcode        warp_inst   share   samples share")
    for op, n in ops.most_common(22):
        print(f"{op:12s} {n:11d}  {n / total_i:6.3f}  {samples[op]:7d} {samples[op] / max(total_s, 1):6.3f}")
    stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
    tot = collections.Counter()
    for r in rows[1:]:
        for c in stall_cols:
            try:
                tot[c] += int(r[h.index(c)] or 0)
            except (ValueError, IndexError):
                pass
    print("stalls:", ", ".join(f"{k[6:]}={v / max(total_s, 1):.2f}" for k, v in tot.most_common(8)))
    if len(sys.argv) > 4:
        by_line(rows, h, sys.argv[4], mangled, top)
    return
    try:
        src = export(rep, kernel, "cuda")[0]
    except Exception:
        return
    h2 = src[0]
    if "# Samples" not in h2:
        return
    s2, i2, c2 = h2.index("# Samples"), h2.index("Instructions Executed"), h2.index("Source")
    a2 = h2.index("Address") if "Address" in h2 else 0
    lines = [(int(r[s2] or 0), int(r[i2] or 0), r[a2], r[c2]) for r in src[1:] if len(r) > max(s2, i2) and (r[s2] or "0").isdigit()]
    lines.sort(reverse=True)
    print("hottest source lines (samples, warp_inst, line, text)")
    for s, n, a, c in lines[:top]:
        print(f"{s:7d} {n:10d}  {a:>6s}  {c.strip()[:110]}")


if __name__ == "__main__":
    main()
