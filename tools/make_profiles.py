#!/usr/bin/env python
"""Turn the artefacts of one tools/gpu_check.sh visit (gpurun_out/*_<tag>.*) into the tracked summaries under
profiles/: launch list (per-kernel share of the bench step), ncu --set full table, source-level hot spots of the
two dominant kernels, and the DRAM traffic per launch that bench.py reports as roofline.traffic.
usage: python tools/make_profiles.py <tag> [round-prefix]"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def launch_summary(path, out):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, rows = r, rows[i + 1:]
            break
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows:
        if len(r) > vi:
            a = agg.setdefault(r[ki], [0, 0.0])
            a[0] += 1
            a[1] += float(r[vi].replace(",", ""))
    ours = {k: v for k, v in agg.items() if "k_dfma" not in k and "at::" not in k}
    tot = sum(v[1] for v in ours.values())
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_pipeline_kernels"])
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            w.writerow([k, v[0], f"{v[1] / 1e3:.1f}", f"{v[1] / 1e3 / v[0]:.1f}", f"{v[1] / tot:.4f}"])
        for k, v in agg.items():
            if k not in ours:
                w.writerow([k + "  (not part of the step: peak microbenchmark / torch fill)", v[0], f"{v[1] / 1e3:.1f}",
                            f"{v[1] / 1e3 / v[0]:.1f}", ""])
    return out


def main():
    tag = sys.argv[1]
    pre = sys.argv[2] if len(sys.argv) > 2 else "r1"
    go = os.path.join(ROOT, "gpurun_out")
    prof = os.path.join(ROOT, "profiles")
    os.makedirs(prof, exist_ok=True)
    print(launch_summary(os.path.join(go, f"launches_{tag}.csv"), os.path.join(prof, f"{pre}_launch_list_summary.csv")))
    rep = os.path.join(go, f"full_{tag}.ncu-rep")
    table = os.path.join(prof, f"{pre}_ncu_full_kernels.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, table], check=True)
    # traffic per launch (bytes) of each kernel, first occurrence
    rows = list(csv.reader(open(table)))
    h = rows[0]
    name_i = 0
    rd = [i for i, c in enumerate(h) if c.startswith("dram_read")][0]
    wr = [i for i, c in enumerate(h) if c.startswith("dram_write")][0]
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    traffic = {}
    for r in rows[1:]:
        if r[name_i] in traffic:
            continue
        ur = unit[h[rd].split("[")[1].rstrip("]")]
        uw = unit[h[wr].split("[")[1].rstrip("]")]
        traffic[r[name_i]] = {"dram_read_bytes": float(r[rd]) * ur, "dram_write_bytes": float(r[wr]) * uw,
                              "time_us": float(r[1])}
    json.dump({"_source": f"ncu --set full capture {os.path.basename(rep)} (per launch, cfg2 bench step)",
               "kernels": traffic}, open(os.path.join(prof, f"{pre}_traffic.json"), "w"), indent=1)
    # source-level views of the dominant kernels of the cfg2 step: the fused I+K kernel of the large-|w| region
    # (second k_ipass launch of a step) and the classifier; hot lines (ncu_hot) and cost by inline call path (ncu_tree)
    build = os.path.join(ROOT, "sciport_rs_b200", "csrc", "build")
    for kern, obj, mangled, launch, label in (("k_ipass", "k_fused.o", "k_ipassILi1ELb1ELi6ELi1", 1, "fused_asyi_k"),
                                              ("k_classify", "k_pipeline.o", "k_classifyILi6ELi1", 0, "classify")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, kern, "30",
                              os.path.join(build, obj), str(launch), mangled], capture_output=True, text=True).stdout
        open(os.path.join(prof, f"{pre}_hot_{label}.txt"), "w").write(out)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_tree.py"), rep, kern,
                              os.path.join(build, obj), mangled, "5", str(2 * launch), "0.015"],
                             capture_output=True, text=True).stdout
        open(os.path.join(prof, f"{pre}_tree_{label}.txt"), "w").write(out)
    print("profiles written")


if __name__ == "__main__":
    main()
