#!/usr/bin/env python
"""Turn the artefacts of one tools/gpu_r2_profile.sh visit (gpurun_out/*_<tag>*) into the tracked summaries under
profiles/ (round 2, headline workload cfg3 = ive + kve pair over 2^28 points):
  r2_launch_list_summary.csv  ncu launch list of the bench command: per-kernel share of the step
  r2_ncu_full_kernels.csv     ncu --set full table of every pipeline kernel of one steady-state chunk (2^24 points)
  r2_traffic.json             DRAM bytes per launch of each kernel (bench.py's roofline.traffic)
  r2_lanes_<kernel>.txt       where each dominant kernel loses lanes (tools/ncu_lanes.py)
  r2_hot_<kernel>.txt         opcode mix, stall reasons, hottest lines (tools/ncu_hot.py)
usage: python tools/make_profiles_r2.py <tag>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_profiles import launch_summary  # noqa: E402

# bench.py's kernel names (spb profile records of the pair call) -> ncu kernel names of the same launches
BENCH_NAMES = {
    "IK/ipass_debye": ["k_debye<"],
    "IK/ipass_wrsk": ["k_ipass<3,"],
    "IK/kpass": ["k_kpass<"],
    "IK/classify": ["k_classify<"],
    "IK/ipass_mlri": ["k_ipass<2,"],
    "IK/ipass_asyi": ["k_ipass<1,"],
    "IK/ipass_series": ["k_ipass<0,"],
    "IK/ipass_uni1": ["k_ipass<4,"],
    "IK/ipass_uni2": ["k_ipass<5,"],
    "IK/scan_scatter": ["k_scan(", "k_scatter(", "k_cell_hist(", "k_cell_scan(", "k_cell_scatter("],
    "IK/general": ["k_general("],
}
TOP = (("k_debye", "k_debye.o", "_ZN4spbk7k_debyeILb0ELi7ELi2E", "debye"),
       ("k_ipass", "k_fused.o", "_ZN4spbk7k_ipassILi3ELb0ELi7ELi2E", "wrsk"),
       ("k_kpass", "k_kpass.o", "_ZN4spbk7k_kpassILi7ELi2ELb0E", "kpass"),
       ("k_classify", "k_pipeline.o", "_ZN4spbk10k_classifyILi7ELi2E", "classify"))


def main():
    tag = sys.argv[1]
    go, prof = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
    print(launch_summary(os.path.join(go, f"launches_{tag}.csv"), os.path.join(prof, "r2_launch_list_summary.csv")))
    table = os.path.join(prof, "r2_ncu_full_kernels.csv")
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"),
                    os.path.join(go, f"full_{tag}_chunk_raw.csv"), table], check=True)
    rows = list(csv.reader(open(table)))
    h = rows[0]
    unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
    rd = [i for i, c in enumerate(h) if c.startswith("dram_read")][0]
    wr = [i for i, c in enumerate(h) if c.startswith("dram_write")][0]
    tu = {"us": 1.0, "ms": 1e3, "ns": 1e-3}[h[1].split("[")[1].rstrip("]")]
    per = {}
    for r in rows[1:]:
        if r[0] not in per:
            per[r[0]] = {"dram_read_bytes": float(r[rd]) * unit[h[rd].split("[")[1].rstrip("]")],
                         "dram_write_bytes": float(r[wr]) * unit[h[wr].split("[")[1].rstrip("]")],
                         "time_us": float(r[1]) * tu}
    kernels = {}
    for name, pats in BENCH_NAMES.items():
        # (kernel names with or without the (int) / (bool) casts of the demangler)
        hit = [(k, v) for k, v in per.items() if any(p in k.replace("(int)", "").replace("(bool)", "") for p in pats)]
        if hit:
            kernels[name] = {"bytes": sum(v["dram_read_bytes"] + v["dram_write_bytes"] for _, v in hit),
                             "dram_read_bytes": sum(v["dram_read_bytes"] for _, v in hit),
                             "dram_write_bytes": sum(v["dram_write_bytes"] for _, v in hit),
                             "ncu_time_us": sum(v["time_us"] for _, v in hit), "ncu_kernels": [k for k, _ in hit]}
    json.dump({"_source": f"ncu --set full of one steady-state chunk (2^24 points) of the bench step, visit '{tag}' "
                          "(tools/gpu_r2_profile.sh); dram__bytes_read.sum + dram__bytes_write.sum per launch",
               "kernels": kernels}, open(os.path.join(prof, "r2_traffic.json"), "w"), indent=1)
    rep = os.path.join(go, f"full_{tag}_top.ncu-rep")
    build = os.path.join(ROOT, "sciport_rs_b200", "csrc", "build")
    for kern, obj, mangled, label in TOP:
        o = os.path.join(build, obj)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lanes.py"), rep, kern, o, mangled, "4", "0",
                              "0.01"], capture_output=True, text=True).stdout
        open(os.path.join(prof, f"r2_lanes_{label}.txt"), "w").write(out)
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, kern, "0"],
                             capture_output=True, text=True).stdout
        open(os.path.join(prof, f"r2_hot_{label}.txt"), "w").write(out)
    print("profiles written")


if __name__ == "__main__":
    main()
