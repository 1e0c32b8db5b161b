#!/usr/bin/env python
"""Human-readable digest of a bench.py JSON line.  usage: python tools/show_bench.py gpurun_out/bench_x.json"""
import json
import sys

lines = [l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")]
if not lines:
    print("no JSON line")
    sys.exit(0)
d = json.loads(lines[-1])
print("value %.4g %s  ms/step %.3f  n_gpus %d  launches %s  enqueue %.2f ms" % (
    d["value"], d["unit"], d["ms_per_step"], d["n_gpus"], d.get("gpu_launches"), d.get("host_enqueue_ms_per_step") or 0))
e = d.get("e2e") or {}
print("e2e %.4g  (%.1f ms/step, %.1f GB/s over PCIe; host ceiling h2d %.1f d2h %.1f bidir %.1f GB/s)" % (
    e.get("value", 0), e.get("ms_per_step", 0), e.get("pcie_gbs_achieved", 0),
    e.get("host_pcie", {}).get("sum_h2d_gbs", 0), e.get("host_pcie", {}).get("sum_d2h_gbs", 0),
    e.get("host_pcie", {}).get("sum_bidir_gbs", 0)))
print("fp64 peak %.2f TF  clocks %s" % (d.get("fp64_peak_tflops") or 0, d.get("clocks")))
r = d.get("roofline") or {}
print("roofline: kernel %s frac %s share %.2f whole-step frac %s" % (
    r.get("kernel"), r.get("frac"), r.get("share_of_step") or 0, r.get("whole_step_frac")))
for k in r.get("kernels", []):
    if k["ms"] > 0.05:
        print("   %-20s %9.3f ms %12d pts  frac %s" % (k["name"], k["ms"], k["points"],
                                                       ("%.3f" % k["frac"]) if k.get("frac") else "-"))
print("accuracy:", json.dumps(d.get("accuracy")))
print("cpu_baseline:", json.dumps(d.get("cpu_baseline")))
print("separate_calls:", json.dumps(d.get("separate_calls")))
print("in_process:", json.dumps(d.get("in_process")))
for rk in d.get("ranks") or []:
    print("  rank", json.dumps(rk))
for c in d.get("configs") or []:
    if "error" in c:
        print("  configs error:", c["error"])
        continue
    acc = c.get("accuracy_vs_scipy") or {}
    if "max_norm_err" in acc:
        a = "%.2e" % acc["max_norm_err"]
    else:
        a = " ".join("%s %.2e" % (k, v["max_norm_err"]) for k, v in acc.items())
    print("  %-5s %-36s %9.3f ms %10.3e evals/s  frac %-6s %s  err %s" % (
        c["config"], c["fn"], c["ms"], c["evals_per_s"], ("%.3f" % c["frac"]) if c.get("frac") else "-",
        ("%.0f GB/s" % c["achieved_gbs"]) if "achieved_gbs" in c else "", a))
