#!/usr/bin/env python
"""bench.py — throughput of the batched complex Bessel path on B200 (BASELINE.json metric).

Headline workload = BASELINE config[2], the configuration the metric's target is quoted on ("complex Bessel J/Y/I/K
batches of >= 256M points at >= 60 % FP64 peak", "exponentially-scaled iv/kv with per-element order v in [0,50],
256M points, sharded over 1/2/4/8 B200"): ONE batch of 2^28 points of the index-addressable cfg3 stream (SURVEY.md
§8d: v = 50 u0, z = 200 u1 e^{i pi (2 u2 - 1)}), evaluated as ive(v, z) and kve(v, z) (KODE = 2).  One *step* = both
functions over the whole batch = 2 * 2^28 evals (one eval = one complex output), computed by the pair entry point
spb_bessel_ik (special.ive_kve): one classification, I and K at the shared right-half-plane point, both outputs from
one series where a single-evaluation region applies.  `separate_calls` reports the same step as two single calls
(spb_bessel(I) then spb_bessel(K)).  With N GPUs the batch is sliced
contiguously, rank r owns [r n / N, (r + 1) n / N) (strong scaling; no element is shared, so there is no collective
on the data path: torch.distributed only carries the barrier and the max-over-ranks of the device time).

    python bench.py --gpus N --steps K --warmup W      N > 1: launched by torchrun, one rank per GPU
    python bench.py --impl reference ...               the reference's CPU path stand-in (oracle port, all host
                                                       threads) on a bounded sample of the same workload

`value`: inputs resident in HBM, CUDA events on the launching stream.  `e2e`: the same step through the public API
(special.ive / special.kve_array on pinned HOST buffers -> C ABI -> H2D, kernels, D2H inside the timed region).
`configs`: the other BASELINE configs (cfg1, cfg2, cfg4, cfg5), each with throughput, roofline fraction and max
normalised error against live scipy.special on a strided 2^16 sample (N = 1 only).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2_POINTS = 28
METRIC = "complex_bessel_evals_per_sec"
UNIT = "evals/s"


def workload_name(log2_points):
    return (f"cfg3: ive+kve (KODE=2), per-element order v in [0,50], |z|<=200, one batch of 2^{log2_points} points "
            "sliced contiguously over the GPUs (2 evals per point)")


# algorithmic bytes per eval: per-element order 8 B + z 16 B in, 16 B out
BYTES_PER_EVAL = 40.0


def slice_bounds(n, world):
    """Contiguous slice [i n / G, (i + 1) n / G) of an n-point batch per rank / device: the partition the C ABI uses
    for n_devices = G (capi.cu) and the one a multi-rank job uses to shard one batch.  No element is shared, so
    no collective runs on the data path."""
    return [(n * r // world, n * (r + 1) // world) for r in range(world)]


def max_over_ranks(value, device=None):
    """Timing rule of the bench contract: the slowest rank defines the step.  all_reduce(MAX) over the default
    process group (NCCL on GPUs, gloo in the CPU tests); identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------- synthetic stream on the host (reference arm)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def u01(seed, idx, k):
    """u_k(i) of SURVEY.md §8d: (mix64(seed * 0x9E3779B97F4A7C15 + 4 i + k) >> 11) * 2^-53 (arithmetic mod 2^64)."""
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(4) * idx.astype(np.uint64) + np.uint64(k))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def cfg3_points(idx):
    """(v, z) of the cfg3 stream at the given indices, on the host (same generator as spb_synth on the device)."""
    v = 50.0 * u01(3, idx, 0)
    r = 200.0 * u01(3, idx, 1)
    th = np.pi * (2.0 * u01(3, idx, 2) - 1.0)
    return v, r * np.cos(th) + 1j * r * np.sin(th)


def strided_indices(n_total, n_sample):
    return (np.arange(n_sample, dtype=np.int64) * max(n_total // n_sample, 1)) % n_total


def load_json(*parts):
    try:
        with open(os.path.join(ROOT, *parts)) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, through NVML from a background thread
    (an `nvidia-smi -lms` child process was measured to stall kernel launches by ~30 % on this driver)."""

    def __init__(self, index):
        self.index = index
        self.lines = []  # (sm_mhz, sm_max_mhz, reasons_bitmask)
        self.stop_flag = threading.Event()
        self.thread = None
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nvml = None
            return
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        nv = self.nvml
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.lines.append((float(sm), float(self.max_mhz), int(reasons)))
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def mark(self):
        return len(self.lines)

    def summary(self, lo, hi=None):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        nv = self.nvml
        lines = self.lines[lo:hi]
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [x[0] for x in lines]
        reasons = sorted(k for k, bit in names.items() if any(x[2] & bit for x in lines))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_mhz),
                "samples": len(sm), "reasons": reasons, "how": "NVML, 20 ms period, samples inside the timed region"}

    def stop(self):
        if self.nvml is None:
            return
        self.stop_flag.set()
        self.thread.join(timeout=2)


# ---------------------------------------------------------------- CPU stand-in of the reference path
def run_cpu_port(v, z, threads):
    """ive + kve of the sample on the CPU checker (oracle/libspb_oracle.so: the point-by-point TOMS-644 drivers, one
    thread per core — the stand-in for the reference's rayon-parallel mapv).  Returns seconds."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers  # noqa: E402  (test infrastructure; allowed here as the cpu_baseline / reference leg)
    t0 = time.perf_counter()
    helpers.cpu_bessel("I", 2, v, z, threads=threads)
    helpers.cpu_bessel("K", 2, v, z, threads=threads)
    return time.perf_counter() - t0


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_total = 1 << args.log2_points
    threads = os.cpu_count() or 1
    n_sample = min(max(args.ref_sample, 4096), n_total)
    v, z = cfg3_points(strided_indices(n_total, n_sample))
    run_cpu_port(v[:4096], z[:4096], threads)  # load + warm the library
    for _ in range(args.warmup):
        run_cpu_port(v[: max(n_sample // 8, 4096)], z[: max(n_sample // 8, 4096)], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu_port(v, z, threads)
    dt = time.perf_counter() - t0
    value = 2 * n_sample * args.steps / dt
    sample = (f"{n_sample} evenly strided points of the 2^{args.log2_points}-point cfg3 stream per step, ive+kve "
              f"(throughput of an elementwise map is size-independent)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.log2_points),
                   "note": "CPU stand-in for the reference path: the Rust crate cannot be built here (no rustc, "
                           "un-vendored dependency); oracle/ port of the same TOMS-644 algorithm, one thread per "
                           "host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------- accuracy against live scipy
def scipy_accuracy(fn, kode, v, z, got, ierr=None):
    """Max / p99.9 normalised error of `got` against scipy.special (the reference's declared oracle) on the sample:
    |got - want| / scale with the absolute bound near zeros (scale = companion magnitude where the function can
    vanish by cancellation), points where scipy reports overflow / no result excluded, tolerance policy of
    oracle/scipy_oracle.py."""
    from oracle import scipy_oracle as so  # noqa: E402  (checker leg only)
    want = so.evaluate(fn, kode, v, z)
    # away from overflow / underflow: values in the band where the algorithm flushes to zero (nz != 0) are excluded
    ok = np.isfinite(want.real) & np.isfinite(want.imag) & (np.abs(want) > 1e-280)
    if ierr is not None:
        ok &= (ierr == 0)
    scale = so.scale_near_zeros(fn, kode, v, z, want)
    err = so.mismatch(got, want, scale)[ok]
    tol = so.tolerance(want, z, v)[ok]
    return {"max_norm_err": float(err.max()), "p999_norm_err": float(np.quantile(err, 0.999)),
            "frac_above_1e-13": float((err > so.RTOL).mean()), "within_tolerance": bool(np.all(err <= tol)),
            "points": int(ok.sum())}


def timed(torch, fn, reps, warm=2):
    """Best-of-`reps` device time (ms) of fn() with CUDA events on the current stream."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best


def aggregate_profile(prof):
    """Sum the per-chunk records of a profiled call by kernel name."""
    agg = {}
    for r in prof:
        a = agg.setdefault(r["name"], {"name": r["name"], "ms": 0.0, "points": 0, "launches": 0})
        a["ms"] += r["ms"]
        a["points"] += r["points"]
        a["launches"] += 1
    return sorted(agg.values(), key=lambda r: -r["ms"])


def other_configs(torch, special, dev, fp64_peak, hbm_peak, flops, reps=3):
    """Device-resident throughput, roofline fraction and accuracy against live scipy of BASELINE configs 1, 2, 4, 5."""
    from oracle import scipy_oracle as so  # noqa: E402  (checker leg only)
    import scipy.special as sc
    out = []
    n_acc = 1 << 16

    def sample(t, n):
        idx = torch.arange(0, n, max(n // n_acc, 1), device=dev)[:n_acc]
        return idx, t[idx].cpu().numpy()

    def fp64_entry(cfg, name, evals, ms, flops_key, acc, prof=None, extra=None):
        fpe = flops.get(flops_key, {}).get("flops_per_point")
        e = {"config": cfg, "fn": name, "evals": evals, "ms": ms, "evals_per_s": evals / (ms * 1e-3), "bound": "fp64",
             "flops_per_eval": fpe, "frac": (fpe * evals / (ms * 1e-3) / 1e12 / fp64_peak) if fpe else None,
             "accuracy_vs_scipy": acc}
        if prof:
            top = prof[0]
            tf = flops.get(top["key"], {}).get("flops_per_point")
            e["dominant_kernel"] = {"name": top["name"], "ms": top["ms"], "points": top["points"],
                                    "frac": (tf * top["points"] / (top["ms"] * 1e-3) / 1e12 / fp64_peak)
                                    if tf and top["ms"] > 0 else None}
        if extra:
            e.update(extra)
        out.append(e)

    def hbm_entry(cfg, name, evals, ms, bytes_per_eval, acc, extra=None):
        gbs = bytes_per_eval * evals / (ms * 1e-3) / 1e9
        e = {"config": cfg, "fn": name, "evals": evals, "ms": ms, "evals_per_s": evals / (ms * 1e-3), "bound": "hbm",
             "bytes_per_eval": bytes_per_eval, "achieved_gbs": gbs, "frac": gbs / hbm_peak if hbm_peak else None,
             "accuracy_vs_scipy": acc}
        if extra:
            e.update(extra)
        out.append(e)

    def real_acc(got, want, floor=0.0):
        with np.errstate(all="ignore"):
            err = np.abs(got - want) / np.maximum(np.abs(want), floor)
        err = err[np.isfinite(want) & np.isfinite(err)]
        return {"max_norm_err": float(err.max()), "p999_norm_err": float(np.quantile(err, 0.999)), "points": int(err.size)}

    def profiled(call, prefix):
        prof = aggregate_profile(special.profile_step(call))
        for r in prof:
            # (k_direct runs the large-|w| routine of the region kernel on points in input order: same tally)
            r["key"] = prefix + r["name"].replace("direct_asyi", "ipass_asyi")
        return prof

    # ---- cfg1: gamma, ln|gamma| and J0/J1/Y0/Y1 (real arguments promoted to complex, mod.rs:10) on [0, 100]
    n1 = 1 << 20
    _, z1 = special.synth(1, n1, n_total=n1, device=dev)
    x1 = z1.real.contiguous()
    nbig = 1 << 27
    xb = x1.repeat(nbig // n1)
    ob = torch.empty_like(xb)
    xs_idx, xs = sample(x1, n1)
    for name, f, ref in (("gamma", special.gamma, sc.gamma), ("gammaln", special.gammaln, sc.gammaln)):
        got = f(x1)[xs_idx].cpu().numpy()
        with np.errstate(all="ignore"):
            # ln|gamma| vanishes at x = 1, 2: absolute bound there (scale floor 1)
            acc = real_acc(got, ref(xs), floor=1.0 if name == "gammaln" else 0.0)
        ms_small = timed(torch, lambda: f(x1), reps)
        ms = timed(torch, lambda: f(xb, out=ob), reps)
        hbm_entry("cfg1", name, nbig, ms, 16.0, acc,
                  {"points": nbig, "note": "2^27 real points in [0,100] (the 2^20-point cfg1 sample tiled 128x): "
                   "the size at which the map is bandwidth- rather than launch-bound",
                   "ms_at_2^20": ms_small, "evals_per_s_at_2^20": n1 / (ms_small * 1e-3)})
    del xb, ob
    for fn in ("J", "Y"):
        for order in (0.0, 1.0):
            call = lambda **kw: special.bessel(fn, order, x1, real_input=True, **kw)  # noqa: E731
            ms = timed(torch, lambda: call(want_codes=False), reps)
            got, ierr, _ = call()
            acc = scipy_accuracy(fn, 1, order, xs + 0j, got[xs_idx].cpu().numpy(), ierr[xs_idx].cpu().numpy())
            fp64_entry("cfg1", f"{fn}{int(order)} (real promoted)", n1, ms, f"{fn}/per_eval", acc,
                       extra={"flops_note": "flops per eval of the cfg2 tally (same family and region mix class)"})
    # ---- cfg2: jv + yv, v = 2.5, 4096 x 4096 grid
    n2 = 4096 * 4096
    _, z2 = special.synth(2, n2, n_total=n2, device=dev)
    outs = [torch.empty(n2, dtype=torch.complex128, device=dev) for _ in range(2)]
    ms = timed(torch, lambda: special.bessel_jy(2.5, z2, outs=outs, want_codes=False), max(reps, 5))
    prof = profiled(lambda: special.bessel_jy(2.5, z2, outs=outs, profile=True, want_codes=False), "")
    (gj, ej, _), (gy, ey, _) = special.bessel_jy(2.5, z2)
    idx, zs = sample(z2, n2)
    acc = {"J": scipy_accuracy("J", 1, 2.5, zs, gj[idx].cpu().numpy(), ej[idx].cpu().numpy()),
           "Y": scipy_accuracy("Y", 1, 2.5, zs, gy[idx].cpu().numpy(), ey[idx].cpu().numpy())}
    # whole-step flops: the pair per POINT (two evals)
    fpp = flops.get("JY/per_eval", {}).get("flops_per_point")
    out.append({"config": "cfg2", "fn": "jv+yv pair (v=2.5, 4096x4096 grid)", "evals": 2 * n2, "ms": ms,
                "evals_per_s": 2 * n2 / (ms * 1e-3), "bound": "fp64", "flops_per_point": fpp,
                "frac": (fpp * n2 / (ms * 1e-3) / 1e12 / fp64_peak) if fpp else None,
                "dominant_kernel": (lambda t: {"name": t["name"], "ms": t["ms"], "points": t["points"],
                                               "frac": (flops.get(t["key"], {}).get("flops_per_point", 0) * t["points"] /
                                                        (t["ms"] * 1e-3) / 1e12 / fp64_peak) or None})(prof[0]),
                "accuracy_vs_scipy": acc})
    del outs, gj, gy, z2
    # ---- cfg4: hankel1/2 (v = 2.5), Ai, Bi on |z| log-uniform 1e-3..1e4: the per-GPU share of 2^30 points over 8 GPUs
    n4 = 1 << 27
    _, z4 = special.synth(4, n4, n_total=1 << 30, device=dev)
    o4 = torch.empty(n4, dtype=torch.complex128, device=dev)
    idx, zs = sample(z4, n4)
    for fn in ("H1", "H2"):
        ms = timed(torch, lambda: special.bessel(fn, 2.5, z4, out=o4, want_codes=False), reps)
        prof = profiled(lambda: special.bessel(fn, 2.5, z4, out=o4, profile=True), "cfg4:")
        got, ierr, _ = special.bessel(fn, 2.5, z4, out=o4)
        acc = scipy_accuracy(fn, 1, 2.5, zs, got[idx].cpu().numpy(), ierr[idx].cpu().numpy())
        fp64_entry("cfg4", f"hankel{fn[1]} (v=2.5)", n4, ms, f"cfg4:{fn}/per_eval", acc, prof,
                   {"points": n4, "note": "2^27 points = one GPU's share of the 2^30-point config"})
    for which, wi in (("Ai", 0), ("Bi", 2)):
        ms = timed(torch, lambda: special.airy_raw(which, z4, out=o4, want_codes=False), reps)
        got, ierr, _ = special.airy_raw(which, z4, out=o4)
        g, ie = got[idx].cpu().numpy(), ierr[idx].cpu().numpy()
        want = so.evaluate_airy(wi, 1, zs)
        comp = so.evaluate_airy(2 - wi, 1, zs)
        ok = np.isfinite(want.real) & np.isfinite(want.imag) & (ie == 0)
        # near the zeros on the negative real axis the envelope |Ai| ~ |Bi| is the scale
        scale = np.maximum(np.abs(want), np.where((zs.real < 0) & np.isfinite(np.abs(comp)), np.abs(comp), 0))
        err = so.mismatch(g, want, scale)[ok]
        zeta = (2.0 / 3.0) * np.abs(zs) ** 1.5
        tol = (so.tolerance(want) + 4 * so.EPS * zeta)[ok]
        acc = {"max_norm_err": float(err.max()), "p999_norm_err": float(np.quantile(err, 0.999)),
               "frac_above_1e-13": float((err > so.RTOL).mean()), "within_tolerance": bool(np.all(err <= tol)),
               "points": int(ok.sum())}
        fp64_entry("cfg4", which, n4, ms, f"cfg4:{which}/per_eval", acc,
                   extra={"points": n4, "note": "2^27 points = one GPU's share of the 2^30-point config"})
    del o4, z4
    # ---- cfg5: J_n(z), n = 0..255 per point, output [256][n]: the per-GPU share (2^21 points) of 2^24 points over 8 GPUs
    n5 = 1 << 21
    _, z5 = special.synth(5, n5, n_total=1 << 24, device=dev)
    o5 = torch.empty((256, n5), dtype=torch.complex128, device=dev)
    ms = timed(torch, lambda: special.bessel("J", 0.0, z5, n_orders=256, out=o5, want_codes=False), reps)
    special.bessel("J", 0.0, z5, n_orders=256, out=o5, want_codes=False)
    pts = torch.arange(0, n5, n5 // 256, device=dev)[:256]
    g = o5[:, pts].cpu().numpy()  # [256 orders][256 points] = 2^16 evals
    zs = z5[pts].cpu().numpy()
    orders = np.arange(256, dtype=np.float64)[:, None]
    want = so.evaluate("J", 1, np.broadcast_to(orders, g.shape), np.broadcast_to(zs[None, :], g.shape))
    comp = np.abs(so.evaluate("Y", 1, np.broadcast_to(orders, g.shape), np.broadcast_to(zs[None, :], g.shape)))
    # members below the underflow band of the sweep (|J| < 1e-280) are flushed differently by the two algorithms
    ok = np.isfinite(want.real) & np.isfinite(want.imag) & (np.abs(want) > 1e-280)
    scale = np.maximum(np.abs(want), np.where(np.isfinite(comp) & (orders <= np.abs(zs)[None, :]), comp, 0.0))
    err = so.mismatch(g, want, scale)[ok]
    with np.errstate(all="ignore"):
        tol = (np.maximum(1e-13, 8 * so.EPS * (orders + 256)) + 4 * so.EPS * np.abs(np.log(np.abs(want))))[ok]
    acc = {"max_norm_err": float(err.max()), "p999_norm_err": float(np.quantile(err, 0.999)),
           "frac_above_1e-13": float((err > so.RTOL).mean()), "within_tolerance": bool(np.all(err <= tol)),
           "points": int(ok.sum())}
    hbm_entry("cfg5", "J_0..255 order sweep", 256 * n5, ms, 16.0 + 16.0 / 256, acc,
              {"points": n5, "note": "2^21 points x 256 orders (8 GiB out) = one GPU's share of the 2^24-point config"})
    del o5, z5
    torch.cuda.empty_cache()
    return out


def pcie_probe(torch, dev, world, nbytes=1 << 30):
    """Pinned-host <-> device copy bandwidth of this rank, alone-per-rank timing but all ranks at once (so at N > 1
    the sum over ranks is what the host moves in aggregate).  GB/s: (h2d, d2h, both directions at once)."""
    import torch.distributed as dist
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream(dev)
    res = []
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            if mode in ("h2d", "both"):
                d.copy_(h, non_blocking=True)
            if mode == "d2h":
                h2.copy_(d2, non_blocking=True)
            if mode == "both":
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res.append((2 if mode == "both" else 1) * 3 * nbytes / dt / 1e9)
    del h, h2, d, d2
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log2-points", type=int, default=LOG2_POINTS, help="batch size of the headline workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1/2/4/5 block (N = 1 only anyway)")
    ap.add_argument("--no-in-process", action="store_true", help="skip the one-process n_devices = N leg")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--ref-sample", type=int, default=1 << 22, help="points per step of the CPU legs")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from sciport_rs_b200 import special, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sciport_rs_b200 has no CPU path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    n_total = 1 << args.log2_points
    lo, hi = slice_bounds(n_total, world)[rank]
    n = hi - lo
    topo = special.device_topology(local_rank)

    # FP64 peak of this GPU: the DFMA microbenchmark of the library, sustained for 2 s, clocks sampled
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.lines and sampler.nvml is not None and time.time() - t_wait < 5.0:
        time.sleep(0.05)
    m0 = sampler.mark()
    fp64_peak, _ = special.fp64_peak(local_rank, 2.0)
    peak_clocks = sampler.summary(m0)

    # this rank's contiguous slice of the ONE batch, generated in place on the device (index-addressable stream)
    v, z = special.synth(3, n, n_total=n_total, offset=lo, device=dev)
    out_i = torch.empty(n, dtype=torch.complex128, device=dev)
    out_k = torch.empty(n, dtype=torch.complex128, device=dev)
    launches = [0]

    def step():
        # device-pointer calls only enqueue; nothing here synchronises with the GPU.  No per-element code arrays: the
        # first failing element of each output is tracked in one device word (Result<Array, i32> semantics)
        special.bessel_ik(v, z, scaled=True, outs=[out_i, out_k], want_codes=False)
        launches[0] += _lib.last_timing(want_ms=False)[1]

    def step_separate():
        special.bessel("I", v, z, scaled=True, out=out_i, want_codes=False)
        special.bessel("K", v, z, scaled=True, out=out_k, want_codes=False)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches[0] = 0
    m1 = sampler.mark()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        step()
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host) / args.steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    m2 = sampler.mark()
    launches_timed = launches[0]
    if m2 - m1 < 2:
        t_busy = time.time()
        while time.time() - t_busy < 0.35:
            step()
        torch.cuda.synchronize()
        m2 = sampler.mark()
    clocks = sampler.summary(m1, m2)
    my_ms = ms
    if world > 1:
        ms = max_over_ranks(ms, dev)
        dist.barrier()
    evals_per_step = 2 * n_total
    value = evals_per_step * args.steps / (ms * 1e-3)
    launches_total = int(sum_over_ranks(launches_timed, dev))

    # the same step as two single calls (device-resident), for comparison
    sep_ms = timed(torch, step_separate, 3, warm=1)
    if world > 1:
        sep_ms = max_over_ranks(sep_ms, dev)
    separate_calls = {"value": evals_per_step / (sep_ms * 1e-3), "unit": UNIT, "ms_per_step": sep_ms,
                      "note": "spb_bessel(I, kode 2) then spb_bessel(K, kode 2) over the same slices, best of 3"}
    step()  # (the accuracy leg below reads the pair's outputs)

    # per-kernel times of one profiled step (events between launches; the records of all chunks summed per kernel)
    prof = aggregate_profile(special.profile_step(
        lambda: special.bessel_ik(v, z, scaled=True, outs=[out_i, out_k], want_codes=False, profile=True)))

    # accuracy of this very run against live scipy on a strided 2^16 sample of this rank's slice
    accuracy = None
    if rank == 0:
        try:
            n_acc = 1 << 16
            idx = torch.arange(0, n, max(n // n_acc, 1), device=dev)[:n_acc]
            vs, zs = v[idx].cpu().numpy(), z[idx].cpu().numpy()
            (gi, ei, _), (gk, ek, _) = special.bessel_ik(v[idx].contiguous(), z[idx].contiguous(), scaled=True)
            same = bool(torch.equal(gi, out_i[idx]) and torch.equal(gk, out_k[idx]))
            accuracy = {"ive": scipy_accuracy("I", 2, vs, zs, out_i[idx].cpu().numpy(), ei.cpu().numpy()),
                        "kve": scipy_accuracy("K", 2, vs, zs, out_k[idx].cpu().numpy(), ek.cpu().numpy()),
                        "sample_recomputed_bit_identical": same,
                        "sample": f"{n_acc} evenly strided points of rank 0's slice vs live scipy.special.ive / kve "
                                  "(the reference's declared oracle, tests/special.rs:10-16); tolerance policy "
                                  "oracle/scipy_oracle.py::tolerance"}
        except Exception as e:  # the checker is optional for the timing line
            accuracy = {"error": repr(e)}

    # ---- end to end through the public API: pinned HOST buffers in, host buffers out, copies inside the timed region
    vh = torch.empty(n, dtype=torch.float64).pin_memory()
    zh = torch.empty(n, dtype=torch.complex128).pin_memory()
    vh.copy_(v)
    zh.copy_(z)
    oi_h = torch.empty(n, dtype=torch.complex128).pin_memory()
    ok_h = torch.empty(n, dtype=torch.complex128).pin_memory()
    del out_i, out_k, v, z
    torch.cuda.empty_cache()
    vn, zn, oin, okn = vh.numpy(), zh.numpy(), oi_h.numpy(), ok_h.numpy()

    e2e_errs = {}

    def e2e_step(devices=None, m=None):
        # the public entry points return the reference's Result<Array, i32>: the whole batch is evaluated and the
        # first Err in index order is raised afterwards.  Among 2^28 cfg3 points a handful overflow for real
        # (K_v(z) e^z at |z| ~ 1e-6 with v ~ 49), so the Err is part of the workload: caught and reported here.
        sl = slice(0, m)
        try:
            special.ive_kve(vn[sl], zn[sl], outs=[oin[sl], okn[sl]], devices=devices)
        except special.SpecialError as err:
            e2e_errs["ive_kve"] = {"first_err_index": int(err.index), "ierr": int(err.code)}

    e2e_step(m=min(n, 1 << 23))  # warm the staging buffers
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    dt = time.perf_counter() - t0
    my_e2e = dt
    dt = max_over_ranks(dt, dev)
    h2d = vn.nbytes + zn.nbytes
    d2h = oin.nbytes + okn.nbytes + 2 * 16
    e2e = {"value": evals_per_step * args.e2e_steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(sum_over_ranks(h2d, dev)), "d2h_bytes_per_step": int(sum_over_ranks(d2h, dev)),
           "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps, "first_errors_rank0": e2e_errs,
           "pcie_gbs_achieved": (int(sum_over_ranks(h2d + d2h, dev)) * args.e2e_steps / dt) / 1e9,
           "note": "special.ive_kve on pinned host buffers through spb_bessel_ik: H2D of v and z, kernels, D2H of both "
                   "results and of the 16-byte first-error words inside the timed region; two streams per device, "
                   "chunks of 2^22 points, copies overlapped with compute"}
    pcie = pcie_probe(torch, dev, world)
    pcie_sum = [sum_over_ranks(x, dev) for x in pcie]
    per_rank = {"rank": rank, "gpu": local_rank, "topology": topo, "device_ms_per_step": my_ms / args.steps,
                "e2e_s_per_step": my_e2e / args.e2e_steps, "pcie_h2d_gbs": pcie[0], "pcie_d2h_gbs": pcie[1],
                "pcie_bidir_gbs": pcie[2]}
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, per_rank)
    else:
        gathered = [per_rank]
    e2e["host_pcie"] = {"sum_h2d_gbs": pcie_sum[0], "sum_d2h_gbs": pcie_sum[1], "sum_bidir_gbs": pcie_sum[2],
                        "how": "1 GiB pinned copies, every rank at once; sum over ranks = what the host moves in "
                               "aggregate (the ceiling of e2e at this N)"}

    # ---- the same batch from ONE process through the library's own split (spb_exec.n_devices = N, capi.cu):
    # one host thread + two streams per device, contiguous slices, no collective
    in_process = None
    if world > 1 and not args.no_in_process:
        dist.barrier()
        if rank == 0:
            try:
                m = min(n, 1 << 26)  # rank 0's pinned buffers hold its own slice: use up to 2^26 points of it
                devs = list(range(world))
                e2e_step(devices=devs, m=m)  # warm: contexts, staging buffers and scratch of every device
                t0 = time.perf_counter()
                e2e_step(devices=devs, m=m)
                dt1 = time.perf_counter() - t0
                t0 = time.perf_counter()
                e2e_step(devices=[0], m=m)
                dt0 = time.perf_counter() - t0
                in_process = {"n_devices": world, "points": m, "evals_per_s": 2 * m / dt1,
                              "evals_per_s_one_device": 2 * m / dt0, "speedup": dt0 / dt1,
                              "note": "special.ive_kve with devices=[0..N-1] from rank 0 alone (the other "
                                      "ranks idle at a barrier): spb_bessel_ik slices the batch contiguously across "
                                      "the devices, one host thread and two streams per device"}
            except Exception as e:
                in_process = {"error": repr(e)}
        dist.barrier()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sampler.stop()
        return
    del vh, zh, oi_h, ok_h

    flops = load_json("profiles", "flops_per_eval.json")
    peaks = load_json("MEASURED_PEAKS.json")
    hbm_peak = peaks.get("hbm_gbs") or 6547.8
    # whole step by tallied algorithmic flops
    # (the pair's own tally per POINT: both outputs from one classification / one series where that applies)
    f_ik = flops.get("cfg3:IK/per_eval", {}).get("flops_per_point")
    step_frac = None
    if f_ik:
        step_frac = f_ik * n_total * args.steps / (ms * 1e-3) / 1e12 / (fp64_peak * world)
    # dominant kernel = the one with the largest share of the profiled step (rank 0's slice)
    roofline = None
    if prof:
        top = prof[0]
        fpe = flops.get("cfg3:" + top["name"], {}).get("flops_per_point")
        pts = top["points"]
        out_bytes = 32.0  # the pair stores two results per point
        per_launch_ms = top["ms"] / max(top["launches"], 1)
        achieved = (fpe * pts / (top["ms"] * 1e-3) / 1e12) if (fpe and pts and top["ms"] > 0) else None
        traffic_rec = load_json("profiles", "r2_traffic.json").get("kernels", {}).get(top["name"]) or {}
        traffic = traffic_rec.get("bytes")
        roofline = {
            "bound": "fp64", "kernel": top["name"], "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (achieved / fp64_peak) if achieved else None,
            "traffic": traffic, "traffic_detail": traffic_rec or None,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the committed "
                            "ncu --set full capture of one 2^24-point chunk (profiles/r2_traffic.json), null if not captured",
            "algorithmic_bytes_per_launch": (24.0 + out_bytes) * pts / max(top["launches"], 1),
            "peak_source": "spb_fp64_peak: DFMA microbenchmark of this library, 8 independent chains per thread, all "
                           "SMs, sustained 2 s on this GPU right before the timed region (MEASURED_PEAKS.json has no "
                           "FP64 figure); clocks while it ran: " + json.dumps(peak_clocks),
            "flops_per_point": fpe, "points_per_launch": pts / max(top["launches"], 1),
            "launch_ms": per_launch_ms, "launches_per_step": top["launches"],
            "share_of_step": top["ms"] / max(sum(r["ms"] for r in prof), 1e-9),
            "whole_step_frac": step_frac,
            "whole_step_flops_per_point": f_ik,
            "hbm_view": {"achieved_gbs": 56.0 * n_total * args.steps / (ms * 1e-3) / 1e9,
                         "peak_gbs": hbm_peak * world},
            "kernels": [dict(r, frac=(flops.get("cfg3:" + r["name"], {}).get("flops_per_point", 0) * r["points"] /
                                      (r["ms"] * 1e-3) / 1e12 / fp64_peak) if r["ms"] > 0 else None) for r in prof],
        }

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = min(max(args.ref_sample, 4096), n_total)
        vs, zs = cfg3_points(strided_indices(n_total, n_sample))
        run_cpu_port(vs[:4096], zs[:4096], threads)
        cdt = run_cpu_port(vs, zs, threads)
        cpu_baseline = {"value": 2 * n_sample / cdt, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{n_sample} evenly strided points of the cfg3 stream, ive+kve, {cdt:.2f} s wall on "
                                  f"{threads} threads (oracle/ port of the TOMS-644 drivers)"}

    configs = None
    if world == 1 and not args.no_configs:
        try:
            configs = other_configs(torch, special, dev, fp64_peak, hbm_peak, flops)
        except Exception as e:
            configs = [{"error": repr(e)}]
    sampler.stop()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.log2_points), "points_total": n_total,
                   "points_per_gpu": n_total // world,
                   "l2": "inputs+outputs per step 56 B/point x 2^28/N points >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the launching stream, max over ranks",
                   "path": "binned pipeline, chunks of 2^24 points, spb_bessel_ik (I and K, kode 2, one pass)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_total,
        "kernel_ms_per_step": sum(r["ms"] for r in prof) if prof else None,
        "host_enqueue_ms_per_step": host_enqueue_ms, "fp64_peak_tflops": fp64_peak, "fp64_peak_clocks": peak_clocks,
        "roofline": roofline, "accuracy": accuracy, "cpu_baseline": cpu_baseline, "ranks": gathered,
        "separate_calls": separate_calls, "in_process": in_process, "configs": configs,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
