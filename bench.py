#!/usr/bin/env python
"""bench.py — throughput of the batched complex Bessel path on B200 (BASELINE.json metric).

One *step* = one pass of the hot path over one batch of synthetic input: J_v and Y_v (v = 2.5) over the
4096 x 4096 complex grid |z| <= 200 of BASELINE config[1] ("complex Bessel jv/yv, fixed order v=2.5,
4096x4096 complex grid |z|<=200, single B200"); one *eval* = one complex output, so a step is
2 * 4096^2 evals.  Inputs are generated on the device (spb_synth) and stay resident in HBM for `value`;
`e2e` runs the same step through the host-buffer C-ABI path (pinned host memory -> H2D -> kernels -> D2H).

    python bench.py --gpus N --steps K --warmup W          (N > 1: launched by torchrun, one rank per GPU,
                                                             every rank its own grid: weak scaling, no collective
                                                             on the data path, barrier + max-over-ranks timing)
    python bench.py --impl reference ...                   the reference's CPU path stand-in (oracle port,
                                                             all host threads) on a bounded sample of the same work
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SIDE = 4096
N_POINTS = SIDE * SIDE
ORDER = 2.5
METRIC = "complex_bessel_evals_per_sec"
UNIT = "evals/s"
WORKLOAD = "cfg2: jv+yv, v=2.5, 4096x4096 complex grid |z|<=200 (2 evals per point)"
# algorithmic cost per eval: 16 B in + 16 B out; flops from the op-count tally
BYTES_PER_EVAL = 32.0


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


def traffic_for(kernel_name):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    # kernel names as ncu prints them (template arguments: region, scalar order, family, kode)
    prefix = {"kpass": "void k_kpass<", "ipass_asyi": "void k_ipass<1,", "ipass_mlri": "void k_ipass<2,",
              "classify": "void k_classify<"}
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f)["kernels"]
        want = prefix[kernel_name.split("/")[-1]]
        rec = next(v for k, v in t.items() if k.startswith(want))
        return rec["dram_read_bytes"] + rec["dram_write_bytes"]
    except Exception:
        return None


def bind_to_gpu_numa_node(torch, index):
    """Best effort: run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers
    allocated afterwards (first touch) and the staging copies stay on the PCIe root the GPU is attached to.
    Returns a short description for the JSON line, or None when the topology cannot be read."""
    try:
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"rank bound to NUMA node {node} ({len(cpus)} cpus) of GPU {bdf}"
    except Exception:
        return None


def load_flops_per_eval():
    path = os.path.join(ROOT, "profiles", "flops_per_eval.json")
    try:
        with open(path) as f:
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

    def stop(self):
        if self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        nv = self.nvml
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        sm = [x[0] for x in self.lines]
        reasons = sorted(k for k, bit in names.items() if any(x[2] & bit for x in self.lines))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_mhz),
                "samples": len(sm), "reasons": reasons, "how": "NVML, 20 ms period, samples inside the timed region"}


def cpu_sample_points(n_sample):
    """Evenly strided subsample of the cfg2 grid (same generator as the device: SURVEY.md §8d)."""
    idx = (np.arange(n_sample, dtype=np.int64) * (N_POINTS // n_sample)) % N_POINTS
    j = idx % SIDE
    k = idx // SIDE
    a = 200.0 / np.sqrt(2.0)
    return (-a + 2 * a * (j + 0.5) / SIDE) + 1j * (-a + 2 * a * (k + 0.5) / SIDE)


def run_cpu_port(n_sample, threads, repeats=1):
    """Time the CPU checker (oracle/libspb_oracle.so: the point-by-point TOMS-644 drivers, one thread per core)
    on a bounded sample of the step.  Returns evals/s."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers  # noqa: E402  (test infrastructure; allowed here as the cpu_baseline leg)
    z = cpu_sample_points(n_sample)
    helpers.cpu_bessel("J", 1, ORDER, z[:4096], threads=threads)  # warm
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        helpers.cpu_bessel("J", 1, ORDER, z, threads=threads)
        helpers.cpu_bessel("Y", 1, ORDER, z, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return 2 * n_sample / best, best


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = min(max(args.ref_sample, 4096), N_POINTS)  # default: the whole grid, about 1 s per step on 16 cores
    for _ in range(args.warmup):
        run_cpu_port(1 << 18, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run_cpu_port(n_sample, threads)
    dt = time.perf_counter() - t0
    value = 2 * n_sample * args.steps / dt
    sample = f"{n_sample} evenly strided points of the 4096x4096 grid per step (all of it by default), jv+yv"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU stand-in for the reference path: the Rust crate cannot be built "
                   "here (no rustc); oracle/ port of the same TOMS-644 algorithm, one thread per host core"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--ref-sample", type=int, default=N_POINTS, help="points per step of the reference arm")
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
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: sciport_rs_b200 has no CPU path")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa_note = bind_to_gpu_numa_node(torch, local_rank) if world > 1 else None
    warmup = max(args.warmup, 3)

    # FP64 peak of this GPU (DFMA microbenchmark, part of the library), before the timed region
    fp64_peak, _ = special.fp64_peak(local_rank, 1.0)

    # inputs resident in HBM: this rank's own 4096 x 4096 grid
    v, z = special.synth(2, N_POINTS, n_total=N_POINTS, device=dev)
    if world > 1:
        # every rank works on its own grid (weak scaling): the same disc |z| <= 200 turned by a rank-dependent angle
        z = z * complex(np.cos(0.1 * rank), np.sin(0.1 * rank))
    vt = torch.full((1,), ORDER, dtype=torch.float64, device=dev)

    kernel_ms = [0.0]
    launches = [0]

    # result buffers are owned by the caller and reused every step (the library never allocates results)
    d_outs = [torch.empty(N_POINTS, dtype=torch.complex128, device=dev) for _ in range(2)]
    d_codes = [torch.empty(N_POINTS, dtype=torch.int32, device=dev) for _ in range(4)]

    def step():
        # device-pointer calls only enqueue; nothing here synchronises with the GPU
        # J and Y over the same grid in one pass (spb_bessel_jy: the I pass is shared by the two functions)
        (oj, ej, _), (oy, ey, _) = special.bessel_jy(vt, z, outs=d_outs, codes=d_codes)
        _, l1 = _lib.last_timing(want_ms=False)
        launches[0] += l1
        return oj, oy, ej, ey

    # start the clock sampler first and wait for its first sample: nvidia-smi's start-up takes driver locks that
    # would otherwise stall launches inside the timed region
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.time()
        while not sampler.lines and time.time() - t_wait < 5.0:
            time.sleep(0.05)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    kernel_ms[0] = 0.0
    launches[0] = 0
    n_before = len(sampler.lines)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_host) / args.steps
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches_timed = launches[0]
    if rank == 0 and len(sampler.lines) - n_before < 2:
        # very short timed region: keep the GPU busy a little longer so that at least two samples are under load
        t_busy = time.time()
        while time.time() - t_busy < 0.35:
            step()
        torch.cuda.synchronize()
    if rank == 0:
        sampler.lines = sampler.lines[n_before:]
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        ms = max_over_ranks(ms, dev)
        dist.barrier()
    evals_per_step = 2 * N_POINTS * world
    value = evals_per_step * args.steps / (ms * 1e-3)

    # per-kernel times of one profiled step (events between launches) for the roofline of the dominant kernel
    prof = special.profile_step(lambda: special.bessel_jy(vt, z, profile=True))

    # end to end through the host-buffer C ABI: pinned host memory in, host memory out, copies inside the timed region
    e2e = None
    if rank == 0 or world > 1:
        zh = torch.empty(N_POINTS, dtype=torch.complex128).pin_memory()
        zh.copy_(z)
        zn = zh.numpy()
        # result buffers in pinned host memory (the caller owns all buffers: SURVEY 8b).  The call is the public
        # API a user of the reference's special module makes: jvyv -> Result<(Array, Array), i32>; the first failing
        # element is tracked on the device, so no per-element code arrays travel back.
        outs = [torch.empty(N_POINTS, dtype=torch.complex128).pin_memory().numpy() for _ in range(2)]
        special.jvyv(ORDER, zn, outs=outs)  # warm the staging buffers
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            oj, oy = special.jvyv(ORDER, zn, outs=outs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt, dev)
        h2d = zn.nbytes + 8
        d2h = oj.nbytes + oy.nbytes + 16
        e2e = {"value": evals_per_step * args.e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": args.e2e_steps,
               "numa": numa_note,
               "note": "special.jvyv on pinned host buffers through spb_bessel_jy: H2D of z, kernels, D2H of J and Y "
                       "and of the 16-byte first-error words inside the timed region"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    flops = load_flops_per_eval()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    # dominant kernel = the one with the largest share of the profiled step
    top = max(prof, key=lambda r: r["ms"]) if prof else None
    roofline = None
    if top:
        fpe = flops.get(top["name"], {}).get("flops_per_point")
        pts = top.get("points") or 0
        achieved = (fpe * pts / (top["ms"] * 1e-3) / 1e12) if (fpe and pts) else None
        roofline = {
            "bound": "fp64", "kernel": top["name"], "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": (achieved / fp64_peak) if achieved else None, "traffic": traffic_for(top["name"]),
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this kernel, one launch, from the committed "
                            "ncu --set full capture (profiles/r1_traffic.json)",
            "algorithmic_bytes": BYTES_PER_EVAL * pts * (1.5 if top["name"].endswith(("kpass", "ipass_asyi")) else 1.0),
            "peak_source": "spb_fp64_peak DFMA microbenchmark on this GPU (MEASURED_PEAKS.json has no FP64 figure)",
            "flops_per_point": fpe, "points_per_launch": pts, "launch_ms": top["ms"],
            "share_of_step": top["ms"] / max(sum(r["ms"] for r in prof), 1e-9),
            "hbm_view": {"achieved_gbs": BYTES_PER_EVAL * pts / (top["ms"] * 1e-3) / 1e9 if pts else None,
                         "peak_gbs": peaks.get("hbm_gbs")},
            "kernels": prof,
        }

    # accuracy of this very run against the CPU oracle on a strided parity sample (BASELINE metric: "max rel err"):
    # normalised error |got - want| / max(|J|, |Y|) (absolute bound near zeros), ierr / nz must agree exactly
    accuracy = None
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import helpers  # noqa: E402  (test infrastructure; the checker leg)
        from oracle import scipy_oracle as so  # noqa: E402
        n_acc = 1 << 16
        idx = torch.arange(0, N_POINTS, N_POINTS // n_acc, device=dev)[:n_acc]
        zs = z[idx].cpu().numpy()
        (oj, ej, nj), (oy, ey, ny) = special.bessel_jy(vt, z)
        gj, gy = oj[idx].cpu().numpy(), oy[idx].cpu().numpy()
        wj, ejc, njc = helpers.cpu_bessel("J", 1, ORDER, zs)
        wy, eyc, nyc = helpers.cpu_bessel("Y", 1, ORDER, zs)
        scale = np.maximum(np.abs(wj), np.abs(wy))
        err = np.concatenate([so.mismatch(gj, wj, scale), so.mismatch(gy, wy, scale)])
        codes_ok = (np.array_equal(ej[idx].cpu().numpy(), ejc) and np.array_equal(ey[idx].cpu().numpy(), eyc) and
                    np.array_equal(nj[idx].cpu().numpy(), njc) and np.array_equal(ny[idx].cpu().numpy(), nyc))
        accuracy = {"max_norm_err": float(err.max()), "p999_norm_err": float(np.quantile(err, 0.999)),
                    "tolerance": so.RTOL, "within_tolerance": bool(err.max() <= so.RTOL), "ierr_nz_identical": codes_ok,
                    "sample": f"{n_acc} evenly strided grid points, J and Y, vs the CPU oracle (oracle/)"}
        del oj, oy
    except Exception as e:  # the checker is optional for the timing line
        accuracy = {"error": repr(e)}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_sample = N_POINTS  # the whole 4096 x 4096 grid, jv + yv (about 16 core-seconds)
        cv, cdt = run_cpu_port(n_sample, threads)
        cpu_baseline = {"value": cv, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"all {n_sample} grid points, jv+yv, {cdt:.2f} s wall on {threads} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "points_per_gpu": N_POINTS, "order": ORDER,
                   "l2": "inputs+outputs per step 1.1 GB >> 126 MB L2 (no flush needed)",
                   "timing": "CUDA events on the launching stream, max over ranks",
                   "path": "binned, J and Y through spb_bessel_jy (shared I pass)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches_timed,
        "kernel_ms_per_step": sum(r["ms"] for r in prof) if prof else None,
        "host_enqueue_ms_per_step": host_enqueue_ms, "fp64_peak_tflops": fp64_peak,
        "roofline": roofline, "accuracy": accuracy, "cpu_baseline": cpu_baseline,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
