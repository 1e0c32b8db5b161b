#!/usr/bin/env python
"""Device-resident throughput of every BASELINE config (not only the headline cfg2 that bench.py reports):
   cfg1  gamma, lgamma, J0/J1/Y0/Y1 on 2^20 real points in [0,100]
   cfg2  jv+yv, v=2.5, 4096x4096 grid                      (bench.py's workload)
   cfg3  ive, kve, per-element order in [0,50], |z|<=200   (n points of the 2^28 stream)
   cfg4  hankel1, hankel2 (v=2.5), Ai, Bi on |z| log-uniform 1e-3..1e4
   cfg5  J_n(z), n=0..255 per point, output [256][n]
Prints one JSON line per (config, function): evals/s, ms, per-kernel profile where available.
usage: python tools/bench_configs.py [--n 16777216] [--n5 1048576] [--reps 5] [--cfg 3,4,5]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps, warm=2):
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=1 << 24)
    ap.add_argument("--n5", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cfg", default="1,2,3,4,5")
    args = ap.parse_args()
    import torch
    from sciport_rs_b200 import special
    dev = torch.device("cuda", 0)
    cfgs = [int(c) for c in args.cfg.split(",")]

    def emit(cfg, name, evals, ms, prof=None, extra=None):
        line = {"cfg": cfg, "fn": name, "evals": evals, "ms": ms, "evals_per_s": evals / (ms * 1e-3)}
        if prof:
            line["kernels"] = prof
        if extra:
            line.update(extra)
        print(json.dumps(line), flush=True)

    def bessel_case(cfg, label, fn, v, z, scaled, real_input=False):
        n = z.numel()
        ms = timed(torch, lambda: special.bessel(fn, v, z, scaled=scaled, real_input=real_input), args.reps)
        prof = special.profile_step(lambda: special.bessel(fn, v, z, scaled=scaled, real_input=real_input, profile=True))
        emit(cfg, label, n, ms, prof)

    if 1 in cfgs:
        n = 1 << 20
        _, z1 = special.synth(1, n, n_total=n, device=dev)
        x = z1.real.contiguous()
        for name, f in (("gamma", special.gamma), ("gammaln", special.gammaln), ("sinc", special.sinc)):
            ms = timed(torch, lambda: f(x), args.reps)
            emit(1, name, n, ms, extra={"gbs": 16.0 * n / (ms * 1e-3) / 1e9})
        # the same memory-bound maps at a size that is not launch-bound (2^27 points, 2 GiB of traffic)
        nbig = 1 << 27
        xb = x.repeat(nbig // n)
        for name, f in (("gamma 2^27", special.gamma), ("gammaln 2^27", special.gammaln), ("sinc 2^27", special.sinc)):
            ms = timed(torch, lambda: f(xb), args.reps)
            emit(1, name, nbig, ms, extra={"gbs": 16.0 * nbig / (ms * 1e-3) / 1e9})
        del xb
        for fn in ("J", "Y"):
            for order in (0.0, 1.0):
                bessel_case(1, f"{fn}{int(order)} (real promoted)", fn, order, x, False, real_input=True)
    if 2 in cfgs:
        n = 4096 * 4096
        v, z = special.synth(2, n, n_total=n, device=dev)
        for fn in ("J", "Y"):
            bessel_case(2, f"{fn}v v=2.5", fn, 2.5, z, False)
        ms = timed(torch, lambda: special.bessel_jy(2.5, z), args.reps)
        prof = special.profile_step(lambda: special.bessel_jy(2.5, z, profile=True))
        emit(2, "jv+yv pair", 2 * n, ms, prof)
    if 3 in cfgs:
        v, z = special.synth(3, args.n, n_total=1 << 28, device=dev)
        for fn in ("I", "K"):
            bessel_case(3, f"{fn.lower()}ve per-element v", fn, v, z, True)
        outs = [torch.empty_like(z) for _ in range(2)]
        ms = timed(torch, lambda: special.bessel_ik(v, z, scaled=True, outs=outs, want_codes=False), args.reps)
        prof = special.profile_step(lambda: special.bessel_ik(v, z, scaled=True, outs=outs, want_codes=False, profile=True))
        emit(3, "ive+kve pair", 2 * args.n, ms, prof)
    if 4 in cfgs:
        v, z = special.synth(4, args.n, n_total=1 << 30, device=dev)
        for fn in ("H1", "H2"):
            bessel_case(4, f"hankel{fn[1]} v=2.5", fn, 2.5, z, False)
        for which in ("Ai", "Bi"):
            ms = timed(torch, lambda: special.airy_raw(which, z), args.reps)
            emit(4, which, args.n, ms)
    if 5 in cfgs:
        n = args.n5
        v, z = special.synth(5, n, n_total=1 << 24, device=dev)
        out = torch.empty((256, n), dtype=torch.complex128, device=dev)
        ms = timed(torch, lambda: special.bessel("J", 0.0, z, n_orders=256, out=out), args.reps)
        emit(5, "J_0..255 sweep", 256 * n, ms, extra={"gbs": 16.0 * 256 * n / (ms * 1e-3) / 1e9})


if __name__ == "__main__":
    main()
