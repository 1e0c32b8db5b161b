#!/usr/bin/env python
"""Tally algorithmic FP64 operations per point for every kernel of the binned pipeline with the
op-counting build oracle/libspb_flopcount.so and freeze the means into profiles/flops_per_eval.json.

Counting rule (binding definition of the survey): add/sub/mul = 1, div = 1, sqrt = 1, exp/log/sin/cos/atan/
sinh/cosh = 30, pow = 60.  The classifier is charged prepare() as it runs it (straight-line: every region predicate), each region kernel the geometry of its point (prepare_pass) + its region routine + its share of finish().
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_pipeline_cpu import gen  # noqa: E402

KERNELS = ["classify", "scan_scatter", "ipass_series", "ipass_asyi", "ipass_mlri", "ipass_wrsk", "kpass", "general",
           "monolithic", "ipass_uni1", "ipass_uni2", "ipass_debye"]
NK = len(KERNELS)
FN = {"J": 0, "Y": 1, "I": 2, "K": 3, "H1": 4, "H2": 5, "JY": 6, "IK": 7}


def cfg2_sample(n):
    side = 4096
    idx = (np.arange(n, dtype=np.int64) * (side * side // n)) % (side * side)
    a = 200.0 / np.sqrt(2.0)
    z = (-a + 2 * a * (idx % side + 0.5) / side) + 1j * (-a + 2 * a * (idx // side + 0.5) / side)
    return np.full(n, 2.5), z


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "libspb_flopcount.so"], check=True)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "libspb_flopcount.so"))
    P = ctypes.c_void_p
    lib.spbo_count_ops2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, P, ctypes.c_int64, P, ctypes.c_int64, P, P]
    out = {"_rule": "add/sub/mul/div/sqrt = 1, exp/log/sin/cos/atan/sinh/cosh = 30, pow = 60; per point, "
                    "prepare + region + finish", "_sample": {}}
    rng = np.random.default_rng(1)
    # last field: the single-evaluation routines of the library's default flow (1 = fused I + K in the large-|w|
    # region for K-type families, 2 = Debye-direct region, 4 = K from the Wronskian route)
    samples = {"cfg2": (cfg2_sample(1 << 18), ("J", "Y", "JY"), 1, 7),
               "cfg3": (gen("cfg3", 1 << 17, rng), ("I", "K", "IK"), 2, 7),
               "cfg4": (gen("cfg4", 1 << 17, rng), ("H1", "H2"), 1, 7)}
    for cfg, ((v, z), fams, kode, fuse) in samples.items():
        v = np.ascontiguousarray(v, dtype=np.float64)
        z = np.ascontiguousarray(z, dtype=np.complex128)
        for fn in fams:
            ops = np.zeros(NK)
            pts = np.zeros(NK, np.int64)
            lib.spbo_count_ops2(FN[fn], kode, fuse, v.ctypes.data, 1, z.ctypes.data, z.size, ops.ctypes.data,
                                pts.ctypes.data)
            # J and I families: the fused large-|w| routine does not apply (bit 0 is ignored by the library too)
            total = ops[2:8].sum() + ops[9:NK].sum() + ops[0]
            for k in range(NK):
                if pts[k]:
                    key = f"{fn}/{KERNELS[k]}" if cfg == "cfg2" else f"{cfg}:{fn}/{KERNELS[k]}"
                    out[key] = {"flops_per_point": ops[k] / pts[k], "points": int(pts[k])}
            key = f"{fn}/per_eval" if cfg == "cfg2" else f"{cfg}:{fn}/per_eval"
            out[key] = {"flops_per_point": total / z.size, "points": int(z.size)}
            out["_sample"][f"{cfg}:{fn}"] = (f"{z.size} points, kode={kode}, fused I+K in the large-|w| region: "
                                              f"{bool(fuse & 1)}, Debye-direct region: {bool(fuse & 2)}")
    # Airy (cfg4 distribution): the complete routine per point
    lib.spbo_count_ops_airy.restype = ctypes.c_double
    lib.spbo_count_ops_airy.argtypes = [ctypes.c_int, ctypes.c_int, P, ctypes.c_int64]
    _, z4 = gen("cfg4", 1 << 16, rng)
    z4 = np.ascontiguousarray(z4, dtype=np.complex128)
    for which, name in ((0, "Ai"), (2, "Bi")):
        f = lib.spbo_count_ops_airy(which, 1, z4.ctypes.data, z4.size)
        out[f"cfg4:{name}/per_eval"] = {"flops_per_point": f, "points": int(z4.size)}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "flops_per_eval.json"), "w") as f:
        json.dump(out, f, indent=1)
    for k, val in out.items():
        if not k.startswith("_"):
            print(f"{k:28s} {val['flops_per_point']:10.1f} flops/pt over {val['points']} pts")


if __name__ == "__main__":
    main()
