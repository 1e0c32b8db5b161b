#!/usr/bin/env python
"""Golden vectors for ORDER SEQUENCES (BASELINE config 5: J_n(z), n = 0..255 per point, output [order][point]),
generated from scipy.special one order at a time — the reference's declared oracle
(/root/reference/tests/special.rs:10-16).  Run:  python tests/golden/make_golden_seq.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import scipy_oracle as so  # noqa: E402


def main():
    rng = np.random.default_rng(20260105)
    out = {}
    # config 5 shape: |z| <= 200, all quadrants (+ both axes and tiny / moderate moduli), orders 0..255
    n = 96
    z = 200 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
    z[:8] = [1e-3, 0.5j, -3.0, 2 + 2j, 25.0, -60j, 199.9, -150 + 100j]
    out["z5"] = z
    ks = np.arange(256, dtype=float)[:, None]
    for kode in (1, 2):
        out[f"J5_{kode}"] = so.evaluate("J", kode, ks, z[None, :])
        out[f"Y5_{kode}"] = so.evaluate("Y", kode, ks, z[None, :])  # companion scale near zeros of J
    # short sequences of every family at fractional starting orders
    m = 160
    zs = 10 ** rng.uniform(-2, 2.2, m) * np.exp(1j * np.pi * (2 * rng.random(m) - 1))
    vs = np.where(rng.random(m) < 0.5, 0.3, 2.5) + rng.integers(0, 20, m)
    out["zs"], out["vs"] = zs, vs
    kk = np.arange(12, dtype=float)[:, None]
    for fn in ("J", "Y", "I", "K", "H1", "H2"):
        for kode in (1, 2):
            out[f"{fn}s_{kode}"] = so.evaluate(fn, kode, vs[None, :] + kk, zs[None, :])
    np.savez_compressed(os.path.join(HERE, "seq_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
