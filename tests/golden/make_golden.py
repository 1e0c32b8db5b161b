#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from scipy.special — the reference's own test oracle
(/root/reference/tests/special.rs:10-16,36; tests/signal/common.rs:154-173).  The reference holds no fixed
vectors (all its tests draw unseeded random inputs and compare with live scipy), so these files pin the same
comparison with a fixed seed.  Run:  python tests/golden/make_golden.py   (scipy 1.18.1 was used)."""
import os
import sys

import numpy as np
import scipy
import scipy.signal
import scipy.special as sc

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import scipy_oracle as so  # noqa: E402

CATS = ["", "domain", "overflow", "underflow", "loss", "no_result", "singular", "slow", "other"]


def points(rng):
    vs, zs = [], []

    def add(v, z):
        vs.append(np.broadcast_to(np.asarray(v, float), np.shape(z)).ravel())
        zs.append(np.asarray(z, complex).ravel())

    n = 600
    # BASELINE config shapes, downscaled
    a = 200 / np.sqrt(2)
    add(2.5, rng.uniform(-a, a, n) + 1j * rng.uniform(-a, a, n))                                   # cfg2
    add(50 * rng.random(n), 200 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1)))    # cfg3
    add(2.5, 10 ** (-3 + 7 * rng.random(n)) * np.exp(1j * np.pi * (2 * rng.random(n) - 1)))        # cfg4
    add(rng.integers(0, 2, n).astype(float), 100 * rng.random(n) + 0j)                             # cfg1 (J0/J1/Y0/Y1)
    # every region: orders 0..120 (integer, half-integer, generic), |z| log-uniform, all quadrants + axes
    m = 900
    v = np.where(rng.random(m) < 0.3, rng.integers(0, 120, m), np.where(rng.random(m) < 0.3,
                 rng.integers(0, 120, m) + 0.5, 120 * rng.random(m) ** 2))
    r = 10 ** rng.uniform(-3, 2.6, m)
    th = np.pi * (2 * rng.random(m) - 1)
    axis = rng.integers(0, 8, m)
    th = np.where(axis == 0, 0.0, np.where(axis == 1, np.pi, np.where(axis == 2, np.pi / 2,
                  np.where(axis == 3, -np.pi / 2, th))))
    z = r * np.exp(1j * th)
    z = np.where(axis == 0, r + 0j, np.where(axis == 1, -r + 0j, np.where(axis == 2, 1j * r, np.where(axis == 3, -1j * r, z))))
    add(v, z)
    # reference call patterns: i0 on [0,1) (tests/special.rs:25-30), Kaiser beta in [2,30), besselap poles
    add(0.0, rng.random(100) + 0j)
    add(0.0, 30 * rng.random(100) + 0j)
    for order in (1, 5, 12, 25, 37, 49):
        _, p, _ = scipy.signal.besselap(order, norm="delay")
        for dv in (-0.5, 0.5, 1.5):
            add(order + dv, p)
    # edge cases
    ev = [2.5, 0.0, 2.5, 2.5, 2.5, 2.5, 2.5, 2.5, 3e4, 2.5, 2.5, 2.5, 2.5, 2.5, 0.0, 1.0, 2.5, 60.0, 100.0, 150.0, 250.0, 90.0]
    ez = [0, 0, 800, 800j, -800, 1e-130, 1e-130j, 4e4, 1, 5e7, 3e15, 2e9, -3 + 0j, complex(-3, -0.0), 1e-320, 1e-300,
          700 + 5j, 0.5 + 0.1j, 30 - 20j, 140 + 10j, 200j, -95 + 2j]
    add(np.array(ev), np.array(ez, dtype=complex))
    return np.concatenate(vs), np.concatenate(zs)


def main():
    rng = np.random.default_rng(20260101)
    v, z = points(rng)
    out = {"v": v, "z": z}
    for (fn, kode), f in so.FUNCS.items():
        out[f"{fn}_{kode}"] = so.evaluate(fn, kode, v, z)
        cats = so.error_classes(fn, kode, v, z)
        out[f"{fn}_{kode}_err"] = np.array([CATS.index(c) for c in cats], dtype=np.int8)
        print(fn, kode, "errors:", {c: int((cats == c).sum()) for c in set(cats) if c})
    np.savez_compressed(os.path.join(HERE, "bessel_golden.npz"), **out)

    za = np.concatenate([
        10 ** rng.uniform(-3, 3, 1500) * np.exp(1j * np.pi * (2 * rng.random(1500) - 1)),
        rng.uniform(-30, 30, 300) + 0j, 1j * rng.uniform(-30, 30, 100),
        np.array([0, 1e-320, 1, -1, 1e4, -1e4, 2000j, 1.2e6, 1e-20, 1 + 1e-17j], dtype=complex)])
    air = {"z": za}
    for kode in (1, 2):
        for w, name in enumerate(("Ai", "Aip", "Bi", "Bip")):
            air[f"{name}_{kode}"] = so.evaluate_airy(w, kode, za)
    np.savez_compressed(os.path.join(HERE, "airy_golden.npz"), **air)

    x = np.concatenate([100 * rng.random(2000), rng.uniform(-170, 171, 1000), 1 + 1e-3 * rng.standard_normal(50),
                        2 + 1e-3 * rng.standard_normal(50), 10 ** rng.uniform(-300, 300, 200),
                        np.array([0.0, -0.0, -1.0, -2.0, 171.6, 171.7, -0.5, -170.5, -180.5, np.inf, np.nan, 0.5, 1, 2, 3])])
    zc = np.concatenate([rng.uniform(-50, 50, 1500) + 1j * rng.uniform(-50, 50, 1500),
                         rng.uniform(-8, 8, 1500) + 1j * rng.uniform(-8, 8, 1500),
                         1 + 0.3 * (rng.random(200) - 0.5) + 0.3j * (rng.random(200) - 0.5),
                         2 + 0.3 * (rng.random(200) - 0.5) + 0.3j * (rng.random(200) - 0.5),
                         rng.uniform(-20, 20, 200) + 0j, np.array([0, -1, -2, 1, 2, 0.5, 1e-5j, 200j, -200.5 + 0.1j], dtype=complex)])
    with np.errstate(all="ignore"), sc.errstate(all="ignore"):
        np.savez_compressed(os.path.join(HERE, "gamma_golden.npz"), x=x, gamma=sc.gamma(x), gammaln=sc.gammaln(x),
                            sinc=np.sinc(x), zc=zc, cgamma=sc.gamma(zc), cloggamma=sc.loggamma(zc))
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        f.write(f"scipy {scipy.__version__}\nnumpy {np.__version__}\nseed 20260101\n")
    print("points:", v.size, za.size, x.size, zc.size)


if __name__ == "__main__":
    main()
