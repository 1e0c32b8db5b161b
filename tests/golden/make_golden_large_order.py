#!/usr/bin/env python
"""Generate tests/golden/large_order_truth2.npz: mpmath values (50 + 0.9 |z| digits or more) of the exponentially scaled K, H1, H2, Y
(kve, hankel1e, hankel2e, yve) and of the unscaled / scaled functions at the scipy "loss" points.

Why: scipy 1.18.1's kode = 2 results for orders above fnul = 85.92 are defective in the left half plane (0 with an
underflow flag, or half the value), so tests/test_oracle_golden.py excludes those points from the scipy comparison
(scipy_kode2_defect), and it excludes the points where scipy reports a loss of half the digits (ierr 3).  This
fixture puts an independent oracle behind BOTH exclusions: every order > fnul point of bessel_golden.npz (206),
every "loss" point of it (category CATS.index("loss")), and 120 more seeded points with orders in (86, 300] in all
four quadrants.  large_order_truth.npz (48 points, round 1) stays as it is.

TEST INFRASTRUCTURE.  Needs mpmath.  Seeded; run from the repo root:  python tests/golden/make_golden_large_order.py
"""
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402

FNUL = 85.92135864716213


def _eval(fn, vv, zz):
    """One function at the working precision.  H1 / H2 where they are the recessive solution (upper / lower half
    plane) come from K at the rotated point: J + iY cancels there to exp(-2 |Im z|) of its terms."""
    if fn == "K":
        return mp.besselk(vv, zz)
    if fn == "H1":
        if zz.imag > 0 or (zz.imag == 0 and zz.real > 0):
            return 2 / (mp.pi * 1j) * mp.exp(-1j * mp.pi * vv / 2) * mp.besselk(vv, -1j * zz)  # -pi/2 < arg z <= pi
        return mp.hankel1(vv, zz)
    if fn == "H2":
        if zz.imag < 0 or (zz.imag == 0 and zz.real > 0):
            return 2j / mp.pi * mp.exp(1j * mp.pi * vv / 2) * mp.besselk(vv, 1j * zz)  # -pi < arg z <= pi/2
        return mp.hankel2(vv, zz)
    return {"Y": mp.bessely, "J": mp.besselj, "I": mp.besseli}[fn](vv, zz)


def _scale(fn, zz):
    return {"K": mp.exp(zz), "H1": mp.exp(-1j * zz), "H2": mp.exp(1j * zz), "Y": mp.exp(-abs(zz.imag)),
            "J": mp.exp(-abs(zz.imag)), "I": mp.exp(-abs(zz.real))}[fn]


def truth(v, z, scaled, check=None, digits=None):
    """K, H1, H2, Y, J, I at (v, z); scaled: times e^z, e^{-iz}, e^{iz}, e^{-|Im z|}, e^{-|Im z|}, e^{-|Re z|}.
    mpmath sums hypergeometric series whose terms reach e^{|z|} while the result may be e^{-|z|}: the working
    precision is 50 + 0.9 |z| digits, and a value that differs from `check` (the CPU restatement, only used to
    decide WHEN to recompute) by more than 1e-12 is recomputed with 80 more digits; what is stored is always
    mpmath's own value."""
    out = {k: np.full(len(v), np.nan + 0j) for k in ("K", "H1", "H2", "Y", "J", "I")}
    redo = 0
    for i in range(len(v)):
        for k in out:
            for extra in (0, 80):
                mp.mp.dps = (digits if digits else 50 + int(0.9 * abs(z[i]))) + extra
                zz, vv = mp.mpc(z[i].real, z[i].imag), mp.mpf(v[i])
                try:
                    x = _eval(k, vv, zz) * (_scale(k, zz) if scaled else 1)
                except Exception:
                    break
                a = abs(x)
                if not (a < mp.mpf("1e300") and a > mp.mpf("1e-300")):
                    break
                out[k][i] = complex(x)
                c = None if check is None else check[k][i]
                if c is None or not np.isfinite(c.real) or abs(out[k][i] - c) <= 1e-12 * abs(out[k][i]):
                    break
                redo += extra == 0
    print("recomputed at higher precision:", redo)
    return out


def main():
    g = helpers.load_golden("bessel_golden.npz")
    v, z = g["v"], g["z"]
    big = np.flatnonzero(v > FNUL)
    loss = np.zeros(len(v), bool)
    for fn in ("J", "Y", "I", "K", "H1", "H2"):
        for kode in (1, 2):
            loss |= g[f"{fn}_{kode}_err"] == helpers.CATS.index("loss")
    loss = np.flatnonzero(loss & (z != 0))
    rng = np.random.default_rng(20261021)
    n = 120
    ve = np.where(rng.random(n) < 0.3, rng.integers(87, 300, n).astype(float), rng.uniform(86.0, 300.0, n))
    ze = np.minimum(ve * 10 ** rng.uniform(-0.5, 0.4, n), 320.0) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
    vb, zb = np.concatenate([v[big], ve]), np.concatenate([z[big], ze])

    def port(vv, zz, kode):
        return {k: helpers.cpu_bessel(k, kode, vv, zz)[0] for k in ("K", "H1", "H2", "Y", "J", "I")}

    t2 = truth(vb, zb, True, port(vb, zb, 2))
    # (the "loss" points are real arguments of 4e4 .. 2e9: mpmath switches to asymptotic series there; 60 digits)
    l1 = truth(v[loss], z[loss], False, digits=60)
    l2 = truth(v[loss], z[loss], True, digits=60)
    out = os.path.join(ROOT, "tests", "golden", "large_order_truth2.npz")
    np.savez_compressed(out, v=vb, z=zb, golden_index=np.concatenate([big, np.full(n, -1)]),
                        **{f"{k}_2": t2[k] for k in t2},
                        loss_v=v[loss], loss_z=z[loss], loss_index=loss,
                        **{f"loss_{k}_1": l1[k] for k in l1}, **{f"loss_{k}_2": l2[k] for k in l2})
    print("wrote", out, len(vb), "large-order points,", len(loss), "loss points")


if __name__ == "__main__":
    main()
