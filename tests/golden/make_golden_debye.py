#!/usr/bin/env python
"""Generate tests/golden/debye_truth.npz: 50-digit mpmath values of e^{-|Re z|} I_v(z) and e^{z} K_v(z)

 (a) inside and along the border of the Debye-direct region of the binned pipeline (spb_debye.h: |s| >= S0,
     |s|^3 / v^2 >= P0 with s = sqrt(v^2 + z^2)): cfg3-distributed points, points hugging the border from inside,
     points next to the imaginary axis (where the second exponential of I matters) and next to the real axis,
     orders up to 86 and moduli up to 1000;
 (b) the cfg3 points (strided sample of the 2^28-point stream of BASELINE config 3) where the pipeline and
     scipy.special disagree by more than 8e-14 (normalised): the arbiter between the two.

TEST INFRASTRUCTURE.  Needs mpmath and the built CPU checker (oracle/).  Seeded; run from the repo root:
    python tests/golden/make_golden_debye.py
"""
import os
import sys

import mpmath as mp
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import helpers  # noqa: E402
from oracle import scipy_oracle as so  # noqa: E402

mp.mp.dps = 50
S0, P0 = 36.0, 240.0  # keep in step with SPB_DEBYE_S0 / SPB_DEBYE_P0 (spb_debye.h)


def params(v, z):
    w = np.where(z.real < 0, -z, z)
    s2 = v * v + w * w
    return np.sqrt(np.abs(s2)), np.abs(s2) ** 1.5 / np.maximum(v * v, 1e-300)


def truth(v, z):
    ti = np.empty(len(v), np.complex128)
    tk = np.empty(len(v), np.complex128)
    for i in range(len(v)):
        zz = mp.mpc(z[i].real, z[i].imag)
        vv = mp.mpf(v[i])
        ti[i] = complex(mp.besseli(vv, zz) * mp.exp(-abs(zz.real)))
        tk[i] = complex(mp.besselk(vv, zz) * mp.exp(zz))
    return ti, tk


def main():
    rng = np.random.default_rng(20261020)
    vs, zs, tags = [], [], []

    def take(v, z, tag, n):
        s, p = params(v, z)
        keep = (s >= S0) & (p >= P0) & (v > 1.0)
        idx = np.flatnonzero(keep)[:n]
        vs.append(v[idx]); zs.append(z[idx]); tags.append(np.full(len(idx), tag))

    m = 200000
    # 0: cfg3 distribution
    take(50 * rng.random(m), 200 * rng.random(m) * np.exp(1j * np.pi * (2 * rng.random(m) - 1)), 0, 500)
    # 1: hugging the border from inside (P within 1.3 x P0 or |s| within 1.1 x S0)
    v = 86 * rng.random(m); z = (15 + 300 * rng.random(m)) * np.exp(1j * np.pi * (2 * rng.random(m) - 1))
    s, p = params(v, z)
    sel = ((p < 1.3 * P0) | (s < 1.1 * S0))
    take(v[sel], z[sel], 1, 700)
    # 2: next to the imaginary axis, beyond the turning point (second exponential of I of full size)
    v = 60 * rng.random(m); r = 20 + 400 * rng.random(m)
    th = (np.pi / 2) * rng.choice([-1, 1], m) + 10 ** rng.uniform(-8, -0.7, m) * rng.choice([-1, 1], m)
    take(v, r * np.exp(1j * th), 2, 400)
    # 3: next to the real axis (both signs of Re z) and exactly on it
    v = 86 * rng.random(m); r = 20 + 400 * rng.random(m)
    th = np.pi * rng.integers(0, 2, m) + 10 ** rng.uniform(-10, -1, m) * rng.choice([-1, 0, 1], m)
    take(v, r * np.exp(1j * th), 3, 300)
    # 4: large moduli / orders up to fnul
    take(1 + 85 * rng.random(m), 10 ** rng.uniform(2, 3, m) * np.exp(1j * np.pi * (2 * rng.random(m) - 1)), 4, 300)
    # 5: cfg3 stream points where the pipeline (Debye-direct region on) and scipy disagree most
    idx = np.arange(0, 1 << 28, (1 << 28) // 400000, dtype=np.int64)[:400000]
    v3, z3 = bench.cfg3_points(idx)
    worst = []
    for fn in ("I", "K"):
        got = helpers.cpu_bessel(fn, 2, v3, z3, pipeline=True, fuse=3)[0]
        want = so.evaluate(fn, 2, v3, z3)
        err = so.mismatch(got, want, so.scale_near_zeros(fn, 2, v3, z3, want))
        worst.append(np.flatnonzero(np.isfinite(err) & (err > 8e-14)))
    w = np.unique(np.concatenate(worst))
    vs.append(v3[w]); zs.append(z3[w]); tags.append(np.full(len(w), 5))
    v = np.concatenate(vs); z = np.concatenate(zs); tag = np.concatenate(tags).astype(np.int32)
    ti, tk = truth(v, z)
    out = os.path.join(ROOT, "tests", "golden", "debye_truth.npz")
    np.savez_compressed(out, v=v, z=z, tag=tag, I_2=ti, K_2=tk, cfg3_index=np.concatenate([np.full(len(v) - len(w), -1), idx[w]]))
    print("wrote", out, len(v), "points;", {int(t): int((tag == t).sum()) for t in np.unique(tag)})


if __name__ == "__main__":
    main()
