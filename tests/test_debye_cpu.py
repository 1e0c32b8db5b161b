"""CPU: the Debye-direct region of the binned pipeline (sciport_rs_b200/csrc/amos/spb_debye.h; fourth departure from
the published region map) against 50-digit mpmath truth and against the classic drivers.

tests/golden/debye_truth.npz (generator: tests/golden/make_golden_debye.py) holds e^{-|Re z|} I_v(z) and e^{z} K_v(z)
inside and along the border of the region, next to the imaginary and the real axis, for orders up to 86 and moduli up
to 1000 — and, tag 5, the points of a 400 000-point strided sample of the BASELINE cfg3 stream where the pipeline and
scipy.special disagree by more than 8e-14: mpmath arbitrates between the two."""
import numpy as np
import pytest

import helpers
from helpers import FN
from oracle import scipy_oracle as so
from test_pipeline_cpu import gen

FUSE_ALL = 7  # fused large-|w| routine + Debye-direct region + K from the Wronskian route: the library default


def scaled_companion_scale(fn, v, z, truth_i, truth_k):
    """Scale of the absolute bound near zeros for the exponentially scaled functions: I e^{-|Re z|} against
    K e^{-|Re z|}/pi next to the imaginary axis, K e^{z} against pi I e^{z} in the left half plane."""
    with np.errstate(all="ignore"):
        if fn == "I":
            comp = np.abs(truth_k) * np.exp(-z.real - np.abs(z.real)) / np.pi
            return np.maximum(np.abs(truth_i), np.where(np.isfinite(comp) & (np.abs(z.real) <= 1), comp, 0))
        comp = np.pi * np.abs(truth_i) * np.exp(z.real + np.abs(z.real))
        return np.maximum(np.abs(truth_k), np.where(np.isfinite(comp) & (z.real < 0), comp, 0))


def test_debye_region_vs_mpmath_truth():
    g = helpers.load_golden("debye_truth.npz")
    v, z, tag = g["v"], g["z"], g["tag"]
    for fn in ("I", "K"):
        got, ierr, nz, path = helpers.cpu_bessel(fn, 2, v, z, pipeline=True, fuse=FUSE_ALL)
        classic = helpers.cpu_bessel(fn, 2, v, z)[0]
        want = g[f"{fn}_2"]
        scale = scaled_companion_scale(fn, v, z, g["I_2"], g["K_2"])
        ok = (ierr == 0) & (nz == 0)
        assert ok.mean() > 0.999
        err = so.mismatch(got, want, scale)
        # north-star bound, widened only by the conditioning of the argument itself beyond the BASELINE box
        # (|z| > 200: relative perturbation eps of z moves the value by eps |z|)
        tol = 1e-13 * np.maximum(1.0, np.abs(z) / 200.0)
        bad = ok & (err > tol)
        assert not bad.any(), (fn, int(bad.sum()), v[bad][:3], z[bad][:3], err[bad][:3])
        # the region is exercised: most fixture points take the new routine (values differ from the classic flow)
        changed = (got.view(np.uint64) != classic.view(np.uint64)).reshape(-1, 2).any(1)
        assert changed[tag < 5].mean() > 0.6  # (the rest lies in the large-|z| region, which comes first)
        # tag 5: cfg3 points where pipeline and scipy disagree by > 8e-14 -- the pipeline is within 1e-13 of the
        # truth there (asserted above); record how far scipy is, for DESIGN.md
        w = (tag == 5) & ok
        if w.any():
            sc = so.evaluate(fn, 2, v[w], z[w])
            e_sc = so.mismatch(sc, want[w], scale[w])
            print(f"{fn}: cfg3 worst {int(w.sum())} points: pipeline max {err[w].max():.2e}, scipy max {e_sc.max():.2e}")


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "real", "wide", "far"])
@pytest.mark.parametrize("fn", list(FN))
def test_debye_flow_vs_classic_drivers(fn, cfg):
    """Default flow of the library (both single-evaluation routines on) against the point-by-point drivers: identical
    error classes, values inside the parity tolerance (the routines are different algorithms, so not bit for bit)."""
    rng = np.random.default_rng(helpers.seed_of(fn, cfg, "debye"))
    v, z = gen(cfg, 20000, rng)
    for kode in (1, 2):
        a, ia, na = helpers.cpu_bessel(fn, kode, v, z)
        b, ib, nb, path = helpers.cpu_bessel(fn, kode, v, z, pipeline=True, fuse=FUSE_ALL)
        assert np.array_equal(ia, ib) and np.array_equal(na, nb)
        cname = {"J": "Y", "Y": "J", "I": "K", "K": "I", "H1": "J", "H2": "J"}[fn]
        weight = {"J": 1, "Y": 1, "I": 1 / np.pi, "K": np.pi, "H1": 1, "H2": 1}[fn]
        where = {"J": True, "Y": True, "I": np.abs(z.real) <= 1, "K": z.real < 0, "H1": z.imag < 0,
                 "H2": z.imag > 0}[fn]
        comp = np.abs(helpers.cpu_bessel(cname, kode, v, z)[0]) * weight
        scale = np.maximum(np.abs(a), np.where(np.isfinite(comp) & where, comp, 0))
        ok = (ia == 0) & (na == 0)
        err = so.mismatch(b, a, scale)
        cond = 16.0 if cfg in ("wide", "far") else 0.0
        bad = ok & (err > so.tolerance(a, z, v, conditioning=cond))
        assert not bad.any(), (kode, int(bad.sum()), v[bad][:3], z[bad][:3], err[bad][:3])
    if cfg == "cfg3":
        assert (path == 0).mean() > 0.99


def test_debye_region_is_used_and_bounded():
    """Classifier: the region takes a large share of cfg3, nothing of cfg2 / cfg4 (small fixed order), and never a
    point whose parameters are below the calibrated thresholds."""
    rng = np.random.default_rng(7)
    for cfg, lo, hi in (("cfg3", 0.40, 0.60), ("cfg2", 0.0, 0.0), ("cfg4", 0.0, 0.0)):
        v, z = gen(cfg, 50000, rng)
        for fn in ("I", "K", "J"):
            _, full = helpers.cpu_classify(fn, 2, v, z, FUSE_ALL)
            deb = (full & 7) == 6
            assert lo <= deb.mean() <= hi, (cfg, fn, deb.mean())
            if deb.any():
                w = z if fn != "J" else -1j * z
                w = np.where(w.real < 0, -w, w)
                s2 = v * v + w * w
                assert np.all(np.sqrt(np.abs(s2[deb])) >= 36.0 * (1 - 1e-12))
                assert np.all(np.abs(s2[deb]) ** 1.5 / (v[deb] ** 2) >= 240.0 * (1 - 1e-12))


def test_debye_region_on_the_axes_exactly():
    """Real arguments of J / Y / H put the evaluation point w exactly on the imaginary axis (s^2 = v^2 + w^2 on the
    branch cut of the square root beyond the turning point), imaginary arguments of I / K likewise; real arguments
    of I / K sit on the real axis.  Against live scipy and against points a hair off the axis."""
    rng = np.random.default_rng(17)
    n = 4000
    v = 5 + 45 * rng.random(n)
    r = 60 + 140 * rng.random(n)
    for fn, zs in (("J", r + 0j), ("Y", r + 0j), ("H1", r + 0j), ("H2", -r + 0j), ("I", 1j * r), ("I", -1j * r),
                   ("K", 1j * r), ("I", r + 0j), ("K", r + 0j), ("I", -r + 0j), ("J", -r + 0j)):
        for kode in (1, 2):
            got, ierr, nz, path = helpers.cpu_bessel(fn, kode, v, zs, pipeline=True, fuse=FUSE_ALL)
            _, full = helpers.cpu_classify(fn, kode, v, zs, FUSE_ALL)
            assert ((full & 7) == 6).mean() > 0.3  # the Debye-direct region is exercised on the axis
            want = so.evaluate(fn, kode, v, zs)
            scale = so.scale_near_zeros(fn, kode, v, zs, want)
            if fn in ("H1", "H2"):
                # on the real axis (the border of the half plane where H grows) the two terms of the continuation
                # cancel partially: |J| is the size of the terms (measured against 40-digit mpmath at v = 40.67,
                # z = -116.05: this flow 8e-14, the classic drivers 1.7e-13, scipy 1.3e-12 relative to |H2| = 0.009)
                scale = np.maximum(scale, np.abs(so.evaluate("J", kode, v, zs)))
            ok = (ierr == 0) & (nz == 0) & np.isfinite(want.real) & np.isfinite(want.imag)
            err = so.mismatch(got, want, scale)
            bad = ok & (err > so.tolerance(want, zs, v))
            assert not bad.any(), (fn, kode, int(bad.sum()), v[bad][:2], zs[bad][:2], err[bad][:2])


@pytest.mark.parametrize("fuse", [0, 3, 7], ids=["separate", "single-evaluation routines", "default"])
@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "real", "wide", "far"])
def test_ik_pair_matches_separate_drivers(cfg, fuse):
    """The FAM_IK flow (spb_bessel_ik: I and K from one classification, one I pass, one K pass) against the I and K
    drivers: bit for bit with the region passes alone, to the parity tolerance (identical error classes) with the
    fused large-|w| routine and the Debye-direct region on."""
    rng = np.random.default_rng(helpers.seed_of("ik", cfg))
    v, z = gen(cfg, 20000, rng)
    for kode in (1, 2):
        (i_, ei, ni), (k_, ek, nk) = helpers.cpu_bessel_ik(kode, v, z, fuse=fuse)
        a, ia, na = helpers.cpu_bessel("I", kode, v, z)
        b, ib, nb = helpers.cpu_bessel("K", kode, v, z)
        assert np.array_equal(ei, ia) and np.array_equal(ni, na) and np.array_equal(ek, ib) and np.array_equal(nk, nb)
        if fuse == 0:
            same_i = (i_.view(np.uint64) == a.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(a.real) & np.isnan(i_.real))
            same_k = (k_.view(np.uint64) == b.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(b.real) & np.isnan(k_.real))
            assert same_i.all() and same_k.all()
            continue
        cond = 16.0 if cfg in ("wide", "far") else 0.0
        with np.errstate(all="ignore"):
            si = np.maximum(np.abs(a), np.where((np.abs(z.real) <= 1) & np.isfinite(np.abs(b)), np.abs(b) / np.pi
                                                * (np.exp(-z.real - np.abs(z.real)) if kode == 2 else 1.0), 0))
            sk = np.maximum(np.abs(b), np.where((z.real < 0) & np.isfinite(np.abs(a)), np.pi * np.abs(a)
                                                * (np.exp(z.real + np.abs(z.real)) if kode == 2 else 1.0), 0))
        ok = (ia == 0) & (na == 0) & (ib == 0) & (nb == 0)
        assert np.all(so.mismatch(i_, a, si)[ok] <= so.tolerance(a, z, v, conditioning=cond)[ok])
        assert np.all(so.mismatch(k_, b, sk)[ok] <= so.tolerance(b, z, v, conditioning=cond)[ok])
