"""CPU: pin the oracle restatement (oracle/libspb_oracle.so) against the golden vectors generated from
scipy.special, the reference's own test oracle (/root/reference/tests/special.rs:10-16,36).

Tolerance (oracle/scipy_oracle.py): normalised error <= 1e-13 (BASELINE.json north_star) for values on scale,
4 eps |ln|f|| beyond 1e+-49, measured against max(|f|, |companion|) where the function can vanish by cancellation
(absolute bound near zeros)."""
import numpy as np
import pytest

import helpers
from helpers import CATS, FN
from oracle import scipy_oracle as so

FNUL = 85.92135864716213


def scipy_kode2_defect(fn, kode, v):
    """scipy 1.18.1's exponentially scaled K / H / Y return 0 (flagged underflow) or half the value for orders
    above fnul = 85.92 when the continuation to the left half plane goes through the Airy-type uniform
    expansion; mpmath confirms the unscaled value times the scale factor (tests/golden/large_order_truth.npz)."""
    return (kode == 2) & (v > FNUL) & (fn in ("K", "Y", "H1", "H2"))


def sem_ok(cat_code, ierr, nz, z):
    cat = CATS[cat_code]
    if z == 0:
        return ierr in (0, 1, 2)  # scipy special-cases the origin before calling AMOS
    if cat == "":
        return ierr == 0 and nz == 0
    if cat == "underflow":
        return nz != 0
    return {"loss": ierr == 3, "domain": ierr == 1, "overflow": ierr == 2, "no_result": ierr in (4, 5)}.get(cat, False)


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("fn", list(FN))
def test_bessel_vs_golden(fn, kode):
    g = helpers.load_golden("bessel_golden.npz")
    v, z = g["v"], g["z"]
    want, cats = g[f"{fn}_{kode}"], g[f"{fn}_{kode}_err"]
    got, ierr, nz = helpers.cpu_bessel(fn, kode, v, z, sem=1)
    err = so.mismatch(got, want, so.scale_near_zeros(fn, kode, v, z, want))
    skip = scipy_kode2_defect(fn, kode, v) | (cats == CATS.index("loss"))
    bad = (err > so.tolerance(want, z, v)) & ~skip
    assert not bad.any(), f"{bad.sum()} value mismatches, first at v={v[bad][0]} z={z[bad][0]} err={err[bad][0]:.2e}"
    ok = np.array([sem_ok(c, e, n, zz) for c, e, n, zz in zip(cats, ierr, nz, z)]) | skip
    assert ok.all(), f"{(~ok).sum()} ierr/nz class mismatches, first at v={v[~ok][0]} z={z[~ok][0]}"
    # half-precision region: still close
    loss = (cats == CATS.index("loss")) & np.isfinite(want.real)
    if loss.any():
        assert np.all(err[loss] < 1e-6)


def test_large_order_scaled_vs_mpmath_truth():
    """Where scipy's kode=2 large-order values are defective, the restatement is checked against mpmath."""
    g = helpers.load_golden("large_order_truth.npz")
    for fn in ("K", "H1", "H2", "Y"):
        got, ierr, nz = helpers.cpu_bessel(fn, 2, g["v"], g["z"])
        want = g[f"{fn}_2"]
        err = np.abs(got - want) / np.abs(want)
        assert np.all(err < 5e-13), f"{fn}: {err.max():.2e}"


def large_order_truth_errors(evaluate):
    """Normalised errors of `evaluate(fn, kode, v, z) -> (values, ierr, nz)` against tests/golden/large_order_truth2.npz
    (mpmath at 50 + 0.9 |z| digits or more, generator tests/golden/make_golden_large_order.py): every order > fnul
    point of bessel_golden.npz plus 120 seeded points with orders up to 300 (scaled functions), and the points where
    scipy reports a loss of half the digits (both scalings).  Yields (label, err over the points both sides resolve)."""
    g = helpers.load_golden("large_order_truth2.npz")
    v, z = g["v"], g["z"]
    comp_of = {"K": ("I", np.pi), "Y": ("J", 1.0), "H1": ("J", 1.0), "H2": ("J", 1.0), "J": ("Y", 1.0), "I": ("K", 1 / np.pi)}
    for fn in ("K", "H1", "H2", "Y", "J", "I"):
        got, ierr, nz = evaluate(fn, 2, v, z)
        want = g[f"{fn}_2"]
        ok = np.isfinite(want.real) & (ierr == 0) & (nz == 0)
        assert ok.sum() >= 0.7 * len(v), (fn, int(ok.sum()))
        # absolute bound where the function is a difference of two larger terms (continuation / Hankel combination):
        # the companion's truth, put on the scale of fn's scaling factor, is the size of those terms
        cname, wgt = comp_of[fn]
        comp = np.abs(g[f"{cname}_2"]) * wgt
        with np.errstate(all="ignore"):
            if fn == "K":
                comp = np.where(z.real < 0, comp * np.exp(2 * z.real), 0.0)  # e^{z} * e^{|Re z|} I
            elif fn == "I":
                comp = np.where(np.abs(z.real) <= 1, comp * np.exp(-2 * np.abs(z.real)), 0.0)
            elif fn in ("H1", "H2"):
                grow = z.imag < 0 if fn == "H1" else z.imag > 0
                comp = np.where(grow, 0.0, comp * np.exp(-2 * np.abs(z.imag)))  # recessive side only: no cancellation
        scale = np.maximum(np.abs(want), np.where(np.isfinite(comp), comp, 0.0))
        with np.errstate(all="ignore"):
            err = np.abs(got - want) / scale
        # measured: at most 1.9e-13 for orders 86 .. 300 (the tolerance policy would allow 6e-13 .. 2.7e-12 there)
        yield f"{fn}_2", err[ok], np.minimum(so.tolerance(want, z, v)[ok], 3e-13)
    lv, lz = g["loss_v"], g["loss_z"]
    for kode in (1, 2):
        for fn in ("K", "H1", "H2", "Y", "J", "I"):
            got, ierr, nz = evaluate(fn, kode, lv, lz)
            want = g[f"loss_{fn}_{kode}"]
            ok = np.isfinite(want.real) & np.isin(ierr, (0, 3)) & (nz == 0)
            with np.errstate(all="ignore"):
                err = np.abs(got - want) / np.abs(want)
            yield f"loss {fn}_{kode}", err[ok], np.full(int(ok.sum()), 1e-13)


def test_large_orders_and_loss_points_vs_mpmath_truth():
    """The two classes of points the scipy comparison skips (scipy_kode2_defect, category "loss") against mpmath."""
    for label, err, tol in large_order_truth_errors(lambda fn, kode, v, z: helpers.cpu_bessel(fn, kode, v, z)):
        assert np.all(err <= tol), f"{label}: max {err.max():.2e} ({int((err > tol).sum())} of {err.size} above tolerance)"


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("which", [0, 1, 2, 3])
def test_airy_vs_golden(which, kode):
    g = helpers.load_golden("airy_golden.npz")
    z = g["z"]
    want = g[f"{('Ai', 'Aip', 'Bi', 'Bip')[which]}_{kode}"]
    got, ierr, nz = helpers.cpu_airy(which, kode, z)
    nanw = np.isnan(want.real)
    assert np.all(np.isin(ierr[nanw], (1, 2, 4, 5))), "oracle NaN must carry a failing ierr"
    ok = ~nanw & (ierr == 0)
    # conditioning: the phase of the result is Im zeta = Im (2/3) z^(3/2); one rounding of zeta moves it by eps |zeta|
    zeta = np.abs(2.0 / 3.0 * z * np.sqrt(z))
    tol = 1e-13 + 4 * so.EPS * zeta
    err = so.mismatch(got, want)
    # near the zeros on the negative real axis use the envelope |Ai| ~ |Bi| as scale
    comp = g[f"{('Bi', 'Bip', 'Ai', 'Aip')[which]}_{kode}"]
    scale = np.maximum(np.abs(want), np.where((z.real < 0) & np.isfinite(np.abs(comp)), np.abs(comp), 0))
    err = np.minimum(err, so.mismatch(got, want, scale))
    bad = ok & (err > tol)
    assert not bad.any(), f"{bad.sum()} mismatches, first z={z[bad][0]} err={err[bad][0]:.2e}"
    half = ~nanw & (ierr == 3)
    assert np.all(err[half] < 1e-6)


def test_gamma_family_vs_golden():
    g = helpers.load_golden("gamma_golden.npz")
    x = g["x"]
    for which, key, tol in ((0, "gamma", 1e-13), (2, "sinc", 1e-15)):
        got = helpers.cpu_gamma(which, x)
        want = g[key]
        with np.errstate(all="ignore"):
            err = np.abs(got - want) / np.abs(want)
        err[(got == want) | (np.isnan(got) & np.isnan(want))] = 0
        # reflection: sin(pi x) carries a relative error of pi |x| eps
        bound = tol + np.where(x < 0, 4 * so.EPS * np.pi * np.abs(x), 0.0)
        assert np.all(err <= bound), f"{key}: {np.nanmax(err - bound):.2e} at x={x[np.nanargmax(err - bound)]}"
    got = helpers.cpu_gamma(1, x)
    want = g["gammaln"]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    err[(got == want) | (np.isnan(got) & np.isnan(want))] = 0
    sub = np.abs(x) < 1e-300  # scipy treats subnormals as 0
    assert np.all(err[~sub] <= 1e-13), f"gammaln {np.nanmax(err[~sub]):.2e}"
    zc = g["zc"]
    got = helpers.cpu_gamma(11, zc)
    want = g["cloggamma"]
    d = got - want
    err = np.abs(d) / np.maximum(np.abs(want), 1.0)
    err[np.isnan(want.real) & np.isnan(got.real)] = 0
    assert np.all(err <= 1e-13), f"loggamma {np.nanmax(err):.2e}"
    got = helpers.cpu_gamma(10, zc)
    want = g["cgamma"]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.abs(want)
    err[(got == want) | (np.isnan(want.real) & np.isnan(got.real))] = 0
    bound = 1e-13 + 4 * so.EPS * np.abs(g["cloggamma"])  # gamma = exp(loggamma)
    bound[~np.isfinite(bound)] = 1e-13
    assert np.all(err <= bound), f"cgamma {np.nanmax(err - bound):.2e}"
