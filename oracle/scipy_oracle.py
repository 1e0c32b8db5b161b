"""oracle/scipy_oracle.py — TEST INFRASTRUCTURE ONLY.

The reference's arithmetic lives in the un-vendored crate complex-bessel-rs 1.2.0
(/root/reference/Cargo.toml:16), so parity is anchored where the reference anchors it itself: its tests
compare against scipy.special through an embedded CPython (/root/reference/tests/special.rs:10-16,36;
tests/signal/common.rs:154-173).  scipy's complex Bessel / Airy ufuncs are AMOS (TOMS 644), the algorithm
family the dependency wraps.  This module maps (family, kode) to the scipy ufunc, records the sf_error class
per element, and holds the tolerance policy shared by tests/ and bench.py's checker.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / parity legs may import this module.
"""
import numpy as np
import scipy.special as sc

FUNCS = {
    ("J", 1): sc.jv, ("J", 2): sc.jve, ("Y", 1): sc.yv, ("Y", 2): sc.yve,
    ("I", 1): sc.iv, ("I", 2): sc.ive, ("K", 1): sc.kv, ("K", 2): sc.kve,
    ("H1", 1): sc.hankel1, ("H1", 2): sc.hankel1e, ("H2", 1): sc.hankel2, ("H2", 2): sc.hankel2e,
}
COMPANION = {"J": "Y", "Y": "J", "I": "K", "K": "I", "H1": "H2", "H2": "H1"}
# sf_error category -> AMOS ierr (underflow is reported through nz, not ierr)
SF_TO_IERR = {"domain": 1, "overflow": 2, "loss": 3, "no_result": 4, "underflow": 0, "singular": 1, "slow": 5}

EPS = 2.220446049250313e-16
#: headline tolerance of BASELINE.json's north_star: max relative error away from zeros / overflow
RTOL = 1e-13


def evaluate(fn, kode, v, z):
    """scipy value of family `fn` (complex path, even for real input: mod.rs:10 promotes with `.into()`)."""
    z = np.asarray(z, dtype=np.complex128)
    with np.errstate(all="ignore"), sc.errstate(all="ignore"):
        return FUNCS[(fn, kode)](np.asarray(v, dtype=np.float64), z)


def evaluate_airy(which, kode, z):
    z = np.asarray(z, dtype=np.complex128)
    with np.errstate(all="ignore"), sc.errstate(all="ignore"):
        return (sc.airy if kode == 1 else sc.airye)(z)[which]


def error_classes(fn, kode, v, z):
    """Per-element sf_error category ('' when none) by evaluating one element at a time (slow; fixtures only)."""
    f = FUNCS[(fn, kode)]
    v = np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(z))
    out = []
    for vi, zi in zip(np.ravel(v), np.ravel(z)):
        cat = ""
        try:
            with sc.errstate(all="raise"), np.errstate(all="ignore"):
                f(vi, complex(zi))
        except sc.SpecialFunctionError as e:
            msg = str(e)
            if "floating point" in msg:
                # IEEE flag raised by a harmless intermediate (scipy's ufunc loop reports it after the call);
                # the AMOS status itself was clean, otherwise its own error would have been raised first
                out.append("")
                continue
            for key in ("domain", "overflow", "underflow", "loss", "no result", "singular", "slow"):
                if key in msg:
                    cat = key.replace(" ", "_")
                    break
            else:
                cat = "other"
        out.append(cat)
    return np.array(out).reshape(np.shape(z))


def scale_near_zeros(fn, kode, v, z, want):
    """Magnitude the absolute bound near zeros is stated against.  Where a function can vanish by
    cancellation, the scale is the size of the terms that cancel:
      J, Y     max(|J|, |Y|)                      (same envelope everywhere)
      I        max(|I|, |K|/pi) for |Re z| <= 1   (zeros only on the imaginary axis)
      K        max(|K|, pi |I|) for Re z < 0      (zeros only in the left half plane)
      H1 / H2  max(|H|, |J|) in the half plane where H grows (Im z < 0 for H1, > 0 for H2)
    elsewhere it is |f| itself (plain relative error)."""
    z = np.asarray(z, dtype=np.complex128)
    aw = np.abs(want)

    def comp(name, weight, where):
        c = np.abs(evaluate(name, kode, v, z)) * weight
        c = np.where(np.isfinite(c) & where, c, 0.0)
        return np.maximum(aw, c)

    if fn == "J":
        return comp("Y", 1.0, np.ones(z.shape, bool))
    if fn == "Y":
        return comp("J", 1.0, np.ones(z.shape, bool))
    if fn == "I":
        return comp("K", 1.0 / np.pi, np.abs(z.real) <= 1.0)
    if fn == "K":
        return comp("I", np.pi, z.real < 0)
    if fn == "H1":
        return comp("J", 1.0, z.imag < 0)
    return comp("J", 1.0, z.imag > 0)


FNUL = 85.92135864716213


def uniform_region(v, z):
    """Points whose value comes from the uniform (Olver) asymptotic expansions: order above fnul, or |z| above
    fnul where the large-|z| expansion does not apply (2|z| < v^2, v > 1).  There the result is
    phi * exp(-zeta1 + zeta2) * sum with |zeta| ~ max(|z|, v), so its phase carries eps * max(|z|, v)."""
    v = np.asarray(v, dtype=np.float64)
    az = np.abs(z)
    return (v > FNUL) | ((az > FNUL) & (2 * az < v * v) & (v > 1))


def tolerance(want, z=None, v=None, rtol=RTOL, conditioning=0.0):
    """Per-element bound on the normalised error: `rtol` (1e-13, BASELINE.json north_star) for values on scale.
    Conditioning terms widen it where one rounding of an exponent / phase already exceeds rtol:
    * |f| beyond 1e+-49: the result is exp(x), |x| > 112; bound 4 eps |ln|f|| ("away from overflow");
    * |z| > 200 (outside the BASELINE configs' box): bound rtol |z| / 200;
    * uniform-asymptotic region (see uniform_region): bound 32 eps max(|z|, v) (the exponent zeta ~ max(|z|, v) is
      formed in about thirty rounded operations; at |z| = 1253 that is 9e-12);
    * forward / backward recurrences over the order: bound 16 eps v.
    `conditioning` = C > 0 (adversarial distributions outside the BASELINE boxes only) adds the algorithm's own
    error model, C eps max(|z|, v): the published error estimate of TOMS 644 is eps * 10^S with
    S = max(1, |log10 |z||, |log10 v|), i.e. argument reduction costs a factor max(|z|, v); two correct
    implementations differ by a few times that (observed: 1.4e-13 at |z| = 82, v = 15.6; 1.1e-13 at |z| = 37,
    v = 25.6; the tests use C = 16)."""
    with np.errstate(all="ignore"):
        lf = np.abs(np.log(np.abs(want)))
    lf = np.where(np.isfinite(lf), lf, 0.0)
    tol = np.maximum(rtol, 4 * EPS * lf)
    if z is not None:
        tol = tol * np.maximum(1.0, np.abs(z) / 200.0)
        if v is not None:
            vv = np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(z))
            tol = np.maximum(tol, np.where(uniform_region(vv, z), 32 * EPS * np.maximum(np.abs(z), vv), 0.0))
            tol = np.maximum(tol, 16 * EPS * vv)  # order recurrences accumulate ~v roundings (matters above v = 28)
            if conditioning > 0:
                tol = np.maximum(tol, conditioning * EPS * np.maximum(np.abs(z), vv))
    return tol


def mismatch(got, want, scale=None):
    """Normalised error |got - want| / scale with NaN==NaN and equal infinities counted as exact."""
    got = np.asarray(got)
    want = np.asarray(want)
    scale = np.abs(want) if scale is None else scale
    d = np.abs(got - want)
    with np.errstate(all="ignore"):
        r = d / scale
    # a NaN on either component marks "no value": both sides must agree on that, nothing more
    want_nan = np.isnan(want.real) | np.isnan(want.imag)
    got_nan = np.isnan(got.real) | np.isnan(got.imag)
    r = np.where(want_nan & got_nan, 0.0, r)
    exact = (got.real == want.real) & (got.imag == want.imag)
    r = np.where(exact, 0.0, r)
    r = np.where(np.isnan(r), np.inf, r)
    return r
