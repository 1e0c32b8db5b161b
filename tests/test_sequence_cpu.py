"""CPU: pin the order-sequence form of the oracle restatement (out[k][i] = F_{v_i+k}(z_i), BASELINE config 5)
against tests/golden/seq_golden.npz (scipy.special, one order at a time).

Tolerance: a sequence is produced by ONE evaluation at the top (or bottom) of the range followed by the
three-term recurrence over the order (backward from the top member for I / J, forward for K / H), so a member
carries up to ~n_orders roundings relative to scipy's independent per-order evaluation: normalised error
<= max(1e-13, 8 eps (v + n_orders)) + 4 eps |ln|f|| (both sides are that far
from the mpmath truth at the small-|z| / high-order corner), against max(|J|, |Y|) for J and Y (absolute bound near
zeros).  For the config-5 sweep (256 orders) that is 4.5e-13."""
import numpy as np
import pytest

import helpers
from oracle import scipy_oracle as so


def seq_tolerance(want, v0, n_orders):
    k = np.zeros((n_orders, 1)) + np.asarray(v0, dtype=float)[None, :] + n_orders
    with np.errstate(all="ignore"):
        lf = np.abs(np.log(np.abs(want)))
    lf = np.where(np.isfinite(lf), lf, 0.0)
    return np.maximum(1e-13, 8 * so.EPS * k) + 4 * so.EPS * lf


def check_seq(got, ierr, nz, want, comp, v0, label, z=None):
    """comp = companion values (Y for J, J for Y): the scale near zeros is max(|f|, |comp|), applied only where
    zeros can occur (order <= |z|; beyond the turning point J and Y have no zeros and plain relative error holds)."""
    n_orders = want.shape[0]
    scale = np.abs(want)
    if comp is not None:
        c = np.where(np.isfinite(np.abs(comp)), np.abs(comp), 0.0)
        if z is not None:
            order = np.arange(n_orders, dtype=float)[:, None] + np.asarray(v0, dtype=float)[None, :]
            c = np.where(order <= np.abs(z)[None, :], c, 0.0)
        scale = np.maximum(scale, c)
    err = so.mismatch(got, want, scale)
    # Underflow band: below exp(-alim) = 1.8e-289 the reference flushes a member to zero or keeps it depending on
    # rounding noise in its negligible component (the `uchk` test of the algorithm: J_177(-3) = 4e-292 is
    # flushed, J_176(-3) = 5e-290 is kept); the write-once sweep flushes on the modulus at exp(-elim) = 3.9e-305.
    # Both count as underflow: inside the band either a zero or the value itself is accepted.
    BAND = 2e-288
    live = np.isfinite(want.real) & np.isfinite(want.imag) & (np.abs(want) >= BAND) & (ierr[None, :] == 0)
    bad = live & (err > seq_tolerance(want, v0, n_orders))
    assert not bad.any(), f"{label}: {bad.sum()} mismatches, worst {np.where(live, err, 0).max():.2e}"
    band = np.isfinite(want.real) & (np.abs(want) < BAND) & (ierr[None, :] == 0)
    assert np.all(np.abs(got[band]) < 10 * BAND), f"{label}: members in the underflow band must stay there"
    inband = band & (np.abs(want) > 1e-300) & (got != 0)
    assert np.all(err[inband] <= 1e-9), f"{label}: kept members of the underflow band must still be accurate"
    return np.where(live, err, 0).max()


@pytest.mark.parametrize("kode", [1, 2])
def test_j_sweep_0_255_vs_golden(kode):
    g = helpers.load_golden("seq_golden.npz")
    z = g["z5"]
    got, ierr, nz = helpers.cpu_bessel_seq("J", kode, 0.0, z, 256)
    assert np.all(ierr == 0)
    check_seq(got, ierr, nz, g[f"J5_{kode}"], g[f"Y5_{kode}"], np.zeros(z.size), f"J sweep kode={kode}", z=z)
    # nz counts the leading (highest-order) members that underflowed
    assert np.array_equal(nz, (got == 0).sum(axis=0))
    # the classic sequence driver (array form, re-reading its members) agrees with the write-once sweep
    classic, ierr_c, nz_c = helpers.cpu_bessel_seq("J", kode, 0.0, z, 256, classic=True)
    check_seq(got, ierr, nz, classic, g[f"Y5_{kode}"], np.zeros(z.size), f"sweep vs classic kode={kode}", z=z)


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("fn", ["J", "I"])
def test_sweep_device_flow_matches_classic_flow(fn, kode):
    """The CUDA sweep kernels take the K pair of the Wronskian normalisation from the large-|w| series (|w| >= rl) or
    from a dense bknu() pre-pass (k_sequence.cu) instead of calling bknu() inside the sweep: same members to a few
    ulp of the normalising factor, identical codes, on the config-5 stream and on fractional start orders."""
    rng = np.random.default_rng(helpers.seed_of("sweep-device-flow", fn, kode))
    n = 3000
    r = 200.0 * rng.random(n)
    th = np.pi * (2.0 * rng.random(n) - 1.0)
    z = r * np.exp(1j * th)
    for v0, orders in ((0.0, 256), (0.3, 64), (7.5, 40)):
        a, ierr_a, nz_a = helpers.cpu_bessel_seq(fn, kode, v0, z, orders)
        b, ierr_b, nz_b = helpers.cpu_bessel_seq(fn, kode, v0, z, orders, device_flow=True)
        assert np.array_equal(ierr_a, ierr_b) and np.array_equal(nz_a, nz_b)
        assert np.array_equal(a == 0, b == 0)
        scale = np.maximum(np.abs(a), np.abs(a).max(axis=0, keepdims=True) * 1e-300)
        on = np.abs(a) > 0
        err = np.abs(a - b)[on] / scale[on]
        assert err.max() <= 64 * so.EPS, f"{fn} kode {kode} v0 {v0}: {err.max():.2e}"
        # both K routines are in play
        assert (np.abs(z) < 21.7).any() and (np.abs(z) > 21.9).any()


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("fn", ["J", "Y", "I", "K", "H1", "H2"])
def test_short_sequences_vs_golden(fn, kode):
    g = helpers.load_golden("seq_golden.npz")
    z, v = g["zs"], g["vs"]
    got, ierr, nz = helpers.cpu_bessel_seq(fn, kode, v, z, 12)
    comp = g[f"{so.COMPANION[fn]}s_{kode}"] if fn in ("J", "Y") else None
    check_seq(got, ierr, nz, g[f"{fn}s_{kode}"], comp, v, f"{fn} kode={kode}", z=z)


def test_sequence_member_zero_equals_single_order():
    """Member 0 of a length-1 'sequence' is the single-order driver bit for bit."""
    g = helpers.load_golden("seq_golden.npz")
    z, v = g["zs"], g["vs"]
    for fn in ("J", "I", "K", "H1", "H2"):
        a, ea, na = helpers.cpu_bessel_seq(fn, 1, v, z, 1)
        b, eb, nb = helpers.cpu_bessel(fn, 1, v, z)
        assert np.array_equal(a[0].view(np.float64), b.view(np.float64), equal_nan=True)
        assert np.array_equal(ea, eb) and np.array_equal(na, nb)
