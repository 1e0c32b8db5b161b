"""CPU: the lean large-|zeta| Airy routines (sciport_rs_b200/csrc/amos/spb_airy_lean.h, the flow of the GPU kernels
k_airy_large / k_airy_sorted) against the complete routines (error codes identical; Bi values bit for bit, Ai values
to rounding: K from the Hankel series instead of the Miller algorithm) and against live scipy.special.airy / airye
on the distribution of BASELINE config 4 (|z| log-uniform 1e-3 .. 1e4, all quadrants) and on the axes."""
import numpy as np
import pytest

import helpers
from oracle import scipy_oracle as so


def cfg4_points(n, rng):
    z = 10 ** (-3 + 7 * rng.random(n)) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
    m = n // 10
    z[:m] = np.abs(z[:m])                      # positive real axis
    z[m:2 * m] = -np.abs(z[m:2 * m])           # negative real axis (zeros of Ai, Bi)
    z[2 * m:3 * m] = 1j * z[2 * m:3 * m].real  # imaginary axis
    return z


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("which", [0, 1, 2, 3], ids=["Ai", "Aip", "Bi", "Bip"])
def test_lean_airy_flow_vs_complete_routines_and_scipy(which, kode):
    rng = np.random.default_rng(helpers.seed_of("airy-lean", which, kode))
    z = cfg4_points(60000, rng)
    got, ierr, nz, path = helpers.cpu_airy_flow(which, kode, z)
    ref, ierr_r, nz_r = helpers.cpu_airy(which, kode, z)
    assert path.mean() > 0.35  # |z| >= 10.2: the lean routines carry their share of a log-uniform batch
    assert np.array_equal(ierr, ierr_r) and np.array_equal(nz, nz_r)
    same = (got.view(np.uint64) == ref.view(np.uint64)).reshape(-1, 2).all(1)
    if which >= 2:
        assert same.all()  # Bi: the region driver of I enters the large-|zeta| expansion anyway
    want = so.evaluate_airy(which, kode, z)
    comp = so.evaluate_airy((which + 2) % 4, kode, z)
    ok = (ierr == 0) & (nz == 0) & np.isfinite(want.real) & np.isfinite(want.imag)
    # near the zeros on the negative real axis the envelope |Ai| ~ |Bi| is the scale
    with np.errstate(all="ignore"):
        scale = np.maximum(np.abs(want), np.where((z.real < 0) & np.isfinite(np.abs(comp)), np.abs(comp), 0))
    zeta = (2.0 / 3.0) * np.abs(z) ** 1.5
    tol = so.tolerance(want) + 4 * so.EPS * zeta
    err = so.mismatch(got, want, scale)
    assert np.all(err[ok] <= tol[ok]), float(np.max((err - tol)[ok]))
    # and against the complete routine: a few ulp of the conditioning of zeta
    err_r = so.mismatch(got, ref, np.maximum(scale, np.abs(ref)))
    assert np.all(err_r[ok] <= 1e-14 + 4 * so.EPS * zeta[ok])
