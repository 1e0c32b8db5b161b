"""CPU: the restructured flow the GPU kernels run (prepare -> I pass -> K pass -> finish, general path on any
hand-over; sciport_rs_b200/csrc/amos/spb_family.h) must reproduce the point-by-point drivers bit for bit; the
fused I + K routine of the large-|w| region, which computes K differently, to a few ulp with identical codes."""
import numpy as np
import pytest

import helpers
from helpers import FN


def gen(cfg, n, rng):
    if cfg == "cfg2":
        a = 200 / np.sqrt(2)
        return np.full(n, 2.5), rng.uniform(-a, a, n) + 1j * rng.uniform(-a, a, n)
    if cfg == "cfg3":
        return 50 * rng.random(n), 200 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
    if cfg == "cfg4":
        return np.full(n, 2.5), 10 ** (-3 + 7 * rng.random(n)) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
    if cfg == "real":
        return rng.integers(0, 2, n).astype(float), (100 * rng.random(n)).astype(np.complex128)
    if cfg == "far":
        # large moduli in every direction, half of them hugging the axes (small |Re w| or small |Im w| at large |w|):
        # the corner where the over/underflow pre-test bound of prepare() is tightest
        r = 10 ** rng.uniform(1.5, 4.5, n)
        th = np.pi * (2 * rng.random(n) - 1)
        snap = rng.integers(0, 4, n)
        near = np.round(th / (np.pi / 2)) * (np.pi / 2) + 10 ** rng.uniform(-6, -0.5, n) * rng.choice([-1, 1], n)
        th = np.where(snap < 2, near, th)
        return 10 ** rng.uniform(-1, 1.9, n), r * np.exp(1j * th)
    v = 10 ** rng.uniform(-2, 2.3, n)
    return v, 10 ** rng.uniform(-3, 3, n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))


def close_to_drivers(fn, kode, v, z, got, want, ierr, nz, ulps=16):
    """|got - want| <= ulps eps scale, scale = the larger of |f| and the term it can cancel against (the absolute
    bound near zeros used by every parity test)."""
    cname = {"Y": "J", "K": "I", "H1": "J", "H2": "J"}[fn]
    weight = {"Y": 1, "K": np.pi, "H1": 1, "H2": 1}[fn]
    where = {"Y": True, "K": z.real < 0, "H1": z.imag < 0, "H2": z.imag > 0}[fn]
    comp = np.abs(helpers.cpu_bessel(cname, kode, v, z)[0]) * weight
    scale = np.maximum(np.abs(want), np.where(np.isfinite(comp) & where, comp, 0))
    ok = (ierr == 0) & (nz == 0) & np.isfinite(scale) & (scale > 0)
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / scale
    same = got.view(np.uint64).reshape(-1, 2) == want.view(np.uint64).reshape(-1, 2)
    return bool(np.all((err[ok] <= ulps * 2.220446049250313e-16) | same.all(1)[ok]))


@pytest.mark.parametrize("fuse", [False, True], ids=["separate", "fused"])
@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "real", "wide", "far"])
@pytest.mark.parametrize("fn", list(FN))
def test_pipeline_matches_drivers(fn, cfg, fuse):
    """fuse: points needing I and K in the large-|w| region of I go through the fused routine (shared 1/w,
    1/sqrt(w), sincos(Im w), exp(-Re w); K from the series the I routine sums anyway) like in the fused region
    kernel.  The separate flow reproduces the drivers bit for bit, the fused one to 16 eps with identical codes."""
    rng = np.random.default_rng(helpers.seed_of(fn, cfg))
    v, z = gen(cfg, 20000, rng)
    for kode in (1, 2):
        a, ia, na = helpers.cpu_bessel(fn, kode, v, z)
        b, ib, nb, path = helpers.cpu_bessel(fn, kode, v, z, pipeline=True, fuse=fuse)
        if fuse and fn not in ("J", "I"):
            # the fused routine takes K from the large-|w| series (the all-positive sum of the I routine) where
            # the drivers run the Miller algorithm: same codes, values to a few ulp of the cancelling terms
            assert np.array_equal(ia, ib) and np.array_equal(na, nb)
            assert close_to_drivers(fn, kode, v, z, b, a, ia, na)
            continue
        assert np.array_equal(a.view(np.uint64), b.view(np.uint64))
        assert np.array_equal(ia, ib) and np.array_equal(na, nb)
    if cfg == "cfg2":
        assert (path == 0).all(), "every cfg2 point must take the binned path"
    if cfg == "cfg3":
        assert (path == 0).mean() > 0.99, "cfg3 (orders <= 50, |z| <= 200) must stay on the binned path"


def test_golden_points_through_pipeline():
    g = helpers.load_golden("bessel_golden.npz")
    for fn in FN:
        for kode in (1, 2):
            a, ia, na = helpers.cpu_bessel(fn, kode, g["v"], g["z"])
            b, ib, nb, _ = helpers.cpu_bessel(fn, kode, g["v"], g["z"], pipeline=True)
            same = (a.view(np.uint64) == b.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(a.real) & np.isnan(b.real))
            assert same.all() and np.array_equal(ia, ib) and np.array_equal(na, nb)


@pytest.mark.parametrize("fuse", [False, True], ids=["separate", "fused"])
@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "real", "wide"])
def test_jy_pair_matches_separate_drivers(cfg, fuse):
    """The shared-I-pass J/Y pair must give exactly what the J and Y drivers give separately."""
    rng = np.random.default_rng(helpers.seed_of("jy", cfg))
    v, z = gen(cfg, 20000, rng)
    for kode in (1, 2):
        (j, ej, nj), (y, ey, ny) = helpers.cpu_bessel_jy(kode, v, z, fuse=fuse)
        a, ia, na = helpers.cpu_bessel("J", kode, v, z)
        b, ib, nb = helpers.cpu_bessel("Y", kode, v, z)
        assert np.array_equal(j.view(np.uint64), a.view(np.uint64)) and np.array_equal(ej, ia) and np.array_equal(nj, na)
        assert np.array_equal(ey, ib) and np.array_equal(ny, nb)
        if fuse:  # K from the large-|w| series instead of the Miller algorithm (see test_pipeline_matches_drivers)
            assert close_to_drivers("Y", kode, v, z, y, b, ib, nb)
        else:
            assert np.array_equal(y.view(np.uint64), b.view(np.uint64))


@pytest.mark.parametrize("fuse", [False, True], ids=["separate", "fused"])
@pytest.mark.parametrize("fn", list(FN) + ["JY"])
def test_fast_accept_of_the_classifier_agrees_with_prepare(fn, fuse):
    """spb::fast_classify (|z|^2 and a few comparisons, no square root) may only accept a point that prepare()
    classifies identically; it must take the bulk of the large-|z| configs and decline near every boundary."""
    rates = {}
    for cfg in ("cfg2", "cfg3", "cfg4", "real", "wide", "far", "edges"):
        rng = np.random.default_rng(helpers.seed_of("fast", fn, cfg))
        if cfg == "edges":
            # moduli hugging the thresholds the fast accept replaces: rl, 32767, 2 sqrt(fnu + 1), fnu^2 / 2
            n = 40000
            v = rng.choice([0.0, 0.5, 1.0, 2.5, 7.25, 30.0, 85.0, 85.92135864716213, 86.0], n)
            r0 = rng.choice(4, n)
            base = np.where(r0 == 0, 21.784271729432426, np.where(r0 == 1, 32767.0,
                            np.where(r0 == 2, 2 * np.sqrt(v + 1), v * v / 2)))
            r = np.maximum(base, 1e-3) * (1 + rng.choice([-1, 1], n) * 10 ** rng.uniform(-16, -6, n))
            th = rng.choice([0.0, np.pi / 2, np.pi, -np.pi / 2, 1.0, -2.0, 3.0], n) + rng.choice([0, 1e-12, -1e-9], n)
            z = r * np.exp(1j * th)
            z = np.where(rng.random(n) < 0.2, z.real + 0j, z)  # exactly on the real axis as well
        else:
            v, z = gen(cfg, 50000, rng)
        for kode in (1, 2):
            fast, full = helpers.cpu_classify(fn, kode, v, z, fuse)
            acc = fast >= 0
            assert np.array_equal(fast[acc], full[acc]), f"{cfg} kode={kode}: fast accept disagrees with prepare()"
            rates[cfg] = acc.mean()
    assert rates["cfg2"] > 0.95, rates
