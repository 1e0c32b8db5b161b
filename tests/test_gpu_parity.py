"""GPU: parity of the CUDA path (through the C ABI, libsciport_b200.so) with the oracle.

Checkers: (1) the committed golden vectors from scipy.special; (2) the CPU oracle restatement on seeded samples
of every BASELINE config; (3) size-independent properties at full config size.  Tolerance policy as in
tests/test_oracle_golden.py (1e-13 normalised, conditioning-aware beyond 1e+-49, absolute near zeros)."""
import numpy as np
import pytest

import helpers
from helpers import FN, CATS
from oracle import scipy_oracle as so
from test_oracle_golden import scipy_kode2_defect, sem_ok
from test_pipeline_cpu import gen

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def special():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sciport_rs_b200 import special as sp
    return sp


def to_np(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("fn", list(FN))
def test_golden_device_pointers(special, fn, kode):
    g = helpers.load_golden("bessel_golden.npz")
    v, z = g["v"], g["z"]
    want, cats = g[f"{fn}_{kode}"], g[f"{fn}_{kode}_err"]
    zt = torch.from_numpy(z).cuda()
    vt = torch.from_numpy(v).cuda()
    got, ierr, nz = special.bessel(fn, vt, zt, scaled=(kode == 2), semantics="scipy")
    got, ierr, nz = to_np(got), to_np(ierr), to_np(nz)
    err = so.mismatch(got, want, so.scale_near_zeros(fn, kode, v, z, want))
    skip = scipy_kode2_defect(fn, kode, v) | (cats == CATS.index("loss"))
    bad = (err > so.tolerance(want, z, v)) & ~skip
    assert not bad.any(), f"{bad.sum()} mismatches, first v={v[bad][0]} z={z[bad][0]} err={err[bad][0]:.2e}"
    ok = np.array([sem_ok(c, e, n, zz) for c, e, n, zz in zip(cats, ierr, nz, z)]) | skip
    assert ok.all(), f"{(~ok).sum()} ierr/nz mismatches, first v={v[~ok][0]} z={z[~ok][0]}"


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "real", "wide", "far"])
@pytest.mark.parametrize("fn", list(FN))
def test_vs_cpu_oracle_and_monolithic(special, fn, cfg):
    """Binned CUDA path vs LIVE scipy.special (the reference's declared oracle; primary check), vs the CPU port
    (second check: values and error classes) and vs the monolithic CUDA path."""
    rng = np.random.default_rng(helpers.seed_of(fn, cfg, "gpu"))
    v, z = gen(cfg, 200000, rng)
    zt, vt = torch.from_numpy(z).cuda(), torch.from_numpy(v).cuda()
    for kode in (1, 2):
        got, ierr, nz = special.bessel(fn, vt, zt, scaled=(kode == 2))
        mono, ierr_m, nz_m = special.bessel(fn, vt, zt, scaled=(kode == 2), path="monolithic")
        g, m = to_np(got), to_np(mono)
        want, ierr_c, nz_c = helpers.cpu_bessel(fn, kode, v, z)
        # absolute bound near zeros: scale by the terms that cancel (same rule as so.scale_near_zeros)
        cname = {"J": "Y", "Y": "J", "I": "K", "K": "I", "H1": "J", "H2": "J"}[fn]
        weight = {"J": 1, "Y": 1, "I": 1 / np.pi, "K": np.pi, "H1": 1, "H2": 1}[fn]
        where = {"J": True, "Y": True, "I": np.abs(z.real) <= 1, "K": z.real < 0, "H1": z.imag < 0,
                 "H2": z.imag > 0}[fn]
        comp = np.abs(helpers.cpu_bessel(cname, kode, v, z)[0]) * weight
        scale = np.maximum(np.abs(want), np.where(np.isfinite(comp) & where, comp, 0))
        err = so.mismatch(g, want, scale)
        okv = (ierr_c == 0) & (nz_c == 0)
        # the adversarial distributions (outside the BASELINE boxes) carry the algorithm's own conditioning term
        cond = 16.0 if cfg in ("wide", "far") else 0.0
        bad = okv & (err > so.tolerance(want, z, v, conditioning=cond))
        assert not bad.any(), f"kode={kode}: {bad.sum()} mismatches, first v={v[bad][0]} z={z[bad][0]} err={err[bad][0]:.2e}"
        # primary check: live scipy (independent of the product's headers), same scale and tolerance policy;
        # skipped where scipy has no value (nan / inf), is known to be defective (kode 2, order > fnul) or has
        # lost half its digits by its own account (|z| or v > 32767: sf_error "loss")
        ws = so.evaluate(fn, kode, v, z)
        oks = okv & np.isfinite(ws.real) & np.isfinite(ws.imag) & ~scipy_kode2_defect(fn, kode, v)
        oks &= (np.abs(z) < 32767.0) & (v < 32767.0)
        errs = so.mismatch(g, ws, np.maximum(scale, np.abs(ws)))
        bads = oks & (errs > so.tolerance(ws, z, v, conditioning=cond))
        assert not bads.any(), (f"kode={kode}: {bads.sum()} mismatches vs scipy, first v={v[bads][0]} z={z[bads][0]} "
                                f"err={errs[bads][0]:.2e}")
        # the two CUDA paths run the same region math, but nvcc contracts FMAs per kernel, so they agree to
        # rounding (same tolerance), not bit for bit; their error classes must be identical
        errm = so.mismatch(g, m, scale)
        badm = okv & (errm > so.tolerance(want, z, v, conditioning=cond))
        assert not badm.any(), f"kode={kode}: binned vs monolithic {errm[badm].max():.2e}"
        assert torch.equal(ierr, ierr_m) and torch.equal(nz, nz_m)
        # error classes: identical up to points sitting exactly on an over/underflow threshold
        ie = to_np(ierr)
        diff = (ie != ierr_c) | (to_np(nz) != nz_c)
        assert diff.sum() <= 2, f"{diff.sum()} ierr/nz differences vs CPU oracle"


def test_host_pointer_path_matches_device_path(special):
    rng = np.random.default_rng(11)
    v, z = gen("cfg3", 300000, rng)
    for fn in ("J", "K", "Y"):
        a, ia, na = special.bessel(fn, v, z)  # numpy in: staged through pinned copies inside the library
        b, ib, nb = special.bessel(fn, torch.from_numpy(v).cuda(), torch.from_numpy(z).cuda())
        assert np.array_equal(a.view(np.uint64), to_np(b).view(np.uint64))
        assert np.array_equal(ia, to_np(ib)) and np.array_equal(na, to_np(nb))
    # scalar order broadcast (v_stride = 0) and a ragged length.  The broadcast form runs kernel instances that keep
    # the order's phase factors in shared memory, the per-element form computes them per point: same formulas,
    # separately compiled, so agreement is to a few ulp rather than bit for bit
    a, _, _ = special.bessel("I", 0.75, z[:12345])
    b, _, _ = special.bessel("I", np.full(12345, 0.75), z[:12345])
    assert np.all(np.abs(a - b) <= 16 * so.EPS * np.abs(a))


def test_empty_and_tiny_batches(special):
    out, ierr, nz = special.bessel("J", 2.5, np.zeros(0, np.complex128))
    assert out.size == 0
    out, ierr, nz = special.bessel("J", 2.5, np.array([3 + 4j]))
    assert abs(out[0] - (3.967307975128131 + 4.355646112647439j)) < 1e-13 * abs(out[0])


def test_reference_surface_i0_kv_kve_sinc(special):
    """special::i0 / kv / kve / sinc with the reference's conventions (mod.rs:8-13, kv.rs:11-41, trig.rs:4-13)."""
    import scipy.special as sc
    rng = np.random.default_rng(5)
    # tests/special.rs:23-46: |i0(x)| vs scipy.special.i0, x in [0,1), epsilon 1e-14
    for _ in range(20):
        x = rng.random(rng.integers(2, 100))
        got = np.abs(special.i0(x))
        assert np.allclose(got, sc.i0(x), rtol=1e-14, atol=1e-14)
    x = 30 * rng.random(1000)  # Kaiser-window range (fir_filter_design_windows.rs:247)
    assert np.allclose(np.abs(special.i0(x)), sc.i0(x), rtol=1e-14, atol=0)
    # kv.rs doctest: kve(1,1) == kv(1,1) * exp(1) exactly
    assert special.kve(1.0, 1.0) == special.kv(1.0, 1.0) * np.exp(1.0 + 0j)
    # kv.rs:51-58 smoke inputs
    assert special.kv(-3.5, 1000.0) == 0
    z3 = 1.0 / complex(1.2, 0.3)
    assert abs(special.kve(3.5, z3) - sc.kve(3.5, z3)) < 5e-13 * abs(sc.kve(3.5, z3))
    # NaN z is mapped to 0 (kv.rs:12-14) where K is singular -> Err -> unwrap panics (kv.rs:24)
    with pytest.raises(special.SpecialPanic):
        special.kv(2.5, complex(np.nan, 0))
    assert special.kv(np.nan, 1.0) == special.kv(0.0, 1.0)  # NaN v -> 0 (kv.rs:16-18)
    # first Err in index order wins (mod.rs:11-12)
    with pytest.raises(special.SpecialError) as e:
        special.iv(2.5, np.array([1.0, 800.0, 1.0, 900.0], dtype=complex))
    assert e.value.index == 1 and e.value.code == 2
    xs = np.concatenate([[0.0], rng.uniform(-20, 20, 1000)])
    assert np.allclose(special.sinc(xs), np.sinc(xs), rtol=1e-15, atol=1e-17)
    # besselap poles: kve at half-integer orders in the left half plane (bessel.rs:121-142), 1e-6 in the reference
    g = helpers.load_golden("bessel_golden.npz")
    sel = (g["z"].real < 0) & (np.abs(g["v"] % 1 - 0.5) < 1e-12) & (g["v"] < 51)
    got, ierr, _ = special.bessel("K", g["v"][sel], g["z"][sel], scaled=True)
    ref = sc.kve(g["v"][sel], g["z"][sel])
    scale = np.maximum(np.abs(ref), np.pi * np.abs(sc.ive(g["v"][sel], g["z"][sel])))
    assert np.all(np.abs(got - ref) <= 1e-13 * scale)


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("which", ["Ai", "Aip", "Bi", "Bip"])
def test_airy_golden(special, which, kode):
    g = helpers.load_golden("airy_golden.npz")
    z = g["z"]
    want = g[f"{which}_{kode}"]
    got, ierr, nz = special.airy_raw(which, torch.from_numpy(z).cuda(), scaled=(kode == 2), semantics="scipy")
    got, ierr = to_np(got), to_np(ierr)
    w = ("Ai", "Aip", "Bi", "Bip").index(which)
    cpu, ierr_c, _ = helpers.cpu_airy(w, kode, z)
    assert np.array_equal(ierr, ierr_c)
    nanw = np.isnan(want.real)
    assert np.all(np.isnan(got.real[nanw]))
    ok = ~nanw & (ierr == 0)
    zeta = np.abs(2.0 / 3.0 * z * np.sqrt(z))
    tol = 1e-13 + 4 * so.EPS * zeta
    comp = g[f"{('Bi', 'Bip', 'Ai', 'Aip')[w]}_{kode}"]
    scale = np.maximum(np.abs(want), np.where((z.real < 0) & np.isfinite(np.abs(comp)), np.abs(comp), 0))
    err = so.mismatch(got, want, scale)
    assert np.all(err[ok] <= tol[ok]), f"{np.max((err - tol)[ok]):.2e}"


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("which", ["Ai", "Bi"])
def test_airy_config_scale_vs_scipy(special, which, kode):
    """BASELINE config 4 at scale for the Airy companions: 200 000 points of the cfg4 stream (|z| log-uniform
    1e-3..1e4, all quadrants) against LIVE scipy.special.airy / airye, values and error classes.  |z| reaches 1e4,
    ten times beyond the golden fixture: the large-|zeta| kernels, the loss-of-significance band (ierr 3) and the
    argument limit (ierr 4 / scipy NaN) all occur."""
    n = 200_000
    _, z_dev = special.synth(4, n, n_total=1 << 30, device=torch.device("cuda", 0))
    z = to_np(z_dev)
    got, ierr, nz = special.airy_raw(which, z_dev, scaled=(kode == 2), semantics="scipy")
    got, ierr = to_np(got), to_np(ierr)
    w = ("Ai", "Aip", "Bi", "Bip").index(which)
    want = so.evaluate_airy(w, kode, z)
    comp = so.evaluate_airy(2 - w, kode, z)
    # error classes: the CPU port of the drivers is the reference for the codes (scipy folds them into NaN / inf)
    _, ierr_c, _ = helpers.cpu_airy(w, kode, z)
    assert np.array_equal(ierr, ierr_c)
    assert set(np.unique(ierr)) >= {0}, "the stream must contain regular points"
    lost = np.isnan(want.real)
    assert np.all(np.isnan(got.real[lost & (ierr != 0)]))
    # (ierr 3 = more than half of the significant digits lost to the argument reduction: a value is still returned)
    ok = np.isfinite(want.real) & np.isfinite(want.imag) & ((ierr == 0) | (ierr == 3))
    assert ok.mean() > 0.6 and np.any(ierr == 3)
    zeta = (2.0 / 3.0) * np.abs(z) ** 1.5
    tol = so.tolerance(want) + 4 * so.EPS * zeta
    # near the zeros on the negative real axis the envelope |Ai| ~ |Bi| is the scale (north star: absolute bound there)
    scale = np.maximum(np.abs(want), np.where((z.real < 0) & np.isfinite(np.abs(comp)), np.abs(comp), 0))
    if which == "Bi":
        # Bi also vanishes along the rays arg z = +-pi/3 (its complex zeros): Bi(z) = e^{i pi/6} Ai(z w) + e^{-i pi/6}
        # Ai(z conj w), w = e^{2 pi i/3}, and the two terms cancel there; their moduli are the scale
        om = np.exp(2j * np.pi / 3)
        with np.errstate(all="ignore"):
            terms = np.abs(so.evaluate_airy(0, 1, z * om)) + np.abs(so.evaluate_airy(0, 1, z * np.conj(om)))
            if kode == 2:
                terms = terms * np.exp(-np.abs((2.0 / 3.0 * z * np.sqrt(z)).real))
        scale = np.maximum(scale, np.where(np.isfinite(terms), terms, 0))
    err = so.mismatch(got, want, scale)
    assert np.all(err[ok] <= tol[ok]), f"{which} kode {kode}: {np.max((err - tol)[ok]):.2e} over the tolerance"
    # inside |z| <= 100 (|zeta| <= 667) the plain 1e-13 of the north star holds up to the conditioning term
    inner = ok & (np.abs(z) <= 100.0)
    assert np.all(err[inner] <= 1e-13 + 4 * so.EPS * zeta[inner])


def test_gamma_family(special):
    g = helpers.load_golden("gamma_golden.npz")
    x = g["x"]
    got = to_np(special.gamma(torch.from_numpy(x).cuda()))
    want = g["gamma"]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.abs(want)
    err[(got == want) | (np.isnan(got) & np.isnan(want))] = 0
    assert np.all(err <= 1e-13 + np.where(x < 0, 4 * so.EPS * np.pi * np.abs(x), 0.0))
    got = special.gammaln(x)  # host pointers
    want = g["gammaln"]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.maximum(np.abs(want), 1.0)
    err[(got == want) | (np.isnan(got) & np.isnan(want))] = 0
    assert np.all(err[np.abs(x) >= 1e-300] <= 1e-13)
    zc = g["zc"]
    got = to_np(special.loggamma(torch.from_numpy(zc).cuda()))
    want = g["cloggamma"]
    d = got - want
    err = np.abs(d) / np.maximum(np.abs(want), 1.0)
    err[np.isnan(want.real) & np.isnan(got.real)] = 0
    assert np.all(err <= 1e-13)
    got = to_np(special.gamma(torch.from_numpy(zc).cuda()))
    want = g["cgamma"]
    with np.errstate(all="ignore"):
        err = np.abs(got - want) / np.abs(want)
    err[(got == want) | (np.isnan(want.real) & np.isnan(got.real))] = 0
    bound = 1e-13 + 4 * so.EPS * np.abs(g["cloggamma"])
    bound[~np.isfinite(bound)] = 1e-13
    assert np.all(err <= bound)


def test_first_error_and_synth(special):
    ierr = torch.zeros(1 << 20, dtype=torch.int32, device="cuda")
    assert special.first_error(ierr) == (-1, 0)
    ierr[777777] = 3
    ierr[12345] = 2
    assert special.first_error(ierr) == (12345, 2)
    v1, z1 = special.synth(3, 1 << 16, n_total=1 << 20, offset=4096, device="cuda:0")
    v2, z2 = special.synth(3, 1 << 20, n_total=1 << 20, offset=0, device="cuda:0")
    assert torch.equal(z1, z2[4096:4096 + (1 << 16)]) and torch.equal(v1, v2[4096:4096 + (1 << 16)])
    assert float(z2.abs().max()) <= 200.0 and float(v2.max()) <= 50.0
    vh, zh = special.synth(3, 1000, n_total=1 << 20, offset=4096)  # host buffers
    assert np.array_equal(zh, to_np(z1[:1000]))


def test_full_size_cfg2_properties(special):
    """BASELINE config[1] at full size (4096 x 4096): properties that need no oracle.
    (a) conjugate symmetry f(conj z) = conj f(z): the grid is symmetric under im -> -im;
    (b) no element fails; (c) a strided sample agrees with the CPU oracle."""
    side = 4096
    n = side * side
    v, z = special.synth(2, n, n_total=n, device="cuda:0")
    for fn in ("J", "Y"):
        out, ierr, nz = special.bessel(fn, 2.5, z)
        assert special.first_error(ierr) == (-1, 0)
        grid = out.view(side, side)
        mirrored = torch.flip(grid, dims=[0]).conj()
        rel = (grid - mirrored).abs() / grid.abs()
        assert float(rel.max()) <= 1e-13, float(rel.max())
        idx = torch.arange(0, n, 4099, device="cuda:0")
        zs = to_np(z[idx])
        want, _, _ = helpers.cpu_bessel(fn, 1, 2.5, zs)
        comp, _, _ = helpers.cpu_bessel("Y" if fn == "J" else "J", 1, 2.5, zs)
        err = so.mismatch(to_np(out[idx]), want, np.maximum(np.abs(want), np.abs(comp)))
        assert np.all(err <= 1e-13), err.max()


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "real", "wide"])
def test_jy_pair_matches_separate_calls(special, cfg):
    """spb_bessel_jy (shared I pass) vs spb_bessel(J) and spb_bessel(Y): same region math, so agreement to rounding
    (FMA contraction differs per kernel) and identical error classes."""
    rng = np.random.default_rng(helpers.seed_of("jy-gpu", cfg))
    v, z = gen(cfg, 200000, rng)
    zt, vt = torch.from_numpy(z).cuda(), torch.from_numpy(v).cuda()
    for scaled in (False, True):
        (j, ej, nj), (y, ey, ny) = special.bessel_jy(vt, zt, scaled=scaled)
        a, ia, na = special.bessel("J", vt, zt, scaled=scaled)
        b, ib, nb = special.bessel("Y", vt, zt, scaled=scaled)
        assert torch.equal(ej, ia) and torch.equal(nj, na) and torch.equal(ey, ib) and torch.equal(ny, nb)
        ja, ya, aa, ba = to_np(j), to_np(y), to_np(a), to_np(b)
        scale = np.maximum(np.abs(aa), np.abs(ba))
        ok = (to_np(ia) == 0) & (to_np(ib) == 0) & np.isfinite(scale)
        # a point may take different routes in the two calls (the pair is classified like Y: binned in one family and
        # general in the other, or the Debye-direct region in one and the Miller / Wronskian route in the other):
        # both results are within the parity tolerance of the truth, so they agree to that tolerance, not to rounding;
        # outside the BASELINE boxes the difference of route shows the algorithm's own conditioning, eps max(|z|, v)
        cond = 16.0 if cfg == "wide" else 0.0
        tolj = so.tolerance(aa, z, v, conditioning=cond) if cfg in ("wide", "cfg3") else np.full(z.shape, 2e-14)
        assert np.all(so.mismatch(ja, aa, scale)[ok] <= tolj[ok])
        assert np.all(so.mismatch(ya, ba, scale)[ok] <= so.tolerance(ba, z, v, conditioning=cond)[ok])
    # host-pointer path of the pair
    (jh, ejh, _), (yh, eyh, _) = special.bessel_jy(v[:50000], z[:50000])
    (jd, _, _), (yd, _, _) = special.bessel_jy(vt[:50000], zt[:50000])
    assert np.array_equal(jh.view(np.uint64), to_np(jd).view(np.uint64))
    assert np.array_equal(yh.view(np.uint64), to_np(yd).view(np.uint64))


# ---------------------------------------------------------------- order sequences (BASELINE config 5)
def test_j_sweep_0_255_golden_and_oracle(special):
    """out[k][i] = J_k(z_i), k = 0..255: golden (scipy per order), CPU oracle on a cfg5 sample, host-pointer path."""
    from test_sequence_cpu import check_seq
    g = helpers.load_golden("seq_golden.npz")
    z = g["z5"]
    for kode in (1, 2):
        got, ierr, nz = special.bessel("J", 0.0, torch.from_numpy(z).cuda(), scaled=(kode == 2), n_orders=256)
        assert got.shape == (256, z.size)
        check_seq(to_np(got), to_np(ierr), to_np(nz), g[f"J5_{kode}"], g[f"Y5_{kode}"], np.zeros(z.size),
                  f"J sweep kode={kode}", z=z)
    v, zs = special.synth(5, 20000, n_total=1 << 24)
    dev, ierr, nz = special.bessel("J", 0.0, torch.from_numpy(zs).cuda(), n_orders=256)
    dev = to_np(dev)
    want, ierr_c, nz_c = helpers.cpu_bessel_seq("J", 1, 0.0, zs, 256)
    assert np.array_equal(to_np(ierr), ierr_c) and np.array_equal(to_np(nz), nz_c)
    # same algorithm on both sides, different FMA contraction: agreement to a few roundings of the recurrence
    ycomp = helpers.cpu_bessel_seq("Y", 1, 0.0, zs, 256)[0]
    check_seq(dev, ierr_c, nz_c, want, ycomp, np.zeros(zs.size), "J sweep vs CPU oracle", z=zs)
    # zeros (underflowed leading members) sit in the same slots
    assert np.array_equal(dev == 0, want == 0)
    host, ierr_h, nz_h = special.bessel("J", 0.0, zs, n_orders=256)  # numpy in: staged, strided copy back
    assert np.array_equal(host.view(np.uint64), dev.view(np.uint64))
    assert np.array_equal(ierr_h, ierr_c) and np.array_equal(nz_h, nz_c)
    swept = special.jn_sequence(256, zs[:100])
    assert np.array_equal(swept.view(np.uint64), dev[:, :100].copy().view(np.uint64))


@pytest.mark.parametrize("fn", ["J", "Y", "I", "K", "H1", "H2"])
def test_short_sequences_golden(special, fn):
    from test_sequence_cpu import check_seq
    g = helpers.load_golden("seq_golden.npz")
    z, v = g["zs"], g["vs"]
    for kode in (1, 2):
        got, ierr, nz = special.bessel(fn, torch.from_numpy(v).cuda(), torch.from_numpy(z).cuda(),
                                       scaled=(kode == 2), n_orders=12)
        comp = g[f"{so.COMPANION[fn]}s_{kode}"] if fn in ("J", "Y") else None
        check_seq(to_np(got), to_np(ierr), to_np(nz), g[f"{fn}s_{kode}"], comp, v, f"{fn} kode={kode}", z=z)
        _, ierr_c, nz_c = helpers.cpu_bessel_seq(fn, kode, v, z, 12)
        assert np.array_equal(to_np(ierr), ierr_c) and np.array_equal(to_np(nz), nz_c)


def test_sweep_recurrence_property_large(special):
    """Size-independent property on 2^20 cfg5 points x 256 orders (4 GiB of results): the three-term
    recurrence J_{k-1}(z) + J_{k+1}(z) = (2k/z) J_k(z) holds for every interior member."""
    n = 1 << 20
    v, z = special.synth(5, n, n_total=1 << 24, device="cuda:0")
    out, ierr, nz = special.bessel("J", 0.0, z, n_orders=256)
    assert special.first_error(ierr) == (-1, 0)
    worst = 0.0
    for k0 in range(1, 255, 32):  # slabs of orders to bound temporary memory
        k1 = min(k0 + 32, 255)
        k = torch.arange(k0, k1, device="cuda:0", dtype=torch.float64)[:, None]
        mid = out[k0:k1]
        res = out[k0 - 1:k1 - 1] + out[k0 + 1:k1 + 1] - (2.0 * k / z[None, :]) * mid
        scale = out[k0 - 1:k1 - 1].abs() + out[k0 + 1:k1 + 1].abs() + ((2.0 * k / z[None, :]) * mid).abs()
        ok = scale > 1e-290
        worst = max(worst, float((res.abs() / scale.clamp_min(1e-300))[ok].max()))
    assert worst <= 1e-14, worst


def test_first_error_tracked_on_device(special):
    """`collect::<Result<_, i32>>` (mod.rs:11-12) without per-element code arrays: the kernels keep the smallest
    failing (index, ierr) of the call in one device word (spb_last_first_error)."""
    n = (1 << 22) + 12345  # more than one host chunk
    z = np.full(n, 1.0 + 0.5j)
    z[n - 7] = 900.0        # I overflows (ierr 2)
    z[4200000] = 800.0      # ... and here, earlier, in the second staged chunk
    out, ierr, nz = special.bessel("I", 2.5, z, want_codes=False)
    assert ierr is None and nz is None
    assert special.last_first_error() == (4200000, 2)
    zt = torch.from_numpy(z).cuda()
    special.bessel("I", 2.5, zt, want_codes=False)
    assert special.last_first_error() == (4200000, 2)
    full, ierr_full, _ = special.bessel("I", 2.5, zt)  # with codes: same answer from the array
    assert special.first_error(ierr_full) == (4200000, 2)
    assert special.last_first_error() == (4200000, 2)
    z[4200000] = 1.0
    z[n - 7] = 1.0
    special.bessel("I", 2.5, z, want_codes=False)
    assert special.last_first_error() == (-1, 0)
    # pair family: separate words for J and Y (Y is singular at the origin, J is not)
    zz = np.array([1 + 1j, 0j, 2 + 0j, 0j])
    special.bessel_jy(2.5, zz, want_codes=False)
    assert special.last_first_error(0) == (-1, 0)
    assert special.last_first_error(1)[0] == 1
    with pytest.raises(special.SpecialError) as e:
        special.jvyv(2.5, zz)
    assert e.value.index == 1
    # Airy: no result beyond |z| ~ 1.05e6
    za = np.array([1.0, 2.0, 2e6, 3.0], dtype=complex)
    special.airy_raw("Ai", za)
    assert special.last_first_error() == (2, 4)
    # order sweeps: ierr is per point
    special.bessel("K", 0.5, np.array([1.0, 0.0, 2.0], dtype=complex), n_orders=20, want_codes=False)
    assert special.last_first_error() == (1, 1)


def test_multi_device_slices_match_single_device(special):
    """spb_exec.n_devices = G: contiguous slices, one host thread + two streams per device, no collective.
    Results (and the first-error index, which is global) must not depend on G."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(21)
    v, z = gen("cfg3", 3_000_017, rng)
    z[2_900_000] = 0.0  # K is singular at the origin: ierr 1 in the second slice
    one, e1, n1 = special.bessel("K", v, z, scaled=True)
    idx1 = special.last_first_error()
    g = torch.cuda.device_count()
    many, eg, ng = special.bessel("K", v, z, scaled=True, devices=list(range(g)))
    assert special.last_first_error() == idx1 == (2_900_000, 1)
    assert np.array_equal(one.view(np.uint64), many.view(np.uint64))
    assert np.array_equal(e1, eg) and np.array_equal(n1, ng)
    sw1, _, _ = special.bessel("J", 0.0, z[:200_001], n_orders=32, want_codes=False)
    swg, _, _ = special.bessel("J", 0.0, z[:200_001], n_orders=32, want_codes=False, devices=list(range(g)))
    assert np.array_equal(sw1.view(np.uint64), swg.view(np.uint64))
    # the J/Y pair, Airy and the elementwise companions take the same split
    devs = list(range(g))
    (j1, _, _), (y1, _, _) = special.bessel_jy(2.5, z)
    (jg, _, _), (yg, _, _) = special.bessel_jy(2.5, z, devices=devs)
    assert np.array_equal(j1.view(np.uint64), jg.view(np.uint64)) and np.array_equal(y1.view(np.uint64), yg.view(np.uint64))
    zz = z * 0.05  # |z| <= 10: no overflow of the unscaled functions anywhere
    zz[2_900_001] = -1.0e7  # beyond the argument range of the Airy routines: ierr 4 in the second slice
    for which in ("Ai", "Bi"):
        a1, ea1, na1 = special.airy_raw(which, zz)
        i1 = special.last_first_error()
        ag, eag, nag = special.airy_raw(which, zz, devices=devs)
        assert special.last_first_error() == i1 and i1[0] == 2_900_001
        same = (a1.view(np.uint64) == ag.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(a1.real) & np.isnan(ag.real))
        assert same.all() and np.array_equal(ea1, eag) and np.array_equal(na1, nag)
    x = np.abs(z.real) + 0.25
    for f in (special.gamma, special.gammaln, special.sinc):
        assert np.array_equal(f(x).view(np.uint64), f(x, devices=devs).view(np.uint64))
    for f in (special.gamma, special.loggamma):
        a, b = f(z * 0.3), f(z * 0.3, devices=devs)
        same = (a.view(np.uint64) == b.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(a.real) & np.isnan(b.real))
        assert same.all()
    assert "bdf=" in special.device_topology(g - 1)


def test_negative_orders_inside_the_c_abi(special):
    """The reference calls kv(-3.5, 1000) and unwraps (kv.rs:53-54): spb_bessel itself must evaluate negative orders
    (reflection on the device), for host and device pointers alike; SPB_RAW_NEG_ORDERS restores ierr = 1."""
    import ctypes
    from sciport_rs_b200 import _lib
    v = np.array([-3.5], dtype=np.float64)
    z = np.array([1000.0 + 0j])
    out = np.empty(1, np.complex128)
    ierr = np.full(1, -7, np.int32)
    nz = np.full(1, -7, np.int32)
    ex = _lib.make_exec()
    _lib.check(_lib.lib.spb_bessel(_lib.FN["K"], 1, v.ctypes.data, 0, z.ctypes.data, out.ctypes.data, ierr.ctypes.data,
                                   nz.ctypes.data, 1, 1, ctypes.byref(ex)))
    assert out[0] == 0 and ierr[0] == 0 and nz[0] == 1  # K_3.5(1000) = 0 by underflow, not an Err
    ex = _lib.make_exec(flags=_lib.RAW_NEG_ORDERS)
    _lib.check(_lib.lib.spb_bessel(_lib.FN["K"], 1, v.ctypes.data, 0, z.ctypes.data, out.ctypes.data, ierr.ctypes.data,
                                   nz.ctypes.data, 1, 1, ctypes.byref(ex)))
    assert ierr[0] == 1
    # device inputs take the same route (large batch: binned K path with the order folded at load)
    rng = np.random.default_rng(3)
    n = 100000
    vv = -12 * rng.random(n)
    zz = 30 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1)) + 0.05
    for fn in ("K", "J", "I", "H1"):
        a, ea, _ = special.bessel(fn, vv, zz)
        b, eb, _ = special.bessel(fn, torch.from_numpy(vv).cuda(), torch.from_numpy(zz).cuda())
        assert np.array_equal(ea, to_np(eb)) and (ea == 0).all()
        assert np.all(np.abs(a - to_np(b)) <= 1e-13 * np.maximum(np.abs(a), 1e-300))
    (j, ej, _), (y, ey, _) = special.bessel_jy(vv, zz)
    a, _, _ = special.bessel("J", vv, zz)
    b, _, _ = special.bessel("Y", vv, zz)
    assert np.array_equal(j.view(np.uint64), a.view(np.uint64)) and np.array_equal(y.view(np.uint64), b.view(np.uint64))


def test_negative_orders_by_reflection(special):
    """Orders < 0: reflection formulae evaluated inside the library, checked against scipy."""
    import scipy.special as sc
    rng = np.random.default_rng(9)
    n = 4000
    v = np.where(rng.random(n) < 0.3, -rng.integers(0, 12, n).astype(float), -12 * rng.random(n))
    v[::7] = np.abs(v[::7])  # mixed signs in one call
    z = 30 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1)) + 0.05
    pairs = {"jv": (special.jv, sc.jv, sc.yv), "yv": (special.yv, sc.yv, sc.jv), "iv": (special.iv, sc.iv, sc.kv),
             "kv": (special.kv_array, sc.kv, None), "h1": (special.hankel1, sc.hankel1, None),
             "h2": (special.hankel2, sc.hankel2, None), "ive": (special.ive, sc.ive, sc.kve),
             "jve": (special.jve, sc.jve, sc.yve)}
    for name, (ours, ref, comp) in pairs.items():
        got = ours(v, z)
        want = ref(v, z)
        scale = np.abs(want)
        if comp is not None:
            with np.errstate(all="ignore"):
                cw = np.abs(comp(np.abs(v), z)) * (np.exp(-z.real - np.abs(z.real)) if name == "ive" else 1.0)
            scale = np.maximum(scale, np.where(np.isfinite(cw), cw, 0))
            scale = np.maximum(scale, np.abs(ref(np.abs(v), z)))
        ok = np.isfinite(want.real) & (scale > 1e-280) & (scale < 1e280)
        err = np.abs(got - want)[ok] / scale[ok]
        assert err.max() <= 2e-13, (name, err.max())


def test_full_size_cfg3_properties(special):
    """BASELINE config 3 at its full size (2^28 = 268 435 456 points, per-element order in [0, 50], |z| <= 200, scaled
    I and K; 17 chunks of 2^24): size-independent properties.
    (a) no element fails; (b) f(conj z) = conj f(z) on a slab; (c) a strided sample agrees with the CPU oracle;
    (d) the Wronskian I_v K_{v+1} + I_{v+1} K_v = 1/z on a slab (ties the I and K kernels together)."""
    n = 1 << 28
    v, z = special.synth(3, n, n_total=n, device="cuda:0")
    idx = torch.arange(0, n, 262147, device="cuda:0")
    zs, vs = to_np(z[idx]), to_np(v[idx])
    for fn in ("I", "K"):
        out, _, _ = special.bessel(fn, v, z, scaled=True, want_codes=False)
        first, code = special.last_first_error()
        if first >= 0:
            # among 2^28 points a few sit at the exponent limits for real (|z| ~ 1e-5 at order 49: K overflows);
            # the first failing element must fail in the CPU oracle with the same code
            zf, vf = to_np(z[first:first + 1]), to_np(v[first:first + 1])
            _, ierr_c, _ = helpers.cpu_bessel(fn, 2, vf, zf)
            assert ierr_c[0] == code and abs(zf[0]) < 1e-3, (first, code, zf, vf)
        want, _, _ = helpers.cpu_bessel(fn, 2, vs, zs)
        comp = np.abs(helpers.cpu_bessel("K" if fn == "I" else "I", 2, vs, zs)[0])
        comp = comp * (np.exp(-2 * np.abs(zs.real)) / np.pi if fn == "I" else np.pi * np.exp(2 * np.abs(zs.real)))
        where = (np.abs(zs.real) <= 1) if fn == "I" else (zs.real < 0)
        scale = np.maximum(np.abs(want), np.where(np.isfinite(comp) & where, comp, 0))
        err = so.mismatch(to_np(out[idx]), want, scale)
        assert np.all(err <= so.tolerance(want, zs, vs)), err.max()
        slab = slice(5 << 24, (5 << 24) + (1 << 22))
        mirrored, _, _ = special.bessel(fn, v[slab], z[slab].conj(), scaled=True, want_codes=False)
        rel = (mirrored.conj() - out[slab]).abs() / out[slab].abs()
        rel = torch.where(torch.isfinite(rel), rel, torch.zeros_like(rel))  # 0/0 where both underflowed
        assert float(rel.max()) <= 1e-13, float(rel.max())
        del out, mirrored
    slab = slice(9 << 24, (9 << 24) + (1 << 22))
    zz, vv = z[slab], v[slab]
    zz = torch.where(zz.real < 0, -zz, zz)  # right half plane: plain Wronskian
    i0_, _, _ = special.bessel("I", vv, zz, scaled=True, want_codes=False)
    i1_, _, _ = special.bessel("I", vv + 1, zz, scaled=True, want_codes=False)
    k0_, _, _ = special.bessel("K", vv, zz, scaled=True, want_codes=False)
    k1_, _, _ = special.bessel("K", vv + 1, zz, scaled=True, want_codes=False)
    # scaled: I e^{-Re z}, K e^{z}  ->  product carries e^{i Im z}
    w = (i0_ * k1_ + i1_ * k0_) * zz * torch.exp(-1j * zz.imag)
    terms = (i0_ * k1_).abs() + (i1_ * k0_).abs()
    resid = (w - 1).abs() / (terms * zz.abs()).clamp_min(1.0)
    assert float(resid.max()) <= 2e-13, float(resid.max())


def test_full_size_cfg5_per_gpu_share(special):
    """BASELINE config 5, the share of one of 8 GPUs: 2^21 points x 256 orders = 8 GiB of results.  Every interior
    member satisfies the three-term recurrence and the top/bottom rows agree with single-order calls."""
    n = 1 << 21
    v, z = special.synth(5, n, n_total=1 << 24, offset=3 << 21, device="cuda:0")
    out, _, _ = special.bessel("J", 0.0, z, n_orders=256, want_codes=False)
    assert special.last_first_error() == (-1, 0)
    single0, _, _ = special.bessel("J", 0.0, z, want_codes=False)
    y0, _, _ = special.bessel("Y", 0.0, z, want_codes=False)
    scale0 = torch.maximum(single0.abs(), y0.abs())
    assert float(((out[0] - single0).abs() / scale0).max()) <= 4.5e-13
    worst = 0.0
    for k0 in range(1, 255, 16):
        k1 = min(k0 + 16, 255)
        k = torch.arange(k0, k1, device="cuda:0", dtype=torch.float64)[:, None]
        mid = (2.0 * k / z[None, :]) * out[k0:k1]
        res = out[k0 - 1:k1 - 1] + out[k0 + 1:k1 + 1] - mid
        scale = out[k0 - 1:k1 - 1].abs() + out[k0 + 1:k1 + 1].abs() + mid.abs()
        ok = scale > 1e-280
        worst = max(worst, float((res.abs() / scale.clamp_min(1e-300))[ok].max()))
    assert worst <= 1e-14, worst


def test_calls_on_different_streams_share_scratch_safely(special):
    """Device-pointer calls only enqueue; the per-device scratch is shared, so a call on another stream must be
    ordered behind the previous one by the library itself."""
    rng = np.random.default_rng(33)
    v, z = gen("cfg3", 1 << 20, rng)
    zt, vt = torch.from_numpy(z).cuda(), torch.from_numpy(v).cuda()
    ref_k, _, _ = special.bessel("K", vt, zt, scaled=True)
    ref_j, _, _ = special.bessel("J", 2.5, zt)
    torch.cuda.synchronize()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for rep in range(4):
        with torch.cuda.stream(s1):
            a, _, _ = special.bessel("K", vt, zt, scaled=True)
        with torch.cuda.stream(s2):
            b, _, _ = special.bessel("J", 2.5, zt)
        outs.append((a, b))
    torch.cuda.synchronize()
    for a, b in outs:
        assert torch.equal(a.view(torch.float64), ref_k.view(torch.float64))
        assert torch.equal(b.view(torch.float64), ref_j.view(torch.float64))
    host, _, _ = special.bessel("K", v, z, scaled=True)  # a host-buffer call right behind queued device work
    assert np.array_equal(host.view(np.uint64), to_np(ref_k).view(np.uint64))


@pytest.mark.parametrize("order", [0.0, 0.3, 2.5, 7.25, 30.0])
@pytest.mark.parametrize("fn", ["Y", "K", "H1", "H2"])
def test_fused_large_w_kernel_scalar_order(special, fn, order):
    """Every K-type point in the large-|w| region of I goes through the fused I + K kernel (k_fused.cu: K from the
    series the I routine sums), in its scalar-order and its per-element-order instantiation.  Both must match
    the CPU oracle (whose drivers run the Miller algorithm / closed form for K there), agree with each other to
    rounding and report identical codes."""
    for cfg in ("cfg2", "cfg4", "far"):
        rng = np.random.default_rng(helpers.seed_of("fused", fn, cfg, str(order)))
        _, z = gen(cfg, 100000, rng)
        v = np.full(z.shape, order)
        zt = torch.from_numpy(z).cuda()
        for kode in (1, 2):
            got, ierr, nz = special.bessel(fn, order, zt, scaled=(kode == 2))  # scalar-order instantiation
            sep, ierr_s, nz_s = special.bessel(fn, torch.from_numpy(v).cuda(), zt, scaled=(kode == 2))
            assert torch.equal(ierr, ierr_s) and torch.equal(nz, nz_s)
            g, s = to_np(got), to_np(sep)
            want, ierr_c, nz_c = helpers.cpu_bessel(fn, kode, v, z)
            cname = {"Y": "J", "K": "I", "H1": "J", "H2": "J"}[fn]
            weight = {"Y": 1, "K": np.pi, "H1": 1, "H2": 1}[fn]
            where = {"Y": True, "K": z.real < 0, "H1": z.imag < 0, "H2": z.imag > 0}[fn]
            comp = np.abs(helpers.cpu_bessel(cname, kode, v, z)[0]) * weight
            scale = np.maximum(np.abs(want), np.where(np.isfinite(comp) & where, comp, 0))
            okv = (ierr_c == 0) & (nz_c == 0)
            # the north star's 1e-13 on the BASELINE orders; higher scalar orders and the adversarial distribution
            # carry the algorithm's own conditioning term (see so.tolerance: two correct implementations differ by
            # a few eps max(|z|, v); measured here 1.6e-13 at v = 30, z = -71.3+0.39i, both 1e-13 from mpmath)
            tol = so.tolerance(want, z, v, conditioning=16.0 if (cfg == "far" or order > 2.5) else 0.0)
            bad = okv & (so.mismatch(g, want, scale) > tol)
            assert not bad.any(), f"{cfg} kode={kode}: {bad.sum()} mismatches, first z={z[bad][0]}"
            assert not (okv & (so.mismatch(g, s, scale) > tol)).any()
            diff = (to_np(ierr) != ierr_c) | (to_np(nz) != nz_c)
            assert diff.sum() <= 2


def test_compaction_of_homogeneous_and_ragged_tiles(special):
    """Tiles whose 2048 points share one class are compacted by the straight index write-out of k_scatter, mixed
    tiles by the staged scan: a batch made of long constant runs, one mixed stretch and a ragged tail must come
    back in the original order either way."""
    rng = np.random.default_rng(5)
    runs = [np.full(5000, 150.0 + 20.0j), np.full(4096, 0.5 - 0.25j), np.full(3000, -40.0 + 3.0j)]
    _, mixed = gen("cfg4", 6000, rng)
    z = np.concatenate(runs[:2] + [mixed] + runs[2:] + [np.full(777, 9.0 + 9.0j)])
    zt = torch.from_numpy(z).cuda()
    for fn in ("J", "Y", "K"):
        got, ierr, nz = special.bessel(fn, 2.5, zt)
        want, ierr_c, nz_c = helpers.cpu_bessel(fn, 1, 2.5, z)
        comp = np.abs(helpers.cpu_bessel({"J": "Y", "Y": "J", "K": "I"}[fn], 1, 2.5, z)[0]) * (np.pi if fn == "K" else 1)
        scale = np.maximum(np.abs(want), np.where(np.isfinite(comp), comp, 0))
        okv = (ierr_c == 0) & (nz_c == 0)
        assert np.array_equal(to_np(ierr), ierr_c) and np.array_equal(to_np(nz), nz_c)
        assert np.all(so.mismatch(to_np(got), want, scale)[okv] <= so.tolerance(want, z, np.full(z.shape, 2.5))[okv])
    (j, _, _), (y, _, _) = special.bessel_jy(2.5, zt)
    a, _, _ = special.bessel("J", 2.5, zt)
    assert np.all(np.abs(to_np(j) - to_np(a)) <= 1e-13 * np.maximum(np.abs(to_np(a)), np.abs(to_np(y))))


def test_device_transcendentals(special):
    """exp / exp(-x) (the two members of the device exp pair), sin and cos as the region kernels evaluate them,
    against extended-precision references: <= 2 ulp over the ranges the kernels produce, exact limits outside."""
    rng = np.random.default_rng(77)
    x = np.concatenate([rng.uniform(-708, 708, 200000), rng.uniform(-2, 2, 100000), 10 ** rng.uniform(-300, -1, 1000),
                        np.array([0.0, -0.0, 709.7, -745.0, 800.0, -800.0, 1e308, -1e308, np.inf, -np.inf, np.nan])])
    xl = x.astype(np.longdouble)
    xt = torch.from_numpy(x).cuda()
    for which, ref in (("exp", np.exp(xl)), ("expm", np.exp(-xl))):
        got = to_np(special.devmath(which, xt))
        want = ref.astype(np.float64)
        fin = np.isfinite(want) & (want > 1e-300)
        ulp = np.abs((got[fin].astype(np.longdouble) - ref[fin]) / np.spacing(want[fin]).astype(np.longdouble))
        assert ulp.max() <= 2.0, (which, float(ulp.max()))
        assert np.array_equal(np.isnan(got), np.isnan(x))
        sign = 1.0 if which == "exp" else -1.0
        assert np.all(got[sign * x >= 710] == np.inf) and np.all(got[sign * x <= -746] == 0.0)
        # subnormal results: a few units of the last (subnormal) place
        sub = np.isfinite(want) & (want <= 1e-300) & (np.abs(x) < 746)
        assert np.all(np.abs(got[sub] - want[sub]) <= 4 * np.spacing(np.maximum(want[sub], 5e-324)))
    y = np.concatenate([rng.uniform(-40000, 40000, 200000), rng.uniform(-3, 3, 100000), rng.uniform(-1e15, 1e15, 2000)])
    yl = y.astype(np.longdouble)
    yt = torch.from_numpy(y).cuda()
    small = np.abs(y) < 1048576
    for which, ref in (("sin", np.sin(yl)), ("cos", np.cos(yl))):
        got = to_np(special.devmath(which, yt))
        # absolute error in units of the last place of 1 (the phase factors multiply values of modulus ~ 1); the
        # longdouble reference reduces huge arguments with a 64-bit pi, so only the kernels' range is compared tightly
        err = np.abs(got.astype(np.longdouble) - ref)
        assert float(err[small].max()) <= 2 * 2.220446049250313e-16, (which, float(err[small].max()))
        assert np.all(np.abs(got) <= 1.0)
    # log: <= 1.5 ulp on positive normal arguments (tight around 1, where the result is small), library semantics
    # for zero, negative, subnormal, infinite and NaN arguments
    w = np.concatenate([10 ** rng.uniform(-307, 308, 200000), 1 + rng.uniform(-0.5, 0.5, 100000),
                        1 + 10 ** rng.uniform(-16, -1, 20000) * rng.choice([-1, 1], 20000),
                        np.array([1.0, 2.0, 0.5, np.sqrt(2.0), 1.7976931348623157e308, 2.2250738585072014e-308])])
    got = to_np(special.devmath("log", torch.from_numpy(w).cuda()))
    ref = np.log(w.astype(np.longdouble))
    want = ref.astype(np.float64)
    ulp = np.abs((got.astype(np.longdouble) - ref) / np.spacing(np.abs(want) + 5e-324).astype(np.longdouble))
    assert got[0 + 320000] == 0.0 and float(ulp.max()) <= 1.5, float(ulp.max())
    odd = np.array([0.0, -0.0, -1.0, 5e-324, 1e-310, np.inf, np.nan])
    got = to_np(special.devmath("log", torch.from_numpy(odd).cuda()))
    with np.errstate(all="ignore"):
        want = np.log(odd)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(want)], want[~np.isnan(want)])


def test_host_generator_matches_device_stream(special):
    """bench.py's numpy restatement of the SURVEY 8d generator (used by the CPU legs) addresses the same stream as
    spb_synth on the device."""
    import bench
    n_total = 1 << 28
    off = 123_456_789
    v, z = special.synth(3, 4096, n_total=n_total, offset=off, device="cuda:0")
    vh, zh = bench.cfg3_points(np.arange(off, off + 4096, dtype=np.int64))
    assert np.array_equal(to_np(v), vh)
    assert np.all(np.abs(to_np(z) - zh) <= 4 * so.EPS * np.abs(zh))


def test_scan_with_several_rounds_per_bin():
    """k_scan scans 8192 tile counters per round: a chunk of 2^25 points (16384 tiles, SPB_CHUNK_LOG2 = 25) makes
    every bin take two rounds, with the carry handed from round to round; an adversarial input puts points of every
    region into both halves.  Runs in a fresh process (the chunk size is read once) and must reproduce the default
    chunking bit for bit."""
    import subprocess, sys, os
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
from sciport_rs_b200 import special
n = (1 << 25) + 12345
v, z = special.synth(3, n, n_total=1 << 28, device="cuda:0")
# second half: another distribution (log-uniform moduli), so the bin populations of the two rounds differ
_, z4 = special.synth(4, n // 2, n_total=1 << 30, device="cuda:0")
z[n // 2: n // 2 + z4.numel()] = z4
out = {}
for fn in ("I", "K"):
    o, e, k = special.bessel(fn, v, z, scaled=True)
    out[fn] = (o.cpu().numpy().view(np.uint64), e.cpu().numpy(), k.cpu().numpy())
np.savez(sys.argv[1], **{f"{fn}_{i}": a for fn, t in out.items() for i, a in enumerate(t)})
""" % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        res = {}
        for tag, log2 in (("a", "24"), ("b", "25")):
            env = dict(os.environ, SPB_CHUNK_LOG2=log2)
            r = subprocess.run([sys.executable, "-c", code, os.path.join(d, tag + ".npz")], env=env, capture_output=True,
                               text=True, timeout=900)
            assert r.returncode == 0, r.stderr[-2000:]
            res[tag] = np.load(os.path.join(d, tag + ".npz"))
        for k in res["a"].files:
            assert np.array_equal(res["a"][k], res["b"][k]), k


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "wide"])
def test_ik_pair_matches_single_calls_and_scipy(special, cfg):
    """spb_bessel_ik (I and K in one pass) against spb_bessel(I) / spb_bessel(K) and live scipy: identical error
    classes, values to the parity tolerance (a point may take another region routine in the pair)."""
    rng = np.random.default_rng(helpers.seed_of("ik-gpu", cfg))
    v, z = gen(cfg, 200000, rng)
    zt, vt = torch.from_numpy(z).cuda(), torch.from_numpy(v).cuda()
    cond = 16.0 if cfg == "wide" else 0.0
    for scaled in (False, True):
        kode = 2 if scaled else 1
        (i_, ei, ni), (k_, ek, nk) = special.bessel_ik(vt, zt, scaled=scaled)
        a, ia, na = special.bessel("I", vt, zt, scaled=scaled)
        b, ib, nb = special.bessel("K", vt, zt, scaled=scaled)
        assert torch.equal(ei, ia) and torch.equal(ni, na) and torch.equal(ek, ib) and torch.equal(nk, nb)
        for fn, got, single, ie, nzz in (("I", i_, a, ia, na), ("K", k_, b, ib, nb)):
            g, sg = to_np(got), to_np(single)
            want = so.evaluate(fn, kode, v, z)
            scale = so.scale_near_zeros(fn, kode, v, z, want)
            ok = (to_np(ie) == 0) & (to_np(nzz) == 0) & np.isfinite(want.real) & np.isfinite(want.imag)
            ok &= ~scipy_kode2_defect(fn, kode, v) & (np.abs(z) < 32767.0) & (v < 32767.0)
            tol = so.tolerance(want, z, v, conditioning=cond)
            assert np.all(so.mismatch(g, want, scale)[ok] <= tol[ok]), (fn, scaled)
            assert np.all(so.mismatch(g, sg, np.maximum(scale, np.abs(sg)))[ok] <= tol[ok]), (fn, scaled)
    # host buffers (staged chunks) give the same bits as device pointers
    (i_h, _, _), (k_h, _, _) = special.bessel_ik(v[:50000], z[:50000], scaled=True)
    (i_d, _, _), (k_d, _, _) = special.bessel_ik(vt[:50000], zt[:50000], scaled=True)
    for h, d in ((i_h, to_np(i_d)), (k_h, to_np(k_d))):
        same = (h.view(np.uint64) == d.view(np.uint64)).reshape(-1, 2).all(1) | (np.isnan(h.real) & np.isnan(d.real))
        assert same.all()
    if cfg == "cfg2":
        i2, k2 = special.ive_kve(v[:50000], z[:50000])  # raises like Result<_, i32>; none fails on this grid
        assert np.array_equal(i2.view(np.uint64), i_h.view(np.uint64)) and np.array_equal(k2.view(np.uint64), k_h.view(np.uint64))


@pytest.mark.parametrize("kode", [1, 2])
@pytest.mark.parametrize("fn", ["J", "Y", "I", "K", "H1", "H2"])
def test_direct_kernel_is_bit_identical_to_the_classified_path(special, fn, kode):
    """Broadcast-order jobs run k_direct ahead of the classifier (k_direct.cu): groups of 32 consecutive points that
    all pass the fast accept are evaluated in input order.  It runs the region kernel's own routine, so values AND
    codes must be bit-identical to path='classified' (SPB_PATH_NO_DIRECT: every point through the classifier) —
    on a grid with a ragged tail (most groups accepted, the disc around the origin declined), on long constant runs
    mixed with unrelated points (tiles with finished and unfinished groups side by side), and on a batch of
    unrelated points (every warp gives up after its first groups)."""
    rng = np.random.default_rng(helpers.seed_of("direct", fn, str(kode)))
    side = 700
    x = np.linspace(-100.0, 100.0, side)
    grid = (x[None, :] + 1j * x[:, None]).reshape(-1)[: side * side - 13]
    _, mixed = gen("cfg4", 70000, rng)
    runs = np.concatenate([np.full(5000, 150.0 + 20.0j), mixed[:3000], np.full(4111, -40.0 + 3.0j), mixed[3000:6000],
                           np.full(33, 55.0 - 60.0j), np.array([np.nan + 0j, 0j, np.inf + 0j]), np.full(2050, 0.3 + 99.0j)])
    # whole groups that pass the fast accept but that the fused routine hands back (unscaled K / H beyond the
    # exponent range of exp(-Re w), orders <= 2 have no pre-test): k_direct must leave them to the pipeline
    far = np.concatenate([np.full(4096, -800.0 + 5.0j), np.full(4096, 900.0 + 1.0j), np.full(2048, 3.0 + 900.0j),
                          np.full(2048, 3.0 - 900.0j), np.full(2048, -760.0 - 740.0j), np.full(1000, 30000.0 + 1.0j)])
    for label, z in (("grid", grid), ("runs", runs), ("unrelated", mixed), ("far", far)):
        zt = torch.from_numpy(np.ascontiguousarray(z)).cuda()
        for order in (0.0, 2.5, 17.0):
            a, ea, na = special.bessel(fn, order, zt, scaled=(kode == 2))
            b, eb, nb = special.bessel(fn, order, zt, scaled=(kode == 2), path="classified")
            assert torch.equal(ea, eb) and torch.equal(na, nb), (label, order)
            assert np.array_equal(to_np(a).view(np.uint64), to_np(b).view(np.uint64)), (label, order)
    # real arguments promoted to complex (spb_bessel_real) take the same kernel
    xr = torch.from_numpy(np.linspace(0.0, 3000.0, 300001)).cuda()
    a, ea, _ = special.bessel(fn, 1.0, xr, scaled=(kode == 2), real_input=True)
    b, eb, _ = special.bessel(fn, 1.0, xr, scaled=(kode == 2), real_input=True, path="classified")
    assert torch.equal(ea, eb) and np.array_equal(to_np(a).view(np.uint64), to_np(b).view(np.uint64))


def test_direct_kernel_pairs_profile_and_scipy(special):
    """The pair entry points through k_direct: bit-identical to the classified path, most of a gridded batch taken
    by the direct kernel (profile record), values against live scipy."""
    import scipy.special as sc
    side = 1024
    x = np.linspace(-120.0, 120.0, side)
    z = (x[None, :] + 1j * x[:, None]).reshape(-1)
    zt = torch.from_numpy(z).cuda()
    for entry, kw in ((special.bessel_jy, {}), (special.bessel_ik, {"scaled": True})):
        (a, ea, _), (b, eb, _) = entry(2.5, zt, **kw)
        (c, ec, _), (d, ed, _) = entry(2.5, zt, path="classified", **kw)
        assert torch.equal(ea, ec) and torch.equal(eb, ed)
        assert np.array_equal(to_np(a).view(np.uint64), to_np(c).view(np.uint64))
        assert np.array_equal(to_np(b).view(np.uint64), to_np(d).view(np.uint64))
    prof = special.profile_step(lambda: special.bessel_jy(2.5, zt, profile=True))
    direct = sum(r["points"] for r in prof if r["name"].endswith("direct_asyi"))
    assert direct > 0.9 * z.size, (direct, z.size)
    (j, _, _), (y, _, _) = special.bessel_jy(2.5, zt)
    idx = np.arange(0, z.size, 97)
    wj, wy = sc.jv(2.5, z[idx]), sc.yv(2.5, z[idx])
    scale = np.maximum(np.abs(wj), np.abs(wy))
    assert np.all(np.abs(to_np(j)[idx] - wj) <= 1e-13 * scale) and np.all(np.abs(to_np(y)[idx] - wy) <= 1e-13 * scale)


def test_scalar_negative_order_of_k_on_the_binned_path(special):
    """K_{-v} = K_v with a BROADCAST negative order: the order is folded where it is loaded, also for the per-CTA
    order phase of the scalar-order kernels (left half plane: the continuation uses it)."""
    rng = np.random.default_rng(11)
    _, z = gen("cfg2", 150000, rng)
    zt = torch.from_numpy(z).cuda()
    for kode in (1, 2):
        a, ea, _ = special.bessel("K", -3.5, zt, scaled=(kode == 2))
        b, eb, _ = special.bessel("K", 3.5, zt, scaled=(kode == 2))
        assert torch.equal(ea, eb) and np.array_equal(to_np(a).view(np.uint64), to_np(b).view(np.uint64))


def test_large_orders_and_loss_points_vs_mpmath_truth_gpu(special):
    """The points the scipy comparison has to skip (kode-2 orders above fnul, scipy's "loss" category) against the
    mpmath fixture tests/golden/large_order_truth2.npz, through the C ABI with device pointers."""
    from test_oracle_golden import large_order_truth_errors

    def evaluate(fn, kode, v, z):
        got, ierr, nz = special.bessel(fn, torch.from_numpy(v).cuda(), torch.from_numpy(z).cuda(), scaled=(kode == 2))
        return to_np(got), to_np(ierr), to_np(nz)

    for label, err, tol in large_order_truth_errors(evaluate):
        assert np.all(err <= tol), f"{label}: max {err.max():.2e} ({int((err > tol).sum())} of {err.size} above tolerance)"


@pytest.mark.parametrize("fn", ["J", "Y", "I", "K", "H1", "H2"])
def test_broadcast_order_kernels_agree_with_per_element_orders(special, fn):
    """The kernels of a broadcast-order job take the order-only values of ln Gamma (power series, Miller / Neumann
    normalisation, Temme series of K) from a per-CTA cache (OrderPhase::gl_*, order_phase_scalar).  A scalar order and
    the same order passed per element must report the same codes and agree to rounding (the two are different
    instantiations of the same routines, so the compiler's FMA contraction may differ in the last bits) — on a batch
    that lives in the small- and mid-|z| regions where those routines run."""
    rng = np.random.default_rng(helpers.seed_of("glcache", fn))
    _, z = gen("cfg4", 40000, rng)
    zt = torch.from_numpy(z).cuda()
    for order in (0.0, 0.3, 2.5, 7.25, 30.0):
        v = np.full(z.size, order)
        vt = torch.from_numpy(v).cuda()
        for kode in (1, 2):
            a, ea, na = special.bessel(fn, order, zt, scaled=(kode == 2))
            b, eb, nb = special.bessel(fn, vt, zt, scaled=(kode == 2))
            assert torch.equal(ea, eb) and torch.equal(na, nb), (order, kode)
            a, b = to_np(a), to_np(b)
            ok = (to_np(ea) == 0) & (to_np(na) == 0) & np.isfinite(b.real) & np.isfinite(b.imag)
            scale = so.scale_near_zeros(fn, kode, v, z, b)
            err = so.mismatch(a, b, scale)[ok]
            # (a few roundings of an O(v) exponent / recurrence: measured 75 eps at v = 30)
            assert err.max() <= max(64.0, 6.0 * order) * so.EPS, (order, kode, float(err.max()))
