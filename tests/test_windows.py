"""Kaiser / Kaiser-Bessel-derived / Lanczos windows against scipy.signal.windows with the
reference's own tolerance: assert_relative_eq!(..., epsilon = 1e-14) in
/root/reference/tests/signal/fir_filter_design_windows.rs:243-281 (an absolute 1e-14, or relative f64::EPSILON)."""
import numpy as np
import pytest

import helpers


def close(got, want):
    d = np.abs(got - want)
    return np.all((d <= 1e-14) | (d <= 2.220446049250313e-16 * np.maximum(np.abs(got), np.abs(want))))


def cases():
    g = helpers.load_golden("window_golden.npz")
    return g, [(i, int(m), float(b), bool(s)) for i, (m, b, s) in enumerate(g["cases"])]


def test_cpu_checker_vs_scipy_windows():
    g, cs = cases()
    for i, m, beta, sym in cs:
        assert close(helpers.cpu_kaiser(m, beta, sym), g[f"kaiser_{i}"]), (m, beta, sym)
        assert close(helpers.cpu_lanczos(m, sym), g[f"lanczos_{i}"]), (m, sym)


@pytest.mark.gpu
def test_gpu_windows_vs_scipy_and_checker():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    from sciport_rs_b200 import windows
    g, cs = cases()
    for i, m, beta, sym in cs:
        k = windows.kaiser(m, beta, sym)
        assert k.shape == (m,)
        assert close(k, g[f"kaiser_{i}"]), (m, beta, sym)
        assert close(k, helpers.cpu_kaiser(m, beta, sym))
        assert close(windows.lanczos(m, sym), g[f"lanczos_{i}"]), (m, sym)
        me = m + (m % 2)
        if me >= 2:
            assert close(windows.kaiser_bessel_derived(me, beta), g[f"kbd_{i}"]), (me, beta)
    with pytest.raises(ValueError):
        windows.kaiser_bessel_derived(7, 5.0)
    # device-resident output, get_window selector (sym = not fftbins), a GPU-sized window with a size-independent
    # property: symmetry w[n] = w[M-1-n] and the unit peak at the centre
    big = windows.get_window(("kaiser", 12.5), 1 << 22, fftbins=False, device="cuda:0")
    assert float((big - torch.flip(big, dims=[0])).abs().max()) <= 1e-15
    assert abs(float(big.max()) - 1.0) <= 1e-12 and float(big.min()) > 0
    per = windows.get_window(("kaiser", 12.5), 4096, fftbins=True)
    assert close(per, windows.kaiser(4097, 12.5, True)[:-1])
    assert close(windows.get_window("lanczos", 513, fftbins=False), windows.lanczos(513, True))


def all_cases():
    g = helpers.load_golden("window_all_golden.npz")
    return g, [(i, int(m), bool(s), float(a), float(sig), float(p)) for i, (m, s, a, sig, p) in enumerate(g["cases"])]


def window_calls(mod, i, m, sym, alpha, sig, p):
    """(golden key, value) for every elementwise window of windows.rs:154-622 at the parameters of case i."""
    simple = ["boxcar", "triang", "blackman", "hamming", "hann", "bartlett", "flattop", "parzen", "bohman",
              "blackmanharris", "nuttall", "barthann", "cosine"]
    for name in simple:
        yield f"{name}_{i}", getattr(mod, name)(m, sym)
    yield f"exponential_{i}", mod.exponential(m, None, None, sym)
    yield f"exponential_tau_{i}", mod.exponential(m, None if sym else 0.3 * m, sig, sym)
    yield f"tukey_{i}", mod.tukey(m, alpha, sym)
    yield f"taylor_{i}", mod.taylor(m, None, None, None, sym)
    yield f"gaussian_{i}", mod.gaussian(m, sig, sym)
    yield f"general_gaussian_{i}", mod.general_gaussian(m, p, sig, sym)
    yield f"general_cosine_{i}", mod.general_cosine(m, [0.3, 0.4, 0.2, 0.07, 0.02, 0.01], sym)
    yield f"general_hamming_{i}", mod.general_hamming(m, alpha, sym)


def test_numpy_oracle_of_the_elementwise_windows_vs_scipy():
    """oracle/windows_oracle.py (the numpy restatement of windows.rs:154-622) pinned against scipy.signal.windows at
    the reference tests' sizes and tolerance."""
    from oracle import windows_oracle as wo
    g, cs = all_cases()
    for case in cs:
        for key, got in window_calls(wo, *case):
            assert got.shape == g[key].shape and close(got, g[key]), key
        i, m, sym = case[:3]
        assert close(wo.taylor(m, 6, 45.0, True, sym), g[f"taylor_n6_{i}"])


@pytest.mark.gpu
def test_gpu_elementwise_windows_vs_scipy_and_oracle():
    """The fused k_window<KIND> kernels through spb_window against the scipy golden vectors (the reference tests'
    oracle and tolerance, fir_filter_design_windows.rs:29-320) and the numpy restatement; get_window selector; KBD with
    its running sum on the device."""
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    import ctypes
    from sciport_rs_b200 import windows, _lib
    from oracle import windows_oracle as wo
    g, cs = all_cases()
    for case in cs:
        for key, got in window_calls(windows, *case):
            assert got.shape == g[key].shape and close(got, g[key]), key
        i, m, sym = case[:3]
        # norm = True through the raw ABI (the mirror ignores it like the reference)
        out = np.empty(m)
        pa = (ctypes.c_double * 3)(6.0, 45.0, 1.0)
        ex = _lib.make_exec()
        _lib.check(_lib.lib.spb_window(windows.KIND["taylor"], m, 1 if sym else 0, pa, 3, out.ctypes.data, ctypes.byref(ex)))
        assert close(out, g[f"taylor_n6_{i}"]), ("taylor norm", m, sym)
    # selector: every type with sym = not fftbins; device-resident output equals the host-staged one
    for w in ("hann", ("tukey", 0.3), ("taylor", 5, 35.0, True), ("gaussian", 7.5), ("general_gaussian", 0.8, 9.0),
              ("exponential", None, 3.0), ("general_cosine", [0.5, 0.3, 0.2]), ("kaiser", 9.0), "bohman", "parzen"):
        a = windows.get_window(w, 1001, fftbins=True)
        b = windows.get_window(w, 1001, fftbins=True, device="cuda:0").cpu().numpy()
        assert np.array_equal(a, b), w
        name = w if isinstance(w, str) else w[0]
        if name in ("hann", "bohman", "parzen"):
            assert close(a, getattr(wo, name)(1001, False))
    with pytest.raises(NotImplementedError):
        windows.get_window(("dpss", 2.5), 64)
    # tukey's degenerate parameters (windows.rs:423-427)
    assert np.array_equal(windows.tukey(100, 0.0), np.ones(100))
    assert close(windows.tukey(100, 1.0), windows.hann(100))
    # a GPU-sized window with size-independent properties: symmetry and the COLA sum of the periodic Hann window
    big = windows.hann(1 << 24, sym=False, device="cuda:0")
    assert float((big[: 1 << 23] + big[1 << 23:] - 1.0).abs().max()) <= 1e-14
    kbd = windows.kaiser_bessel_derived(1 << 20, 11.0, device="cuda:0")
    assert float((kbd[: 1 << 19] ** 2 + kbd[1 << 19:] ** 2 - 1.0).abs().max()) <= 1e-13  # Princen-Bradley condition
