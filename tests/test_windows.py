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
