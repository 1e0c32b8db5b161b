"""Bessel-filter prototypes through the batched K evaluations against scipy.signal.besselap, with
the reference's own bound: poles and gain to 1e-6 relative / absolute
(/root/reference/tests/signal/common.rs:90-117, tests/signal/bessel.rs:57-71, orders up to 49)."""
import numpy as np
import pytest

import helpers
from sciport_rs_b200 import filter_design as fd


def cpu_kve(v, z):
    return helpers.cpu_bessel("K", 2, v, z)[0]


def check(res, g, norm, orders, tol=1e-6):
    for (z, p, k), n in zip(res, orders):
        want_p, want_k = g[f"p_{norm}_{n}"], float(g[f"k_{norm}_{n}"])
        assert z.size == 0 and p.shape == want_p.shape
        # pole order may differ: match as sets sorted by (imag, real)
        a = p[np.lexsort((p.real, p.imag))]
        b = want_p[np.lexsort((want_p.real, want_p.imag))]
        assert np.all(np.abs(a - b) <= tol * np.maximum(1.0, np.abs(b))), (norm, n, np.abs(a - b).max())
        assert abs(k - want_k) <= tol * max(1.0, abs(want_k)), (norm, n, k, want_k)
    return True


@pytest.mark.parametrize("norm", ["phase", "delay", "mag"])
def test_host_logic_with_cpu_checker(norm):
    """All 50 orders designed together; K values from the CPU checker (no GPU needed)."""
    g = helpers.load_golden("besselap_golden.npz")
    orders = list(range(0, 50))
    check(fd.besselap_many(orders, norm=norm, kve=cpu_kve), g, norm, orders)


def test_roots_are_accurate_far_beyond_the_reference_bound():
    g = helpers.load_golden("besselap_golden.npz")
    orders = [1, 2, 5, 12, 25, 37, 49]
    res = fd.besselap_many(orders, norm="delay", kve=cpu_kve)
    check(res, g, "delay", orders, tol=1e-11)


@pytest.mark.gpu
def test_gpu_besselap_all_orders():
    torch = pytest.importorskip("torch")
    assert torch.cuda.is_available()
    g = helpers.load_golden("besselap_golden.npz")
    orders = list(range(0, 50))
    for norm in ("phase", "delay", "mag"):
        check(fd.besselap_many(orders, norm=norm), g, norm, orders)
    z, p, k = fd.besselap(4, "phase")
    assert p.size == 4 and np.all(p.real < 0)
