"""Bessel-filter prototype design on the batched K kernels: host-side mirror of
/root/reference/src/signal/filter_design/bessel.rs — ``besselap`` (:42-81), ``_bessel_zeros`` (:110-157),
``_aberth`` (:159-201), ``_campos_zeros`` (:204-236), ``_norm_factor`` (:83-99), ``_falling_factorial`` (:101-108).

The reference finds the zeros of the reverse Bessel polynomial of ONE order at a time, calling the scalar
``kve`` four times per root and iteration (f = kve(n + 1/2, 1/x); f' from kve at n - 1/2, n + 1/2, n + 3/2).  Here
every order of a request is iterated at once: one batched call of the K kernel per iteration evaluates the three
orders at every root of every requested filter order (scaled K, KODE 2, in the left half plane: the analytic-
continuation branch of the path).  The host keeps only the O(total roots) bookkeeping of the Aberth sums.

Differences from the reference, both without effect on the converged result: the Aberth sweep updates all roots
of a sweep together (the reference updates them one by one and stops as soon as ONE root moves by less than
1e-16); the Newton polish that follows runs to |dx| <= 4 eps |x| for every root.  The reference's own test bound
for these poles is 1e-6 (tests/signal/common.rs:90-117, tests/signal/bessel.rs:57-71).
"""
import numpy as np


def _gpu_kve(v, z):
    from . import special
    vals, _, _ = special.bessel("K", v, z, scaled=True, want_codes=False)
    idx, code = special.last_first_error(0)
    if idx >= 0:
        raise special.SpecialError(code, idx)
    return vals


def _campos_zeros(n):
    """Starting points (Campos-Calderon approximation), bessel.rs:204-236."""
    if n == 1:
        return np.array([-1 + 0j])
    s = np.polyval([0, 0, 2, 0, -3, 1][::-1], n)
    b3 = np.polyval([16, -8][::-1], n) / s
    b2 = np.polyval([-24, -12, 12][::-1], n) / s
    b1 = np.polyval([8, 24, -12, -2][::-1], n) / s
    b0 = np.polyval([0, -6, 0, 5, -1][::-1], n) / s
    r = np.polyval([0, 0, 2, 1][::-1], n)
    a1 = np.polyval([-6, -6][::-1], n) / r
    a2 = 6 / r
    k = np.arange(1, n + 1, dtype=float)
    x = np.polyval([0, a1, a2][::-1], k)
    y = np.polyval([b0, b1, b2, b3][::-1], k)
    return x + 1j * y


def _falling_factorial(x, n):
    y = 1.0
    for i in range(x - n + 1, x + 1):
        y *= i
    return y


def bessel_zeros(orders, kve=None, max_sweeps=200):
    """Roots of the reverse Bessel polynomials of every order in `orders` (the 1/x of the Bessel-filter poles),
    all orders iterated together.  Returns a list of complex arrays, one per order (bessel.rs:110-157).
    `kve(v_array, z_array)` defaults to the CUDA K kernel; tests inject the CPU checker."""
    kve = kve or _gpu_kve
    orders = [int(o) for o in orders]
    live = [o for o in orders if o > 0]
    if not live:
        return [np.zeros(0, complex) for _ in orders]
    x = np.concatenate([_campos_zeros(o) for o in live])
    order_of = np.concatenate([np.full(o, o, dtype=float) for o in live])
    seg_start = np.cumsum([0] + live[:-1])
    seg = np.concatenate([np.full(o, i) for i, o in enumerate(live)])

    def f_and_fp(x):
        # one batched K call: orders n - 1/2, n + 1/2, n + 3/2 at 1/x for every root of every filter order
        w = 1.0 / x
        v = np.concatenate([order_of - 0.5, order_of + 0.5, order_of + 1.5])
        k = kve(v, np.concatenate([w, w, w]))
        km, k0, kp = np.split(np.asarray(k), 3)
        x2 = x * x
        return k0, km / (2 * x2) - k0 / x2 + kp / (2 * x2)

    eps = np.finfo(float).eps
    for _ in range(max_sweeps):  # Aberth-Ehrlich, bessel.rs:159-201
        f, fp = f_and_fp(x)
        ssum = np.empty_like(x)
        for i, o in enumerate(live):
            xs = x[seg_start[i]:seg_start[i] + o]
            d = xs[:, None] - xs[None, :]
            np.fill_diagonal(d, np.inf)
            ssum[seg_start[i]:seg_start[i] + o] = (1.0 / d).sum(axis=1)
        new = x + f / (f * ssum - fp)
        if not np.all(np.isfinite(new)):
            raise ArithmeticError("failed to converge in aberth method")
        done = np.abs(new - x) <= 1e-13 * np.abs(new)
        x = new
        if done.all():
            break
    else:
        raise ArithmeticError("failed to converge in aberth method")
    for _ in range(50):  # Newton polish of every root (bessel.rs:145-149)
        f, fp = f_and_fp(x)
        step = f / fp
        x = x - step
        if np.all(np.abs(step) <= 4 * eps * np.abs(x)):
            break
    out, pos = {}, 0
    for o in live:
        xs = x[pos:pos + o]
        pos += o
        out[o] = (xs + np.conj(xs[::-1])) / 2  # average complex conjugates (bessel.rs:151-154)
    return [out.get(o, np.zeros(0, complex)) for o in orders]


def _norm_factor(p, k):
    """Frequency at which |k / prod(jw - p)| = 1/sqrt(2) (secant from 1.5, bessel.rs:83-99)."""
    def cutoff(w):
        return abs(k / np.prod(1j * w - p)) - 1 / np.sqrt(2)

    x0, x1 = 1.5, 1.5 * (1 + 1e-4)
    f0, f1 = cutoff(x0), cutoff(x1)
    for _ in range(100):
        if f1 == f0:
            break
        x2 = x1 - f1 * (x1 - x0) / (f1 - f0)
        x0, f0, x1, f1 = x1, f1, x2, cutoff(x2)
        if abs(x1 - x0) <= 1e-15 * abs(x1):
            break
    return x1


def besselap_many(orders, norm="phase", kve=None):
    """(z, p, k) analog Bessel prototypes for every order in `orders`, designed together."""
    if norm not in ("phase", "delay", "mag"):
        raise ValueError("norm must be 'phase', 'delay' or 'mag'")
    zeros = bessel_zeros(orders, kve=kve)
    res = []
    for n, xs in zip(orders, zeros):
        n = int(n)
        if n == 0:
            res.append((np.zeros(0), np.zeros(0, complex), 1.0))
            continue
        a_last = np.floor(_falling_factorial(2 * n, n) / 2.0 ** n)
        p = 1.0 / xs
        k = 1.0
        if norm in ("delay", "mag"):
            k = a_last
            if norm == "mag":
                nf = _norm_factor(p, k)
                p = p / nf
                k = nf ** (-n) * a_last
        else:
            p = p * 10 ** (-np.log10(a_last) / n)
        p = p + 0.0  # normalize_zeros: -0.0 -> +0.0 (tools/complex.rs:4-11)
        res.append((np.zeros(0), p, float(k)))
    return res


def besselap(order, norm="phase", kve=None):
    """signal::besselap(order, norm) (bessel.rs:42-81)."""
    return besselap_many([order], norm=norm, kve=kve)[0]
