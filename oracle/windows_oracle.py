"""oracle/windows_oracle.py — TEST INFRASTRUCTURE ONLY.

numpy restatement of the elementwise window functions of the reference,
/root/reference/src/signal/fir_filter_design/windows.rs (each function cites the lines it follows), used by tests/ to
check the fused CUDA kernel k_window<KIND> (sciport_rs_b200/csrc/k_window.cu).  Pinned against scipy.signal.windows —
the oracle of the reference's own tests (/root/reference/tests/signal/fir_filter_design_windows.rs:29-320) — through
the committed golden vectors tests/golden/window_all_golden.npz (tests/test_windows.py).

Conventions shared by all of them (windows.rs:624-656): lengths <= 1 give ones; sym = False ("periodic") evaluates
the symmetric window of length m + 1 and drops its last point."""
import numpy as np


def _guard(m):
    return np.ones(max(int(m), 0))


def _extend(m, sym):
    return (m, False) if sym else (m + 1, True)


def _trunc(w, needed):
    return w[:-1] if needed else w


def boxcar(m, sym=True):  # windows.rs:154-165
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    return _trunc(np.ones(me), t)


def triang(m, sym=True):  # windows.rs:167-196
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(1, (me + 1) // 2 + 1, dtype=np.float64)
    if me % 2 == 0:
        w = (2 * n - 1.0) / me
        w = np.r_[w, w[::-1]]
    else:
        w = 2 * n / (me + 1.0)
        w = np.r_[w, w[-2::-1]]
    return _trunc(w, t)


def general_cosine(m, a, sym=True):  # windows.rs:198-218
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    fac = np.linspace(-np.pi, np.pi, me)
    w = np.zeros(me)
    for k, ak in enumerate(a):
        w = w + ak * np.cos(k * fac)
    return _trunc(w, t)


def blackman(m, sym=True):  # windows.rs:220-228
    return general_cosine(m, [0.42, 0.50, 0.08], sym)


def general_hamming(m, alpha, sym=True):  # windows.rs:230-236
    return general_cosine(m, [alpha, 1.0 - alpha], sym)


def hamming(m, sym=True):  # windows.rs:238-243
    return general_hamming(m, 0.54, sym)


def hann(m, sym=True):  # windows.rs:245-250
    return general_hamming(m, 0.5, sym)


def bartlett(m, sym=True):  # windows.rs:252-270
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(me, dtype=np.float64)
    w = np.where(n <= (me - 1) / 2.0, 2.0 * n / (me - 1), 2.0 - 2.0 * n / (me - 1))
    return _trunc(w, t)


def flattop(m, sym=True):  # windows.rs:272-285
    return general_cosine(m, [0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368], sym)


def parzen(m, sym=True):  # windows.rs:301-320
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(-(me - 1) / 2.0, (me - 1) / 2.0 + 0.5, 1.0)
    na = n[n < -(me - 1) / 4.0]
    nb = n[np.abs(n) <= (me - 1) / 4.0]
    wa = 2 * (1 - np.abs(na) / (me / 2.0)) ** 3.0
    wb = 1 - 6 * (np.abs(nb) / (me / 2.0)) ** 2.0 + 6 * (np.abs(nb) / (me / 2.0)) ** 3.0
    return _trunc(np.r_[wa, wb, wa[::-1]], t)


def bohman(m, sym=True):  # windows.rs:322-340
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    fac = np.abs(np.linspace(-1, 1, me)[1:-1])
    w = (1 - fac) * np.cos(np.pi * fac) + 1.0 / np.pi * np.sin(np.pi * fac)
    return _trunc(np.r_[0, w, 0], t)


def blackmanharris(m, sym=True):  # windows.rs:342-348
    return general_cosine(m, [0.35875, 0.48829, 0.14128, 0.01168], sym)


def nuttall(m, sym=True):  # windows.rs:350-356
    return general_cosine(m, [0.3635819, 0.4891775, 0.1365995, 0.0106411], sym)


def barthann(m, sym=True):  # windows.rs:358-373
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(me, dtype=np.float64)
    fac = np.abs(n / (me - 1.0) - 0.5)
    return _trunc(0.62 - 0.48 * fac + 0.38 * np.cos(2 * np.pi * fac), t)


def cosine(m, sym=True):  # windows.rs:375-387
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    return _trunc(np.sin(np.pi / me * (np.arange(me) + 0.5)), t)


def exponential(m, center=None, tau=None, sym=True):  # windows.rs:389-417
    tau = 1.0 if tau is None else tau
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    if center is None:
        center = (me - 1) / 2.0
    n = np.arange(me, dtype=np.float64)
    return _trunc(np.exp(-np.abs(n - center) / tau), t)


def tukey(m, alpha=None, sym=True):  # windows.rs:419-455
    alpha = 0.5 if alpha is None else alpha
    if m <= 1:
        return _guard(m)
    if alpha <= 0:
        return np.ones(m)
    if alpha >= 1.0:
        return hann(m, sym)
    me, t = _extend(m, sym)
    n = np.arange(me, dtype=np.float64)
    width = int(np.floor(alpha * (me - 1) / 2.0))
    n1, n2, n3 = n[: width + 1], n[width + 1: me - width - 1], n[me - width - 1:]
    w1 = 0.5 * (1 + np.cos(np.pi * (-1 + 2.0 * n1 / alpha / (me - 1))))
    w3 = 0.5 * (1 + np.cos(np.pi * (-2.0 / alpha + 1 + 2.0 * n3 / alpha / (me - 1))))
    return _trunc(np.r_[w1, np.ones(n2.shape), w3], t)


def taylor(m, nbar=None, sll=None, norm=None, sym=True):  # windows.rs:457-508 (the reference ignores `norm`: :466)
    nbar = 4 if nbar is None else int(nbar)
    sll = 30.0 if sll is None else sll
    norm = bool(norm)
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    b = 10 ** (sll / 20.0)
    a = np.arccosh(b) / np.pi
    s2 = nbar ** 2 / (a ** 2 + (nbar - 0.5) ** 2)
    ma = np.arange(1, nbar, dtype=np.float64)
    fm = np.zeros(nbar - 1)
    signs = np.where(np.arange(nbar - 1) % 2 == 0, 1.0, -1.0)
    m2 = ma * ma
    for mi in range(len(ma)):
        numer = signs[mi] * np.prod(1 - m2[mi] / s2 / (a ** 2 + (ma - 0.5) ** 2))
        denom = 2 * np.prod(1 - m2[mi] / m2[:mi]) * np.prod(1 - m2[mi] / m2[mi + 1:])
        fm[mi] = numer / denom

    def wf(n):
        return 1 + 2 * fm @ np.cos(2 * np.pi * ma[:, None] * (n - me / 2.0 + 0.5) / me)

    w = wf(np.arange(me, dtype=np.float64))
    if norm:
        w = w / wf(np.array([(me - 1) / 2.0]))
    return _trunc(w, t)


def gaussian(m, std, sym=True):  # windows.rs:591-605
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(me, dtype=np.float64) - (me - 1.0) / 2.0
    return _trunc(np.exp(-(n ** 2) / (2 * std * std)), t)


def general_gaussian(m, p, sig, sym=True):  # windows.rs:607-622
    if m <= 1:
        return _guard(m)
    me, t = _extend(m, sym)
    n = np.arange(me, dtype=np.float64) - (me - 1.0) / 2.0
    return _trunc(np.exp(-0.5 * np.abs(n / sig) ** (2 * p)), t)
