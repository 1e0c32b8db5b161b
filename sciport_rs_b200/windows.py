"""Host-side mirror of the reference's window module, /root/reference/src/signal/fir_filter_design/windows.rs: every
window function (:154-622) and the ``get_window`` selector (:116-152), same names, argument order and conventions
(``sym=True`` symmetric, ``sym=False`` periodic = first m points of the symmetric window of length m + 1; lengths
<= 1 give ones; the Kaiser-Bessel derived window is defined for even lengths only).

Everything is evaluated by the CUDA library (``spb_window``: one fused kernel per window kind, index -> argument ->
value -> store; Kaiser / Kaiser-Bessel derived / Lanczos sit on the special-function path: I_0 through the complex
routine, sinc).  Nothing here computes a window on the CPU."""
import ctypes

import numpy as np

from ._lib import lib, check, make_exec

try:
    import torch
except Exception:  # pragma: no cover
    torch = None

KIND = {"boxcar": 0, "triang": 1, "general_cosine": 2, "bartlett": 3, "parzen": 4, "bohman": 5, "barthann": 6,
        "cosine": 7, "exponential": 8, "tukey": 9, "taylor": 10, "gaussian": 11, "general_gaussian": 12, "kaiser": 13,
        "lanczos": 14, "kaiser_bessel_derived": 15}


def _window(kind, m, sym, params=(), device=None):
    m = int(m)
    if m < 0:
        raise ValueError("window length must be non-negative")
    pa = (ctypes.c_double * max(len(params), 1))(*[float(x) for x in params])
    if device is not None:
        dev = torch.device(device)
        out = torch.empty(m, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ex = make_exec(device_ptrs=True, stream=torch.cuda.current_stream(dev).cuda_stream, devices=[dev.index or 0])
            check(lib.spb_window(KIND[kind], m, 1 if sym else 0, pa, len(params), out.data_ptr(), ctypes.byref(ex)))
        return out
    out = np.empty(m, dtype=np.float64)
    ex = make_exec()
    check(lib.spb_window(KIND[kind], m, 1 if sym else 0, pa, len(params), out.ctypes.data, ctypes.byref(ex)))
    return out


def boxcar(m, sym=True, device=None):
    """windows.rs:154-165"""
    return _window("boxcar", m, sym, device=device)


def triang(m, sym=True, device=None):
    """windows.rs:167-196"""
    return _window("triang", m, sym, device=device)


def general_cosine(m, a, sym=True, device=None):
    """sum_k a_k cos(k x), x = linspace(-pi, pi, m) (windows.rs:198-218)."""
    a = [float(x) for x in np.ravel(a)]
    if len(a) > 16:
        raise ValueError("general_cosine takes at most 16 coefficients")
    return _window("general_cosine", m, sym, a, device=device)


def blackman(m, sym=True, device=None):
    """windows.rs:220-228"""
    return general_cosine(m, [0.42, 0.50, 0.08], sym, device=device)


def general_hamming(m, alpha, sym=True, device=None):
    """windows.rs:230-236"""
    return general_cosine(m, [alpha, 1.0 - alpha], sym, device=device)


def hamming(m, sym=True, device=None):
    """windows.rs:238-243"""
    return general_hamming(m, 0.54, sym, device=device)


def hann(m, sym=True, device=None):
    """windows.rs:245-250"""
    return general_hamming(m, 0.5, sym, device=device)


def bartlett(m, sym=True, device=None):
    """windows.rs:252-270"""
    return _window("bartlett", m, sym, device=device)


def flattop(m, sym=True, device=None):
    """windows.rs:272-285"""
    return general_cosine(m, [0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368], sym, device=device)


def parzen(m, sym=True, device=None):
    """windows.rs:301-320"""
    return _window("parzen", m, sym, device=device)


def bohman(m, sym=True, device=None):
    """windows.rs:322-340"""
    return _window("bohman", m, sym, device=device)


def blackmanharris(m, sym=True, device=None):
    """windows.rs:342-348"""
    return general_cosine(m, [0.35875, 0.48829, 0.14128, 0.01168], sym, device=device)


def nuttall(m, sym=True, device=None):
    """windows.rs:350-356"""
    return general_cosine(m, [0.3635819, 0.4891775, 0.1365995, 0.0106411], sym, device=device)


def barthann(m, sym=True, device=None):
    """windows.rs:358-373"""
    return _window("barthann", m, sym, device=device)


def cosine(m, sym=True, device=None):
    """windows.rs:375-387"""
    return _window("cosine", m, sym, device=device)


def exponential(m, center=None, tau=None, sym=True, device=None):
    """exp(-|n - center| / tau), center = (M - 1)/2 and tau = 1 by default (windows.rs:389-417)."""
    return _window("exponential", m, sym, [float("nan") if center is None else center, 1.0 if tau is None else tau],
                   device=device)


def tukey(m, alpha=None, sym=True, device=None):
    """Tapered cosine; alpha = 0.5 by default, alpha <= 0 gives ones, alpha >= 1 a Hann window (windows.rs:419-455)."""
    return _window("tukey", m, sym, [0.5 if alpha is None else alpha], device=device)


def taylor(m, nbar=None, sll=None, norm=None, sym=True, device=None):
    """Taylor window, nbar = 4 and sll = 30 by default (windows.rs:457-508).  The reference accepts ``norm`` and
    ignores it (``_norm``, windows.rs:466; its test compares with ``norm=False``): so does this mirror."""
    nbar = 4 if nbar is None else int(nbar)
    return _window("taylor", m, sym, [nbar, 30.0 if sll is None else sll, 0.0], device=device)


def lanczos(m, sym=True, device=None):
    """sinc(2 n / (M - 1) - 1) (windows.rs:509-538)."""
    return _window("lanczos", m, sym, device=device)


def kaiser(m, beta, sym=True, device=None):
    """w[n] = |I0(beta sqrt(1 - ((n - alpha)/alpha)^2)) / I0(beta)|, alpha = (M - 1)/2 (windows.rs:540-558)."""
    return _window("kaiser", m, sym, [beta], device=device)


def kaiser_bessel_derived(m, beta, sym=True, device=None):
    """Square root of the normalised running sum of a Kaiser window of length m/2 + 1, mirrored
    (windows.rs:560-589).  Only even m; ``sym`` is accepted and ignored like in the reference.  Kaiser window and
    running sum (a device scan) both run in the library."""
    m = int(m)
    if m < 1:
        return np.empty(0) if device is None else torch.empty(0, dtype=torch.float64, device=device)
    if m % 2 == 1:
        raise ValueError("Kaiser-Bessel Derived windows are only defined for even number of points")
    return _window("kaiser_bessel_derived", m, True, [beta], device=device)


def gaussian(m, std_dev, sym=True, device=None):
    """windows.rs:591-605"""
    return _window("gaussian", m, sym, [std_dev], device=device)


def general_gaussian(m, power, width, sym=True, device=None):
    """windows.rs:607-622"""
    return _window("general_gaussian", m, sym, [power, width], device=device)


_SIMPLE = {"boxcar": boxcar, "triang": triang, "blackman": blackman, "hamming": hamming, "hann": hann,
           "bartlett": bartlett, "flattop": flattop, "parzen": parzen, "bohman": bohman,
           "blackmanharris": blackmanharris, "nuttall": nuttall, "barthann": barthann, "cosine": cosine,
           "lanczos": lanczos}


def get_window(window, nx, fftbins=True, device=None):
    """get_window(WindowType, nx, fftbins) (windows.rs:116-152): ``window`` is a name or a tuple (name, params...),
    sym = not fftbins.  Dpss and Chebwin are ``todo!()`` in the reference (windows.rs:148-149) and raise here too."""
    sym = not fftbins
    if isinstance(window, str):
        name, args = window, ()
    else:
        name, args = window[0], tuple(window[1:])
    if name in _SIMPLE:
        return _SIMPLE[name](nx, sym, device=device)
    if name == "exponential":
        return exponential(nx, *(args + (None, None))[:2], sym=sym, device=device)
    if name == "tukey":
        return tukey(nx, args[0] if args else None, sym, device=device)
    if name == "taylor":
        a = (args + (None, None, None))[:3]
        return taylor(nx, a[0], a[1], a[2], sym, device=device)
    if name == "kaiser":
        return kaiser(nx, args[0], sym, device=device)
    if name in ("kaiser_bessel_derived", "kbd"):
        return kaiser_bessel_derived(nx, args[0], sym, device=device)
    if name == "gaussian":
        return gaussian(nx, args[0], sym, device=device)
    if name == "general_cosine":
        return general_cosine(nx, args[0], sym, device=device)
    if name == "general_gaussian":
        return general_gaussian(nx, args[0], args[1], sym, device=device)
    if name in ("dpss", "chebwin"):
        raise NotImplementedError(f"window {name!r} is todo!() in the reference (windows.rs:148-149)")
    raise ValueError(f"unknown window type {name!r}")
