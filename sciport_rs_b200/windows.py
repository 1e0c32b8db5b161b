"""Host-side mirror of the reference's window functions that sit on the special-function path
(/root/reference/src/signal/fir_filter_design/windows.rs): ``kaiser`` (:540-558), ``kaiser_bessel_derived``
(:560-589), ``lanczos`` (:509-538) and the ``get_window`` selector for them (:116-152).  Same names, argument
order and conventions (``sym=True`` symmetric, ``sym=False`` periodic; lengths <= 1 give ones; the derived
window is defined for even lengths only).  All special-function work runs in the CUDA library."""
import ctypes

import numpy as np

from ._lib import lib, check, make_exec

try:
    import torch
except Exception:  # pragma: no cover
    torch = None


def _window(call, m, device):
    m = int(m)
    if m < 0:
        raise ValueError("window length must be non-negative")
    if device is not None:
        dev = torch.device(device)
        out = torch.empty(m, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            ex = make_exec(device_ptrs=True, stream=torch.cuda.current_stream(dev).cuda_stream, devices=[dev.index or 0])
            check(call(out.data_ptr(), ctypes.byref(ex)))
        return out
    out = np.empty(m, dtype=np.float64)
    ex = make_exec()
    check(call(out.ctypes.data, ctypes.byref(ex)))
    return out


def kaiser(m, beta, sym=True, device=None):
    """w[n] = |I0(beta sqrt(1 - ((n - alpha)/alpha)^2)) / I0(beta)|, alpha = (M - 1)/2 (windows.rs:540-558)."""
    return _window(lambda p, ex: lib.spb_kaiser(int(m), float(beta), 1 if sym else 0, p, ex), m, device)


def lanczos(m, sym=True, device=None):
    """sinc(2 n / (M - 1) - 1) (windows.rs:509-538)."""
    return _window(lambda p, ex: lib.spb_lanczos(int(m), 1 if sym else 0, p, ex), m, device)


def kaiser_bessel_derived(m, beta, sym=True, device=None):
    """Square root of the normalised running sum of a Kaiser window of length m/2 + 1, mirrored
    (windows.rs:560-589).  Only even m; ``sym`` is accepted and ignored like in the reference."""
    m = int(m)
    if m < 1:
        return np.empty(0) if device is None else torch.empty(0, dtype=torch.float64, device=device)
    if m % 2 == 1:
        raise ValueError("Kaiser-Bessel Derived windows are only defined for even number of points")
    k = kaiser(m // 2 + 1, beta, True, device=device)
    if device is None:
        c = np.cumsum(k)
        half = np.sqrt(c[:-1] / c[-1])
        return np.concatenate([half, half[::-1]])
    c = torch.cumsum(k, 0)
    half = torch.sqrt(c[:-1] / c[-1])
    return torch.cat([half, torch.flip(half, dims=[0])])


def get_window(window, nx, fftbins=True, device=None):
    """get_window(WindowType, nx, fftbins) (windows.rs:116-152) for the window types on the special-function
    path: ("kaiser", beta), ("kaiser_bessel_derived", beta), "lanczos".  sym = not fftbins."""
    sym = not fftbins
    if isinstance(window, str):
        name, args = window, ()
    else:
        name, args = window[0], tuple(window[1:])
    if name == "kaiser":
        return kaiser(nx, args[0], sym, device=device)
    if name in ("kaiser_bessel_derived", "kbd"):
        return kaiser_bessel_derived(nx, args[0], sym, device=device)
    if name == "lanczos":
        return lanczos(nx, sym, device=device)
    raise NotImplementedError(f"window {name!r} does not use the special-function path (out of scope, SURVEY 8)")
