"""Host-side mirror of sciport-rs's `special` module, batched.

Reference surface (same names, argument order and error behaviour):

* ``i0(x)``      /root/reference/src/special/mod.rs:8-13   real array -> complex array, first ``Err(i32)`` in
                 index order wins (``collect::<Result<_, i32>>``) -> raises :class:`SpecialError`
* ``kv(v, z)``   /root/reference/src/special/kv.rs:11-25   NaN z -> 0, NaN v -> 0, ``unwrap()`` panics on Err
                 -> raises :class:`SpecialPanic`
* ``kve(v, z)``  /root/reference/src/special/kv.rs:39-41   ``kv(v, z) * exp(z)`` (post-multiplication)
* ``sinc(x)``    /root/reference/src/special/trig.rs:4-13

Widened, batched surface named by BASELINE.json (order first, then z; ``xv`` / ``xve`` pairs like kv / kve):
``jv jve yv yve iv ive kv_array kve_array hankel1 hankel1e hankel2 hankel2e airy airye gamma gammaln loggamma``.

Inputs may be numpy arrays (host buffers: the library stages them through pinned async copies) or torch CUDA
tensors (device pointers: zero copies, launched on the current torch stream).  Everything is computed by the
CUDA library; nothing here evaluates a special function on the CPU.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import (lib, check, make_exec, FN, AIRY, SEM_AMOS, SEM_SCIPY, RAW_NEG_ORDERS, PATH_MONOLITHIC, PROFILE,
                   PATH_NO_DIRECT)

try:  # torch is plumbing only (device memory / streams)
    import torch
except Exception:  # pragma: no cover
    torch = None


class SpecialError(Exception):
    """``Err(code)`` of the reference's ``Result<_, i32>`` (first failing element, index order)."""

    def __init__(self, code, index):
        super().__init__(f"special function failed with ierr={code} at index {index}")
        self.code = code
        self.index = index


class SpecialPanic(RuntimeError):
    """The reference ``unwrap()``s the scalar kv result (kv.rs:24) and panics on Err."""


def _is_torch(a):
    return torch is not None and isinstance(a, torch.Tensor)


def _stream_for(t):
    return torch.cuda.current_stream(t.device).cuda_stream


_PROFILE_LOG = None


def profile_step(fn):
    """Run `fn` (which calls bessel(..., profile=True)) and return the per-kernel records, with the family
    prefixed: [{"name": "J/ipass_asyi", "ms": .., "points": ..}, ...]."""
    global _PROFILE_LOG
    _PROFILE_LOG = []
    try:
        fn()
        return _PROFILE_LOG
    finally:
        _PROFILE_LOG = None


def bessel(fn, v, z, scaled=False, semantics="amos", path="binned", real_input=False, devices=None, profile=False,
           out=None, ierr=None, nz=None, n_orders=1, want_codes=True, raw_neg_orders=False):
    """Evaluate one family over a batch.  Returns ``(values, ierr, nz)``.

    ``want_codes=False`` passes NULL for the per-element ``ierr`` / ``nz`` arrays (they come back as ``None``):
    the first failing element is still available through :func:`last_first_error`, which is all the
    reference's ``Result<Array, i32>`` needs, and 8 bytes per element of device-to-host traffic are saved.

    ``n_orders > 1`` evaluates the order sequence v, v+1, ..., v+n_orders-1 per point; ``values`` then has
    shape ``(n_orders, n)`` (order-major, BASELINE config 5), ``ierr`` / ``nz`` stay per point.

    fn: 'J','Y','I','K','H1','H2'; v: scalar or array broadcast against z; z: complex128 (or float64 with
    ``real_input=True``, promoted like ``v.into()`` in mod.rs:10).  Host (numpy) calls may pass preallocated
    ``out`` / ``ierr`` / ``nz`` arrays (pinned memory avoids the per-call page registration).
    """
    kode = 2 if scaled else 1
    flags = (SEM_SCIPY if semantics == "scipy" else SEM_AMOS) | (PATH_MONOLITHIC if path == "monolithic" else 0)
    if path == "classified":  # binned, but every point through the classifier (no direct large-|z| kernel)
        flags |= PATH_NO_DIRECT
    if profile:
        flags |= PROFILE
    if raw_neg_orders:
        flags |= RAW_NEG_ORDERS
    if _is_torch(z):
        if not z.is_cuda:
            raise ValueError("torch inputs must be CUDA tensors (use numpy arrays for host buffers)")
        if ierr is not None or nz is not None:
            raise ValueError("preallocated ierr / nz buffers are only taken for host (numpy) inputs")
        if devices is not None:
            raise ValueError("device tensors live on one device: `devices` is for host (numpy) inputs")
        zc = z.resolve_conj().contiguous()  # a lazy conjugate view shares storage with the unconjugated tensor
        want = torch.float64 if real_input else torch.complex128
        if zc.dtype != want:
            zc = zc.to(want)
        n = zc.numel()
        if _is_torch(v):
            vt = v.to(device=zc.device, dtype=torch.float64).contiguous()
            if vt.numel() == 1:
                stride = 0
            else:
                if vt.numel() != n:
                    raise ValueError("v and z must have the same number of elements")
                stride = 1
        else:
            vt = torch.full((1,), float(v), dtype=torch.float64, device=zc.device)
            stride = 0
        oshape = tuple(zc.shape) if n_orders == 1 else (n_orders, n)
        if out is None:
            out = torch.empty(oshape, dtype=torch.complex128, device=zc.device)
        assert out.is_contiguous() and out.numel() == n * n_orders and out.dtype == torch.complex128
        ierr = torch.empty(zc.shape, dtype=torch.int32, device=zc.device) if want_codes else None
        nz = torch.empty(zc.shape, dtype=torch.int32, device=zc.device) if want_codes else None
        with torch.cuda.device(zc.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(zc), flags=flags, devices=[zc.device.index])
            f = lib.spb_bessel_real if real_input else lib.spb_bessel
            check(f(FN[fn], kode, vt.data_ptr(), stride, zc.data_ptr(), out.data_ptr(),
                    ierr.data_ptr() if want_codes else None, nz.data_ptr() if want_codes else None, n, n_orders,
                    ctypes.byref(ex)))
        if profile and _PROFILE_LOG is not None:
            for r in _lib.last_profile():
                r["name"] = f"{fn}/{r['name']}"
                _PROFILE_LOG.append(r)
        return out, ierr, nz
    za = np.ascontiguousarray(z, dtype=np.float64 if real_input else np.complex128)
    n = za.size
    va = np.asarray(v, dtype=np.float64)
    if va.size == 1:
        va = np.ascontiguousarray(va.reshape(1))
        stride = 0
    else:
        va = np.ascontiguousarray(np.broadcast_to(va, za.shape))
        stride = 1
    # caller-provided (e.g. pinned) result buffers are filled in place, like the C ABI
    oshape = za.shape if n_orders == 1 else (n_orders, n)
    out = np.empty(oshape, dtype=np.complex128) if out is None else out
    assert out.dtype == np.complex128 and out.size == n * n_orders and out.flags.c_contiguous
    if want_codes:
        ierr = np.empty(za.shape, dtype=np.int32) if ierr is None else ierr
        nz = np.empty(za.shape, dtype=np.int32) if nz is None else nz
        assert ierr.dtype == np.int32 and nz.dtype == np.int32 and ierr.size == n and nz.size == n
    else:
        ierr = nz = None
    ex = make_exec(device_ptrs=False, flags=flags, devices=devices)
    f = lib.spb_bessel_real if real_input else lib.spb_bessel
    check(f(FN[fn], kode, va.ctypes.data, stride, za.ctypes.data, out.ctypes.data,
            ierr.ctypes.data if want_codes else None, nz.ctypes.data if want_codes else None, n, n_orders,
            ctypes.byref(ex)))
    return out, ierr, nz


def bessel_ik(v, z, scaled=False, semantics="amos", profile=False, outs=None, codes=None, want_codes=True,
              devices=None, path="binned"):
    """I_v(z) and K_v(z) in one pass (spb_bessel_ik): returns ((I, ierr_i, nz_i), (K, ierr_k, nz_k)); scaled=True
    gives the ive / kve pair.  K in the left half plane is continued with I at the same right-half-plane point and
    the large-|z| / Debye-direct regions produce both functions from one series, so the pair costs far less than
    the two single calls."""
    return bessel_jy(v, z, scaled, semantics, profile, outs, codes, want_codes, devices, _entry="ik", path=path)


def ive_kve(v, z, outs=None, devices=None):
    """(ive(v, z), kve(v, z)) computed together (BASELINE config 3); raises on the first Err of either."""
    (i, _, _), (k, _, _) = bessel_ik(v, z, scaled=True, outs=outs, want_codes=False, devices=devices)
    for which in (0, 1):
        idx, code = last_first_error(which)
        if idx >= 0:
            raise SpecialError(code, idx)
    return i, k


def iv_kv(v, z, outs=None, devices=None):
    """(iv(v, z), kv(v, z)) computed together; raises on the first Err of either."""
    (i, _, _), (k, _, _) = bessel_ik(v, z, scaled=False, outs=outs, want_codes=False, devices=devices)
    for which in (0, 1):
        idx, code = last_first_error(which)
        if idx >= 0:
            raise SpecialError(code, idx)
    return i, k


def bessel_jy(v, z, scaled=False, semantics="amos", profile=False, outs=None, codes=None, want_codes=True,
              devices=None, _entry="jy", path="binned"):
    """J_v(z) and Y_v(z) in one pass (spb_bessel_jy): returns ((J, ierr_j, nz_j), (Y, ierr_y, nz_y)).
    Both functions need I_v at the same rotated point, so the pair shares classification and the I pass."""
    entry = lib.spb_bessel_jy if _entry == "jy" else lib.spb_bessel_ik
    tag = "JY" if _entry == "jy" else "IK"
    kode = 2 if scaled else 1
    flags = (SEM_SCIPY if semantics == "scipy" else SEM_AMOS) | (PROFILE if profile else 0)
    if path == "classified":
        flags |= PATH_NO_DIRECT
    if _is_torch(z):
        zc = z.resolve_conj().contiguous().to(torch.complex128)
        n = zc.numel()
        if _is_torch(v):
            vt = v.to(device=zc.device, dtype=torch.float64).contiguous()
            if vt.numel() != 1 and vt.numel() != n:
                raise ValueError("v and z must have the same number of elements")
            stride = 0 if vt.numel() == 1 else 1
        else:
            vt = torch.full((1,), float(v), dtype=torch.float64, device=zc.device)
            stride = 0
        # caller-provided device buffers are filled in place (the C ABI never allocates results)
        if outs is None:
            outs = [torch.empty(zc.shape, dtype=torch.complex128, device=zc.device) for _ in range(2)]
        if not want_codes:
            codes = [None] * 4
        elif codes is None:
            codes = [torch.empty(zc.shape, dtype=torch.int32, device=zc.device) for _ in range(4)]
        cp = [c.data_ptr() if c is not None else None for c in codes]
        with torch.cuda.device(zc.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(zc), flags=flags, devices=[zc.device.index])
            check(entry(kode, vt.data_ptr(), stride, zc.data_ptr(), outs[0].data_ptr(), outs[1].data_ptr(),
                        cp[0], cp[1], cp[2], cp[3], n, ctypes.byref(ex)))
        if profile and _PROFILE_LOG is not None:
            for r in _lib.last_profile():
                r["name"] = f"{tag}/{r['name']}"
                _PROFILE_LOG.append(r)
        return (outs[0], codes[0], codes[2]), (outs[1], codes[1], codes[3])
    za = np.ascontiguousarray(z, dtype=np.complex128)
    va = np.asarray(v, dtype=np.float64)
    if va.size == 1:
        va, stride = np.ascontiguousarray(va.reshape(1)), 0
    else:
        va, stride = np.ascontiguousarray(np.broadcast_to(va, za.shape)), 1
    # caller-provided (e.g. pinned) buffers: outs = [J, Y], codes = [ierr_j, ierr_y, nz_j, nz_y]
    outs = [np.empty(za.shape, np.complex128) for _ in range(2)] if outs is None else outs
    if not want_codes:
        codes = [None] * 4
    elif codes is None:
        codes = [np.empty(za.shape, np.int32) for _ in range(4)]
    cp = [c.ctypes.data if c is not None else None for c in codes]
    ex = make_exec(flags=flags, devices=devices)
    check(entry(kode, va.ctypes.data, stride, za.ctypes.data, outs[0].ctypes.data, outs[1].ctypes.data,
                cp[0], cp[1], cp[2], cp[3], za.size, ctypes.byref(ex)))
    return (outs[0], codes[0], codes[2]), (outs[1], codes[1], codes[3])


def jvyv(v, z, scaled=False, outs=None, devices=None):
    """(jv(v, z), yv(v, z)) computed together; raises on the first Err of either.  No per-element codes are
    produced: the library tracks the first failing element of each output on the device."""
    (j, _, _), (y, _, _) = bessel_jy(v, z, scaled=scaled, outs=outs, want_codes=False, devices=devices)
    for which in (0, 1):
        idx, code = last_first_error(which)
        if idx >= 0:
            raise SpecialError(code, idx)
    return j, y


def last_first_error(which=0):
    """(index, code) of the first failing element of the last bessel / bessel_jy / airy call of this thread
    (-1, 0 if none); which = 1 selects the second output of bessel_jy."""
    idx = ctypes.c_int64()
    code = ctypes.c_int32()
    check(lib.spb_last_first_error(which, ctypes.byref(idx), ctypes.byref(code)))
    return idx.value, code.value


def first_error(ierr):
    """(index, code) of the first non-zero ierr in index order, (-1, 0) if none."""
    idx = ctypes.c_int64()
    code = ctypes.c_int32()
    if _is_torch(ierr):
        t = ierr.contiguous()
        with torch.cuda.device(t.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(t), devices=[t.device.index])
            check(lib.spb_first_error(t.data_ptr(), t.numel(), ctypes.byref(idx), ctypes.byref(code), ctypes.byref(ex)))
    else:
        a = np.ascontiguousarray(ierr, dtype=np.int32)
        ex = make_exec()
        check(lib.spb_first_error(a.ctypes.data, a.size, ctypes.byref(idx), ctypes.byref(code), ctypes.byref(ex)))
    return idx.value, code.value


def _result(vals, ierr, accept_half=True):
    """`collect::<Result<Array, i32>>()`: raise on the first Err in index order (mod.rs:11-12)."""
    idx, code = first_error(ierr)
    if idx >= 0:
        raise SpecialError(code, idx)
    return vals


# ---------------------------------------------------------------- reference surface
def i0(x):
    """special::i0 (mod.rs:8-13): I_0 of a real array through the complex routine; complex result."""
    vals, _, _ = bessel("I", 0.0, x, real_input=True, want_codes=False)
    idx, code = last_first_error(0)
    if idx >= 0:
        raise SpecialError(code, idx)
    return vals


def kv(v, z):
    """special::kv (kv.rs:11-25), scalar: NaN z -> 0, NaN v -> 0, then K_v(z); panics on Err."""
    z = complex(z)
    v = float(v)
    if z != z:
        z = 0j
    if v != v:
        v = 0.0
    # (negative orders: the C ABI folds K_{-v} = K_v on the device; the reference exercises kv(-3.5, 1000),
    # kv.rs:53-54)
    vals, ierr, _ = bessel("K", v, np.array([z], dtype=np.complex128))
    if ierr[0] != 0:
        print(f"{v} {z}")  # kv.rs:20-22
        raise SpecialPanic(f"called `Result::unwrap()` on an `Err` value: {int(ierr[0])}")
    return complex(vals[0])


def kve(v, z):
    """special::kve (kv.rs:39-41): kv(v, z) * exp(z)."""
    return kv(v, z) * np.exp(complex(z))


def sinc(x, devices=None, out=None):
    """special::sinc (trig.rs:4-13)."""
    return _real_map(lib.spb_sinc, x, devices, out)


def devmath(which, x):
    """The device transcendentals the region kernels use ('exp', 'expm' = exp(-x), 'sin', 'cos', 'log'), for accuracy
    tests of the library itself."""
    code = {"exp": 0, "expm": 1, "sin": 2, "cos": 3, "log": 4}[which]
    return _real_map(lambda a, b, n, ex: lib.spb_devmath(code, a, b, n, ex), x)


# ---------------------------------------------------------------- widened batched surface
def _checked(fn, v, z, scaled, **kw):
    vals, _, _ = bessel(fn, v, z, scaled=scaled, want_codes=False, **kw)
    idx, code = last_first_error(0)
    if idx >= 0:
        raise SpecialError(code, idx)
    return vals


def _family(fn, scaled):
    """Batched `xv` / `xve` entry point: orders of either sign (negative orders are evaluated by the reflection
    formulae inside the library, on the device, for host and device inputs alike); raises on the first Err."""
    def f(v, z, **kw):
        return _checked(fn, v, z, scaled, **kw)

    return f


jv, jve = _family("J", False), _family("J", True)
yv, yve = _family("Y", False), _family("Y", True)
iv, ive = _family("I", False), _family("I", True)
kv_array, kve_array = _family("K", False), _family("K", True)
hankel1, hankel1e = _family("H1", False), _family("H1", True)
hankel2, hankel2e = _family("H2", False), _family("H2", True)


def jn_sequence(n_orders, z, v0=0.0, scaled=False):
    """Order sweep J_{v0+k}(z_i), k = 0..n_orders-1, as an (n_orders, n) array (BASELINE config 5); raises on the
    first Err in index order."""
    vals, _, _ = bessel("J", v0, z, scaled=scaled, n_orders=n_orders, want_codes=False)
    idx, code = last_first_error(0)
    if idx >= 0:
        raise SpecialError(code, idx)
    return vals


def airy_raw(which, z, scaled=False, semantics="amos", devices=None, out=None, want_codes=True):
    kode = 2 if scaled else 1
    flags = SEM_SCIPY if semantics == "scipy" else SEM_AMOS
    if _is_torch(z):
        zc = z.resolve_conj().contiguous().to(torch.complex128)
        out = torch.empty_like(zc) if out is None else out
        ierr = torch.empty(zc.shape, dtype=torch.int32, device=zc.device) if want_codes else None
        nz = torch.empty(zc.shape, dtype=torch.int32, device=zc.device) if want_codes else None
        with torch.cuda.device(zc.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(zc), flags=flags, devices=[zc.device.index])
            check(lib.spb_airy(AIRY[which], kode, zc.data_ptr(), out.data_ptr(),
                               ierr.data_ptr() if want_codes else None, nz.data_ptr() if want_codes else None,
                               zc.numel(), ctypes.byref(ex)))
        return out, ierr, nz
    za = np.ascontiguousarray(z, dtype=np.complex128)
    out = np.empty_like(za) if out is None else out
    ierr = np.empty(za.shape, np.int32) if want_codes else None
    nz = np.empty(za.shape, np.int32) if want_codes else None
    ex = make_exec(flags=flags, devices=devices)
    check(lib.spb_airy(AIRY[which], kode, za.ctypes.data, out.ctypes.data, ierr.ctypes.data if want_codes else None,
                       nz.ctypes.data if want_codes else None, za.size, ctypes.byref(ex)))
    return out, ierr, nz


def airy(z, scaled=False):
    """(Ai, Ai', Bi, Bi') like scipy.special.airy; raises on the first Err."""
    outs = []
    for which in ("Ai", "Aip", "Bi", "Bip"):
        vals, ierr, _ = airy_raw(which, z, scaled=scaled)
        outs.append(_result(vals, ierr))
    return tuple(outs)


def airye(z):
    return airy(z, scaled=True)


def _real_map(fn, x, devices=None, out=None):
    if _is_torch(x):
        xc = x.contiguous().to(torch.float64)
        out = torch.empty_like(xc) if out is None else out
        with torch.cuda.device(xc.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(xc), devices=[xc.device.index])
            check(fn(xc.data_ptr(), out.data_ptr(), xc.numel(), ctypes.byref(ex)))
        return out
    xa = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(xa) if out is None else out
    ex = make_exec(devices=devices)
    check(fn(xa.ctypes.data, out.ctypes.data, xa.size, ctypes.byref(ex)))
    return out


def _cplx_map(fn, z, devices=None, out=None):
    if _is_torch(z):
        zc = z.resolve_conj().contiguous().to(torch.complex128)
        out = torch.empty_like(zc) if out is None else out
        with torch.cuda.device(zc.device):
            ex = make_exec(device_ptrs=True, stream=_stream_for(zc), devices=[zc.device.index])
            check(fn(zc.data_ptr(), out.data_ptr(), zc.numel(), ctypes.byref(ex)))
        return out
    za = np.ascontiguousarray(z, dtype=np.complex128)
    out = np.empty_like(za) if out is None else out
    ex = make_exec(devices=devices)
    check(fn(za.ctypes.data, out.ctypes.data, za.size, ctypes.byref(ex)))
    return out


def _is_complex(a):
    if _is_torch(a):
        return a.is_complex()
    return np.iscomplexobj(a)


def gamma(x, devices=None, out=None):
    if _is_complex(x):
        return _cplx_map(lib.spb_cgamma, x, devices, out)
    return _real_map(lib.spb_gamma, x, devices, out)


def gammaln(x, devices=None, out=None):
    """ln|Gamma(x)| for real x."""
    return _real_map(lib.spb_lgamma, x, devices, out)


def loggamma(z, devices=None, out=None):
    """Principal branch of log Gamma(z) for complex z; ln|Gamma(x)| for real x."""
    if _is_complex(z):
        return _cplx_map(lib.spb_clgamma, z, devices, out)
    return _real_map(lib.spb_lgamma, z, devices, out)


def synth(cfg, n, n_total=None, offset=0, seed=None, device=None):
    """Synthetic inputs of BASELINE config `cfg` generated on the GPU.

    Returns (v, z): torch CUDA tensors when `device` is given, numpy arrays otherwise."""
    seed = cfg if seed is None else seed
    n_total = n if n_total is None else n_total
    if device is not None:
        dev = torch.device(device)
        v = torch.empty(n, dtype=torch.float64, device=dev)
        z = torch.empty(n, dtype=torch.complex128, device=dev)
        with torch.cuda.device(dev):
            ex = make_exec(device_ptrs=True, stream=torch.cuda.current_stream(dev).cuda_stream, devices=[dev.index or 0])
            check(lib.spb_synth(cfg, seed, offset, n, n_total, v.data_ptr(), z.data_ptr(), ctypes.byref(ex)))
        return v, z
    v = np.empty(n, np.float64)
    z = np.empty(n, np.complex128)
    ex = make_exec()
    check(lib.spb_synth(cfg, seed, offset, n, n_total, v.ctypes.data, z.ctypes.data, ctypes.byref(ex)))
    return v, z


def device_topology(device=0):
    """'bdf=<pci address> numa=<node> cpus=<local_cpulist>' of a device (what multi-device worker threads bind to)."""
    buf = ctypes.create_string_buffer(512)
    n = lib.spb_device_topology(device, buf, 512)
    if n < 0:
        check(n)
    return buf.value.decode()


def fp64_peak(device=0, seconds=2.0):
    t = ctypes.c_double()
    mhz = ctypes.c_double()
    check(lib.spb_fp64_peak(device, seconds, ctypes.byref(t), ctypes.byref(mhz)))
    return t.value, mhz.value
