"""ctypes binding of libsciport_b200.so (the C ABI in include/sciport_b200.h).

The library is the product: if it cannot be loaded this module raises, and
every compute call raises when no CUDA device is present — there is no CPU
fallback anywhere in this package.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsciport_b200.so")

FN = {"J": 0, "Y": 1, "I": 2, "K": 3, "H1": 4, "H2": 5}
AIRY = {"Ai": 0, "Aip": 1, "Bi": 2, "Bip": 3}
SEM_AMOS, SEM_SCIPY, RAW_NEG_ORDERS, PATH_MONOLITHIC, PROFILE, PATH_NO_DIRECT = 0, 1, 2, 16, 32, 64
KERNEL_NAMES = ["classify", "scan_scatter", "ipass_series", "ipass_asyi", "ipass_mlri", "ipass_wrsk", "kpass",
                "general", "monolithic", "ipass_uni1", "ipass_uni2", "ipass_debye", "direct_asyi"]
E_ARG, E_NO_DEVICE, E_CUDA, E_ALLOC = -1, -2, -3, -4


class SpbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libsciport_b200 error {code}: {msg}")
        self.code = code


class Exec(ctypes.Structure):
    _fields_ = [
        ("n_devices", ctypes.c_int),
        ("devices", ctypes.POINTER(ctypes.c_int)),
        ("device_ptrs", ctypes.c_int),
        ("stream", ctypes.c_void_p),
        ("flags", ctypes.c_int),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). sciport_rs_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    EX = ctypes.POINTER(Exec)
    lib.spb_bessel.argtypes = [I, I, P, L, P, P, P, P, L, I, EX]
    lib.spb_bessel_real.argtypes = [I, I, P, L, P, P, P, P, L, I, EX]
    lib.spb_bessel_jy.argtypes = [I, P, L, P, P, P, P, P, P, P, L, EX]
    lib.spb_bessel_ik.argtypes = [I, P, L, P, P, P, P, P, P, P, L, EX]
    lib.spb_airy.argtypes = [I, I, P, P, P, P, L, EX]
    for name in ("spb_gamma", "spb_lgamma", "spb_cgamma", "spb_clgamma", "spb_sinc"):
        getattr(lib, name).argtypes = [P, P, L, EX]
    lib.spb_devmath.argtypes = [I, P, P, L, EX]
    lib.spb_kaiser.argtypes = [L, ctypes.c_double, I, P, EX]
    lib.spb_lanczos.argtypes = [L, I, P, EX]
    lib.spb_window.argtypes = [I, L, I, P, I, P, EX]
    lib.spb_first_error.argtypes = [P, L, ctypes.POINTER(L), ctypes.POINTER(ctypes.c_int32), EX]
    lib.spb_last_first_error.argtypes = [I, ctypes.POINTER(L), ctypes.POINTER(ctypes.c_int32)]
    lib.spb_synth.argtypes = [I, ctypes.c_uint64, L, L, L, P, P, EX]
    lib.spb_fp64_peak.argtypes = [I, ctypes.c_double, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.spb_last_timing.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(I)]
    lib.spb_last_profile.argtypes = [I, P, P, P]
    lib.spb_device_topology.argtypes = [I, ctypes.c_char_p, I]
    lib.spb_last_error.restype = ctypes.c_char_p
    for name in EXPORTS:
        getattr(lib, name)  # every declared symbol must be exported
    return lib


EXPORTS = [
    "spb_bessel", "spb_bessel_jy", "spb_bessel_ik", "spb_bessel_real", "spb_airy", "spb_gamma", "spb_lgamma", "spb_cgamma", "spb_clgamma",
    "spb_sinc", "spb_devmath", "spb_kaiser", "spb_lanczos", "spb_window", "spb_first_error", "spb_last_first_error", "spb_synth", "spb_fp64_peak", "spb_last_timing", "spb_last_profile", "spb_device_topology", "spb_device_count",
    "spb_version", "spb_last_error",
]

lib = _load()


def check(rc):
    if rc != 0:
        raise SpbError(rc, lib.spb_last_error().decode())


def make_exec(device_ptrs=False, stream=0, flags=0, devices=None):
    ex = Exec()
    if devices:
        arr = (ctypes.c_int * len(devices))(*devices)
        ex.n_devices = len(devices)
        ex.devices = arr
        ex._keep = arr
    else:
        ex.n_devices = 1
        ex.devices = None
    ex.device_ptrs = 1 if device_ptrs else 0
    ex.stream = stream or None
    ex.flags = flags
    return ex


def last_timing(want_ms=True):
    """(kernel_ms, launches) of the last compute call on this thread; want_ms=False does not synchronise."""
    ms = ctypes.c_double()
    n = ctypes.c_int()
    lib.spb_last_timing(ctypes.byref(ms) if want_ms else None, ctypes.byref(n))
    return ms.value, n.value


def last_profile(max_records=4096):
    import numpy as np
    ids = np.zeros(max_records, np.int32)
    ms = np.zeros(max_records, np.float64)
    pts = np.zeros(max_records, np.int64)
    n = lib.spb_last_profile(max_records, ids.ctypes.data, ms.ctypes.data, pts.ctypes.data)
    return [{"name": KERNEL_NAMES[ids[i]], "ms": float(ms[i]), "points": int(pts[i])} for i in range(n)]
