"""Shared helpers for the tests: the CPU checker library (oracle/), golden vectors, tolerance policy."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
FN = {"J": 0, "Y": 1, "I": 2, "K": 3, "H1": 4, "H2": 5}
CATS = ["", "domain", "overflow", "underflow", "loss", "no_result", "singular", "slow", "other"]

_oracle = None


def seed_of(*parts):
    """Deterministic 32-bit seed from strings (Python's hash() is salted per process)."""
    import zlib
    return zlib.crc32("/".join(str(p) for p in parts).encode())


def oracle_lib():
    """oracle/libspb_oracle.so, built on demand with g++ (CPU only)."""
    global _oracle
    if _oracle is None:
        path = os.path.join(ROOT, "oracle", "libspb_oracle.so")
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        lib = ctypes.CDLL(path)
        P, I, L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
        lib.spbo_bessel.argtypes = [I, I, P, L, P, P, P, P, L, I]
        lib.spbo_bessel_sem.argtypes = [I, I, I, P, L, P, P, P, P, L, I]
        lib.spbo_bessel_pipeline.argtypes = [I, I, P, L, P, P, P, P, P, L, I]
        lib.spbo_airy.argtypes = [I, I, P, P, P, P, L, I]
        lib.spbo_airy_flow.argtypes = [I, I, P, P, P, P, P, L, I]
        lib.spbo_bessel_jy_pipeline.argtypes = [I, P, L, P, P, P, P, L, I]
        lib.spbo_bessel_ik_pipeline.argtypes = [I, P, L, P, P, P, P, L, I]
        lib.spbo_bessel_seq.argtypes = [I, I, P, L, P, P, P, P, L, I, I]
        lib.spbo_kaiser.argtypes = [L, ctypes.c_double, I, P]
        lib.spbo_lanczos.argtypes = [L, I, P]
        lib.spbo_gamma.argtypes = [I, P, P, L]
        lib.spbo_cgamma.argtypes = [I, P, P, L]
        _oracle = lib
    return _oracle


def _mode_bits(fuse):
    """fuse: False / True (fused large-|w| I + K routine) or a bitmask (1 = that, 2 = Debye-direct region, 4 = K of
    a Miller + Wronskian point from the Wronskian routine)."""
    f = int(fuse)
    return (0x10000 if f & 1 else 0) | (0x20000 if f & 2 else 0) | (0x40000 if f & 4 else 0)


def cpu_bessel(fn, kode, v, z, sem=0, pipeline=False, threads=8, fuse=False):
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    out = np.empty_like(z)
    ierr = np.empty(z.shape, np.int32)
    nz = np.empty(z.shape, np.int32)
    if pipeline:
        path = np.empty(z.shape, np.int32)
        lib.spbo_bessel_pipeline(FN[fn], kode, v.ctypes.data, 1, z.ctypes.data, out.ctypes.data, ierr.ctypes.data,
                                 nz.ctypes.data, path.ctypes.data, z.size, threads | _mode_bits(fuse))
        return out, ierr, nz, path
    lib.spbo_bessel_sem(FN[fn], kode, sem, v.ctypes.data, 1, z.ctypes.data, out.ctypes.data, ierr.ctypes.data,
                        nz.ctypes.data, z.size, threads)
    return out, ierr, nz


def cpu_bessel_seq(fn, kode, v, z, n_orders, threads=8, classic=False, device_flow=False):
    """Order sequence F_{v+k}(z), k < n_orders, on the CPU checker: ((n_orders, n) values, ierr[n], nz[n]).
    classic=True forces the classic sequence drivers (no write-once sweep) for J and I; device_flow=True runs the
    sweep as the CUDA kernels do (K pair of the normalisation from the large-|w| series / handed in from bknu)."""
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128).ravel()
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    out = np.empty((n_orders, z.size), np.complex128)
    ierr = np.empty(z.size, np.int32)
    nz = np.empty(z.size, np.int32)
    lib.spbo_bessel_seq(FN[fn], kode, v.ctypes.data, 1, z.ctypes.data, out.ctypes.data, ierr.ctypes.data,
                        nz.ctypes.data, z.size, n_orders, threads | (0x10000 if classic else 0) | (0x20000 if device_flow else 0))
    return out, ierr, nz


def cpu_classify(fn, kode, v, z, fuse):
    """(fast, full) class bytes of the classifier: fast = -1 where the fast accept declines."""
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    fast = np.empty(z.shape, np.int8)
    full = np.empty(z.shape, np.int8)
    lib.spbo_classify({**FN, "JY": 6, "IK": 7}[fn], kode, int(fuse), ctypes.c_void_p(v.ctypes.data), ctypes.c_int64(1),
                      ctypes.c_void_p(z.ctypes.data), ctypes.c_void_p(fast.ctypes.data),
                      ctypes.c_void_p(full.ctypes.data), ctypes.c_int64(z.size))
    return fast, full


def cpu_bessel_jy(kode, v, z, threads=8, fuse=False):
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    oj, oy = np.empty_like(z), np.empty_like(z)
    codes = np.empty((4, z.size), np.int32)
    lib.spbo_bessel_jy_pipeline(kode, v.ctypes.data, 1, z.ctypes.data, oj.ctypes.data, oy.ctypes.data, codes.ctypes.data,
                                z.size, threads | _mode_bits(fuse))
    return (oj, codes[0], codes[2]), (oy, codes[1], codes[3])


def cpu_bessel_ik(kode, v, z, threads=8, fuse=7):
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    oi, ok = np.empty_like(z), np.empty_like(z)
    codes = np.empty((4, z.size), np.int32)
    lib.spbo_bessel_ik_pipeline(kode, v.ctypes.data, 1, z.ctypes.data, oi.ctypes.data, ok.ctypes.data, codes.ctypes.data,
                                z.size, threads | _mode_bits(fuse))
    return (oi, codes[0], codes[2]), (ok, codes[1], codes[3])


def cpu_airy(which, kode, z, threads=8):
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    out = np.empty_like(z)
    ierr = np.empty(z.shape, np.int32)
    nz = np.empty(z.shape, np.int32)
    lib.spbo_airy(which, kode, z.ctypes.data, out.ctypes.data, ierr.ctypes.data, nz.ctypes.data, z.size, threads)
    return out, ierr, nz


def cpu_airy_flow(which, kode, z, threads=8):
    """Airy through the flow of the GPU kernels (lean large-|zeta| routines where they apply): values, ierr, nz and
    path (1 = lean routine)."""
    lib = oracle_lib()
    z = np.ascontiguousarray(z, dtype=np.complex128)
    out = np.empty_like(z)
    ierr = np.empty(z.shape, np.int32)
    nz = np.empty(z.shape, np.int32)
    path = np.empty(z.shape, np.int32)
    lib.spbo_airy_flow(which, kode, z.ctypes.data, out.ctypes.data, ierr.ctypes.data, nz.ctypes.data, path.ctypes.data,
                       z.size, threads)
    return out, ierr, nz, path


def cpu_gamma(which, x):
    """which: 0 gamma, 1 gammaln, 2 sinc (real); 10 gamma, 11 loggamma (complex)."""
    lib = oracle_lib()
    if which >= 10:
        z = np.ascontiguousarray(x, dtype=np.complex128)
        out = np.empty_like(z)
        lib.spbo_cgamma(which - 10, z.ctypes.data, out.ctypes.data, z.size)
        return out
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    lib.spbo_gamma(which, x.ctypes.data, out.ctypes.data, x.size)
    return out


def cpu_kaiser(m, beta, sym=True):
    out = np.empty(m, np.float64)
    oracle_lib().spbo_kaiser(m, float(beta), 1 if sym else 0, out.ctypes.data)
    return out


def cpu_lanczos(m, sym=True):
    out = np.empty(m, np.float64)
    oracle_lib().spbo_lanczos(m, 1 if sym else 0, out.ctypes.data)
    return out


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def expected_ierr_ok(cat_code, ierr, nz):
    """Does (ierr, nz) agree with the oracle's sf_error class?  underflow <-> nz != 0; loss <-> ierr 3;
    domain 1; overflow 2; no_result 4 or 5."""
    cat = CATS[cat_code]
    if cat == "":
        return ierr == 0 and nz == 0
    if cat == "underflow":
        return nz != 0 and ierr == 0
    if cat == "loss":
        return ierr == 3
    if cat == "domain":
        return ierr == 1
    if cat == "overflow":
        return ierr == 2
    if cat == "no_result":
        return ierr in (4, 5)
    return False
