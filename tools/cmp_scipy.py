#!/usr/bin/env python
"""Development aid: compare the CPU checker (oracle/libspb_oracle.so) with scipy.special."""
import ctypes, sys, os
import numpy as np
import scipy.special as sc

HERE = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(HERE, "..", "oracle", "libspb_oracle.so"))
P = ctypes.c_void_p
lib.spbo_bessel.argtypes = [ctypes.c_int, ctypes.c_int, P, ctypes.c_int64, P, P, P, P, ctypes.c_int64, ctypes.c_int]
lib.spbo_airy.argtypes = [ctypes.c_int, ctypes.c_int, P, P, P, P, ctypes.c_int64, ctypes.c_int]
FN = dict(J=0, Y=1, I=2, K=3, H1=4, H2=5)
SC = {("J", 1): sc.jv, ("J", 2): sc.jve, ("Y", 1): sc.yv, ("Y", 2): sc.yve, ("I", 1): sc.iv, ("I", 2): sc.ive,
      ("K", 1): sc.kv, ("K", 2): sc.kve, ("H1", 1): sc.hankel1, ("H1", 2): sc.hankel1e, ("H2", 1): sc.hankel2,
      ("H2", 2): sc.hankel2e}


def ours(fn, kode, v, z, threads=8):
    z = np.ascontiguousarray(z, dtype=np.complex128)
    v = np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), z.shape))
    out = np.empty_like(z)
    ierr = np.empty(z.shape, np.int32)
    nz = np.empty(z.shape, np.int32)
    lib.spbo_bessel(FN[fn], kode, v.ctypes.data, 1, z.ctypes.data, out.ctypes.data, ierr.ctypes.data, nz.ctypes.data,
                    z.size, threads)
    return out, ierr, nz


def ours_airy(which, kode, z, threads=8):
    z = np.ascontiguousarray(z, dtype=np.complex128)
    out = np.empty_like(z)
    ierr = np.empty(z.shape, np.int32)
    nz = np.empty(z.shape, np.int32)
    lib.spbo_airy(which, kode, z.ctypes.data, out.ctypes.data, ierr.ctypes.data, nz.ctypes.data, z.size, threads)
    return out, ierr, nz


def relerr(a, b):
    d = np.abs(a - b)
    m = np.abs(b)
    with np.errstate(divide="ignore", invalid="ignore"):
        r = np.where(m > 0, d / m, np.where(d == 0, 0, np.inf))
    both_nan = np.isnan(a) & np.isnan(b)
    r = np.where(both_nan, 0, r)
    same_inf = np.isinf(a.real) & (a.real == b.real) | np.isinf(a.imag) & (a.imag == b.imag)
    r = np.where(same_inf & ~np.isfinite(r), 0, r)
    return r


def report(name, v, z, got, want, ierr, nz, top=5):
    r = relerr(got, want)
    bad = ~np.isfinite(r)
    rr = np.where(bad, np.inf, r)
    idx = np.argsort(-rr)[:top]
    print(f"{name}: n={z.size} max={np.nanmax(rr):.3e} p99.9={np.quantile(rr[np.isfinite(rr)], 0.999) if np.isfinite(rr).any() else np.nan:.3e} "
          f"median={np.median(rr):.2e} nonfinite={bad.sum()} ierr!=0:{(ierr != 0).sum()} nz!=0:{(nz != 0).sum()}")
    vv = np.broadcast_to(v, z.shape)
    for i in idx:
        if rr[i] > 1e-13:
            print(f"   v={vv[i]!r} z={z[i]!r} got={got[i]!r} want={want[i]!r} rel={rr[i]:.2e} ierr={ierr[i]} nz={nz[i]}")


def gen(cfg, n, rng):
    if cfg == "cfg2":
        a = 200 / np.sqrt(2)
        z = rng.uniform(-a, a, n) + 1j * rng.uniform(-a, a, n)
        return np.full(n, 2.5), z
    if cfg == "cfg3":
        v = 50 * rng.random(n)
        z = 200 * rng.random(n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
        return v, z
    if cfg == "cfg4":
        z = 10 ** (-3 + 7 * rng.random(n)) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
        return np.full(n, 2.5), z
    if cfg == "real":
        return np.zeros(n), (100 * rng.random(n)).astype(np.complex128)
    if cfg == "wide":
        v = 10 ** rng.uniform(-2, 2.3, n)
        z = 10 ** rng.uniform(-3, 3, n) * np.exp(1j * np.pi * (2 * rng.random(n) - 1))
        return v, z
    raise ValueError(cfg)


if __name__ == "__main__":
    cfg = sys.argv[1]
    fns = sys.argv[2].split(",")
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
    kodes = [int(k) for k in sys.argv[4].split(",")] if len(sys.argv) > 4 else [1, 2]
    rng = np.random.default_rng(1234)
    v, z = gen(cfg, n, rng)
    for fn in fns:
        for kode in kodes:
            if fn in ("Ai", "Aip", "Bi", "Bip"):
                which = ["Ai", "Aip", "Bi", "Bip"].index(fn)
                got, ierr, nz = ours_airy(which, kode, z)
                with np.errstate(all="ignore"):
                    want = (sc.airy if kode == 1 else sc.airye)(z)[which]
            else:
                got, ierr, nz = ours(fn, kode, v, z)
                with np.errstate(all="ignore"):
                    want = SC[(fn, kode)](v, z)
            report(f"{cfg} {fn} kode={kode}", v, z, got, want, ierr, nz)
