#!/usr/bin/env python
"""Generate sciport_rs_b200/csrc/amos/spb_tables_*.inc from first principles.

Every table the Amos-style region routines need is a list of mathematical
constants; none of them is copied from any source file.  They are recomputed
here with exact rationals (sympy) or 60-digit arithmetic (mpmath):

  GLN[100]   ln Gamma(n), n = 1..100                      (lgamma shortcut)
  CF[22]     B_2k / (2k (2k-1)), k = 1..22                (Stirling series)
  UK[120]    coefficients of the Debye polynomials u_k(t), k = 0..14
             (DLMF 10.41.10-11), highest power of t^2 first, as used by the
             uniform expansion in the order (I/K) and, first 105, by the
             Airy-type expansion (J/Y/H)
  AR[14],BR[14]  (3/2)^k u_k, (3/2)^k v_k of the Airy asymptotic series
             (DLMF 9.7.2)
  GAMA[30]   Maclaurin coefficients in w^2 = 1 - z^2 of zeta / w^2
             (DLMF 10.20.2-3)
  ALFA[180]  Maclaurin coefficients in w^2 of A_s(zeta), s = 1..6  (30 each)
  LG1[32]    (-1)^k (zeta(k)-1)/k, k = 2..33: Maclaurin series of lgamma(1+t) (DLMF 5.7.3)
  BETA[210]  Maclaurin coefficients in w^2 of B_s(zeta), s = 0..6  (30 each)
             (DLMF 10.20.10-13); extracted by a Cauchy integral on |w^2| = r
             because the closed forms have removable singularities at w = 0.

Run:  python tools/gen_tables.py            (writes the three spb_tables_*.inc files)
"""
import os
import sys
from fractions import Fraction

import mpmath as mp
import sympy as sp

mp.mp.dps = 70
HERE = os.path.dirname(os.path.abspath(__file__))
OUTDIR = os.path.join(HERE, "..", "sciport_rs_b200", "csrc", "amos")


def fmt(x):
    return mp.nstr(mp.mpf(x), 21, min_fixed=0, max_fixed=0, strip_zeros=False)


def frac_to_mp(fr):
    return mp.mpf(fr.numerator) / mp.mpf(fr.denominator)


def table(name, vals, per_line=3):
    lines = ["SPB_TABLE(%s, %d) = {" % (name, len(vals))]
    for i in range(0, len(vals), per_line):
        lines.append("    " + ", ".join(fmt(v) for v in vals[i:i + per_line]) + ",")
    lines.append("};")
    return "\n".join(lines)


# ---------------------------------------------------------------- GLN / CF
def gen_gln():
    return [mp.loggamma(n) for n in range(1, 101)]


def gen_cf():
    out = []
    for k in range(1, 23):
        b = sp.bernoulli(2 * k)
        out.append(frac_to_mp(Fraction(int(b.p), int(b.q)) / (2 * k * (2 * k - 1))))
    return out


# ---------------------------------------------------------------- Debye u_k
def debye_polys(kmax):
    t = sp.symbols("t")
    u = [sp.Integer(1)]
    for k in range(kmax):
        uk = u[-1]
        nxt = sp.Rational(1, 2) * t**2 * (1 - t**2) * sp.diff(uk, t) + sp.Rational(1, 8) * sp.integrate(
            (1 - 5 * t**2) * uk, (t, 0, t))
        u.append(sp.expand(nxt))
    return t, u


def gen_uk(kmax=14):
    """u_k(t) = t^k * P_k(t^2); list P_k coefficients, highest power first."""
    t, u = debye_polys(kmax)
    flat = []
    polys = []
    for k, uk in enumerate(u):
        p = sp.Poly(uk, t)
        coeffs = {}
        for (e,), c in p.terms():
            assert (e - k) % 2 == 0 and 0 <= (e - k) // 2 <= k
            coeffs[(e - k) // 2] = Fraction(int(c.p), int(c.q))
        row = [coeffs.get(j, Fraction(0)) for j in range(k, -1, -1)]
        polys.append(row)
        flat.extend(row)
    return flat, polys


# ---------------------------------------------------------------- Airy u_k, v_k
def airy_uv(kmax):
    u = [Fraction(1)]
    v = [Fraction(1)]
    for k in range(1, kmax + 1):
        # u_k = (2k+1)(2k+3)...(6k-1) / (216^k k!)
        num = 1
        for j in range(2 * k + 1, 6 * k, 2):
            num *= j
        den = 216**k
        for j in range(1, k + 1):
            den *= j
        uk = Fraction(num, den)
        u.append(uk)
        v.append(uk * Fraction(6 * k + 1, 1 - 6 * k))
    return u, v


# ---------------------------------------------------------------- zeta(w^2), A_s, B_s
def U_polys(kmax):
    """Debye polynomials as python callables on mp numbers (variable p)."""
    t, u = debye_polys(kmax)
    fs = []
    for uk in u:
        p = sp.Poly(uk, t)
        terms = [(e, mp.mpf(int(c.p)) / mp.mpf(int(c.q))) for (e,), c in p.terms()]
        fs.append(terms)
    return fs


def eval_terms(terms, x):
    return mp.fsum(c * x**e for e, c in terms)


def zeta_of_w2(x):
    """zeta as an analytic function of x = w^2 = 1 - z^2 near x = 0:
    (2/3) zeta^(3/2) = atanh(w) - w."""
    w = mp.sqrt(x)
    f = mp.atanh(w) - w  # = w^3 (1/3 + x/5 + ...)
    # zeta = x * ( (3/2) f / w^3 )^(2/3), analytic branch (real positive at x>0)
    g = mp.mpf(3) / 2 * f / (w**3)
    return x * mp.exp(mp.mpf(2) / 3 * mp.log(g))


def AB_at(x, smax, Uterms, au, av):
    """A_s(zeta), B_s(zeta), s=0..smax at x = w^2 (|x| not tiny), DLMF 10.20.10-11."""
    w = mp.sqrt(x)
    p = 1 / w  # (1 - z^2)^(-1/2)
    zeta = zeta_of_w2(x)
    sq = mp.sqrt(zeta)
    z32 = zeta * sq  # zeta^(3/2) with the principal branch consistent with atanh(w)-w
    # consistency: (2/3) zeta^(3/2) must equal atanh(w) - w
    target = mp.mpf(3) / 2 * (mp.atanh(w) - w)
    if abs(z32 - target) > abs(z32 + target):
        z32 = -z32
        sq = -sq
    A = []
    B = []
    for s in range(smax + 1):
        a = mp.mpf(0)
        for j in range(0, 2 * s + 1):
            a += frac_to_mp(av[j]) * mp.mpf(3 / 2) ** 0 * (mp.mpf(3) / 2) ** j * z32 ** (-j) * eval_terms(Uterms[2 * s - j], p)
        b = mp.mpf(0)
        for j in range(0, 2 * s + 2):
            b += frac_to_mp(au[j]) * (mp.mpf(3) / 2) ** j * z32 ** (-j) * eval_terms(Uterms[2 * s - j + 1], p)
        A.append(a)
        B.append(-b / sq)
    return A, B


def taylor_by_cauchy(fvals_on_circle, r, ncoef):
    """fvals sampled at x_m = r exp(2 pi i m / M); returns c_k, k < ncoef."""
    M = len(fvals_on_circle)
    out = []
    for k in range(ncoef):
        s = mp.mpc(0)
        for m, fv in enumerate(fvals_on_circle):
            s += fv * mp.expjpi(-2 * mp.mpf(k) * m / M)
        out.append((s / M / r**k).real)
    return out


def gen_gama_alfa_beta(K=30, SMAX=6, M=256, r=mp.mpf("0.55")):
    Uterms = U_polys(2 * SMAX + 2)
    au, av = airy_uv(2 * SMAX + 2)
    xs = [r * mp.expjpi(2 * mp.mpf(m) / M) for m in range(M)]
    gvals = [zeta_of_w2(x) / x for x in xs]
    gama = taylor_by_cauchy(gvals, r, K)
    Avals = [[] for _ in range(SMAX + 1)]
    Bvals = [[] for _ in range(SMAX + 1)]
    for x in xs:
        A, B = AB_at(x, SMAX, Uterms, au, av)
        for s in range(SMAX + 1):
            Avals[s].append(A[s])
            Bvals[s].append(B[s])
    alfa = []
    for s in range(1, SMAX + 1):
        alfa.extend(taylor_by_cauchy(Avals[s], r, K))
    beta = []
    for s in range(0, SMAX + 1):
        beta.extend(taylor_by_cauchy(Bvals[s], r, K))
    return gama, alfa, beta


HEADER = ("// %s — GENERATED by tools/gen_tables.py; do not edit.\n"
          "// Mathematical constants recomputed from their definitions (see the generator's docstring).\n"
          "// Included inside a function body so the arrays are function-local statics (host + device).\n")


def write(name, tables):
    path = os.path.join(OUTDIR, name)
    with open(path, "w") as f:
        f.write(HEADER % name + "\n".join(tables) + "\n")
    print("wrote", os.path.normpath(path))


def gen_fastmath():
    """Coefficients of the device exp / sincos in amos/spb_common.h (namespace fm): near-minimax (Chebyshev
    interpolation) fits of P(u) = (sin(sqrt u)/sqrt u - 1)/u and Q(u) = (cos(sqrt u) - 1 + u/2)/u^2 on
    0 <= u <= 1.02 (pi/4)^2, highest power first, and the Taylor coefficients of exp split into its even part
    cosh r - 1 (EXPE: 1/12! .. 1/2!) and its odd part sinh r / r - 1 (EXPO: 1/13! .. 1/3!).
    `python tools/gen_tables.py --fastmath` prints them for comparison with the header."""
    q = (mp.pi / 4) ** 2 * mp.mpf("1.02")
    P = lambda u: (mp.sin(mp.sqrt(u)) / mp.sqrt(u) - 1) / u
    Q = lambda u: (mp.cos(mp.sqrt(u)) - 1 + u / 2) / (u * u)
    sinc = mp.chebyfit(P, [mp.mpf("1e-30"), q], 7)
    cosc = mp.chebyfit(Q, [mp.mpf("1e-30"), q], 7)
    expe = [1 / mp.factorial(k) for k in range(12, 1, -2)]
    expo = [1 / mp.factorial(k) for k in range(13, 2, -2)]
    # log: (2 atanh(s)/s - 2)/s^2 as a polynomial in u = s^2 on u <= 1.02 ((sqrt 2 - 1)/(sqrt 2 + 1))^2
    L = lambda u: (2 * mp.atanh(mp.sqrt(u)) / mp.sqrt(u) - 2) / u
    logc = mp.chebyfit(L, [mp.mpf("1e-30"), mp.mpf("0.02944") * mp.mpf("1.02")], 7)
    return sinc, cosc, expe, expo, logc


def main():
    if "--fastmath" in sys.argv:
        for name, t in zip(("SINC", "COSC", "EXPE", "EXPO", "LOGC"), gen_fastmath()):
            print(name, ", ".join("%.17e" % float(c) for c in t))
        return
    gln = gen_gln()
    cf = gen_cf()
    write("spb_tables_gln.inc", [table("GLN", gln), table("CF", cf)])
    uk, _ = gen_uk(14)
    write("spb_tables_unik.inc", [table("UK", [frac_to_mp(c) for c in uk])])
    au, av = airy_uv(13)
    ar = [frac_to_mp(au[k] * Fraction(3, 2) ** k) for k in range(14)]
    br = [frac_to_mp(av[k] * Fraction(3, 2) ** k) for k in range(14)]
    # lgamma(1+t) = -log1p(t) + t(1-euler) + sum_{k>=2} (-1)^k (zeta(k)-1) t^k / k   (DLMF 5.7.3)
    lg1 = [(-1) ** k * (mp.zeta(k) - 1) / k for k in range(2, 34)]
    # loggamma(1+t) = -euler t + sum_{k>=2} (-1)^k zeta(k) t^k / k for |t| <= 0.2 (complex Taylor branch)
    lgz = [(-1) ** k * mp.zeta(k) / k for k in range(2, 25)]
    write("spb_tables_gamma.inc", [table("LG1", lg1), table("LGZ", lgz)])
    if "--fast" in sys.argv:
        return
    gama, alfa, beta = gen_gama_alfa_beta()
    write("spb_tables_unhj.inc", [table("AR", ar), table("BR", br), table("UC", [frac_to_mp(c) for c in uk[:105]]),
                                  table("GAMA", gama), table("ALFA", alfa), table("BETA", beta)])


if __name__ == "__main__":
    main()
