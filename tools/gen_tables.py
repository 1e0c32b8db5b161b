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


if __name__ == "__main__" and "--debye" not in sys.argv:
    main()


# ---------------------------------------------------------------- Debye-direct region (spb_debye.h)
def gen_debye_series(kmax=14):
    """Rearrangement of  sum_{k<=kmax} P_k(t^2) / s^k,  t^2 = v^2 / s^2  (u_k(t) = t^k P_k(t^2)),  by powers of 1/s:
         sum_m Q_m(v^2) / s^m,   Q_m(v^2) = sum_j c_{m-2j, j} (v^2)^j   (j <= k = m - 2j <= kmax)
    with REAL polynomials Q_m in the real variable v^2.  Writes the coefficient table DQ (Horner order per m, highest
    power of v^2 first) and the straight-line evaluation code (even and odd m as two Horner chains in y = 1/s^2)."""
    _, polys = gen_uk(kmax)  # polys[k] = coefficients of P_k, highest power of t^2 first (degree k)
    c = {}
    for k, row in enumerate(polys):
        for pos, val in enumerate(row):
            j = k - pos
            if val != 0:
                c[(k, j)] = val
    table_vals, lines = [], []
    mmax = 3 * kmax
    present = {}
    for m in range(1, mmax + 1):
        js = sorted([j for j in range(0, m // 3 + 1) if (m - 2 * j, j) in c and m - 2 * j <= kmax], reverse=True)
        if not js:
            continue
        jmin, jmax = js[-1], js[0]
        expr = None
        for j in range(jmax, jmin - 1, -1):
            idx = len(table_vals)
            table_vals.append(c.get((m - 2 * j, j), Fraction(0)))
            expr = f"real(DQ[{idx}])" if expr is None else f"fma_({expr}, v2, real(DQ[{idx}]))"
        if jmin > 0:
            expr = f"({expr}) * p{jmin}"
        present[m] = expr
    lines.append("// GENERATED by tools/gen_tables.py (gen_debye_series); do not edit.")
    lines.append("// in: real v2 (= v^2), cplx rs (= 1/s), cplx y (= 1/s^2); out: cplx ev (even m, incl. the leading 1), cplx od (odd m)")
    maxj = max(j for (_, j) in c)
    lines.append("const real p1 = v2;")
    for j in range(2, maxj + 1):
        a, b = j // 2, j - j // 2
        lines.append(f"const real p{j} = p{a} * p{b};")
    for parity, name in ((0, "ev"), (1, "od")):
        ms = sorted([m for m in present if m % 2 == parity], reverse=True)
        first = True
        for m in ms:
            if first:
                lines.append(f"cplx {name}({present[m]}, real(0.0));")
                prev = m
                first = False
                continue
            for _ in range((prev - m) // 2 - 1):  # an absent m in between: multiply only
                lines.append(f"{name} = {name} * y;")
            lines.append(f"{name} = horner_step({name}, y, {present[m]});  // m = {m}")
            prev = m
        if parity == 0:
            for _ in range(prev // 2 - 1):
                lines.append(f"{name} = {name} * y;")
            lines.append(f"{name} = horner_step({name}, y, real(1.0));  // m = 0")
        else:
            for _ in range((prev - 1) // 2):
                lines.append(f"{name} = {name} * y;")
            lines.append(f"{name} = {name} * rs;")
    return table_vals, lines


def write_debye():
    vals, lines = gen_debye_series(14)
    with open(os.path.join(OUTDIR, "spb_tables_debye.inc"), "w") as f:
        f.write("// spb_tables_debye.inc — GENERATED by tools/gen_tables.py (gen_debye_series); do not edit.\n"
                "// Coefficients c_{k,j} of the Debye polynomials u_k(t) = t^k sum_j c_{k,j} t^{2j} (DLMF 10.41.10-11),\n"
                "// regrouped by m = k + 2j (Horner order in v^2 per m).  Namespace-scope table: __constant__ on the device.\n")
        f.write("SPB_CONST_TABLE(DQ, %d) = {\n" % len(vals))
        for i in range(0, len(vals), 3):
            f.write("    " + ", ".join(fmt(frac_to_mp(v)) for v in vals[i:i + 3]) + ",\n")
        f.write("};\n")
    with open(os.path.join(OUTDIR, "spb_debye_series.inc"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("debye series:", len(vals), "coefficients,", len(lines), "statements")


def gen_atan_fit(deg=13):
    """Near-minimax fit (Chebyshev interpolation) of (atan(sqrt(u))/sqrt(u) - 1)/u on 0 <= u <= tan(pi/8)^2."""
    hi = mp.tan(mp.pi / 8) ** 2

    def f(u):
        if u == 0:
            return -mp.mpf(1) / 3
        r = mp.sqrt(u)
        return (mp.atan(r) / r - 1) / u
    n = deg + 1
    nodes = [hi / 2 * (1 + mp.cos(mp.pi * (2 * i + 1) / (2 * n))) for i in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, x in enumerate(nodes):
        for j in range(n):
            A[i, j] = x ** j
        b[i] = f(x)
    sol = mp.lu_solve(A, b)
    coef = [sol[j] for j in range(n)]
    # report the fit error
    worst = 0
    for i in range(2001):
        u = hi * i / 2000
        p = sum(coef[j] * u ** j for j in range(n))
        worst = max(worst, abs(p - f(u)) * u)  # error relative to atan(r)/r ~ 1
    return coef[::-1], worst


if __name__ == "__main__" and "--debye" in sys.argv:
    write_debye()
    co, w = gen_atan_fit(int(os.environ.get("ATAN_DEG", "13")))
    print("ATANC (highest power first), fit error %s:" % mp.nstr(w, 3))
    print(", ".join("%.17e" % float(c) for c in co))
