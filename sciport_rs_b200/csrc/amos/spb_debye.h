// spb_debye.h — I_v(w) and K_v(w), Re w >= 0, by the uniform (Debye-type) expansion evaluated AT THE ORDER ITSELF
// (DLMF 10.41.3-4), with the second exponential of I restored near the imaginary axis.
//
// The published region map (TOMS 644, spb_binu.h) uses the uniform expansions only for orders above
// fnul = 85.9: below, it either runs the Miller / Wronskian route (|w| <= fnul) or raises the order to fnul,
// evaluates two expansions there and recurs backward over up to 86 orders (|w| > fnul).  The bound behind that
// choice is uniform in w.  For a GIVEN w the expansion parameter is much larger than the order: with
//     s = sqrt(v^2 + w^2),  t = v / s,  u_k(t) = t^k P_k(t^2)   (P_k of degree k, DLMF 10.41.10-11)
// the k-th term of  sum u_k(t) / v^k  is  P_k(t^2) / s^k, so that
//   * for |t| <= 1 the series behaves like the large-|w| Hankel series in 1/|s| (|s| >= |w|-ish), and
//   * for |t| >  1 (towards the turning points w = +-iv) like a series in v^2 / |s|^3.
// Fourth departure from the published map (after the three of DESIGN.md section 2): wherever
//     |s| >= SPB_DEBYE_S0   and   |s|^3 / v^2 >= SPB_DEBYE_P0
// (calibrated against 60-digit mpmath, tests/golden/make_golden_debye.py; measured error <= 3e-14 inside the
// region) ONE evaluation of the series gives
//     I_v(w) = e^{+v eta} / sqrt(2 pi s) * sum      P_k / s^k   +  second exponential,
//     K_v(w) = e^{-v eta} * sqrt(pi / 2s) * sum (-1)^k P_k / s^k,
//     v eta  = s + v ln(w / (v + s)),
// at a few hundred flops, where the routes above cost 2500 (Miller ratios + K by recurrence + Wronskian) to 2700
// (two raised-order expansions + backward recurrence).  Both functions share every P_k(t^2).
//
// Second exponential.  The dominant series alone is the expansion of -+(i/pi) K_v(w e^{-+i pi}); exactly
//     I_v(w) = -+(i/pi) K_v(w e^{-+i pi})  +-  (i/pi) e^{+-i pi v} K_v(w)        (upper signs: Im w > 0),
// DLMF 10.34.2.  The last term is negligible for Re(v eta) >> 1 and of the same size as the first one on the
// imaginary axis beyond the turning point (|w| > v, the oscillatory region of J_v); the large-|w| routine (asyi)
// keeps the analogous term of the Hankel expansion.  With it the expansion holds up to and on the imaginary axis.
#pragma once
#include "spb_common.h"

// Namespace-scope coefficient tables: constant memory on the device (a DFMA then reads the coefficient as a
// constant-bank operand; as a literal it costs two moves into a register pair first: 17 % of the instructions of
// the first version of this kernel), a plain static array on the host.
#if defined(__CUDA_ARCH__)
#define SPB_CONST_TABLE(name, n) static __constant__ double name[n]
#elif defined(__CUDACC__)
#define SPB_CONST_TABLE(name, n) static __constant__ double name[n]
#else
#define SPB_CONST_TABLE(name, n) static const double name[n]
#endif

namespace spb {
namespace dby {
#include "spb_tables_debye.inc"
// (atan(r)/r - 1)/r^2 on r^2 <= tan(pi/8)^2, highest power first (tools/gen_tables.py --debye; fit error 2e-22)
SPB_CONST_TABLE(ATANC, 14) = {1.19595677955128283e-02, -2.72748188162988958e-02, 3.72343513633974776e-02,
                              -4.29426924398485510e-02, 4.75462140340531039e-02, -5.26245398479938398e-02,
                              5.88230463938551468e-02,  -6.66666434483427356e-02, 7.69230761627277687e-02,
                              -9.09090908929153091e-02, 1.11111111110904534e-01,  -1.42857142857141489e-01,
                              1.99999999999999983e-01,  -3.33333333333333315e-01};
}  // namespace dby

#ifndef SPB_DEBYE_S0
#define SPB_DEBYE_S0 36.0
#endif
#ifndef SPB_DEBYE_P0
#define SPB_DEBYE_P0 240.0
#endif

// acc * y + q with a real q: two fused multiply-adds per component
SPB_HD cplx horner_step(cplx acc, cplx y, real q) {
    return cplx(fma_(acc.re, y.re, fma_(-acc.im, y.im, q)), fma_(acc.re, y.im, acc.im * y.re));
}

// atan2(y, x) for arguments on scale (neither both zero nor near the ends of the exponent range): the quotient
// min/max through the branch-free reciprocal, one reduction at tan(pi/8) and a degree-13 polynomial in r^2
// (<= 1 ulp of pi/4 absolute, <= 2 ulp relative near 0); quadrant by selection.  The library form on the host.
SPB_HD real atan2_onscale(real y, real x) {
#if defined(__CUDA_ARCH__)
    const double ax = fabs(x), ay = fabs(y);
    const double mx = ax > ay ? ax : ay, mn = ax > ay ? ay : ax;
    double a = mn * rcp_onscale(mx);  // [0, 1]
    const bool hi = a > 0.41421356237309503;
    // (a - 1)/(a + 1) in [-0.4143, 0] for a in (tan(pi/8), 1]
    const double r = hi ? (a - 1.0) * rcp_onscale(a + 1.0) : a;
    const double u = r * r;
    double p = dby::ATANC[0];
#pragma unroll
    for (int k = 1; k < 14; ++k) p = fma(p, u, dby::ATANC[k]);
    double t = fma(r * u, p, r);
    t = hi ? t + 0.78539816339744828 + 3.061616997868383e-17 : t;  // + pi/4 (head + tail)
    t = ay > ax ? 1.5707963267948966 - t : t;
    t = x < 0.0 ? 3.141592653589793 - t : t;
    return y < 0.0 ? -t : t;
#else
    return atan2(y, x);
#endif
}

// Is (w, v) inside the calibrated region?  Squares only: m = |v^2 + w^2|^2 = |s|^4;
//   |s| >= S0  <=>  m >= S0^4 ;   |s|^3 / v^2 >= P0  <=>  m^3 >= (P0 v^2)^4
SPB_HD bool debye_ok(cplx w, real fnu) {
    // (straight-line: the classifier evaluates it for every point of the iterative regions)
    const real v2 = fnu * fnu;
    const real sr = v2 + (w.re - w.im) * (w.re + w.im), si = real(2.0) * w.re * w.im;
    const real m = sr * sr + si * si;
    const real s04 = real(SPB_DEBYE_S0 * SPB_DEBYE_S0 * SPB_DEBYE_S0 * SPB_DEBYE_S0);
    real q = real(SPB_DEBYE_P0) * v2;
    q = q * q;
    return (fnu > real(1.0)) & (m >= s04) & (m < real(1.0e60)) & (m * m * m >= q * q);
}

// need_i / need_k: which of the two functions the caller wants (the other output is left untouched).
// kode 1: I_v(w), K_v(w); kode 2: e^{-Re w} I_v(w), e^{w} K_v(w).  Returns 0, or < 0 when the point has to go to
// the general path (exponent off the middle scaling level).
//
// Series.  sum_{k<=14} P_k(t^2)/s^k with t^2 = v^2/s^2 is regrouped by powers of 1/s (tools/gen_tables.py,
// gen_debye_series):  sum_m Q_m(v^2)/s^m  with REAL polynomials Q_m of the real variable v^2 (119 coefficients in
// all), summed as two Horner chains in y = 1/s^2 (even m, odd m).  Even + odd is the series of I, even - odd the
// alternating series of K.  Against the term-by-term form (complex Horner in t^2 per term, convergence exit): a
// third of the arithmetic, no data-dependent trip count (the lanes of a warp never part), and independent chains
// for the pipeline to overlap.  All 14 terms are always summed; inside the calibrated region they decrease
// monotonically, the 14th is below 3e-14.
SPB_FN int debye_ik(cplx w, real fnu, int kode, bool need_i, bool need_k, cplx &ival, cplx &kval, real az,
                    const OrderPhase *oph = nullptr) {
    using dby::DQ;
    const real alim = SPB_ALIM;
    const real rt2pi = 0.398942280401432678;   // 1/sqrt(2 pi)
    const real rthpi = 1.25331413731550025;    // sqrt(pi/2)
    (void)az;
    const real v2 = fnu * fnu;
    // s^2 = v^2 + w^2, s = sqrt(s^2) with Re s >= 0, 1/s = conj(s)/|s|^2, 1/sqrt(s): selections, no branches
    const cplx s2(v2 + (w.re - w.im) * (w.re + w.im), real(2.0) * w.re * w.im);
    cplx s, rs, phi;
    {
#if defined(__CUDA_ARCH__)
        const double m2 = fma(s2.re, s2.re, s2.im * s2.im);       // |s|^4, on scale (>= S0^4)
        const double mi = rsqrt_onscale(m2);
        double d = m2 * mi;                                       // |s|^2
        d = fma(fma(-d, d, m2), 0.5 * mi, d);
        const double rd = fma(mi, fma(-d, mi, 1.0), mi);          // 1/|s|^2
        const double tq = 0.5 * (d + fabs(s2.re));
        const double ri = rsqrt_onscale(tq);
        double r = tq * ri;
        r = fma(fma(-r, r, tq), 0.5 * ri, r);                     // sqrt((|s^2| + |Re s^2|)/2)
        const double q = fabs(s2.im) * (0.5 * ri);                // |Im s^2| / (2 r)
        // Re s^2 >= 0: s = (r, +-q); Re s^2 < 0: s = (q, +-r); sign of Im s = sign of Im s^2, on the branch cut
        // (Im s^2 = 0, real arguments of J / Y / H beyond the turning point) the sign of Im w: s ~ w for large |w|
        const bool pos = s2.re >= 0.0;
        const bool neg_im = (s2.im < 0.0) || (s2.im == 0.0 && w.im < 0.0);
        const double sre = pos ? r : q, sim = pos ? q : r;
        s = cplx(sre, neg_im ? -sim : sim);
        rs = cplx(s.re * rd, -s.im * rd);
        // 1/sqrt(s) = conj(sqrt s)/|s|,  sqrt s = (g, Im s/(2g)),  g = sqrt((|s| + Re s)/2),  |s| = sqrt(d)
        const double di = rsqrt_onscale(d);                       // 1/|s|
        double as = d * di;
        as = fma(fma(-as, as, d), 0.5 * di, as);                  // |s|
        const double tg = 0.5 * (as + s.re);
        const double gi = rsqrt_onscale(tg);
        double g = tg * gi;
        g = fma(fma(-g, g, tg), 0.5 * gi, g);
        phi = cplx(g * di, -(s.im * (0.5 * gi)) * di);
#else
        const real d = cabs_(s2);  // |s|^2
        s = csqrt_mod(s2, d);
        if (s.re == real(0.0) && w.im < real(0.0)) s.im = -rabs(s.im);  // branch cut: see the device form
        const real rd = real(1.0) / d;
        rs = cplx(s.re * rd, -s.im * rd);
        phi = csqrt_(rs);
#endif
    }
    const cplx y = rs * rs;
#include "spb_debye_series.inc"
    const cplx sp = ev + od, sm = ev - od;
    // exponent  E = v eta - w = v^2 / (s + w) - v ln((v + s) / w):  s - w without cancellation (Re s, Re w >= 0);
    // ln|(v + s)/w| from the squared moduli, arg from (v + s) conj(w)
    const cplx den = s + w;
    const real rdd = rcp_fast(den.re * den.re + den.im * den.im);
    const cplx e1(v2 * den.re * rdd, -v2 * den.im * rdd);
    const cplx sv(s.re + fnu, s.im);
    const real w2 = w.re * w.re + w.im * w.im;
    const real lnm = real(0.5) * log_((sv.re * sv.re + sv.im * sv.im) * rcp_fast(w2));
    const real ang = atan2_onscale(sv.im * w.re - sv.re * w.im, sv.re * w.re + sv.im * w.im);
    const cplx e(e1.re - fnu * lnm, e1.im - fnu * ang);
    // exponentials: kode 2 carries e^{E.re} (I) and e^{-E.re} (K); kode 1 e^{+-(E.re + Re w)}
    const real tt = (kode == 2) ? e.re : e.re + w.re;
    if (!(rabs(tt) < alim)) return -9;  // leaves the middle scaling level: general path
    real ep, em;
    exp_pair(tt, ep, em);
    real se, ce, sw, cw;
    sincos_(e.im, se, ce);
    sincos_(w.im, sw, cw);
    // e^{i (E.im + Im w)}: phase of the dominant term of I in both scalings
    const cplx ph_i(ce * cw - se * sw, se * cw + ce * sw);
    if (need_k) {
        // e^{-i E.im} (kode 2: e^{w} K) or e^{-i (E.im + Im w)} (kode 1)
        const cplx ph_k = (kode == 2) ? cplx(ce, -se) : conj(ph_i);
        kval = (phi * rthpi) * (sm * (ph_k * em));
    }
    if (need_i) {
        cplx acc = sp * (ph_i * ep);
        // second exponential: +-i e^{+-i pi v} e^{-(E + w)} sum_minus  (times e^{-Re w} for kode 2); relative size
        // e^{-2 Re(v eta)} = e^{-2 (E.re + Re w)}: dropped below 1e-17
        const real reta = e.re + w.re;
        if (w.im != real(0.0) && reta < real(40.0)) {
            OrderPhase local;
            if (!oph) {
                local = order_phase(fnu);
                oph = &local;
            }
            // e^{+-i pi v} = (-1)^inu (cp +- i sp)
            const bool up = w.im > real(0.0);
            cplx rot(oph->cp, up ? oph->sp : -oph->sp);
            if (oph->inu & 1) rot = -rot;
            // +-i
            rot = up ? mul_i(rot) : mul_mi(rot);
            const real f2 = (kode == 2) ? em * exp_(real(-2.0) * w.re) : em;
            acc = acc + sm * ((rot * conj(ph_i)) * f2);
        }
        ival = (phi * rt2pi) * acc;
    }
    return 0;
}

}  // namespace spb
