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

namespace spb {

#ifndef SPB_DEBYE_S0
#define SPB_DEBYE_S0 36.0
#endif
#ifndef SPB_DEBYE_P0
#define SPB_DEBYE_P0 240.0
#endif

// Is (w, v) inside the calibrated region?  Squares only: m = |v^2 + w^2|^2 = |s|^4;
//   |s| >= S0  <=>  m >= S0^4 ;   |s|^3 / v^2 >= P0  <=>  m^3 >= (P0 v^2)^4
SPB_HD bool debye_ok(cplx w, real fnu) {
    if (!(fnu > real(1.0))) return false;
    const real v2 = fnu * fnu;
    const real sr = v2 + (w.re - w.im) * (w.re + w.im), si = real(2.0) * w.re * w.im;
    const real m = sr * sr + si * si;
    const real s04 = real(SPB_DEBYE_S0 * SPB_DEBYE_S0 * SPB_DEBYE_S0 * SPB_DEBYE_S0);
    if (!(m >= s04) || !(m < real(1.0e60))) return false;
    real q = real(SPB_DEBYE_P0) * v2;
    q = q * q;
    return m * m * m >= q * q;
}

// need_i / need_k: which of the two functions the caller wants (the other output is left untouched).
// kode 1: I_v(w), K_v(w); kode 2: e^{-Re w} I_v(w), e^{w} K_v(w).  Returns 0, or < 0 when the point has to go to
// the general path (exponent off the middle scaling level, series not converged).
SPB_FN int debye_ik(cplx w, real fnu, int kode, bool need_i, bool need_k, cplx &ival, cplx &kval, real az,
                    const OrderPhase *oph = nullptr) {
#include "spb_tables_unik.inc"
    const real tol = SPB_TOL, alim = SPB_ALIM;
    const real rt2pi = 0.398942280401432678;   // 1/sqrt(2 pi)
    const real rthpi = 1.25331413731550025;    // sqrt(pi/2)
    const real v2 = fnu * fnu;
    // s^2 = v^2 + w^2, s = sqrt(s^2) with Re s >= 0; 1/s = conj(s)/|s|^2
    const cplx s2(v2 + (w.re - w.im) * (w.re + w.im), real(2.0) * w.re * w.im);
    const real d = cabs_(s2);  // |s|^2
    cplx s = csqrt_mod(s2, d);
    // On the imaginary axis beyond the turning point (Re w = 0, |Im w| > v: real arguments of J, Y, H) s^2 is a
    // negative real number and its square root sits on the branch cut: s is the continuation of the values next to
    // the axis, where Im s^2 = 2 Re w Im w has the sign of Im w (s ~ w for large |w|)
    if (s.re == real(0.0) && w.im < real(0.0)) s.im = -rabs(s.im);
    const real rd = rcp_fast(d);
    const cplx rs(s.re * rd, -s.im * rd);
    // t^2 = v^2 / s^2 = v^2 conj(s^2) / |s|^4
    const real q = v2 * rd * rd;
    const cplx x(s2.re * q, -s2.im * q);
    // sum of P_k(t^2) / s^k, k = 1..14 (table UK: u_k coefficients highest power of t^2 first, u_0 = 1 at UK[0]);
    // both loops unrolled so that every table index is a compile-time constant; all-plus and alternating sums
    cplx sp(1.0, 0.0), sm(1.0, 0.0), pw(1.0, 0.0);
    real last = 1.0;
    int l = 1;
    SPB_UNROLL
    for (int k = 1; k <= 14; ++k) {
        cplx pk(real(UK[l]), 0.0);
        SPB_UNROLL
        for (int j = 1; j <= k; ++j) pk = pk * x + real(UK[l + j]);
        l += k + 1;
        pw = pw * rs;
        const cplx term = pw * pk;
        sp = sp + term;
        sm = (k & 1) ? sm - term : sm + term;
        last = rabs(term.re) + rabs(term.im);
        if (last < tol) break;
    }
    // Inside the calibrated region the 14th term is below 3e-14 (an asymptotic series: the sum is then as accurate
    // as its last term); anything larger means the point does not belong here
    if (!(last < real(1.0e-13))) return -2;
    // exponent  E = v eta - w = v^2 / (s + w) - v ln((v + s) / w):  s - w without cancellation (Re s, Re w >= 0)
    const real raz = rcp_fast(az);
    const cplx rw = cplx(w.re * raz, -w.im * raz) * raz;
    const cplx den = s + w;
    const real rdd = rcp_fast(den.re * den.re + den.im * den.im);
    const cplx e1(v2 * den.re * rdd, -v2 * den.im * rdd);
    const cplx zn = (s + fnu) * rw;
    const cplx lg = clog_(zn);
    const cplx e(e1.re - fnu * lg.re, e1.im - fnu * lg.im);
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
    // 1/sqrt(s) = sqrt(1/s)
    const cplx phi = csqrt_(rs);
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
