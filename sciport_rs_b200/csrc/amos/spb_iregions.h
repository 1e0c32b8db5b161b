// spb_iregions.h — the four "elementary" regions of the modified Bessel
// function I_fnu(z), Re z >= 0 (TOMS 644 region logic):
//   seri  : ascending power series                (|z| small or order large)
//   asyi  : large-|z| Hankel-type asymptotic      (|z| >= rl)
//   mlri  : Miller backward recurrence, normalised by a Neumann series
//   rati  : ratios I_{v+1}/I_v by backward continued fraction (for the
//           Wronskian normalisation in spb_wrsk)
// All routines fill y[0..n-1] with I_{fnu+k}(z) (kode=1) or
// exp(-|Re z|) I_{fnu+k}(z) (kode=2) and return the AMOS "nz" convention:
// >=0 number of underflowed leading members, <0 failure / hand-over code.
#pragma once
#include "spb_common.h"
#include "spb_gamln.h"

namespace spb {

// ---------------------------------------------------------------------------
// Power series.  nz = -k means: k members underflowed and |z|^2/4 exceeds the
// remaining order, so the caller must finish with another method.
template <class Y>
SPB_FN int seri(cplx z, real fnu, int kode, int n, Y y, real tol, real elim, real alim,
                real az_known = real(-1.0), const real *gl = nullptr) {
    int nz = 0;
    real az = az_known >= real(0.0) ? az_known : cabs_(z);
    if (az == real(0.0)) {
        y[0] = cplx(fnu == real(0.0) ? 1.0 : 0.0, 0.0);
        for (int i = 1; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return 0;
    }
    const real arm = real(1.0e3) * SPB_D1MACH1;
    const real rtr1 = sqrt(arm);
    real crscr = 1.0;
    int iflag = 0;
    if (az < arm) {
        nz = n;
        if (fnu == real(0.0)) nz -= 1;
        y[0] = cplx(fnu == real(0.0) ? 1.0 : 0.0, 0.0);
        for (int i = 1; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return nz;
    }
    cplx hz = z * real(0.5);
    cplx cz(0.0, 0.0);
    if (az > rtr1) cz = hz * hz;
    real acz = cabs_(cz);
    int nn = n;
    cplx ck = clog_(hz);
    cplx w[2];
    real ss = 1.0, ascle = 0.0;
    for (;;) {  // restart point after dropping an underflowed top member
        real dfnu = fnu + real(nn - 1);
        real fnup = dfnu + real(1.0);
        // underflow test on the leading factor (z/2)^v / Gamma(v+1)
        cplx ak1 = ck * dfnu;
        real ak = (gl && nn == 1) ? gl[0] : gamln(fnup);  // (nn == 1: fnup = fnu + 1, see order_gln)
        ak1.re -= ak;
        if (kode == 2) ak1.re -= z.re;
        bool underflow = !(ak1.re > -elim);
        if (!underflow) {
            if (!(ak1.re > -alim)) {
                iflag = 1;
                ss = real(1.0) / tol;
                crscr = tol;
                ascle = arm * ss;
            }
            real aa = exp_(ak1.re);
            if (iflag == 1) aa *= ss;
            cplx coef = cis_(ak1.im) * aa;
            real atol = tol * acz / fnup;
            int il = nn < 2 ? nn : 2;
            for (int i = 1; i <= il; ++i) {
                dfnu = fnu + real(nn - i);
                fnup = dfnu + real(1.0);
                cplx s1(1.0, 0.0);
                if (!(acz < tol * fnup)) {
                    cplx t(1.0, 0.0);
                    real akk = fnup + real(2.0);
                    real s = fnup;
                    real a2 = 2.0;
                    do {
                        real rs = rcp_fast(s);  // s = k (fnu + k), k >= 1
                        t = (t * cz) * rs;
                        s1 = s1 + t;
                        s += akk;
                        akk += real(2.0);
                        a2 = a2 * acz * rs;
                    } while (a2 > atol);
                }
                cplx s2 = s1 * coef;
                w[i - 1] = s2;
                if (iflag != 0 && uchk(s2, ascle, tol) != 0) {
                    underflow = true;
                    break;
                }
                y[nn - i] = s2 * crscr;
                if (i != il) coef = cdiv(coef, hz) * dfnu;
            }
        }
        if (underflow) {
            nz += 1;
            y[nn - 1] = cplx(0.0, 0.0);
            if (acz > dfnu) return -nz;
            nn -= 1;
            if (nn == 0) return nz;
            continue;
        }
        break;
    }
    if (nn <= 2) return nz;
    // backward recurrence for the remaining members of the sequence
    int k = nn - 2;  // 1-based index of the next member to fill
    real ak = real(k);
    cplx rz = crecip(z) * real(2.0);
    int ib = 3;
    if (iflag == 1) {
        // recur on scaled values until they are back on scale
        cplx s1 = w[0], s2 = w[1];
        int l;
        bool onscale = false;
        for (l = 3; l <= nn; ++l) {
            cplx c = s2;
            s2 = s1 + (rz * c) * (ak + fnu);
            s1 = c;
            cplx c2 = s2 * crscr;
            y[k - 1] = c2;
            ak -= real(1.0);
            k -= 1;
            if (cabs_(c2) > ascle) {
                onscale = true;
                break;
            }
        }
        if (!onscale) return nz;
        ib = l + 1;
        if (ib > nn) return nz;
    }
    for (int i = ib; i <= nn; ++i) {
        y[k - 1] = (rz * y[k]) * (ak + fnu) + y[k + 1];
        ak -= real(1.0);
        k -= 1;
    }
    return nz;
}

// ---------------------------------------------------------------------------
// Large-|z| asymptotic expansion.  nz = -1 overflow, -2 no convergence.
//
// The term ratio of the Hankel series is (4 v^2 - (2j-1)^2) / (j * 8z); the
// reciprocal 1/(8z) is formed once and the integer reciprocals 1/j come from a
// table, so the inner loop is one complex multiply and two real multiplies per
// term (no division, no modulus).  The second exponential exp(-2z) is below
// half an ulp of the leading sum once 2 Re z > 80 and is skipped there.
template <class Y>
SPB_FN int asyi(cplx z, real fnu, int kode, int n, Y y, real rl, real tol, real elim, real alim,
                real az_known = real(-1.0), const OrderPhase *oph = nullptr, const WShared *ws = nullptr,
                cplx *ksum = nullptr) {
    static const double RJ[48] = {
        0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
        1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
        1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31,
        1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41,
        1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47};
    const real rt2pi = 0.398942280401432678;  // 1/sqrt(2 pi)
    real az = az_known >= real(0.0) ? az_known : cabs_(z);
    const real arm = real(1.0e3) * SPB_D1MACH1;
    const real rtr1 = sqrt(arm);
    int il = n < 2 ? n : 2;
    real dfnu = fnu + real(n - il);
    // 1/z through the modulus (on scale for any |z| >= rl); ws: the quantities shared with the K routine
    real raz = ws ? ws->raz : real(1.0) / az;
    cplx rcz = ws ? ws->rcz : w_recip(z, raz);
    // leading factor sqrt(1/(2 pi z)) * exp(z)
    cplx ak1 = (ws ? ws->rsq : w_rsqrt(z, az, raz)) * rt2pi;
    cplx cz = z;
    if (kode == 2) cz.re = 0.0;
    if (rabs(cz.re) > elim) return -1;
    real sny, csy;  // sin, cos of Im z: phase of exp(z) and, doubled, of the second exponential
    if (ws) {
        sny = ws->sn;
        csy = ws->cs;
    } else {
        sincos_(z.im, sny, csy);
    }
    real dnu2 = dfnu + dfnu;
    int koded = 1;
    if (!(rabs(cz.re) > alim && n > 2)) {
        koded = 0;
        // (unscaled: exp(Re z), which the shared block has formed together with the exp(-Re z) of the K routine)
        real zm = (ws && ws->have_em && kode != 2) ? ws->ep : exp_(cz.re);
        ak1 = ak1 * cplx(zm * csy, zm * sny);
    }
    real fdn = 0.0;
    if (dnu2 > rtr1) fdn = dnu2 * dnu2;
    // for imaginary z the error test is made relative to the first reciprocal
    // power, the leading term of the imaginary part
    real aez = real(8.0) * az;
    real raez = real(0.125) * raz;
    cplx rez = rcz * real(0.125);  // 1/(8z)
    real s = tol * raez;
    int jl = (int)(rl + rl) + 2;
    cplx p1(0.0, 0.0);
    const bool second = (z.im != real(0.0)) && (z.re + z.re < real(80.0));
    cplx e2(0.0, 0.0);
    if (z.im != real(0.0)) {
        // exp(i pi (1/2 + fnu + n - il)) with the integer part folded out
        // sin, cos of pi (fnu - inu) come from the order phase (one sincos per distinct order, shared with the
        // rotations of the combine step)
        int inu = (int)fnu;
        OrderPhase local;
        if (!oph) {
            local = order_phase(fnu);
            oph = &local;
        }
        inu = inu + n - il;
        real sn = oph->sp, cs = oph->cp;
        real ak = -sn;
        real bk = cs;
        if (z.im < real(0.0)) bk = -bk;
        p1 = cplx(ak, bk);
        if (inu % 2 != 0) p1 = -p1;
        if (second) {
            // exp(-2z) = exp(-Re z)^2 (cos 2y - i sin 2y): the K routine needs the same exp(-Re z)
            real em = (ws && ws->have_em) ? ws->em : exp_(-z.re);
            real e2m = em * em;
            e2 = cplx(e2m * ((csy - sny) * (csy + sny)), -e2m * ((sny + sny) * csy));
        }
    }
    bool noconv = false;
    for (int k = 1; k <= il; ++k) {
        real sqk = fdn - real(1.0);
        real atol = s * rabs(sqk);
        real sgn = 1.0;
        cplx cs1(1.0, 0.0), cs2(1.0, 0.0), ck(1.0, 0.0);
        real ak = 0.0, aa = 1.0;
        bool conv = false;
        for (int j = 1; j <= jl; ++j) {
            real t = sqk * real(RJ[j]);
            ck = (ck * rez) * t;
            cs2 = cs2 + ck;
            sgn = -sgn;
            cs1 = cs1 + ck * sgn;
            aa = aa * (rabs(t) * raez);
            ak += real(8.0);
            sqk -= ak;
            if (aa <= atol) {
                conv = true;
                break;
            }
        }
        // (no return here: a `return` right after the data-dependent loop would move the reconvergence point of its
        // exits to the end of the routine, and the lanes of a warp would run the rest of it in as many passes as
        // they have distinct series lengths -- measured: the tail below ran with 6 of 32 lanes)
        noconv = noconv || !conv;
        // the same terms with all signs positive are the large-|z| series of K (member of order fnu + n - 1 last)
        if (ksum) *ksum = cs2;
        cplx s2 = cs1;
        if (second) s2 = s2 + (e2 * p1) * cs2;
        fdn = fdn + real(8.0) * dfnu + real(4.0);
        p1 = -p1;
        y[n - il + k - 1] = s2 * ak1;
    }
    if (noconv) {
        // a series that did not converge: no member is returned (as when the routine left at the failing series)
        for (int i = 0; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return -2;
    }
    if (n <= 2) return 0;
    int nn = n;
    int k = nn - 2;
    real ak = real(k);
    cplx rz = rcz * real(2.0);
    for (int i = 3; i <= nn; ++i) {
        y[k - 1] = (rz * y[k]) * (ak + fnu) + y[k + 1];
        ak -= real(1.0);
        k -= 1;
    }
    if (koded == 0) return 0;
    cplx ckk = cexp_(cz);
    for (int i = 0; i < nn; ++i) y[i] = y[i] * ckk;
    return 0;
}

// ---------------------------------------------------------------------------
// `steps` steps of the backward recurrence of the Miller algorithm with the running sum of the normalising
// (Neumann) relation.  On entry and on exit p2 is the newest iterate, p1 the one before.  Two steps per round with
// the iterates swapping roles (no register moves, see rati_walk); an odd count starts with one step of the plain
// form.  Same values as the one-step loop, step for step.
SPB_HD void mlri_back(int steps, cplx &p1, cplx &p2, cplx rz, real &fkk, real fnf, real tfnf, real &bk, cplx &sum) {
    if (steps <= 0) return;
    if (steps & 1) {
        const cplx pt = p2;
        p2 = p1 + (rz * pt) * (fkk + fnf);
        p1 = pt;
        const real ac = bk * (real(1.0) - tfnf * rcp_fast(fkk + tfnf));  // 1 <= fkk + tfnf < 1e5
        sum = sum + p1 * (ac + bk);
        bk = ac;
        fkk -= real(1.0);
    }
    for (int i = steps >> 1; i > 0; --i) {
        p1 = p1 + (rz * p2) * (fkk + fnf);  // newest p2 -> newest p1
        real ac = bk * (real(1.0) - tfnf * rcp_fast(fkk + tfnf));
        sum = sum + p2 * (ac + bk);
        bk = ac;
        fkk -= real(1.0);
        p2 = p2 + (rz * p1) * (fkk + fnf);  // newest p1 -> newest p2
        ac = bk * (real(1.0) - tfnf * rcp_fast(fkk + tfnf));
        sum = sum + p1 * (ac + bk);
        bk = ac;
        fkk -= real(1.0);
    }
}

// ---------------------------------------------------------------------------
// Miller algorithm normalised by a Neumann series.  nz = -2: no convergence.
template <class Y>
SPB_FN int mlri(cplx z, real fnu, int kode, int n, Y y, real tol, real az_known = real(-1.0),
                const real *gl = nullptr) {
    const real scle = SPB_D1MACH1 / tol;
    real az = az_known >= real(0.0) ? az_known : cabs_(z);
    int iaz = (int)az;
    int ifnu = (int)fnu;
    int inu = ifnu + n - 1;
    real at = real(iaz) + real(1.0);
    cplx rcz = crecip(z);
    cplx ck = rcz * at;
    cplx rz = rcz * real(2.0);
    real raz = real(1.0) / az;
    cplx p1(0.0, 0.0), p2(1.0, 0.0);
    real ack = (at + real(1.0)) * raz;
    real rho = ack + sqrt(ack * ack - real(1.0));
    real rho2 = rho * rho;
    real tst = (rho2 + rho2) / ((rho2 - real(1.0)) * (rho - real(1.0)));
    tst = tst / tol;
    // relative truncation error index for the series
    real ak = at;
    int i;
    bool found = false;
    for (i = 1; i <= 80; ++i) {
        cplx pt = p2;
        p2 = p1 - ck * pt;
        p1 = pt;
        ck = ck + rz;
        // |p2| > tst ak^2 compared on the squares (magnitudes stay below 1e25 here): no modulus per step
        real ap2 = p2.re * p2.re + p2.im * p2.im;
        real th = tst * ak * ak;
        if (ap2 > th * th) {
            found = true;
            break;
        }
        ak += real(1.0);
    }
    // No early return between the data-dependent loops: a `return` here would move the reconvergence point of
    // the loop exits to the end of the routine, and lanes leaving the walk at different indices would run
    // the (long) backward recurrence below in separate passes.  A failed walk is carried in `fail` instead.
    bool fail = !found;
    if (fail) i = 80;
    i += 1;
    int k = 0;
    if (inu >= iaz) {
        // relative truncation error for the ratios
        p1 = cplx(0.0, 0.0);
        p2 = cplx(1.0, 0.0);
        at = real(inu) + real(1.0);
        ck = rcz * at;
        ack = at * raz;
        tst = sqrt(ack / tol);
        // two loops instead of one with a "second time" flag (see rati): the refinement of the test value runs
        // once per point with the warp reconverged
        found = false;
        bool first = false;
        for (k = 1; k <= 80; ++k) {
            cplx pt = p2;
            p2 = p1 - ck * pt;
            p1 = pt;
            ck = ck + rz;
            real ap2 = p2.re * p2.re + p2.im * p2.im;
            if (ap2 < tst * tst) continue;
            first = true;
            break;
        }
        if (first) {
            real ap = sqrt(p2.re * p2.re + p2.im * p2.im);
            ack = cabs_(ck);
            real flam = ack + sqrt(ack * ack - real(1.0));
            real fkap = ap / cabs_(p1);
            rho = rmin(flam, fkap);
            tst = tst * sqrt(rho / (rho * rho - real(1.0)));
            for (k = k + 1; k <= 80; ++k) {
                cplx pt = p2;
                p2 = p1 - ck * pt;
                p1 = pt;
                ck = ck + rz;
                real ap2 = p2.re * p2.re + p2.im * p2.im;
                if (ap2 < tst * tst) continue;
                found = true;
                break;
            }
        }
        if (!found) {
            fail = true;
            k = 80;
        }
    }
    // backward recurrence and sum of the normalising relation
    k += 1;
    int kk = (i + iaz) > (k + inu) ? (i + iaz) : (k + inu);
    real fkk = real(kk);
    p1 = cplx(0.0, 0.0);
    p2 = cplx(scle, 0.0);  // scaled start value
    real fnf = fnu - real(ifnu);
    real tfnf = fnf + fnf;
    real bk = gamln(fkk + tfnf + real(1.0)) - gamln(fkk + real(1.0)) - (gl ? gl[2] : gamln(tfnf + real(1.0)));
    bk = exp_(bk);
    cplx sum(0.0, 0.0);
    int km = kk - inu;
    mlri_back(km, p1, p2, rz, fkk, fnf, tfnf, bk, sum);
    y[n - 1] = p2;
    for (int j = 2; j <= n; ++j) {
        cplx pt = p2;
        p2 = p1 + (rz * pt) * (fkk + fnf);
        p1 = pt;
        real a = real(1.0) - tfnf * rcp_fast(fkk + tfnf);  // 1 <= fkk + tfnf < 1e5
        real ac = bk * a;
        sum = sum + p1 * (ac + bk);
        bk = ac;
        fkk -= real(1.0);
        y[n - j] = p2;
    }
    mlri_back(ifnu, p1, p2, rz, fkk, fnf, tfnf, bk, sum);
    cplx pt = z;
    if (kode == 2) pt.re = 0.0;
    cplx lrz = clog_(rz);
    cplx q = pt - lrz * fnf;
    real ap = gl ? gl[1] : gamln(real(1.0) + fnf);
    q.re -= ap;
    // exp(q)/(sum+p2) without squaring large magnitudes
    p2 = p2 + sum;
    ap = cabs_(p2);
    real rap = real(1.0) / ap;
    cplx e = cexp_(q) * rap;
    cplx cnorm = e * (conj(p2) * rap);
    if (fail) {
        for (int j = 0; j < n; ++j) y[j] = cplx(0.0, 0.0);
        return -2;
    }
    for (int j = 0; j < n; ++j) y[j] = y[j] * cnorm;
    return 0;
}

// ---------------------------------------------------------------------------
// Forward magnitude walk of the ratio algorithm: p_next = p1 - t1 p2, t1 += rz, until the squared modulus of the
// iterate before the newest exceeds `tests`.  On entry and on exit p2 is the newest iterate, p1 the one before it,
// ap2s = |p2|^2 and ap1s = |p1|^2.  Two steps per round with the iterates swapping roles: the rotation
// (pt = p2, p2 = new, p1 = pt) of the one-step form was a fifth of the instructions of the Wronskian kernel as
// register moves; the values are those of the one-step loop, step for step.
SPB_HD void rati_walk(cplx &p1, cplx &p2, cplx &t1, cplx rz, real &ap1s, real &ap2s, real tests, int &k) {
    for (;;) {
        // newest p2 -> newest p1
        k += 1;
        p1 = p1 - t1 * p2;
        t1 = t1 + rz;
        const real a1 = p1.re * p1.re + p1.im * p1.im;
        if (ap2s > tests) {
            // leave with the roles restored: (previous, newest) = (p2, p1)
            const cplx pt = p1;
            p1 = p2;
            p2 = pt;
            ap1s = ap2s;
            ap2s = a1;
            return;
        }
        // newest p1 -> newest p2
        k += 1;
        p2 = p2 - t1 * p1;
        t1 = t1 + rz;
        ap1s = a1;
        ap2s = p2.re * p2.re + p2.im * p2.im;
        if (ap1s > tests) return;
    }
}

// ---------------------------------------------------------------------------
// Ratios cy[k] = I_{fnu+k+1}(z)/I_{fnu+k}(z), k = 0..n-1, by backward
// recurrence started from an index chosen by a forward magnitude walk.
template <class Y>
SPB_FN void rati(cplx z, real fnu, int n, Y cy, real tol, real az_known = real(-1.0)) {
    const real rt2 = 1.41421356237309515;
    real az = az_known >= real(0.0) ? az_known : cabs_(z);
    int inu = (int)fnu;
    int idnu = inu + n - 1;
    int magz = (int)az;
    real amagz = real(magz + 1);
    real fdnu = real(idnu);
    real fnup = rmax(amagz, fdnu);
    int id = idnu - magz - 1;
    int k = 1;
    cplx rz = crecip(z) * real(2.0);
    cplx t1 = rz * fnup;
    cplx p2 = -t1;
    cplx p1(1.0, 0.0);
    t1 = t1 + rz;
    if (id > 0) id = 0;
    real ap2 = cabs_(p2);
    real ap1 = cabs_(p1);
    // scale test values by ap1 so that an overflow does not occur prematurely
    real arg = (ap2 + ap2) / (ap1 * tol);
    real test1 = sqrt(arg);
    real test = test1;
    real rap1 = real(1.0) / ap1;
    p1 = p1 * rap1;
    p2 = p2 * rap1;
    ap2 = ap2 * rap1;
    // the walk compares squared moduli (values stay below 1e20 here): one square root only when a test fires.
    // Two loops instead of one loop with a "second time" flag: the refinement of the test value (four square roots,
    // two divisions) runs once per point, and inside a single loop every lane of a warp would reach it at its own
    // iteration -- the block then executes once per DISTINCT iteration index in the warp instead of once (measured:
    // a fifth of the instructions of the Wronskian kernel).  The exits of the first loop reconverge before it.
    real ap2s = ap2 * ap2, ap1s, tests = test * test;
    rati_walk(p1, p2, t1, rz, ap1s, ap2s, tests, k);
    {
        real ak = cabs_(t1) * real(0.5);
        real flam = ak + sqrt(ak * ak - real(1.0));
        real rho = rmin(sqrt(ap2s / ap1s), flam);
        test = test1 * sqrt(rho / (rho * rho - real(1.0)));
        tests = test * test;
    }
    rati_walk(p1, p2, t1, rz, ap1s, ap2s, tests, k);
    ap2 = sqrt(ap2s);
    int kk = k + 1 - id;
    real ak = real(kk);
    real t1r = ak;
    real dfnu = fnu + real(n - 1);
    p1 = cplx(real(1.0) / ap2, 0.0);
    p2 = cplx(0.0, 0.0);
    // two steps per round, the iterates swapping roles (no register moves); an odd step count starts with one step
    if (kk & 1) {
        const cplx pt = p1;
        p1 = pt * (rz * (dfnu + t1r)) + p2;
        p2 = pt;
        t1r -= real(1.0);
    }
    for (int i = kk >> 1; i > 0; --i) {
        p2 = p1 * (rz * (dfnu + t1r)) + p2;
        t1r -= real(1.0);
        p1 = p2 * (rz * (dfnu + t1r)) + p1;
        t1r -= real(1.0);
    }
    if (p1.re == real(0.0) && p1.im == real(0.0)) p1 = cplx(tol, tol);
    cy[n - 1] = cdiv(p2, p1);
    if (n == 1) return;
    k = n - 1;
    ak = real(k);
    t1r = ak;
    cplx cdfnu = rz * fnu;
    for (int i = 2; i <= n; ++i) {
        cplx pt = cdfnu + rz * t1r + cy[k];
        ak = cabs_(pt);
        if (ak == real(0.0)) {
            pt = cplx(tol, tol);
            ak = tol * rt2;
        }
        real rak = real(1.0) / ak;
        cy[k - 1] = conj(pt) * rak * rak;
        t1r -= real(1.0);
        k -= 1;
    }
}

}  // namespace spb
