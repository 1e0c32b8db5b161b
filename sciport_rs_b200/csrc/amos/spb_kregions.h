// spb_kregions.h — modified Bessel function K_fnu(z) in the right half plane
// (TOMS 644 region logic, SURVEY.md A.3 "ZBKNU"):
//   |z| <= 2           : Temme-type series for K_dnu, K_dnu+1, |dnu| <= 1/2
//   |z| >  2, |dnu|=1/2: closed form sqrt(pi/2z) e^-z
//   |z| >  2           : Miller algorithm for K (trip count from the cos(pi dnu) test)
//   then forward recurrence in the order under three-level scaling.
// plus the helpers for underflow handling (kscl, s1s2).
#pragma once
#include "spb_common.h"
#include "spb_gamln.h"

namespace spb {

SPB_HD void sinhcosh(cplx z, cplx &sh, cplx &ch) {
    real shr = sinh(z.re), chr = cosh(z.re);
    real sn = sin(z.im), cn = cos(z.im);
    sh = cplx(shr * cn, chr * sn);
    ch = cplx(chr * cn, shr * sn);
}

// Set K members to zero on underflow, carry the recurrence on scaled values
// until two consecutive members are on scale; returns nz, leaves the on-scale
// members in y scaled by 1/tol.
template <class Y>
SPB_FN int kscl(cplx zr, real fnu, int n, Y y, cplx rz, real ascle, real tol, real elim) {
    int nz = 0, ic = 0;
    int nn = n < 2 ? n : 2;
    cplx cy[2];
    for (int i = 1; i <= nn; ++i) {
        cplx s1 = y[i - 1];
        cy[i - 1] = s1;
        real as = cabs_(s1);
        real acs = -zr.re + log(as);
        nz += 1;
        y[i - 1] = cplx(0.0, 0.0);
        if (acs < -elim) continue;
        cplx cs = clog_(s1) - zr;
        real str = exp(cs.re) / tol;
        cs = cplx(str * cos(cs.im), str * sin(cs.im));
        if (uchk(cs, ascle, tol) != 0) continue;
        y[i - 1] = cs;
        ic = i;
        nz -= 1;
    }
    if (n == 1) return nz;
    if (ic <= 1) {
        y[0] = cplx(0.0, 0.0);
        nz = 2;
    }
    if (n == 2) return nz;
    if (nz == 0) return nz;
    real fn = fnu + real(1.0);
    cplx ck = rz * fn;
    cplx s1 = cy[0], s2 = cy[1];
    real helim = real(0.5) * elim;
    real elm = exp(-elim);
    real celmr = elm;
    cplx zd = zr;
    // find two consecutive on-scale values; rescale the recurrence if s2 grows beyond exp(elim/2)
    int kk = 0;
    bool two = false;
    for (int i = 3; i <= n; ++i) {
        kk = i;
        cplx cs = s2;
        s2 = ck * cs + s1;
        s1 = cs;
        ck = ck + rz;
        real as = cabs_(s2);
        real alas = log(as);
        real acs = -zd.re + alas;
        nz += 1;
        y[i - 1] = cplx(0.0, 0.0);
        bool bad = (acs < -elim);
        if (!bad) {
            cs = clog_(s2) - zd;
            real str = exp(cs.re) / tol;
            cs = cplx(str * cos(cs.im), str * sin(cs.im));
            if (uchk(cs, ascle, tol) != 0) bad = true;
            if (!bad) {
                y[i - 1] = cs;
                nz -= 1;
                if (ic == kk - 1) {
                    two = true;
                    break;
                }
                ic = kk;
                continue;
            }
        }
        if (alas < helim) continue;
        zd.re -= elim;
        s1 = s1 * celmr;
        s2 = s2 * celmr;
    }
    if (two) {
        nz = kk - 2;
    } else {
        nz = n;
        if (ic == n) nz = n - 1;
    }
    for (int i = 0; i < nz; ++i) y[i] = cplx(0.0, 0.0);
    return nz;
}

// Underflow test for the sum K-term (s1) + I-term (s2) in the continuation
// formulae; for kode=2 the two can be of comparable magnitude.
SPB_FN int s1s2(cplx zr, cplx &s1, cplx &s2, real ascle, real alim, int &iuf) {
    int nz = 0;
    real as1 = cabs_(s1);
    real as2 = cabs_(s2);
    if ((s1.re != real(0.0) || s1.im != real(0.0)) && as1 != real(0.0)) {
        // s1 exp(-2 zr), dropped when ln|s1| - 2 Re zr < -alim.  The binary exponent of |s1| brackets ln|s1| within
        // ln 2, which decides the test without a logarithm except in a band of that width; the product form
        // replaces exp(log s1 - 2 zr) (one exp and one sincos instead of log, atan2, exp, sincos) wherever
        // exp(-2 Re zr) itself is on scale.
        const real x2 = zr.re + zr.re;
        const real thr = x2 - alim;
        const real ln2 = 0.69314718055994531;
        const int e = expo_((double)as1) - 1022;  // 2^(e-1) <= |s1| < 2^e (e = -1022 also for subnormals)
        bool drop;
        if (real(e) * ln2 + real(1.0e-9) < thr) {
            drop = true;
        } else if (real(e - 1) * ln2 - real(1.0e-9) >= thr) {
            drop = false;
        } else {
            drop = (-x2 + log(as1) < -alim);
        }
        cplx s1d = s1;
        s1 = cplx(0.0, 0.0);
        as1 = 0.0;
        if (!drop) {
            if (x2 < real(700.0)) {
                const real em2 = exp_(-x2);
                const cplx ph = cis_(-(zr.im + zr.im));
                s1 = s1d * cplx(em2 * ph.re, em2 * ph.im);
            } else {
                cplx c1 = clog_(s1d) - zr - zr;
                s1 = cexp_(c1);
            }
            as1 = cabs_(s1);
            iuf += 1;
        }
    }
    real aa = rmax(as1, as2);
    if (aa > ascle) return nz;
    s1 = cplx(0.0, 0.0);
    s2 = cplx(0.0, 0.0);
    nz = 1;
    iuf = 0;
    return nz;
}

// K_{fnu+k}(z), k=0..n-1, Re z >= 0.  Returns nz (>=0), -1 overflow/invalid, -2 no convergence.
// The body reports a failed start-index search of the Miller algorithm through `noconv` instead of
// returning on the spot: a `return` between the data-dependent loops would move the reconvergence point of
// their exits to the end of the routine and leave a warp split for the rest of it.
//
// SIMPLE = true is the form the binned K pass uses for points whose over/underflow pre-test is known to be a
// no-op: the value stays on the middle scaling level throughout, so the three-level machinery is compiled out
// and a point that would leave the middle level after all is reported through `handover` (the caller sends it
// to the general path).  On the middle level every scaling factor is exactly 1, so both forms round identically.
template <bool SIMPLE, class Y>
SPB_FN int bknu_body(cplx z, real fnu, int kode, int n, Y y, real tol, real elim, real alim, real az_known,
                     bool &noconv, bool &handover, const WShared *ws = nullptr) {
    const int kmax = 30;
    const real r1 = 2.0;
    const real dpi = SPB_PI;
    const real rthpi = 1.25331413731550025;  // sqrt(pi/2)
    const real spi = 1.90985931710274403;    // 6/pi
    const real hpi = SPB_HPI;
    const real fpi = 1.89769999331517738;
    const real tth = 6.66666666666666666e-01;
    // Maclaurin coefficients of the regularised difference (1/Gamma(1-d)-1/Gamma(1+d))/(2d)
    const real cc[8] = {5.77215664901532861e-01,  -4.20026350340952355e-02, -4.21977345555443367e-02,
                        7.21894324666309954e-03,  -2.15241674114950973e-04, -2.01348547807882387e-05,
                        1.13302723198169588e-06,  6.11609510448141582e-09};
    real caz = az_known >= real(0.0) ? az_known : cabs_(z);
    real csclr = real(1.0) / tol;
    real crscr = tol;
    // three-level scaling constants, selected (not indexed: a dynamically indexed array would live in local
    // memory on the device) by the level k = 1, 2, 3
    const real bry0 = real(1.0e3) * SPB_D1MACH1 / tol;
    const real bry1 = real(1.0) / bry0;
    auto CSSR = [&](int k) -> real { return k == 1 ? csclr : (k == 2 ? real(1.0) : crscr); };
    auto CSRR = [&](int k) -> real { return k == 1 ? crscr : (k == 2 ? real(1.0) : csclr); };
    auto BRY = [&](int k) -> real { return k == 1 ? bry0 : (k == 2 ? bry1 : real(SPB_D1MACH2)); };
    int nz = 0;
    int iflag = 0;
    int koded = kode;
    // ws: 1/|z|, 1/z, 1/sqrt(z), sincos(Im z), exp(-Re z) already computed for the I routine at the same point
    // (fused I+K region kernel, |z| >= rl there): 2 (1/z) is the same number either way
    real rcaz = ws ? ws->raz : real(1.0) / caz;
    cplx rz = ws ? ws->rcz * real(2.0) : cplx(z.re * rcaz, -z.im * rcaz) * (rcaz + rcaz);
    int inu = (int)(fnu + real(0.5));
    real dnu = fnu - real(inu);
    real dnu2 = 0.0;
    cplx s1, s2, ck, coef, zd, p1, p2;
    cplx cy0(0.0, 0.0), cy1(0.0, 0.0);  // two-slot buffer of the iflag = 1 path
    int kflag = 2;
    real fhs = 0.0, fk = 0.0, ak = 0.0;
    bool half = (rabs(dnu) == real(0.5));
    if (!half && rabs(dnu) > tol) dnu2 = dnu * dnu;

    int fin = 240;          // 240: plain store, 270: store through the underflow scaler
    bool skip_fwd = false;  // single member already final in s1
    bool miller = true;
    // ---------------------------------------------------------------- |z| <= 2
    if (!half && !(caz > r1)) {
        real fc = 1.0;
        cplx smu = clog_(rz);
        cplx fmu = smu * dnu;
        cplx csh, cch;
        sinhcosh(fmu, csh, cch);
        if (dnu != real(0.0)) {
            fc = dnu * dpi;
            fc = fc / sin(fc);
            smu = csh * (real(1.0) / dnu);
        }
        real a2 = real(1.0) + dnu;
        // t1 = 1/Gamma(1-dnu), t2 = 1/Gamma(1+dnu) via Gamma(1-d)Gamma(1+d) = pi d / sin(pi d)
        real t2 = exp_(-gamln(a2));
        real t1 = real(1.0) / (t2 * fc);
        real g1;
        if (rabs(dnu) > real(0.1)) {
            g1 = (t1 - t2) / (dnu + dnu);
        } else {
            real akk = 1.0;
            real s = cc[0];
            for (int k = 2; k <= 8; ++k) {
                akk = akk * dnu2;
                real tm = cc[k - 1] * akk;
                s = s + tm;
                if (rabs(tm) < tol) break;
            }
            g1 = -s;
        }
        real g2 = (t1 + t2) * real(0.5);
        cplx f = (cch * g1 + smu * g2) * fc;
        cplx st = cexp_(fmu);
        cplx p = st * (real(0.5) / t2);
        cplx q = cdiv(cplx(0.5, 0.0), st) * (real(1.0) / t1);
        s1 = f;
        s2 = p;
        real akk = 1.0, a1 = 1.0;
        cplx ckk(1.0, 0.0);
        real bk = real(1.0) - dnu2;
        if (inu == 0 && n == 1) {
            // K_fnu(z), 0 <= fnu < 1/2, single member
            if (!(caz < tol)) {
                cplx cz = (z * z) * real(0.25);
                real t1s = real(0.25) * caz * caz;
                SPB_NO_UNROLL
                do {
                    f = (f * akk + p + q) * rcp_fast(bk);
                    p = p * rcp_fast(akk - dnu);
                    q = q * rcp_fast(akk + dnu);
                    real rak = rcp_fast(akk);
                    ckk = (ckk * cz) * rak;
                    s1 = s1 + ckk * f;
                    a1 = a1 * t1s * rak;
                    bk = bk + akk + akk + real(1.0);
                    akk = akk + real(1.0);
                } while (a1 > tol);
            }
            y[0] = s1;
            if (koded != 1) y[0] = s1 * cexp_(z);
            return nz;
        }
        // K_dnu and K_dnu+1 for the forward recurrence
        if (!(caz < tol)) {
            cplx cz = (z * z) * real(0.25);
            real t1s = real(0.25) * caz * caz;
            SPB_NO_UNROLL
            do {
                f = (f * akk + p + q) * rcp_fast(bk);
                p = p * rcp_fast(akk - dnu);
                q = q * rcp_fast(akk + dnu);
                real rak = rcp_fast(akk);
                ckk = (ckk * cz) * rak;
                s1 = s1 + ckk * f;
                s2 = s2 + ckk * (p - f * akk);
                a1 = a1 * t1s * rak;
                bk = bk + akk + akk + real(1.0);
                akk = akk + real(1.0);
            } while (a1 > tol);
        }
        kflag = 2;
        a1 = fnu + real(1.0);
        real aks = a1 * rabs(smu.re);
        if (aks > alim) {
            if (SIMPLE)
                handover = true;
            else
                kflag = 3;
        }
        real str = CSSR(kflag);
        p2 = s2 * str;
        s2 = p2 * rz;
        s1 = s1 * str;
        if (koded != 1) {
            cplx e = cexp_(z);
            s1 = s1 * e;
            s2 = s2 * e;
        }
        miller = false;
    }

    if (miller) {
        // iflag=1: an underflow of exp(-z) is anticipated; proceed as kode=2 and test for
        // on-scale values during the forward recurrence
        // rthpi / sqrt(z) = rthpi conj(sqrt z) / |z|
        coef = (ws ? ws->rsq : w_rsqrt(z, caz, rcaz)) * rthpi;
        kflag = 2;
        if (koded != 2) {
            if (z.re > alim && SIMPLE) handover = true;
            if (z.re > alim && !SIMPLE) {
                koded = 2;
                iflag = 1;
                kflag = 2;
            } else {
                real str = ((ws && ws->have_em) ? ws->em : exp_(-z.re)) * CSSR(kflag);
                real sn, cs;
                if (ws) {
                    sn = ws->sn;
                    cs = ws->cs;
                } else {
                    sincos_(z.im, sn, cs);
                }
                cplx st(str * cs, -str * sn);
                coef = coef * st;
            }
        }
        bool closed = half;
        if (!closed) {
            ak = rabs(cos(dpi * dnu));
            if (ak == real(0.0)) closed = true;
        }
        if (!closed) {
            fhs = rabs(real(0.25) - dnu2);
            if (fhs == real(0.0)) closed = true;
        }
        if (closed) {
            // half-odd-integer order: K_{1/2} = K_{-1/2} = coef
            s1 = coef;
            s2 = coef;
        } else {
            // Miller algorithm for |z| > 2.  r2 = f(e) separates the forward-walk
            // estimate of the backward start index from the closed-form estimate.
            real t1 = real(52.0) * SPB_R1M5 * real(3.321928094);
            t1 = rmax(t1, real(12.0));
            t1 = rmin(t1, real(60.0));
            real t2 = tth * t1 - real(6.0);
            if (z.re == real(0.0)) {
                t1 = hpi;
            } else {
                t1 = rabs(atan(z.im / z.re));
            }
            bool have_fk = false;
            if (!(t2 > caz)) {
                // forward recurrence loop when |z| >= r2
                real etest = ak / (dpi * caz * tol);
                fk = 1.0;
                if (etest < real(1.0)) {
                    have_fk = true;
                } else {
                    real fks = 2.0;
                    real ckr = caz + caz + real(2.0);
                    real p1r = 0.0, p2r = 1.0;
                    bool ok = false;
                    SPB_NO_UNROLL
                    for (int i = 1; i <= kmax; ++i) {
                        real a = fhs * rcp_fast(fks);  // 2 <= fks, fk + 1 < 1e4
                        real cbr = ckr * rcp_fast(fk + real(1.0));
                        real ptr = p2r;
                        p2r = cbr * p2r - p1r * a;
                        p1r = ptr;
                        ckr += real(2.0);
                        fks = fks + fk + fk + real(2.0);
                        fhs = fhs + fk + fk;
                        fk += real(1.0);
                        real str = rabs(p2r) * fk;
                        if (etest < str) {
                            ok = true;
                            break;
                        }
                    }
                    if (!ok) noconv = true;
                    fk += spi * t1 * sqrt(t2 / caz);
                    fhs = rabs(real(0.25) - dnu2);
                    have_fk = true;
                }
            }
            if (!have_fk) {
                // backward index for |z| < r2
                real a2 = sqrt(caz);
                real akk = fpi * ak / (tol * sqrt(a2));
                real aa = real(3.0) * t1 / (real(1.0) + caz);
                real bb = real(14.7) * t1 / (real(28.0) + caz);
                akk = (log_(akk) + caz * cos(aa) / (real(1.0) + real(0.008) * caz)) / cos(bb);
                fk = real(0.12125) * akk * akk / caz + real(1.5);
            }
            // backward recurrence loop of the Miller algorithm
            int k = (int)fk;
            fk = real(k);
            real fks = fk * fk;
            p1 = cplx(0.0, 0.0);
            p2 = cplx(tol, 0.0);
            cplx cs = p2;
            // One step: a = (fks + fk) / (a1 + fhs) and rak = 2 / (fk + 1) from ONE reciprocal (a division costs as
            // much as a dozen multiplications on the device; the Miller iteration is self-normalising, so the two
            // extra roundings do not matter); next = (newest cb - previous) a.  Two steps per round with the
            // iterates swapping roles (the rotation pt = p2, p2 = next, p1 = pt of the one-step loop was a fifth
            // of the K pass as register moves); an odd count starts with one step.  Same values, step for step.
            auto miller_step = [&](const cplx &newest, cplx &previous) {
                const real a1 = fks - fk;
                const real d1 = a1 + fhs, d2 = fk + real(1.0);
                const real rr = rcp_fast(d1 * d2);  // 1 <= d1 d2 < 1e9
                const real a = (fks + fk) * (d2 * rr);
                const real rak = real(2.0) * (d1 * rr);
                const cplx cb = cplx(fk + z.re, z.im) * rak;
                previous = (newest * cb - previous) * a;  // (now the newest)
                cs = cs + previous;
                fks = a1 - fk + real(1.0);
                fk -= real(1.0);
            };
            if (k & 1) {
                miller_step(p2, p1);
                const cplx pt = p1;
                p1 = p2;
                p2 = pt;
            }
            SPB_NO_UNROLL
            for (int i = k >> 1; i > 0; --i) {
                miller_step(p2, p1);
                miller_step(p1, p2);
            }
            // (p2/cs) = (p2/|cs|) (conj(cs)/|cs|) for better scaling
            real tm = cabs_(cs);
            real rtm = real(1.0) / tm;
            cplx s1n = p2 * rtm;
            cplx csn = conj(cs) * rtm;
            s1 = (coef * s1n) * csn;
            if (inu == 0 && n == 1) {
                zd = z;
                fin = (iflag == 1) ? 270 : 240;
                skip_fwd = true;
            } else {
                // (p1/p2) = (p1/|p2|) (conj(p2)/|p2|) for scaling
                tm = cabs_(p2);
                rtm = real(1.0) / tm;
                cplx p1n = p1 * rtm;
                cplx p2n = conj(p2) * rtm;
                cplx pt = p1n * p2n;
                cplx st(dnu + real(0.5) - pt.re, -pt.im);
                st = cdiv(st, z);
                st.re += real(1.0);
                s2 = st * s1;
            }
        }
    }

    // ---------------------------------------------------------------- forward recurrence
    if (!skip_fwd) {
        real str = dnu + real(1.0);
        ck = rz * str;
        if (n == 1) inu -= 1;
        int inub = 1;
        zd = z;
        if (inu > 0) {
            bool regular = true;
            if (!SIMPLE && iflag == 1) {
                // iflag=1: forward recurrence on scaled values until two members are on scale
                real helim = real(0.5) * elim;
                real elm = exp(-elim);
                real celmr = elm;
                real ascle = bry0;
                int ic = -1;
                int j = 2;
                bool two = false;
                int i;
                for (i = 1; i <= inu; ++i) {
                    cplx st = s2;
                    s2 = st * ck + s1;
                    s1 = st;
                    ck = ck + rz;
                    real as = cabs_(s2);
                    real alas = log(as);
                    real p2r = -zd.re + alas;
                    bool bad = (p2r < -elim);
                    if (!bad) {
                        cplx lg = clog_(s2) - zd;
                        real p1m = exp(lg.re) / tol;
                        cplx pp(p1m * cos(lg.im), p1m * sin(lg.im));
                        if (uchk(pp, ascle, tol) != 0) bad = true;
                        if (!bad) {
                            j = 3 - j;
                            if (j == 1) cy0 = pp; else cy1 = pp;
                            if (ic == i - 1) {
                                two = true;
                                break;
                            }
                            ic = i;
                            continue;
                        }
                    }
                    if (alas < helim) continue;
                    zd.re -= elim;
                    s1 = s1 * celmr;
                    s2 = s2 * celmr;
                }
                if (!two) {
                    if (n == 1) s1 = s2;
                    fin = 270;
                    regular = false;
                } else {
                    kflag = 1;
                    inub = i + 1;
                    s2 = (j == 1) ? cy0 : cy1;
                    j = 3 - j;
                    s1 = (j == 1) ? cy0 : cy1;
                    if (inub > inu) {
                        if (n == 1) s1 = s2;
                        fin = 240;
                        regular = false;
                    }
                }
            }
            if (regular) {
                real p1r = CSRR(kflag);
                real ascle = BRY(kflag);
                if (SIMPLE) {
                    // middle scaling level throughout: the plain recurrence, two steps per round with s1 and s2
                    // swapping roles (no register moves); a value that leaves the level hands the point over
                    int left = inu - inub + 1;
                    if (left > 0 && (left & 1)) {
                        const cplx st = s2;
                        s2 = ck * st + s1;
                        s1 = st;
                        ck = ck + rz;
                        if (rmax(rabs(s2.re), rabs(s2.im)) > ascle) handover = true;
                    }
                    SPB_NO_UNROLL
                    for (int i = left >> 1; i > 0; --i) {
                        s1 = ck * s2 + s1;
                        ck = ck + rz;
                        if (rmax(rabs(s1.re), rabs(s1.im)) > ascle) handover = true;
                        s2 = ck * s1 + s2;
                        ck = ck + rz;
                        if (rmax(rabs(s2.re), rabs(s2.im)) > ascle) handover = true;
                    }
                    inub = inu + 1;  // (the general loop below has nothing left to do)
                }
                SPB_NO_UNROLL
                for (int i = inub; i <= inu; ++i) {
                    cplx st = s2;
                    s2 = ck * st + s1;
                    s1 = st;
                    ck = ck + rz;
                    if (SIMPLE) {
                        if (rmax(rabs(s2.re), rabs(s2.im)) > ascle) handover = true;
                        continue;
                    }
                    if (kflag >= 3) continue;
                    cplx pp = s2 * p1r;
                    real p2m = rmax(rabs(pp.re), rabs(pp.im));
                    if (p2m <= ascle) continue;
                    kflag += 1;
                    ascle = BRY(kflag);
                    s1 = s1 * p1r;
                    s2 = pp;
                    real sc = CSSR(kflag);
                    s1 = s1 * sc;
                    s2 = s2 * sc;
                    p1r = CSRR(kflag);
                }
                if (n == 1) s1 = s2;
                fin = 240;
            }
        } else {
            if (n == 1) s1 = s2;
            fin = (iflag == 1) ? 270 : 240;
        }
    }

    // ---------------------------------------------------------------- store / finish the sequence
    int kk;
    if (!SIMPLE && fin == 270) {
        y[0] = s1;
        if (n != 1) y[1] = s2;
        nz = kscl(zd, fnu, n, y, rz, bry0, tol, elim);
        inu = n - nz;
        if (inu <= 0) return nz;
        kk = nz + 1;
        s1 = y[kk - 1];
        y[kk - 1] = s1 * CSRR(1);
        if (inu == 1) return nz;
        kk = nz + 2;
        s2 = y[kk - 1];
        y[kk - 1] = s2 * CSRR(1);
        if (inu == 2) return nz;
        ck = rz * (fnu + real(kk - 1));
        kflag = 1;
    } else {
        real str = CSRR(kflag);
        y[0] = s1 * str;
        if (n == 1) return nz;
        y[1] = s2 * str;
        if (n == 2) return nz;
        kk = 2;
    }
    kk += 1;
    if (kk > n) return nz;
    {
        real p1r = CSRR(kflag);
        real ascle = BRY(kflag);
        for (int i = kk; i <= n; ++i) {
            cplx pp = s2;
            s2 = ck * pp + s1;
            s1 = pp;
            ck = ck + rz;
            pp = s2 * p1r;
            y[i - 1] = pp;
            if (kflag >= 3) continue;
            real p2m = rmax(rabs(pp.re), rabs(pp.im));
            if (p2m <= ascle) continue;
            kflag += 1;
            ascle = BRY(kflag);
            s1 = s1 * p1r;
            s2 = pp;
            real sc = CSSR(kflag);
            s1 = s1 * sc;
            s2 = s2 * sc;
            p1r = CSRR(kflag);
        }
    }
    return nz;
}

template <class Y>
SPB_FN int bknu(cplx z, real fnu, int kode, int n, Y y, real tol, real elim, real alim,
                real az_known = real(-1.0)) {
    bool noconv = false, handover = false;
    int nz = bknu_body<false>(z, fnu, kode, n, y, tol, elim, alim, az_known, noconv, handover);
    if (noconv) {
        for (int k = 0; k < n; ++k) y[k] = cplx(0.0, 0.0);
        return -2;
    }
    return nz;
}

// K_fnu(z), one member, for a point of the binned K pass (see bknu_body).  Returns 0, or < 0: hand the point to
// the general path (-2 no convergence, -9 the value left the middle scaling level).
SPB_FN int bknu_simple(cplx z, real fnu, int kode, cplx &val, real tol, real elim, real alim, real az_known,
                       const WShared *ws = nullptr) {
    bool noconv = false, handover = false;
    cplx y[1];
    y[0] = cplx(0.0, 0.0);
    int nz = bknu_body<true>(z, fnu, kode, 1, y, tol, elim, alim, az_known, noconv, handover, ws);
    val = y[0];
    if (noconv) return -2;
    if (handover || nz != 0) return -9;
    return 0;
}

// K_fnu(z) and K_fnu+1(z) for a binned point of the Miller + Wronskian region of I (both pre-tests known to be
// no-ops): the lean form of the routine, hand-over codes as in bknu_simple.
SPB_FN int bknu_simple2(cplx z, real fnu, int kode, cplx *cw, real tol, real elim, real alim, real az_known) {
    bool noconv = false, handover = false;
    cw[0] = cplx(0.0, 0.0);
    cw[1] = cplx(0.0, 0.0);
    int nz = bknu_body<true>(z, fnu, kode, 2, cw, tol, elim, alim, az_known, noconv, handover);
    if (noconv) return -2;
    if (handover || nz != 0) return -9;
    return 0;
}

}  // namespace spb
