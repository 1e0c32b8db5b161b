// spb_binu.h — region driver for I_fnu(z), Re z >= 0 (TOMS 644 "ZBINU",
// SURVEY.md A.3) and the composite regions it dispatches to:
//   uoik : over/underflow pre-test from the leading uniform-asymptotic term
//   wrsk : Miller ratios (rati) normalised by the Wronskian with K (bknu)
//   uni1 / uni2 / buni : uniform asymptotic expansions for large order
#pragma once
#include "spb_common.h"
#include "spb_iregions.h"
#include "spb_kregions.h"
#include "spb_uniform.h"
#include "spb_airy.h"
#include "spb_debye.h"

namespace spb {

// ikflg = 1: I sequence (leading members may be zeroed, nuf = how many);
// ikflg = 2: K sequence; nuf = -1 signals overflow.
template <class Y>
SPB_FN int uoik(cplx z, real fnu, int kode, int ikflg, int n, Y y, real tol, real elim, real alim) {
    const real aic = 1.265512123484645396;  // ln(2 sqrt(pi))
    int nuf = 0;
    int nn = n;
    cplx zr = z;
    if (z.re < real(0.0)) zr = -z;
    cplx zb = zr;
    real yy = zb.im;
    real ax = rabs(z.re) * real(1.7321);
    real ay = rabs(yy);
    int iform = 1;
    if (ay > ax) iform = 2;
    real gnu = rmax(fnu, real(1.0));
    if (ikflg != 1) {
        real fnn = real(nn);
        real gnn = fnu + fnn - real(1.0);
        gnu = rmax(gnn, fnn);
    }
    // only magnitudes of arg and phi and the real parts of zeta1, zeta2, zb are needed
    cplx phi, arg(0.0, 0.0), zeta1, zeta2, sum, asum, bsum, cz, zn(0.0, 0.0);
    cplx cwrk[16];
    real aarg = 0.0;
    if (iform != 2) {
        int init = 0;
        unik(zr, gnu, ikflg, 1, tol, init, phi, zeta1, zeta2, sum, cwrk);
        cz = zeta2 - zeta1;
    } else {
        zn = cplx(zr.im, -zr.re);
        if (!(yy > real(0.0))) zn.re = -zn.re;
        unhj(zn, gnu, 1, tol, phi, arg, zeta1, zeta2, asum, bsum);
        cz = zeta2 - zeta1;
        aarg = cabs_(arg);
    }
    if (kode != 1) cz = cz - zb;
    if (ikflg != 1) cz = -cz;
    real aphi = cabs_(phi);
    real rcz = cz.re;
    // overflow test
    if (rcz > elim) return -1;
    bool zero_all = false;
    if (!(rcz < alim)) {
        rcz = rcz + log(aphi);
        if (iform == 2) rcz = rcz - real(0.25) * log(aarg) - aic;
        if (rcz > elim) return -1;
    } else {
        // underflow test
        if (rcz < -elim) {
            zero_all = true;
        } else if (!(rcz > -alim)) {
            rcz = rcz + log(aphi);
            if (iform == 2) rcz = rcz - real(0.25) * log(aarg) - aic;
            if (!(rcz > -elim)) {
                zero_all = true;
            } else {
                real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
                cz = cz + clog_(phi);
                if (iform != 1) {
                    cz = cz - clog_(arg) * real(0.25);
                    cz.re -= aic;
                }
                real axx = exp_(rcz) / tol;
                real ayy = cz.im;
                cplx czz = cis_(ayy) * axx;
                if (uchk(czz, ascle, tol) != 0) zero_all = true;
            }
        }
    }
    if (zero_all) {
        for (int i = 0; i < nn; ++i) y[i] = cplx(0.0, 0.0);
        return nn;
    }
    if (ikflg == 2) return nuf;
    if (n == 1) return nuf;
    // set underflows on the I sequence from the top
    for (;;) {
        gnu = fnu + real(nn - 1);
        if (iform != 2) {
            int init = 0;
            unik(zr, gnu, ikflg, 1, tol, init, phi, zeta1, zeta2, sum, cwrk);
            cz = zeta2 - zeta1;
        } else {
            unhj(zn, gnu, 1, tol, phi, arg, zeta1, zeta2, asum, bsum);
            cz = zeta2 - zeta1;
            aarg = cabs_(arg);
        }
        if (kode != 1) cz = cz - zb;
        aphi = cabs_(phi);
        rcz = cz.re;
        bool under = false;
        if (rcz < -elim) {
            under = true;
        } else {
            if (rcz > -alim) return nuf;
            rcz = rcz + log(aphi);
            if (iform == 2) rcz = rcz - real(0.25) * log(aarg) - aic;
            if (!(rcz > -elim)) {
                under = true;
            } else {
                real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
                cz = cz + clog_(phi);
                if (iform != 1) {
                    cz = cz - clog_(arg) * real(0.25);
                    cz.re -= aic;
                }
                real axx = exp_(rcz) / tol;
                real ayy = cz.im;
                cplx czz = cis_(ayy) * axx;
                if (uchk(czz, ascle, tol) != 0) under = true;
                if (!under) return nuf;
            }
        }
        y[nn - 1] = cplx(0.0, 0.0);
        nn -= 1;
        nuf += 1;
        if (nn == 0) return nuf;
    }
}

// I by ratios from rati normalised through the Wronskian with K_fnu, K_fnu+1.
// SIMPLE: the point is a binned one (both over/underflow pre-tests are no-ops), so K_fnu, K_fnu+1 stay on the
// middle scaling level and the lean form of the K routine applies (it reports a point that leaves the level
// after all, and the caller hands it to the general path); both forms round identically.
template <bool SIMPLE = false, class Y>
SPB_FN int wrsk(cplx zr, real fnu, int kode, int n, Y y, real tol, real elim, real alim,
                real az_known = real(-1.0), cplx *k_out = nullptr) {
    cplx cw[2];
    int nw = SIMPLE ? bknu_simple2(zr, fnu, kode, cw, tol, elim, alim, az_known)
                    : bknu(zr, fnu, kode, 2, cw, tol, elim, alim, az_known);
    // K_fnu(zr) itself, for callers that need K at the same point (K-type families: the K pass would compute the
    // very same value again)
    if (k_out) *k_out = cw[0];
    // a failed K evaluation is reported at the end (no return between the data-dependent loops of bknu and
    // rati, so that their exits reconverge here)
    const bool kfail = nw != 0;
    rati(zr, fnu, n, y, tol, az_known);
    // recur forward on I(fnu+1,z) = R(fnu,z) I(fnu,z)
    cplx cinu(1.0, 0.0);
    if (kode != 1) cinu = cis_(zr.im);
    // K values can sit near either exponent extreme: scale the normalisation
    real acw = cabs_(cw[1]);
    real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
    real csclr = 1.0;
    if (!(acw > ascle)) {
        csclr = real(1.0) / tol;
    } else {
        ascle = real(1.0) / ascle;
        if (!(acw < ascle)) csclr = tol;
    }
    cplx c1 = cw[0] * csclr;
    cplx c2 = cw[1] * csclr;
    cplx st = y[0];
    // cinu = cinu (conj(ct)/|ct|) (1/|ct|) avoids squaring |ct|
    cplx pt = st * c1 + c2;
    cplx ct = zr * pt;
    real act = cabs_(ct);
    real ract = real(1.0) / act;
    ct = conj(ct) * ract;
    pt = cinu * ract;
    cinu = pt * ct;
    y[0] = cinu * csclr;
    for (int i = 1; i < n; ++i) {
        cinu = st * cinu;
        st = y[i];
        y[i] = cinu * csclr;
    }
    if (kfail) {
        for (int i = 0; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return (nw == -2) ? -2 : -1;
    }
    return 0;
}

// Uniform expansion of I for |arg z| <= pi/3 (Debye type).
template <class Y>
SPB_FN int uni1(cplx z, real fnu, int kode, int n, Y y, int &nlast, real fnul, real tol, real elim,
                real alim) {
    int nz = 0;
    int nd = n;
    nlast = 0;
    // values with exponents between alim and elim are carried scaled by tol^{+-1}
    real cscl = real(1.0) / tol;
    real crsc = tol;
    // scaling levels by selection, not by indexing (a dynamically indexed array lives in local memory on the device)
    const real bry0 = real(1.0e3) * SPB_D1MACH1 / tol;
    auto CSSR = [&](int k) -> real { return k == 1 ? cscl : (k == 2 ? real(1.0) : crsc); };
    auto CSRR = [&](int k) -> real { return k == 1 ? crsc : (k == 2 ? real(1.0) : cscl); };
    auto BRY = [&](int k) -> real { return k == 1 ? bry0 : (k == 2 ? real(1.0) / bry0 : real(SPB_D1MACH2)); };
    cplx phi, zeta1, zeta2, sum, s1, s2;
    cplx cwrk[16];
    cplx cy0(0.0, 0.0), cy1(0.0, 0.0);
    // over/underflow check on the first member
    real fn = rmax(fnu, real(1.0));
    int init = 0;
    unik(z, fn, 1, 1, tol, init, phi, zeta1, zeta2, sum, cwrk, false);
    if (kode == 1) {
        s1 = zeta2 - zeta1;
    } else {
        cplx st = z + zeta2;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = st - zeta1;
    }
    real rs1 = s1.re;
    if (rabs(rs1) > elim) {
        if (rs1 > real(0.0)) return -1;
        for (int i = 0; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return n;
    }
    int iflag = 2;
    for (;;) {  // label 30
        int nn = nd < 2 ? nd : 2;
        bool redo = false;
        for (int i = 1; i <= nn; ++i) {
            fn = fnu + real(nd - i);
            init = 0;
            unik(z, fn, 1, 0, tol, init, phi, zeta1, zeta2, sum, cwrk, false);
            if (kode == 1) {
                s1 = zeta2 - zeta1;
            } else {
                cplx st = z + zeta2;
                real rast = fn / cabs_(st);
                st = conj(st) * rast * rast;
                s1 = cplx(st.re - zeta1.re, st.im - zeta1.im + z.im);
            }
            // test for underflow and overflow
            rs1 = s1.re;
            bool bad = rabs(rs1) > elim;
            if (!bad) {
                if (i == 1) iflag = 2;
                if (!(rabs(rs1) < alim)) {
                    // refine the test and scale
                    real aphi = cabs_(phi);
                    rs1 = rs1 + log_(aphi);
                    if (rabs(rs1) > elim) bad = true;
                    if (!bad) {
                        if (i == 1) iflag = 1;
                        if (rs1 >= real(0.0) && i == 1) iflag = 3;
                    }
                }
            }
            if (!bad) {
                // scale s1 if |s1| < ascle
                s2 = phi * sum;
                real e = exp_(s1.re) * CSSR(iflag);
                cplx st = cis_(s1.im) * e;
                s2 = s2 * st;
                if (iflag == 1 && uchk(s2, bry0, tol) != 0) bad = true;
            }
            if (bad) {
                // set underflow and update parameters
                if (rs1 > real(0.0)) return -1;
                y[nd - 1] = cplx(0.0, 0.0);
                nz += 1;
                nd -= 1;
                if (nd == 0) return nz;
                int nuf = uoik(z, fnu, kode, 1, nd, y, tol, elim, alim);
                if (nuf < 0) return -1;
                nd -= nuf;
                nz += nuf;
                if (nd == 0) return nz;
                fn = fnu + real(nd - 1);
                if (fn >= fnul) {
                    redo = true;
                    break;
                }
                nlast = nd;
                return nz;
            }
            if (i == 1) cy0 = s2; else cy1 = s2;
            y[nd - i] = s2 * CSRR(iflag);
        }
        if (!redo) break;
    }
    if (nd <= 2) return nz;
    cplx rz = crecip(z) * real(2.0);
    s1 = cy0;
    s2 = cy1;
    real c1r = CSRR(iflag);
    real ascle = BRY(iflag);
    int k = nd - 2;
    fn = real(k);
    for (int i = 3; i <= nd; ++i) {
        cplx c2 = s2;
        s2 = s1 + (rz * c2) * (fnu + fn);
        s1 = c2;
        c2 = s2 * c1r;
        y[k - 1] = c2;
        k -= 1;
        fn -= real(1.0);
        if (iflag >= 3) continue;
        real c2m = rmax(rabs(c2.re), rabs(c2.im));
        if (c2m <= ascle) continue;
        iflag += 1;
        ascle = BRY(iflag);
        s1 = s1 * c1r;
        s2 = c2;
        s1 = s1 * CSSR(iflag);
        s2 = s2 * CSSR(iflag);
        c1r = CSRR(iflag);
    }
    return nz;
}

// Uniform expansion of I in the rest of the right half plane, through the
// Airy-type expansion of J(fnu, zn), zn = -+ i z.
template <class Y>
SPB_FN int uni2(cplx z, real fnu, int kode, int n, Y y, int &nlast, real fnul, real tol, real elim,
                real alim) {
    const real hpi = SPB_HPI;
    const real aic = 1.265512123484645396;
    const real cipr[4] = {1.0, 0.0, -1.0, 0.0};
    const real cipi[4] = {0.0, 1.0, 0.0, -1.0};
    int nz = 0;
    int nd = n;
    nlast = 0;
    real cscl = real(1.0) / tol;
    real crsc = tol;
    // scaling levels by selection, not by indexing (a dynamically indexed array lives in local memory on the device)
    const real bry0 = real(1.0e3) * SPB_D1MACH1 / tol;
    auto CSSR = [&](int k) -> real { return k == 1 ? cscl : (k == 2 ? real(1.0) : crsc); };
    auto CSRR = [&](int k) -> real { return k == 1 ? crsc : (k == 2 ? real(1.0) : cscl); };
    auto BRY = [&](int k) -> real { return k == 1 ? bry0 : (k == 2 ? real(1.0) / bry0 : real(SPB_D1MACH2)); };
    // zn is in the right half plane after rotation by +-i
    cplx zn(z.im, -z.re);
    cplx zb = z;
    real cidi = -1.0;
    int inu = (int)fnu;
    real ang = hpi * (fnu - real(inu));
    cplx c2 = cis_(ang);
    real car = c2.re, sar = c2.im;
    int in = inu + n - 1;
    in = in % 4;
    c2 = c2 * cplx(cipr[in], cipi[in]);
    if (!(z.im > real(0.0))) {
        zn.re = -zn.re;
        zb.im = -zb.im;
        cidi = -cidi;
        c2.im = -c2.im;
    }
    cplx phi, arg, zeta1, zeta2, asum, bsum, s1, s2;
    cplx cy0(0.0, 0.0), cy1(0.0, 0.0);
    // over/underflow check on the first member
    real fn = rmax(fnu, real(1.0));
    unhj(zn, fn, 1, tol, phi, arg, zeta1, zeta2, asum, bsum);
    if (kode == 1) {
        s1 = zeta2 - zeta1;
    } else {
        cplx st = zb + zeta2;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = st - zeta1;
    }
    real rs1 = s1.re;
    if (rabs(rs1) > elim) {
        if (rs1 > real(0.0)) return -1;
        for (int i = 0; i < n; ++i) y[i] = cplx(0.0, 0.0);
        return n;
    }
    int iflag = 2;
    for (;;) {  // label 40
        int nn = nd < 2 ? nd : 2;
        bool redo = false;
        for (int i = 1; i <= nn; ++i) {
            fn = fnu + real(nd - i);
            unhj(zn, fn, 0, tol, phi, arg, zeta1, zeta2, asum, bsum);
            if (kode == 1) {
                s1 = zeta2 - zeta1;
            } else {
                cplx st = zb + zeta2;
                real rast = fn / cabs_(st);
                st = conj(st) * rast * rast;
                s1 = cplx(st.re - zeta1.re, st.im - zeta1.im + rabs(z.im));
            }
            rs1 = s1.re;
            bool bad = rabs(rs1) > elim;
            if (!bad) {
                if (i == 1) iflag = 2;
                if (!(rabs(rs1) < alim)) {
                    real aphi = cabs_(phi);
                    real aarg = cabs_(arg);
                    rs1 = rs1 + log_(aphi) - real(0.25) * log_(aarg) - aic;
                    if (rabs(rs1) > elim) bad = true;
                    if (!bad) {
                        if (i == 1) iflag = 1;
                        if (rs1 >= real(0.0) && i == 1) iflag = 3;
                    }
                }
            }
            if (!bad) {
                // scale s1 to keep intermediate arithmetic on scale near exponent extremes
                int nai, ndai, idum;
                cplx ai = airy(arg, 0, 2, nai, idum);
                cplx dai = airy(arg, 1, 2, ndai, idum);
                cplx st = dai * bsum + ai * asum;
                s2 = phi * st;
                real e = exp_(s1.re) * CSSR(iflag);
                cplx ex = cis_(s1.im) * e;
                s2 = s2 * ex;
                if (iflag == 1 && uchk(s2, bry0, tol) != 0) bad = true;
            }
            if (bad) {
                if (rs1 > real(0.0)) return -1;
                // set underflow and update parameters
                y[nd - 1] = cplx(0.0, 0.0);
                nz += 1;
                nd -= 1;
                if (nd == 0) return nz;
                int nuf = uoik(z, fnu, kode, 1, nd, y, tol, elim, alim);
                if (nuf < 0) return -1;
                nd -= nuf;
                nz += nuf;
                if (nd == 0) return nz;
                fn = fnu + real(nd - 1);
                if (fn < fnul) {
                    nlast = nd;
                    return nz;
                }
                in = (inu + nd - 1) % 4;
                c2 = cplx(car, sar) * cplx(cipr[in], cipi[in]);
                if (!(z.im > real(0.0))) c2.im = -c2.im;
                redo = true;
                break;
            }
            if (!(z.im > real(0.0))) s2.im = -s2.im;
            s2 = s2 * c2;
            if (i == 1) cy0 = s2; else cy1 = s2;
            y[nd - i] = s2 * CSRR(iflag);
            c2 = c2 * cplx(0.0, cidi);
        }
        if (!redo) break;
    }
    if (nd <= 2) return nz;
    cplx rz = crecip(z) * real(2.0);
    s1 = cy0;
    s2 = cy1;
    real c1r = CSRR(iflag);
    real ascle = BRY(iflag);
    int k = nd - 2;
    fn = real(k);
    for (int i = 3; i <= nd; ++i) {
        cplx c = s2;
        s2 = s1 + (rz * c) * (fnu + fn);
        s1 = c;
        c = s2 * c1r;
        y[k - 1] = c;
        k -= 1;
        fn -= real(1.0);
        if (iflag >= 3) continue;
        real c2m = rmax(rabs(c.re), rabs(c.im));
        if (c2m <= ascle) continue;
        iflag += 1;
        ascle = BRY(iflag);
        s1 = s1 * c1r;
        s2 = c;
        s1 = s1 * CSSR(iflag);
        s2 = s2 * CSSR(iflag);
        c1r = CSRR(iflag);
    }
    return nz;
}

// I for |z| > fnul or fnu+n-1 > fnul: evaluate at an order raised by nui and
// recur backward; nlast != 0 leaves the low orders for another method.
template <class Y>
SPB_FN int buni(cplx z, real fnu, int kode, int n, Y y, int nui, int &nlast, real fnul, real tol,
                real elim, real alim, int iform_known = 0) {
    int nz = 0;
    real ax = rabs(z.re) * real(1.7321);
    real ay = rabs(z.im);
    int iform = 1;
    if (ay > ax) iform = 2;
    if (iform_known != 0) iform = iform_known;  // region kernels compile one form only
    int nw;
    if (nui == 0) {
        if (iform == 2)
            nw = uni2(z, fnu, kode, n, y, nlast, fnul, tol, elim, alim);
        else
            nw = uni1(z, fnu, kode, n, y, nlast, fnul, tol, elim, alim);
        if (nw < 0) return (nw == -2) ? -2 : -1;
        return nw;
    }
    real fnui = real(nui);
    real dfnu = fnu + real(n - 1);
    real gnu = dfnu + fnui;
    cplx cy[2];
    if (iform == 2)
        nw = uni2(z, gnu, kode, 2, cy, nlast, fnul, tol, elim, alim);
    else
        nw = uni1(z, gnu, kode, 2, cy, nlast, fnul, tol, elim, alim);
    if (nw < 0) return (nw == -2) ? -2 : -1;
    if (nw != 0) {
        nlast = n;
        return nz;
    }
    real str = cabs_(cy[0]);
    // scaled backward recurrence
    real bry[3];
    bry[0] = real(1.0e3) * SPB_D1MACH1 / tol;
    bry[1] = real(1.0) / bry[0];
    bry[2] = bry[1];
    int iflag = 2;
    real ascle = bry[1];
    real csclr = 1.0;
    if (!(str > bry[0])) {
        iflag = 1;
        ascle = bry[0];
        csclr = real(1.0) / tol;
    } else if (!(str < bry[1])) {
        iflag = 3;
        ascle = bry[2];
        csclr = tol;
    }
    real cscrr = real(1.0) / csclr;
    cplx s1 = cy[1] * csclr;
    cplx s2 = cy[0] * csclr;
    cplx rz = crecip(z) * real(2.0);
    for (int i = 1; i <= nui; ++i) {
        cplx st = s2;
        s2 = (rz * st) * (dfnu + fnui) + s1;
        s1 = st;
        fnui -= real(1.0);
        if (iflag >= 3) continue;
        st = s2 * cscrr;
        real c1m = rmax(rabs(st.re), rabs(st.im));
        if (c1m <= ascle) continue;
        iflag += 1;
        ascle = bry[iflag - 1];
        s1 = s1 * cscrr;
        s2 = st;
        csclr = csclr * tol;
        cscrr = real(1.0) / csclr;
        s1 = s1 * csclr;
        s2 = s2 * csclr;
    }
    y[n - 1] = s2 * cscrr;
    if (n == 1) return nz;
    int nl = n - 1;
    fnui = real(nl);
    int k = nl;
    for (int i = 1; i <= nl; ++i) {
        cplx st = s2;
        s2 = (rz * st) * (fnu + fnui) + s1;
        s1 = st;
        st = s2 * cscrr;
        y[k - 1] = st;
        fnui -= real(1.0);
        k -= 1;
        if (iflag >= 3) continue;
        real c1m = rmax(rabs(st.re), rabs(st.im));
        if (c1m <= ascle) continue;
        iflag += 1;
        ascle = bry[iflag - 1];
        s1 = s1 * cscrr;
        s2 = st;
        csclr = csclr * tol;
        cscrr = real(1.0) / csclr;
        s1 = s1 * csclr;
        s2 = s2 * csclr;
    }
    return nz;
}

// Region ids reported by binu_region() / used by the GPU binning pass.
enum IRegion {
    IREG_SERIES = 0,  // power series
    IREG_ASYI = 1,    // large-|z| asymptotic
    IREG_MLRI = 2,    // Miller + Neumann normalisation (after uoik)
    IREG_WRSK = 3,    // Miller ratios + Wronskian (after uoik)
    IREG_UNI1 = 4,    // uniform asymptotics, Debye-type form (|Im z| <= sqrt(3) |Re z|), after uoik
    IREG_UNI2 = 5,    // uniform asymptotics, Airy-type form, after uoik
    IREG_DEBYE = 6,   // uniform (Debye-type) expansion at the order itself + second exponential (spb_debye.h): binned
                      // pipeline only; carved out of the four regions above where its parameters are large enough
    IREG_COUNT = 7
};

// Where the published algorithm raises the order to fnul and uses the Airy-type uniform expansion although
// the order itself is below fnul (|z| > fnul, 2|z| < order^2, |arg z| > pi/3), the Miller/Wronskian route is
// used instead as long as its recurrence length (about |z|) stays moderate: the choice of fnul as its upper
// limit bounds the sequential cost of the recurrence, not its accuracy, and the Airy-type expansion costs
// about six times more arithmetic per point at these moduli.  One rule for the region map, the drivers and
// the CPU checker.
#define SPB_WRSK_AZ_MAX 257.76407594148639 /* 3 fnul */
SPB_HD bool wrsk_instead_of_uni2(real az, real dfnu, real fnul) {
    return !(dfnu > fnul) && !(az > real(SPB_WRSK_AZ_MAX));
}

// First region binu() enters for a single member (n = 1); series hand-over
// (nz < 0) and uniform "nlast" hand-backs are resolved inside binu itself.
SPB_HD int binu_region(cplx z, real fnu, int n, real rl, real fnul, real az_known = real(-1.0)) {
    real az = az_known >= real(0.0) ? az_known : cabs_(z);
    real dfnu = fnu + real(n - 1);
    if (az <= real(2.0) || !(az * az * real(0.25) > dfnu + real(1.0))) return IREG_SERIES;
    if (!(az < rl)) {
        if (dfnu <= real(1.0) || !(az + az < dfnu * dfnu)) return IREG_ASYI;
    } else if (dfnu <= real(1.0)) {
        return IREG_MLRI;
    }
    if (dfnu > fnul || az > fnul) {
        // same split as buni(): form 2 when |Im z| > 1.7321 |Re z|
        if (!(rabs(z.im) > rabs(z.re) * real(1.7321))) return IREG_UNI1;
        if (wrsk_instead_of_uni2(az, dfnu, fnul)) return IREG_WRSK;
        return IREG_UNI2;
    }
    if (az > rl) return IREG_WRSK;
    return IREG_MLRI;
}

template <class Y>
SPB_FN int binu(cplx z, real fnu, int kode, int n, Y cy, real rl, real fnul, real tol, real elim,
                real alim) {
    int nz = 0;
    real az = cabs_(z);
    int nn = n;
    real dfnu = fnu + real(n - 1);
    int nw;
    bool do_series = (az <= real(2.0)) || !(az * az * real(0.25) > dfnu + real(1.0));
    if (do_series) {
        nw = seri(z, fnu, kode, nn, cy, tol, elim, alim);
        int inw = nw < 0 ? -nw : nw;
        nz += inw;
        nn -= inw;
        if (nn == 0) return nz;
        if (nw >= 0) return nz;
        dfnu = fnu + real(nn - 1);
    }
    for (;;) {
        // label 20
        bool to_miller_direct = false;  // label 70 without the uoik pre-test
        if (!(az < rl)) {
            if (dfnu <= real(1.0) || !(az + az < dfnu * dfnu)) {
                nw = asyi(z, fnu, kode, nn, cy, rl, tol, elim, alim);
                if (nw < 0) return (nw == -2) ? -2 : -1;
                return nz;
            }
        } else if (dfnu <= real(1.0)) {
            to_miller_direct = true;
        }
        if (!to_miller_direct) {
            // over/underflow test on the I sequence for the Miller algorithm
            nw = uoik(z, fnu, kode, 1, nn, cy, tol, elim, alim);
            if (nw < 0) return (nw == -2) ? -2 : -1;
            nz += nw;
            nn -= nw;
            if (nn == 0) return nz;
            dfnu = fnu + real(nn - 1);
            const bool form2 = rabs(z.im) > rabs(z.re) * real(1.7321);
            if ((dfnu > fnul || az > fnul) && !(form2 && wrsk_instead_of_uni2(az, dfnu, fnul))) {
                // raise the order to fnul, compute there and recur backward
                int nui = (int)(fnul - dfnu) + 1;
                if (nui < 0) nui = 0;
                int nlast = 0;
                nw = buni(z, fnu, kode, nn, cy, nui, nlast, fnul, tol, elim, alim);
                if (nw < 0) return (nw == -2) ? -2 : -1;
                nz += nw;
                if (nlast == 0) return nz;
                nn = nlast;
                // label 60
                if (az > rl) {
                    // fall through to the Wronskian branch below
                } else {
                    to_miller_direct = true;
                }
            } else if (!(az > rl)) {
                to_miller_direct = true;
            }
        }
        if (to_miller_direct) {
            nw = mlri(z, fnu, kode, nn, cy, tol);
            if (nw < 0) return (nw == -2) ? -2 : -1;
            return nz;
        }
        // Miller algorithm normalised by the Wronskian: overflow test on the K functions first
        cplx cw[2];
        nw = uoik(z, fnu, kode, 2, 2, cw, tol, elim, alim);
        if (nw < 0) {
            nz = nn;
            for (int i = 0; i < nn; ++i) cy[i] = cplx(0.0, 0.0);
            return nz;
        }
        if (nw > 0) return -1;
        nw = wrsk(z, fnu, kode, nn, cy, tol, elim, alim);
        if (nw < 0) return (nw == -2) ? -2 : -1;
        return nz;
    }
}

}  // namespace spb
