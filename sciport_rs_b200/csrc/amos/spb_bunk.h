// spb_bunk.h — K_fnu(z) for fnu > fnul by the uniform asymptotic expansions
// (TOMS 644 "ZBUNK" -> "ZUNK1"/"ZUNK2"), including the analytic
// continuation to Re z < 0 (mr = +-1) where the I function from the same
// expansion is added.  The batched API evaluates one order per point, so the
// expansions are written for a single member; a sequence is evaluated member
// by member (each with its own expansion) instead of by forward recurrence.
//   unk1 : |Im z| <= sqrt(3) |Re z|   Debye-type expansion (unik)
//   unk2 : otherwise                  Airy-type expansion of H(2)/J at zn = -+ i z (unhj + airy)
#pragma once
#include "spb_common.h"
#include "spb_uniform.h"
#include "spb_kregions.h"
#include "spb_airy.h"

namespace spb {

SPB_FN int unk1_single(cplx z, real fnu, int kode, int mr, cplx &y, real tol, real elim, real alim) {
    const real pi = SPB_PI;
    int nz = 0;
    real cscl = real(1.0) / tol, crsc = tol;
    real cssr[3] = {cscl, 1.0, crsc};
    real csrr[3] = {crsc, 1.0, cscl};
    real bry0 = real(1.0e3) * SPB_D1MACH1 / tol;
    cplx zr = z;
    if (z.re < real(0.0)) zr = -z;
    cplx phi, zeta1, zeta2, sum, s1, s2;
    cplx cwrk[16];
    real fn = fnu;
    int init = 0;
    unik(zr, fn, 2, 0, tol, init, phi, zeta1, zeta2, sum, cwrk);
    if (kode == 1) {
        s1 = zeta1 - zeta2;
    } else {
        cplx st = zr + zeta2;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = zeta1 - st;
    }
    real rs1 = s1.re;
    bool bad = rabs(rs1) > elim;
    int kflag = 2;
    if (!bad && !(rabs(rs1) < alim)) {
        // refine the test and scale
        real aphi = cabs_(phi);
        rs1 = rs1 + log(aphi);
        if (rabs(rs1) > elim) bad = true;
        if (!bad) {
            kflag = 1;
            if (rs1 >= real(0.0)) kflag = 3;
        }
    }
    if (!bad) {
        s2 = phi * sum;
        real e = exp(s1.re) * cssr[kflag - 1];
        s2 = s2 * cplx(e * cos(s1.im), e * sin(s1.im));
        if (kflag == 1 && uchk(s2, bry0, tol) != 0) bad = true;
    }
    if (bad) {
        if (rs1 > real(0.0)) return -1;
        // for Re z < 0 the I function to be added will overflow
        if (z.re < real(0.0)) return -1;
        y = cplx(0.0, 0.0);
        nz = 1;
    } else {
        y = s2 * csrr[kflag - 1];
    }
    if (mr == 0) return nz;
    // analytic continuation for Re z < 0: cspn and csgn are the coefficients of K and I
    nz = 0;
    real sgn = (mr < 0) ? pi : -pi;
    real csgni = sgn;
    int inu = (int)fnu;
    real fnf = fnu - real(inu);
    real ang = fnf * sgn;
    cplx cspn(cos(ang), sin(ang));
    if (inu % 2 == 1) cspn = -cspn;
    // the parameters of the K expansion are reused: only the sum changes sign pattern
    cplx phid, zet1d, zet2d, sumd;
    unik(zr, fn, 1, 0, tol, init, phid, zet1d, zet2d, sumd, cwrk);
    zet1d = zeta1;
    zet2d = zeta2;
    if (kode == 1) {
        s1 = zet2d - zet1d;
    } else {
        cplx st = zr + zet2d;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = st - zet1d;
    }
    rs1 = s1.re;
    bad = rabs(rs1) > elim;
    int iflag = 2;
    if (!bad && !(rabs(rs1) < alim)) {
        real aphi = cabs_(phid);
        rs1 = rs1 + log(aphi);
        if (rabs(rs1) > elim) bad = true;
        if (!bad) {
            iflag = 1;
            if (rs1 >= real(0.0)) iflag = 3;
        }
    }
    if (bad) {
        if (rs1 > real(0.0)) return -1;
        s2 = cplx(0.0, 0.0);
    } else {
        cplx st = phid * sumd;
        s2 = cplx(-csgni * st.im, csgni * st.re);
        real e = exp(s1.re) * cssr[iflag - 1];
        s2 = s2 * cplx(e * cos(s1.im), e * sin(s1.im));
        if (iflag == 1 && uchk(s2, bry0, tol) != 0) s2 = cplx(0.0, 0.0);
    }
    s2 = s2 * csrr[iflag - 1];
    // add the I and K functions
    s1 = y;
    if (kode != 1) {
        int iuf = 0;
        nz += s1s2(zr, s1, s2, bry0, alim, iuf);
    }
    y = s1 * cspn + s2;
    return nz;
}

SPB_FN int unk2_single(cplx z, real fnu, int kode, int mr, cplx &y, real tol, real elim, real alim) {
    const real pi = SPB_PI, hpi = SPB_HPI;
    const real aic = 1.26551212348464539;
    const cplx cr1(1.0, 1.73205080756887729);
    const cplx cr2(-0.5, -8.66025403784438647e-01);
    const real cipr[4] = {1.0, 0.0, -1.0, 0.0};
    const real cipi[4] = {0.0, -1.0, 0.0, 1.0};
    int nz = 0;
    real cscl = real(1.0) / tol, crsc = tol;
    real cssr[3] = {cscl, 1.0, crsc};
    real csrr[3] = {crsc, 1.0, cscl};
    real bry0 = real(1.0e3) * SPB_D1MACH1 / tol;
    cplx zr = z;
    if (z.re < real(0.0)) zr = -z;
    real yy = zr.im;
    cplx zn(zr.im, -zr.re);
    cplx zb = zr;
    int inu = (int)fnu;
    real fnf = fnu - real(inu);
    real ang = -hpi * fnf;
    real car = cos(ang), sar = sin(ang);
    cplx c2(hpi * sar, -hpi * car);
    int kk = inu % 4;
    cplx cs = cr1 * (c2 * cplx(cipr[kk], cipi[kk]));
    if (!(yy > real(0.0))) {
        zn.re = -zn.re;
        zb.im = -zb.im;
    }
    // K is computed from H(2, fnu, -i z) for z in the first quadrant; fourth-quadrant values by conjugation
    cplx phi, arg, zeta1, zeta2, asum, bsum, s1, s2;
    real fn = fnu;
    unhj(zn, fn, 0, tol, phi, arg, zeta1, zeta2, asum, bsum);
    if (kode == 1) {
        s1 = zeta1 - zeta2;
    } else {
        cplx st = zb + zeta2;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = zeta1 - st;
    }
    real rs1 = s1.re;
    bool bad = rabs(rs1) > elim;
    int kflag = 2;
    if (!bad && !(rabs(rs1) < alim)) {
        real aphi = cabs_(phi);
        real aarg = cabs_(arg);
        rs1 = rs1 + log(aphi) - real(0.25) * log(aarg) - aic;
        if (rabs(rs1) > elim) bad = true;
        if (!bad) {
            kflag = 1;
            if (rs1 >= real(0.0)) kflag = 3;
        }
    }
    if (!bad) {
        // scale s1 to keep intermediate arithmetic on scale near exponent extremes
        cplx a2 = arg * cr2;
        int nai, ndai, idum;
        cplx ai = airy(a2, 0, 2, nai, idum);
        cplx dai = airy(a2, 1, 2, ndai, idum);
        cplx st = (dai * bsum) * cr2 + ai * asum;
        s2 = (st * phi) * cs;
        real e = exp(s1.re) * cssr[kflag - 1];
        s2 = s2 * cplx(e * cos(s1.im), e * sin(s1.im));
        if (kflag == 1 && uchk(s2, bry0, tol) != 0) bad = true;
    }
    if (bad) {
        if (rs1 > real(0.0)) return -1;
        if (z.re < real(0.0)) return -1;
        y = cplx(0.0, 0.0);
        nz = 1;
    } else {
        if (!(yy > real(0.0))) s2.im = -s2.im;
        y = s2 * csrr[kflag - 1];
    }
    if (mr == 0) return nz;
    // analytic continuation for Re z < 0
    nz = 0;
    real sgn = (mr < 0) ? pi : -pi;
    real csgni = sgn;
    if (!(yy > real(0.0))) csgni = -csgni;
    int ifn = inu;
    ang = fnf * sgn;
    cplx cspn(cos(ang), sin(ang));
    if (ifn % 2 == 1) cspn = -cspn;
    // cs = coefficient of the J function to get the I function: I(fnu,z) = exp(i fnu pi/2) J(fnu, -i z)
    cs = cplx(sar * csgni, car * csgni);
    int in = ifn % 4;
    cplx cq(cipr[in], cipi[in]);
    cs = cs * conj(cq);
    // the expansion parameters computed for K are reused for I
    if (kode == 1) {
        s1 = zeta2 - zeta1;
    } else {
        cplx st = zb + zeta2;
        real rast = fn / cabs_(st);
        st = conj(st) * rast * rast;
        s1 = st - zeta1;
    }
    rs1 = s1.re;
    bad = rabs(rs1) > elim;
    int iflag = 2;
    if (!bad && !(rabs(rs1) < alim)) {
        real aphi = cabs_(phi);
        real aarg = cabs_(arg);
        rs1 = rs1 + log(aphi) - real(0.25) * log(aarg) - aic;
        if (rabs(rs1) > elim) bad = true;
        if (!bad) {
            iflag = 1;
            if (rs1 >= real(0.0)) iflag = 3;
        }
    }
    if (bad) {
        if (rs1 > real(0.0)) return -1;
        s2 = cplx(0.0, 0.0);
    } else {
        int nai, ndai, idum;
        cplx ai = airy(arg, 0, 2, nai, idum);
        cplx dai = airy(arg, 1, 2, ndai, idum);
        cplx st = dai * bsum + ai * asum;
        s2 = (st * phi) * cs;
        real e = exp(s1.re) * cssr[iflag - 1];
        s2 = s2 * cplx(e * cos(s1.im), e * sin(s1.im));
        if (iflag == 1 && uchk(s2, bry0, tol) != 0) s2 = cplx(0.0, 0.0);
    }
    if (!(yy > real(0.0))) s2.im = -s2.im;
    s2 = s2 * csrr[iflag - 1];
    s1 = y;
    if (kode != 1) {
        int iuf = 0;
        nz += s1s2(zr, s1, s2, bry0, alim, iuf);
    }
    y = s1 * cspn + s2;
    return nz;
}

template <class Y>
SPB_FN int bunk(cplx z, real fnu, int kode, int mr, int n, Y y, real tol, real elim, real alim) {
    real ax = rabs(z.re) * real(1.7321);
    real ay = rabs(z.im);
    int nz = 0;
    for (int i = 0; i < n; ++i) {
        int nw;
        if (ay > ax)
            nw = unk2_single(z, fnu + real(i), kode, mr, y[i], tol, elim, alim);
        else
            nw = unk1_single(z, fnu + real(i), kode, mr, y[i], tol, elim, alim);
        if (nw < 0) return nw;
        nz += nw;
    }
    return nz;
}

}  // namespace spb
