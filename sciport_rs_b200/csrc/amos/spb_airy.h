// spb_airy.h — Airy functions Ai, Ai', Bi, Bi' of complex argument
// (TOMS 644 "ZAIRY"/"ZBIRY"):
//   |z| <= 1 : Maclaurin series
//   else     : Ai  = (1/pi) sqrt(z/3) K_{1/3}(zeta),  Ai' = -(z/(pi sqrt 3)) K_{2/3}(zeta),
//              Bi from I_{+-1/3}(zeta) / I_{+-2/3}(zeta),  zeta = (2/3) z^{3/2};
//              Re z < 0 continues K through the cut with the I function (acai).
// kode = 2 scales Ai by exp(zeta) and Bi by exp(-|Re zeta|).
#pragma once
#include "spb_common.h"
#include "spb_iregions.h"
#include "spb_kregions.h"

namespace spb {

// forward: full I-function region driver (spb_binu.h), used by biry
template <class Y>
SPB_FN int binu(cplx z, real fnu, int kode, int n, Y cy, real rl, real fnul, real tol, real elim,
                real alim);

// Analytic continuation K(fnu, zn e^{mp}) = K(fnu,zn) e^{-mp fnu} - mp I(fnu,zn), mp = i pi mr,
// specialised to the Airy orders (fnu = 1/3, 2/3; n = 1).
template <class Y>
SPB_FN int acai(cplx z, real fnu, int kode, int mr, int n, Y y, real rl, real tol, real elim,
                real alim) {
    int nz = 0;
    cplx zn = -z;
    real az = cabs_(z);
    int nn = n;
    real dfnu = fnu + real(n - 1);
    int nw;
    // failures are carried to the end of the routine (no return between the data-dependent loops of the I and K
    // routines: their exits must reconverge before the next routine starts)
    int fail = 0;
    if (az <= real(2.0) || !(az * az * real(0.25) > dfnu + real(1.0))) {
        nw = seri(zn, fnu, kode, nn, y, tol, elim, alim, az);
    } else if (!(az < rl)) {
        nw = asyi(zn, fnu, kode, nn, y, rl, tol, elim, alim, az);
        if (nw < 0) fail = (nw == -2) ? -2 : -1;
    } else {
        nw = mlri(zn, fnu, kode, nn, y, tol, az);
        if (nw < 0) fail = (nw == -2) ? -2 : -1;
    }
    cplx cy[2];
    nw = bknu(zn, fnu, kode, 1, cy, tol, elim, alim, az);
    if (nw != 0 && fail == 0) fail = (nw == -2) ? -2 : -1;
    real fmr = real(mr);
    real sgn = (fmr < real(0.0)) ? real(SPB_PI) : real(-SPB_PI);
    cplx csgn(0.0, sgn);
    if (kode != 1) {
        real yy = -zn.im;
        csgn = csgn * cis_(yy);
    }
    int inu = (int)fnu;
    real arg = (fnu - real(inu)) * sgn;
    cplx cspn = cis_(arg);
    if (inu % 2 != 0) cspn = -cspn;
    cplx c1 = cy[0];
    cplx c2 = y[0];
    if (kode != 1) {
        int iuf = 0;
        real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
        nw = s1s2(zn, c1, c2, ascle, alim, iuf);
        nz += nw;
    }
    y[0] = cspn * c1 + csgn * c2;
    if (fail != 0) {
        y[0] = cplx(0.0, 0.0);
        return fail;
    }
    return nz;
}

// Maclaurin sums shared by Ai and Bi for |z| <= 1
SPB_FN void airy_series(cplx z, real az, real fid, real tol, cplx &s1, cplx &s2) {
    s1 = cplx(1.0, 0.0);
    s2 = cplx(1.0, 0.0);
    real aa = az * az;
    if (aa < tol / az) return;
    cplx trm1(1.0, 0.0), trm2(1.0, 0.0);
    real atrm = 1.0;
    cplx z3 = (z * z) * z;
    real az3 = az * aa;
    real ak = real(2.0) + fid;
    real bk = real(3.0) - fid - fid;
    real ck = real(4.0) - fid;
    real dk = real(3.0) + fid + fid;
    real d1 = ak * dk;
    real d2 = bk * ck;
    real ad = rmin(d1, d2);
    ak = real(24.0) + real(9.0) * fid;
    bk = real(30.0) - real(9.0) * fid;
    for (int k = 1; k <= 25; ++k) {
        // (d1, d2: products of small loop counters; ad is one of them)
        const real r1 = rcp_fast(d1), r2 = rcp_fast(d2);
        trm1 = (trm1 * z3) * r1;
        s1 = s1 + trm1;
        trm2 = (trm2 * z3) * r2;
        s2 = s2 + trm2;
        atrm = atrm * az3 * rmax(r1, r2);
        d1 += ak;
        d2 += bk;
        ad = rmin(d1, d2);
        if (atrm < tol * ad) break;
        ak += real(18.0);
        bk += real(18.0);
    }
}

// Ai / Ai' for |z| <= 1 (Maclaurin series): the branch of airy() the lean small-|z| kernel runs on its own
SPB_FN cplx airy_small(cplx z, real az, int id, int kode) {
    const real tth = 6.66666666666666667e-01;
    const real c1 = 3.55028053887817240e-01;
    const real c2 = 2.58819403792806799e-01;
    const real tol = SPB_TOL;
    const real fid = real(id);
    cplx ai;
    if (az < tol) {
        real aa = real(1.0e3) * SPB_D1MACH1;
        cplx s1(0.0, 0.0);
        if (id != 1) {
            if (az > aa) s1 = z * c2;
            return cplx(c1 - s1.re, -s1.im);
        }
        ai = cplx(-c2, 0.0);
        aa = sqrt(aa);
        if (az > aa) s1 = (z * z) * real(0.5);
        return ai + s1 * c1;
    }
    cplx s1, s2;
    airy_series(z, az, fid, tol, s1, s2);
    if (id != 1) {
        ai = s1 * c1 - (z * s2) * c2;
    } else {
        ai = -s2 * c2;
        if (az > tol) {
            real cc = c1 / (real(1.0) + fid);
            ai = ai + ((z * z) * s1) * cc;
        }
    }
    if (kode == 1) return ai;
    cplx zta = (z * csqrt_(z)) * tth;
    return ai * cexp_(zta);
}

// Bi / Bi' for |z| <= 1 (Maclaurin series)
SPB_FN cplx biry_small(cplx z, real az, int id, int kode) {
    const real tth = 6.66666666666666667e-01;
    const real c1 = 6.14926627446000736e-01;
    const real c2 = 4.48288357353826359e-01;
    const real tol = SPB_TOL;
    const real fid = real(id);
    cplx bi;
    if (az < tol) return cplx(c1 * (real(1.0) - fid) + fid * c2, 0.0);
    cplx s1, s2;
    airy_series(z, az, fid, tol, s1, s2);
    if (id != 1) {
        bi = s1 * c1 + (z * s2) * c2;
    } else {
        bi = s2 * c2;
        if (az > tol) {
            real cc = c1 / (real(1.0) + fid);
            bi = bi + ((z * z) * s1) * cc;
        }
    }
    if (kode == 1) return bi;
    cplx zta = (z * csqrt_(z)) * tth;
    real eaa = exp_(-rabs(zta.re));
    return bi * eaa;
}

// Ai (id=0) or Ai' (id=1).  Returns value; nz/ierr follow the AMOS convention.
SPB_FN cplx airy(cplx z, int id, int kode, int &nz, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real c1 = 3.55028053887817240e-01;
    const real c2 = 2.58819403792806799e-01;
    const real coef = 1.83776298473930683e-01;  // 1/(pi sqrt 3)
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL;
    ierr = 0;
    nz = 0;
    real az = cabs_(z);
    real fid = real(id);
    cplx ai;
    if (!(az > real(1.0))) return airy_small(z, az, id, kode);
    // |z| > 1
    real fnu = (real(1.0) + fid) / real(3.0);
    real alaz = log(az);
    // range test: |z| > (min(0.5/tol, 2^30-ish))^(2/3) is total loss, its square root half loss
    // min(0.5/tol, (2^31 - 1)/2)^(2/3), a constant of the arithmetic (binary64: the second argument of the min)
    real aa = real(1048575.9996744783);
    if (az > aa) {
        ierr = 4;
        nz = 0;
        return cplx(0.0, 0.0);
    }
    aa = sqrt(aa);
    if (az > aa) ierr = 3;
    cplx csq = csqrt_(z);
    cplx zta = (z * csq) * tth;
    // Re(zta) <= 0 when Re(z) < 0, especially when Im(z) is small
    int iflag = 0;
    real sfac = 1.0;
    real ak = zta.im;
    if (z.re < real(0.0)) {
        zta.re = -rabs(zta.re);
        zta.im = ak;
    }
    if (z.im == real(0.0) && z.re <= real(0.0)) {
        zta.re = 0.0;
        zta.im = ak;
    }
    aa = zta.re;
    cplx cy[2];
    if (aa >= real(0.0) && z.re > real(0.0)) {
        if (kode != 2) {
            // underflow test
            if (aa >= alim) {
                aa = -aa - real(0.25) * alaz;
                iflag = 2;
                sfac = real(1.0) / tol;
                if (aa < -elim) {
                    nz = 1;
                    return cplx(0.0, 0.0);
                }
            }
        }
        nz = bknu(zta, fnu, kode, 1, cy, tol, elim, alim);
    } else {
        if (kode != 2) {
            // overflow test
            if (aa <= -alim) {
                aa = -aa + real(0.25) * alaz;
                iflag = 1;
                sfac = tol;
                if (aa > elim) {
                    ierr = 2;
                    nz = 0;
                    return cplx(0.0, 0.0);
                }
            }
        }
        // the continuation returns exp(zta) K(fnu,zta) on kode=2
        int mr = 1;
        if (z.im < real(0.0)) mr = -1;
        int nn = acai(zta, fnu, kode, mr, 1, cy, rl, tol, elim, alim);
        if (nn < 0) {
            nz = 0;
            ierr = (nn == -1) ? 2 : 5;
            return cplx(0.0, 0.0);
        }
        nz += nn;
    }
    cplx s1 = cy[0] * coef;
    if (iflag == 0) {
        if (id != 1) return csq * s1;
        return -(z * s1);
    }
    s1 = s1 * sfac;
    if (id != 1) {
        s1 = s1 * csq;
        return s1 * (real(1.0) / sfac);
    }
    s1 = -(s1 * z);
    return s1 * (real(1.0) / sfac);
}

// Bi (id=0) or Bi' (id=1)
SPB_FN cplx biry(cplx z, int id, int kode, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real c1 = 6.14926627446000736e-01;
    const real c2 = 4.48288357353826359e-01;
    const real coef = 5.77350269189625765e-01;  // 1/sqrt 3
    const real pi = SPB_PI;
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL, fnul = SPB_FNUL;
    ierr = 0;
    real az = cabs_(z);
    real fid = real(id);
    cplx bi;
    if (!(az > real(1.0))) return biry_small(z, az, id, kode);
    real fnu = (real(1.0) + fid) / real(3.0);
    // min(0.5/tol, (2^31 - 1)/2)^(2/3), a constant of the arithmetic (binary64: the second argument of the min)
    real aa = real(1048575.9996744783);
    if (az > aa) {
        ierr = 4;
        return cplx(0.0, 0.0);
    }
    aa = sqrt(aa);
    if (az > aa) ierr = 3;
    cplx csq = csqrt_(z);
    cplx zta = (z * csq) * tth;
    real sfac = 1.0;
    real ak = zta.im;
    if (z.re < real(0.0)) {
        zta.re = -rabs(zta.re);
        zta.im = ak;
    }
    if (z.im == real(0.0) && z.re <= real(0.0)) {
        zta.re = 0.0;
        zta.im = ak;
    }
    aa = zta.re;
    if (kode != 2) {
        // overflow test
        real bb = rabs(aa);
        if (!(bb < alim)) {
            bb = bb + real(0.25) * log(az);
            sfac = tol;
            if (bb > elim) {
                ierr = 2;
                return cplx(0.0, 0.0);
            }
        }
    }
    real fmr = 0.0;
    if (!(aa >= real(0.0) && z.re > real(0.0))) {
        fmr = pi;
        if (z.im < real(0.0)) fmr = -pi;
        zta = -zta;
    }
    // fmr*fnu is the phase of the analytic continuation of I(fnu, zta)
    cplx cy[2];
    int nz = binu(zta, fnu, kode, 1, cy, rl, fnul, tol, elim, alim);
    // a failure of the first I evaluation is reported after the second one (reconvergence, see acai)
    const int nz_first = nz;
    aa = fmr * fnu;
    cplx s1 = (cy[0] * cis_(aa)) * sfac;
    fnu = (real(2.0) - fid) / real(3.0);
    nz = binu(zta, fnu, kode, 2, cy, rl, fnul, tol, elim, alim);
    cy[0] = cy[0] * sfac;
    cy[1] = cy[1] * sfac;
    // one backward recurrence step gives order -1/3 or -2/3
    cplx s2 = cdiv(cy[0], zta) * (fnu + fnu) + cy[1];
    aa = fmr * (fnu - real(1.0));
    s1 = (s1 + s2 * cis_(aa)) * coef;
    if (nz_first < 0) {
        ierr = (nz_first == -1) ? 2 : 5;
        return cplx(0.0, 0.0);
    }
    if (id != 1) {
        bi = csq * s1;
        return bi * (real(1.0) / sfac);
    }
    bi = z * s1;
    return bi * (real(1.0) / sfac);
}

}  // namespace spb

namespace spb {
// scipy.special.airy value convention: NaN on domain / overflow / no-result codes
SPB_HD cplx airy_semantics(cplx z, cplx val, int ierr) {
    (void)z;
    if (ierr == 1 || ierr == 2 || ierr == 4 || ierr == 5) {
        const real inf = real(1.0) / real(0.0);
        return cplx(inf - inf, inf - inf);
    }
    return val;
}
}  // namespace spb
