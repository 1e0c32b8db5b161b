// spb_airy_lean.h — Airy functions for |zeta| = (2/3)|z|^(3/2) >= rl = 21.8 (|z| >= 10.2) through the large-|zeta|
// routines only.
//
// The published routines (spb_airy.h) reach K_{1/3}, K_{2/3} (Ai, Ai') and I_{+-1/3}, I_{+-2/3} (Bi, Bi') through the
// complete K routine and the complete region driver of I; compiled for the GPU that is 255 registers with spills
// for every point, although for |zeta| >= rl only one region of each is ever entered:
//   * I_fnu(zeta), fnu <= 5/3: the region driver goes straight to the large-|z| expansion (|zeta| >= rl and
//     2 |zeta| >= fnu^2), so `binu` IS `asyi` there — same routine, same arguments, identical results;
//   * K_fnu(zeta): the Hankel series of K has the same coefficients as the one of I, and the fused routine
//     (spb_family.h, departure 3 of DESIGN.md section 2) returns both from one summation; the published routine
//     runs the Miller algorithm here.
// The wrappers below repeat the range / overflow / underflow logic of airy() and biry() line by line and swap in
// those two routines, so error codes are identical; values of Ai differ from the published routine by the rounding
// of a different (shorter) K evaluation.  A point in the narrow bands where the published routine rescales
// (alim <= |Re zeta| <= elim, unscaled Ai) is DECLINED (returns false): the caller runs the complete routine.
// On a log-uniform batch (config 4) 43 % of the points are |z| <= 1 (Maclaurin kernel), 43 % come here.
#pragma once
#include "spb_family.h"

namespace spb {

// Ai (id = 0) / Ai' (id = 1), |z| > 1 with |zeta| >= rl.  false: declined, nothing decided.
SPB_FN bool airy_large(cplx z, int id, int kode, cplx &val, int &nz, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real coef = 1.83776298473930683e-01;  // 1/(pi sqrt 3)
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    ierr = 0;
    nz = 0;
    val = cplx(0.0, 0.0);
    const real az = cabs_(z);
    const real fid = real(id);
    const real fnu = (real(1.0) + fid) / real(3.0);
    real aa = real(1048575.9996744783);  // total loss of significance beyond (airy(): same constant)
    if (az > aa) {
        ierr = 4;
        return true;
    }
    if (az > sqrt(aa)) ierr = 3;
    const cplx csq = csqrt_(z);
    cplx zta = (z * csq) * tth;
    const real ak = zta.im;
    if (z.re < real(0.0)) {
        zta.re = -rabs(zta.re);
        zta.im = ak;
    }
    if (z.im == real(0.0) && z.re <= real(0.0)) {
        zta.re = 0.0;
        zta.im = ak;
    }
    aa = zta.re;
    cplx y;
    if (aa >= real(0.0) && z.re > real(0.0)) {
        // K_fnu(zeta) directly
        if (kode != 2 && aa >= alim) {
            // underflow test of airy(); the band that needs rescaling goes to the complete routine
            if (-aa - real(0.25) * log(az) < -elim) {
                nz = 1;
                return true;
            }
            return false;
        }
        cplx ival, kval;
        int inz = 0, knz = 0;
        if (fused_ik_region(zta, fnu, kode, ival, inz, kval, knz, real(-1.0)) < 0) return false;
        y = kval;
    } else {
        // continuation through the cut with I_fnu: K(zn e^{+-i pi}) = e^{-+i pi fnu} K(zn) -+ i pi I(zn), zn = -zeta
        if (kode != 2 && aa <= -alim) {
            if (-aa + real(0.25) * log(az) > elim) {
                ierr = 2;
                return true;
            }
            return false;
        }
        const int mr = (z.im < real(0.0)) ? -1 : 1;
        const cplx zn = -zta;
        cplx ival, kval;
        int inz = 0, knz = 0;
        if (fused_ik_region(zn, fnu, kode, ival, inz, kval, knz, real(-1.0)) < 0) return false;
        const real sgn = (mr < 0) ? real(SPB_PI) : real(-SPB_PI);
        cplx csgn(0.0, sgn);
        if (kode != 1) csgn = csgn * cis_(-zn.im);
        const cplx cspn = cis_(fnu * sgn);  // int(fnu) = 0 for the Airy orders
        cplx c1 = kval, c2 = ival;
        if (kode != 1) {
            int iuf = 0;
            const real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
            nz += s1s2(zn, c1, c2, ascle, alim, iuf);
        }
        y = cspn * c1 + csgn * c2;
    }
    const cplx s1 = y * coef;
    val = (id != 1) ? csq * s1 : -(z * s1);
    return true;
}

// Bi (id = 0) / Bi' (id = 1), |z| > 1 with |zeta| >= rl: biry() with the region driver of I replaced by the
// large-|z| expansion it would enter anyway (identical results and codes).  Never declines.
SPB_FN cplx biry_large(cplx z, int id, int kode, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real coef = 5.77350269189625765e-01;  // 1/sqrt 3
    const real pi = SPB_PI;
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL;
    ierr = 0;
    const real az = cabs_(z);
    const real fid = real(id);
    real fnu = (real(1.0) + fid) / real(3.0);
    real aa = real(1048575.9996744783);
    if (az > aa) {
        ierr = 4;
        return cplx(0.0, 0.0);
    }
    if (az > sqrt(aa)) ierr = 3;
    const cplx csq = csqrt_(z);
    cplx zta = (z * csq) * tth;
    real sfac = 1.0;
    const real ak = zta.im;
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
        fmr = (z.im < real(0.0)) ? -pi : pi;
        zta = -zta;
    }
    const real azeta = cabs_(zta);
    cplx cy[2];
    cy[0] = cplx(0.0, 0.0);
    cy[1] = cplx(0.0, 0.0);
    int nw = asyi(zta, fnu, kode, 1, cy, rl, tol, elim, alim, azeta);
    const int nz_first = nw < 0 ? ((nw == -2) ? -2 : -1) : 0;
    cplx s1 = (cy[0] * cis_(fmr * fnu)) * sfac;
    fnu = (real(2.0) - fid) / real(3.0);
    asyi(zta, fnu, kode, 2, cy, rl, tol, elim, alim, azeta);
    cy[0] = cy[0] * sfac;
    cy[1] = cy[1] * sfac;
    // one backward recurrence step gives order -1/3 or -2/3
    const cplx s2 = cdiv(cy[0], zta) * (fnu + fnu) + cy[1];
    s1 = (s1 + s2 * cis_(fmr * (fnu - real(1.0)))) * coef;
    if (nz_first < 0) {
        ierr = (nz_first == -1) ? 2 : 5;
        return cplx(0.0, 0.0);
    }
    const cplx bi = (id != 1) ? csq * s1 : z * s1;
    return bi * (real(1.0) / sfac);
}

}  // namespace spb
