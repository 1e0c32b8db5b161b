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

// What the range / overflow / underflow tests of airy() and biry() decide before any function evaluation, for
// |z| > 1 and |zeta| >= rl: AIRY_COMPUTE (evaluate), AIRY_DECLINE (unscaled Ai in a rescaling band: complete
// routine) or AIRY_DONE (the result is 0 with the codes set: total loss of significance, underflow of Ai, overflow).
// The classifier of the Airy kernels calls it too: decided points are stored there and never visited again (on a
// log-uniform batch with |z| up to 1e4 that is 60 % of the points beyond |z| = 10).
enum { AIRY_COMPUTE = 0, AIRY_DECLINE = 1, AIRY_DONE = 2 };
struct AiryPre {
    cplx csq, zta;  // sqrt(z), zeta = (2/3) z sqrt(z) with the sign conventions of the routines
    bool right;     // Re zeta >= 0 and Re z > 0: K_fnu(zeta) directly (Ai) / no continuation phase (Bi)
    int ierr, nz;
};
SPB_FN int airy_pretest(bool bi, cplx z, int kode, AiryPre &p) {
    const real tth = 6.66666666666666667e-01;
    const real elim = SPB_ELIM, alim = SPB_ALIM;
    p.ierr = 0;
    p.nz = 0;
    const real az = cabs_(z);
    real aa = real(1048575.9996744783);  // total loss of significance beyond (airy() / biry(): same constant)
    if (az > aa) {
        p.ierr = 4;
        return AIRY_DONE;
    }
    if (az > sqrt(aa)) p.ierr = 3;
    p.csq = csqrt_(z);
    cplx zta = (z * p.csq) * tth;
    const real ak = zta.im;
    if (z.re < real(0.0)) {
        zta.re = -rabs(zta.re);
        zta.im = ak;
    }
    if (z.im == real(0.0) && z.re <= real(0.0)) {
        zta.re = 0.0;
        zta.im = ak;
    }
    p.zta = zta;
    aa = zta.re;
    p.right = aa >= real(0.0) && z.re > real(0.0);
    if (kode == 2) return AIRY_COMPUTE;
    if (bi) {
        real bb = rabs(aa);
        if (!(bb < alim)) {
            if (bb + real(0.25) * log(az) > elim) {
                p.ierr = 2;
                return AIRY_DONE;
            }
        }
        return AIRY_COMPUTE;  // (the band alim <= |Re zeta| <= elim is evaluated with the scale factor tol)
    }
    if (p.right) {
        if (aa >= alim) {
            if (-aa - real(0.25) * log(az) < -elim) {
                p.nz = 1;
                return AIRY_DONE;
            }
            return AIRY_DECLINE;
        }
    } else if (aa <= -alim) {
        if (-aa + real(0.25) * log(az) > elim) {
            p.ierr = 2;
            return AIRY_DONE;
        }
        return AIRY_DECLINE;
    }
    return AIRY_COMPUTE;
}

// Ai (id = 0) / Ai' (id = 1), |z| > 1 with |zeta| >= rl.  false: declined, nothing decided.
// One call site for the function evaluation: K_fnu (and I_fnu for the continuation through the cut,
// K(w e^{+-i pi}) = e^{-+i pi fnu} K(w) -+ i pi I(w)) at w = zeta or -zeta, so the lanes of a warp that sit in
// different sectors run it together.
SPB_FN bool airy_large(cplx z, int id, int kode, cplx &val, int &nz, int &ierr) {
    const real coef = 1.83776298473930683e-01;  // 1/(pi sqrt 3)
    const real tol = SPB_TOL, alim = SPB_ALIM;
    val = cplx(0.0, 0.0);
    AiryPre p;
    const int what = airy_pretest(false, z, kode, p);
    ierr = p.ierr;
    nz = p.nz;
    if (what == AIRY_DECLINE) return false;
    if (what == AIRY_DONE) return true;
    const real fnu = (real(1.0) + real(id)) / real(3.0);
    const cplx w = p.right ? p.zta : -p.zta;
    cplx ival, kval;
    int inz = 0, knz = 0;
    if (fused_ik_region(w, fnu, kode, ival, inz, kval, knz, real(-1.0)) < 0) return false;
    cplx y = kval;
    if (!p.right) {
        const int mr = (z.im < real(0.0)) ? -1 : 1;
        const real sgn = (mr < 0) ? real(SPB_PI) : real(-SPB_PI);
        cplx csgn(0.0, sgn);
        if (kode != 1) csgn = csgn * cis_(-w.im);
        const cplx cspn = cis_(fnu * sgn);  // int(fnu) = 0 for the Airy orders
        cplx c1 = kval, c2 = ival;
        if (kode != 1) {
            int iuf = 0;
            const real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
            nz += s1s2(w, c1, c2, ascle, alim, iuf);
        }
        y = cspn * c1 + csgn * c2;
    }
    const cplx s1 = y * coef;
    val = (id != 1) ? p.csq * s1 : -(z * s1);
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

// ---- 1 < |z|, |zeta| < rl (keys 1..4 of the Airy classifier: 14 % of a log-uniform batch, and the expensive ones) ----
// The same wrappers with the routines the published drivers enter for |zeta| < rl and the Airy orders:
//   K_fnu(zeta): the K routine (power series for |zeta| <= 2, Miller algorithm above) in its lean single-member
//     form (bknu_simple: the form the binned K pass runs; it declines where the value leaves the middle scaling
//     level, which cannot happen for |zeta| < rl);
//   I_fnu(zeta), fnu <= 5/3: the region driver takes the power series (|zeta| <= 2 or |zeta|^2/4 <= fnu + n) or the
//     Miller algorithm with the Neumann normalisation (its over/underflow pre-test is a no-op for these orders and
//     moduli), so `binu` is `seri` or `mlri` -- same routine, same arguments, identical results.
// No over/underflow handling is needed: |Re zeta| < rl << alim.

// I_{fnu+k}(zeta), k < n, for |zeta| < rl and fnu + n - 1 <= 5/3: the route binu() takes.  Returns binu's codes.
template <class Y>
SPB_FN int airy_i_mid(cplx zta, real fnu, int kode, int n, Y y, real azeta) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    const real dfnu = fnu + real(n - 1);
    if (azeta <= real(2.0) || !(azeta * azeta * real(0.25) > dfnu + real(1.0))) {
        const int nw = seri(zta, fnu, kode, n, y, tol, elim, alim, azeta);
        if (nw != 0) return -9;  // an underflowed member: cannot happen for 1 < |z|; decline if it does
        return 0;
    }
    const int nw = mlri(zta, fnu, kode, n, y, tol, azeta);
    if (nw < 0) return (nw == -2) ? -2 : -1;
    return 0;
}

// Ai / Ai' for 1 < |z|, |zeta| < rl.  false: declined.
SPB_FN bool airy_mid(cplx z, int id, int kode, cplx &val, int &nz, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real coef = 1.83776298473930683e-01;  // 1/(pi sqrt 3)
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    ierr = 0;
    nz = 0;
    const real fid = real(id);
    const real fnu = (real(1.0) + fid) / real(3.0);
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
    cplx y;
    if (zta.re >= real(0.0) && z.re > real(0.0)) {
        if (bknu_simple(zta, fnu, kode, y, tol, elim, alim, real(-1.0)) != 0) return false;
    } else {
        const int mr = (z.im < real(0.0)) ? -1 : 1;
        const cplx zn = -zta;
        const real azn = cabs_(zn);
        cplx iv[1];
        iv[0] = cplx(0.0, 0.0);
        // (both evaluations before any exit: the loops of the two routines reconverge in between)
        const int ni = airy_i_mid(zn, fnu, kode, 1, iv, azn);
        cplx kval;
        const int nk = bknu_simple(zn, fnu, kode, kval, tol, elim, alim, azn);
        if (ni != 0 || nk != 0) return false;
        const real sgn = (mr < 0) ? real(SPB_PI) : real(-SPB_PI);
        cplx csgn(0.0, sgn);
        if (kode != 1) csgn = csgn * cis_(-zn.im);
        const cplx cspn = cis_(fnu * sgn);
        cplx c1 = kval, c2 = iv[0];
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

// Bi / Bi' for 1 < |z|, |zeta| < rl.  false: declined.
SPB_FN bool biry_mid(cplx z, int id, int kode, cplx &val, int &ierr) {
    const real tth = 6.66666666666666667e-01;
    const real coef = 5.77350269189625765e-01;  // 1/sqrt 3
    const real pi = SPB_PI;
    ierr = 0;
    const real fid = real(id);
    real fnu = (real(1.0) + fid) / real(3.0);
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
    real fmr = 0.0;
    if (!(zta.re >= real(0.0) && z.re > real(0.0))) {
        fmr = (z.im < real(0.0)) ? -pi : pi;
        zta = -zta;
    }
    const real azeta = cabs_(zta);
    cplx cy[2];
    cy[0] = cplx(0.0, 0.0);
    cy[1] = cplx(0.0, 0.0);
    const int n1 = airy_i_mid(zta, fnu, kode, 1, cy, azeta);
    cplx s1 = cy[0] * cis_(fmr * fnu);
    fnu = (real(2.0) - fid) / real(3.0);
    const int n2 = airy_i_mid(zta, fnu, kode, 2, cy, azeta);
    if (n1 != 0 || n2 != 0) return false;
    // one backward recurrence step gives order -1/3 or -2/3
    const cplx s2 = cdiv(cy[0], zta) * (fnu + fnu) + cy[1];
    s1 = (s1 + s2 * cis_(fmr * (fnu - real(1.0)))) * coef;
    val = (id != 1) ? csq * s1 : z * s1;
    return true;
}

}  // namespace spb
