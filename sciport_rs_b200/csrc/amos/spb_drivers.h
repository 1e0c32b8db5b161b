// spb_drivers.h — per-point drivers: I, J, K, H1/H2, Y (TOMS 644 "ZBESI",
// "ZBESJ", "ZBESK", "ZBESH", "ZBESY") and the analytic
// continuation of K to the left half plane ("ZACON").
//
// These are the functions the reference reaches through
//   complex_bessel_rs::bessel_i::bessel_i(order, z)   /root/reference/src/special/mod.rs:10
//   complex_bessel_rs::bessel_k::bessel_k(order, z)   /root/reference/src/special/kv.rs:19
// widened to the J/Y/H families named by BASELINE.json.  Each returns the
// value(s) plus the raw (nz, ierr) pair: ierr 1 input, 2 overflow, 3 half
// precision lost, 4 total loss, 5 no convergence.
#pragma once
#include "spb_common.h"
#include "spb_binu.h"
#include "spb_bunk.h"

namespace spb {

// Range limits: the I/K family uses min(0.5/tol, (2^31-1)/2) as in classic
// AMOS; the J/Y/H family uses 0.5/tol alone, matching the oracle
// (scipy.special).
#define SPB_RANGE_IK 1073741823.5
#define SPB_RANGE_JYH 2251799813685248.0 /* 0.5/tol */

// K(fnu, zn e^{i pi mr}) = e^{-i pi mr fnu} K(fnu, zn) - i pi mr I(fnu, zn)
template <class Y>
SPB_FN int acon(cplx z, real fnu, int kode, int mr, int n, Y y, real rl, real fnul, real tol,
                real elim, real alim) {
    const real pi = SPB_PI;
    int nz = 0;
    cplx zn = -z;
    int nn = n;
    int nw = binu(zn, fnu, kode, nn, y, rl, fnul, tol, elim, alim);
    if (nw < 0) return (nw == -2) ? -2 : -1;
    // continuation to the left half plane for the K function
    nn = n < 2 ? n : 2;
    cplx cy[2];
    nw = bknu(zn, fnu, kode, nn, cy, tol, elim, alim);
    if (nw != 0) return (nw == -2) ? -2 : -1;
    cplx s1 = cy[0];
    real fmr = real(mr);
    real sgn = (fmr < real(0.0)) ? pi : -pi;
    cplx csgn(0.0, sgn);
    if (kode != 1) {
        real yy = -zn.im;
        real sn, cs;
        sincos_(yy, sn, cs);
        csgn = csgn * cplx(cs, sn);
    }
    // cspn = exp(i sgn fnu) with the integer part folded out
    const OrderPhase ph = order_phase(fnu);
    cplx cspn(ph.cp, (fmr < real(0.0)) ? ph.sp : -ph.sp);
    if (ph.inu % 2 != 0) cspn = -cspn;
    int iuf = 0;
    cplx c1 = s1;
    cplx c2 = y[0];
    real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
    cplx sc1(0.0, 0.0), sc2(0.0, 0.0);
    if (kode != 1) {
        nw = s1s2(zn, c1, c2, ascle, alim, iuf);
        nz += nw;
        sc1 = c1;
    }
    y[0] = cspn * c1 + csgn * c2;
    if (n == 1) return nz;
    cspn = -cspn;
    cplx s2 = cy[1];
    c1 = s2;
    c2 = y[1];
    if (kode != 1) {
        nw = s1s2(zn, c1, c2, ascle, alim, iuf);
        nz += nw;
        sc2 = c1;
    }
    y[1] = cspn * c1 + csgn * c2;
    if (n == 2) return nz;
    cspn = -cspn;
    cplx rz = crecip(zn) * real(2.0);
    real fn = fnu + real(1.0);
    cplx ck = rz * fn;
    // scale near exponent extremes during the recurrence on K functions
    real cscl = real(1.0) / tol;
    real cscr = tol;
    real cssr[3] = {cscl, 1.0, cscr};
    real csrr[3] = {cscr, 1.0, cscl};
    real bry[3];
    bry[0] = ascle;
    bry[1] = real(1.0) / ascle;
    bry[2] = SPB_D1MACH2;
    real as2 = cabs_(s2);
    int kflag = 2;
    if (!(as2 > bry[0])) {
        kflag = 1;
    } else if (!(as2 < bry[1])) {
        kflag = 3;
    }
    real bscle = bry[kflag - 1];
    s1 = s1 * cssr[kflag - 1];
    s2 = s2 * cssr[kflag - 1];
    real csr = csrr[kflag - 1];
    for (int i = 3; i <= n; ++i) {
        cplx st = s2;
        s2 = ck * st + s1;
        s1 = st;
        c1 = s2 * csr;
        st = c1;
        c2 = y[i - 1];
        if (kode != 1 && iuf >= 0) {
            nw = s1s2(zn, c1, c2, ascle, alim, iuf);
            nz += nw;
            sc1 = sc2;
            sc2 = c1;
            if (iuf == 3) {
                iuf = -4;
                s1 = sc1 * cssr[kflag - 1];
                s2 = sc2 * cssr[kflag - 1];
                st = sc2;
            }
        }
        y[i - 1] = cspn * c1 + csgn * c2;
        ck = ck + rz;
        cspn = -cspn;
        if (kflag >= 3) continue;
        real c1m = rmax(rabs(c1.re), rabs(c1.im));
        if (c1m <= bscle) continue;
        kflag += 1;
        bscle = bry[kflag - 1];
        s1 = s1 * csr;
        s2 = st;
        s1 = s1 * cssr[kflag - 1];
        s2 = s2 * cssr[kflag - 1];
        csr = csrr[kflag - 1];
    }
    return nz;
}

// Multiply y by csgn with the underflow guard used after every rotation.
SPB_HD cplx rot_guard(cplx y, cplx csgn, real rtol, real ascle, real tol) {
    real aa = y.re, bb = y.im;
    // on scale (practically always): the plain product; the factors 1.0 of the guarded form are exact, so both
    // forms round identically
    if (rmax(rabs(aa), rabs(bb)) > ascle) return cplx(aa * csgn.re - bb * csgn.im, aa * csgn.im + bb * csgn.re);
    aa *= rtol;
    bb *= rtol;
    return cplx((aa * csgn.re - bb * csgn.im) * tol, (aa * csgn.im + bb * csgn.re) * tol);
}

// ---------------------------------------------------------------- I
template <class Y>
SPB_FN int besi(cplx z, real fnu, int kode, int n, Y cy, int &ierr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL, fnul = SPB_FNUL;
    ierr = 0;
    int nz = 0;
    if (fnu < real(0.0) || kode < 1 || kode > 2 || n < 1) {
        ierr = 1;
        return 0;
    }
    real az = cabs_(z);
    real fn = fnu + real(n - 1);
    real aa = SPB_RANGE_IK;
    if (az > aa || fn > aa) {
        ierr = 4;
        return 0;
    }
    aa = sqrt(aa);
    if (az > aa || fn > aa) ierr = 3;
    cplx zn = z;
    cplx csgn(1.0, 0.0);
    if (z.re < real(0.0)) {
        zn = -z;
        // csgn = exp(+-i pi fnu) with the integer part folded out
        const OrderPhase ph = order_phase(fnu);
        csgn = cplx(ph.cp, (z.im < real(0.0)) ? -ph.sp : ph.sp);
        if (ph.inu % 2 != 0) csgn = -csgn;
    }
    nz = binu(zn, fnu, kode, n, cy, rl, fnul, tol, elim, alim);
    if (nz < 0) {
        ierr = (nz == -2) ? 5 : 2;
        return 0;
    }
    if (!(z.re < real(0.0))) return nz;
    // analytic continuation to the left half plane
    int nn = n - nz;
    if (nn == 0) return nz;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * rtol * real(1.0e3);
    for (int i = 0; i < nn; ++i) {
        cy[i] = rot_guard(cy[i], csgn, rtol, ascle, tol);
        csgn = -csgn;
    }
    return nz;
}

// ---------------------------------------------------------------- J
template <class Y>
SPB_FN int besj(cplx z, real fnu, int kode, int n, Y cy, int &ierr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL, fnul = SPB_FNUL;
    const real hpi = SPB_HPI;
    ierr = 0;
    int nz = 0;
    if (fnu < real(0.0) || kode < 1 || kode > 2 || n < 1) {
        ierr = 1;
        return 0;
    }
    real az = cabs_(z);
    real fn = fnu + real(n - 1);
    real aa = SPB_RANGE_JYH;
    if (az > aa || fn > aa) {
        ierr = 4;
        return 0;
    }
    aa = sqrt(aa);
    if (az > aa || fn > aa) ierr = 3;
    // csgn = exp(i fnu pi/2) with the integer part folded out
    real cii = 1.0;
    const OrderPhase ph = order_phase(fnu);
    cplx csgn(ph.ch, ph.sh);
    if ((ph.inu / 2) % 2 == 1) csgn = -csgn;
    // zn = -i z is in the right half plane for Im z >= 0
    cplx zn(z.im, -z.re);
    if (z.im < real(0.0)) {
        zn = -zn;
        csgn.im = -csgn.im;
        cii = -cii;
    }
    nz = binu(zn, fnu, kode, n, cy, rl, fnul, tol, elim, alim);
    if (nz < 0) {
        ierr = (nz == -2) ? 5 : 2;
        return 0;
    }
    int nl = n - nz;
    if (nl == 0) return nz;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * rtol * real(1.0e3);
    for (int i = 0; i < nl; ++i) {
        cy[i] = rot_guard(cy[i], csgn, rtol, ascle, tol);
        csgn = cplx(-csgn.im * cii, csgn.re * cii);
    }
    return nz;
}

// ---------------------------------------------------------------- K
template <class Y>
SPB_FN int besk(cplx z, real fnu, int kode, int n, Y cy, int &ierr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL, fnul = SPB_FNUL;
    ierr = 0;
    int nz = 0;
    if ((z.re == real(0.0) && z.im == real(0.0)) || fnu < real(0.0) || kode < 1 || kode > 2 || n < 1) {
        ierr = 1;
        return 0;
    }
    int nn = n;
    real az = cabs_(z);
    real fn = fnu + real(nn - 1);
    real aa = SPB_RANGE_IK;
    if (az > aa || fn > aa) {
        ierr = 4;
        return 0;
    }
    aa = sqrt(aa);
    if (az > aa || fn > aa) ierr = 3;
    // overflow test on the last member of the sequence
    real ufl = SPB_D1MACH1 * real(1.0e3);
    if (az < ufl) {
        ierr = 2;
        return 0;
    }
    int nw;
    if (fn > fnul) {
        // uniform asymptotic expansions for fnu > fnul
        int mr = 0;
        if (z.re < real(0.0)) {
            mr = 1;
            if (z.im < real(0.0)) mr = -1;
        }
        nw = bunk(z, fnu, kode, mr, nn, cy, tol, elim, alim);
        if (nw < 0) {
            ierr = (nw == -1) ? 2 : 5;
            return 0;
        }
        return nz + nw;
    }
    if (fn > real(1.0)) {
        if (fn > real(2.0)) {
            int nuf = uoik(z, fnu, kode, 2, nn, cy, tol, elim, alim);
            if (nuf < 0) {
                ierr = 2;
                return 0;
            }
            nz += nuf;
            nn -= nuf;
            // here nn = n or nn = 0 since nuf = 0, nn or -1
            if (nn == 0) {
                if (z.re < real(0.0)) {
                    ierr = 2;
                    return 0;
                }
                return nz;
            }
        } else if (!(az > tol)) {
            real arg = real(0.5) * az;
            real aln = -fn * log(arg);
            if (aln > elim) {
                ierr = 2;
                return 0;
            }
        }
    }
    if (!(z.re < real(0.0))) {
        nw = bknu(z, fnu, kode, nn, cy, tol, elim, alim);
        if (nw < 0) {
            ierr = (nw == -1) ? 2 : 5;
            return 0;
        }
        return nw;
    }
    // left half plane: K is large there, an earlier underflow means overflow here
    if (nz != 0) {
        ierr = 2;
        return 0;
    }
    int mr = 1;
    if (z.im < real(0.0)) mr = -1;
    nw = acon(z, fnu, kode, mr, nn, cy, rl, fnul, tol, elim, alim);
    if (nw < 0) {
        ierr = (nw == -1) ? 2 : 5;
        return 0;
    }
    return nw;
}

// ---------------------------------------------------------------- H1 / H2
template <class Y>
SPB_FN int besh(cplx z, real fnu, int kode, int m, int n, Y cy, int &ierr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL, fnul = SPB_FNUL;
    const real hpi = SPB_HPI;
    ierr = 0;
    int nz = 0;
    if ((z.re == real(0.0) && z.im == real(0.0)) || fnu < real(0.0) || m < 1 || m > 2 || kode < 1 ||
        kode > 2 || n < 1) {
        ierr = 1;
        return 0;
    }
    int nn = n;
    real fn = fnu + real(nn - 1);
    int mm = 3 - m - m;
    real fmm = real(mm);
    cplx zn(fmm * z.im, -fmm * z.re);  // -i mm z
    real az = cabs_(z);
    real aa = SPB_RANGE_JYH;
    if (az > aa || fn > aa) {
        ierr = 4;
        return 0;
    }
    aa = sqrt(aa);
    if (az > aa || fn > aa) ierr = 3;
    // overflow test on the last member of the sequence
    real ufl = SPB_D1MACH1 * real(1.0e3);
    if (az < ufl) {
        ierr = 2;
        return 0;
    }
    int nw;
    if (fnu > fnul) {
        // uniform asymptotic expansions for fnu > fnul
        int mr = 0;
        if (!((zn.re >= real(0.0)) && (zn.re != real(0.0) || zn.im >= real(0.0) || m != 2))) {
            mr = -mm;
            if (zn.re == real(0.0) && zn.im < real(0.0)) zn = -zn;
        }
        nw = bunk(zn, fnu, kode, mr, nn, cy, tol, elim, alim);
        if (nw < 0) {
            ierr = (nw == -1) ? 2 : 5;
            return 0;
        }
        nz += nw;
    } else {
        bool skip = false;
        if (fn > real(1.0)) {
            if (fn > real(2.0)) {
                int nuf = uoik(zn, fnu, kode, 2, nn, cy, tol, elim, alim);
                if (nuf < 0) {
                    ierr = 2;
                    return 0;
                }
                nz += nuf;
                nn -= nuf;
                if (nn == 0) {
                    if (zn.re < real(0.0)) {
                        ierr = 2;
                        return 0;
                    }
                    return nz;
                }
            } else if (!(az > tol)) {
                real arg = real(0.5) * az;
                real aln = -fn * log(arg);
                if (aln > elim) {
                    ierr = 2;
                    return 0;
                }
            }
        }
        (void)skip;
        if ((zn.re < real(0.0)) || (zn.re == real(0.0) && zn.im < real(0.0) && m == 2)) {
            // left half plane
            int mr = -mm;
            nw = acon(zn, fnu, kode, mr, nn, cy, rl, fnul, tol, elim, alim);
            if (nw < 0) {
                ierr = (nw == -1) ? 2 : 5;
                return 0;
            }
            nz = nw;
        } else {
            // right half plane, xn >= 0 and (xn != 0 or yn >= 0 or m = 1)
            nz = bknu(zn, fnu, kode, nn, cy, tol, elim, alim);
            if (nz < 0) {
                ierr = (nz == -1) ? 2 : 5;
                return 0;
            }
        }
    }
    // H(m,fnu,z) = -fmm (i/hpi) zt^fnu K(fnu, -z zt), zt = exp(-fmm hpi i)
    const real tpi = 0.63661977236758134308;  // 2/pi
    const OrderPhase ph = order_phase(fnu);
    cplx csgn(-tpi * ph.sh, (m == 1) ? -tpi * ph.ch : tpi * ph.ch);
    if ((ph.inu / 2) % 2 == 1) csgn = -csgn;
    real zti = -fmm;
    real rtol = real(1.0) / tol;
    real ascle = ufl * rtol;
    for (int i = 0; i < nn; ++i) {
        cy[i] = rot_guard(cy[i], csgn, rtol, ascle, tol);
        csgn = cplx(-csgn.im * zti, csgn.re * zti);
    }
    return nz;
}

// ---------------------------------------------------------------- Y  (n = 1 per call)
// Y = (H1 - H2) / (2i); kode = 2 scales by exp(-|Im z|).
SPB_FN int besy(cplx z, real fnu, int kode, cplx *cy, int &ierr) {
    const real tol = SPB_TOL, elim = SPB_ELIM;
    const real hci = 0.5;
    ierr = 0;
    if ((z.re == real(0.0) && z.im == real(0.0)) || fnu < real(0.0) || kode < 1 || kode > 2) {
        ierr = 1;
        return 0;
    }
    cplx h1[1], h2[1];
    int ie;
    int nz1 = besh(z, fnu, kode, 1, 1, h1, ie);
    if (ie != 0 && ie != 3) {
        ierr = ie;
        return 0;
    }
    ierr = ie;
    int nz2 = besh(z, fnu, kode, 2, 1, h2, ie);
    if (ie != 0 && ie != 3) {
        ierr = ie;
        return 0;
    }
    ierr = ie;
    int nz = nz1 < nz2 ? nz1 : nz2;
    if (kode != 2) {
        cplx d = h2[0] - h1[0];
        cy[0] = cplx(-d.im * hci, d.re * hci);
        return nz;
    }
    real exr, exi;
    sincos_(z.re, exi, exr);
    real ey = 0.0;
    real tay = rabs(z.im + z.im);
    if (tay < elim) ey = exp_(-tay);
    cplx c1, c2;
    if (z.im < real(0.0)) {
        c1 = cplx(exr, exi);
        c2 = cplx(exr * ey, -exi * ey);
    } else {
        c1 = cplx(exr * ey, exi * ey);
        c2 = cplx(exr, -exi);
    }
    nz = 0;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * rtol * real(1.0e3);
    cplx a = rot_guard(h2[0], c2, rtol, ascle, tol);
    cplx b = rot_guard(h1[0], c1, rtol, ascle, tol);
    cplx st = a - b;
    cy[0] = cplx(-st.im * hci, st.re * hci);
    if (st.re == real(0.0) && st.im == real(0.0) && ey == real(0.0)) nz += 1;
    return nz;
}

}  // namespace spb
