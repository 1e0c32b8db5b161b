// spb_family.h — the six Bessel families re-expressed for the binned GPU
// pipeline:
//
//     prepare(z, v)  ->  one right-half-plane point w, which of I_v(w) / K_v(w)
//                        are needed, and whether the point is "simple"
//     I pass         ->  I_v(w) by exactly one region routine (no fall-through)
//     K pass         ->  K_v(w) by the K region routine
//     finish         ->  rotation / analytic continuation / Hankel difference
//
// A point is "simple" when (a) the over/underflow pre-test (uoik) is provably a
// no-op, (b) its region routine completes without handing over to another
// region, and (c) the order is below the uniform-K threshold.  Everything else
// is routed to the general per-point driver (spb_drivers.h) on the GPU, so both
// paths produce the same numbers; the split only removes divergence.
//
// J and I follow "ZBESJ"/"ZBESI"; K, H1, H2 follow "ZBESK"/"ZBESH" with the
// continuation "ZACON"; Y = (H1 - H2)/(2i) shares ONE (I, K) pair between its
// two Hankel terms because both reduce to the same right-half-plane point.
#pragma once
#include "spb_common.h"
#include "spb_binu.h"
#include "spb_drivers.h"
#include "spb_sweep.h"

namespace spb {

// FAM_JY evaluates J and Y together: both need I_v at the same right-half-plane point, so the I pass is shared.
// FAM_IK evaluates I and K together: K in the left half plane is continued with I at the same right-half-plane point,
// and the single-evaluation routines (large-|w| series, Debye-direct region) produce both from one series.
enum Family { FAM_J = 0, FAM_Y = 1, FAM_I = 2, FAM_K = 3, FAM_H1 = 4, FAM_H2 = 5, FAM_JY = 6, FAM_IK = 7 };

// how finish() combines the pass results
enum Combine {
    CMB_FINAL = 0,    // prepare already produced the result (error / trivial)
    CMB_GENERAL = 1,  // not simple: evaluate with the per-point driver
    CMB_I_ROT = 2,    // J / I: rotate I(w)
    CMB_K_ROT = 3,    // K / H with w in the right half plane: rotate K(w)
    CMB_ACON = 4,     // K / H in the left half plane: continuation from I(w), K(w)
    CMB_Y_SHARED = 5, // Y from one (I(w), K(w)) pair
    CMB_Y_CONJ = 6    // Y on the positive real axis: K(w) and its conjugate
};

// `fuse` argument of prepare() / the classifier: which of the single-evaluation region routines of the binned
// pipeline are enabled (the classic per-point drivers know neither)
enum { FUSE_ASYI = 1,   // I and K of a K-type point in the large-|w| region of I from one Hankel series (fused_ik_region)
       FUSE_DEBYE = 2,  // Debye-direct region (spb_debye.h) for I, K or both
       FUSE_WRSK = 4 };  // K-type points whose I region is Miller + Wronskian: that routine evaluates K_v(w), K_v+1(w)
                         // for its normalisation, so K_v(w) is taken from there instead of from a second evaluation
                         // in the K pass (wrsk_ik_region)

// K-pass regions
enum KRegion { KREG_SERIES = 0, KREG_MILLER = 1, KREG_HALF = 2, KREG_COUNT = 3 };
// kreg of a point whose K_v(w) is evaluated together with I_v(w) by the fused large-|w| routine (no K bin)
enum { KREG_FUSED = -2 };

struct Prep {
    cplx w;       // right-half-plane evaluation point
    int combine;  // Combine
    int ireg;     // IRegion for the I pass, -1 if I is not needed
    int kreg;     // KRegion for the K pass, -1 if K is not needed
    int ierr;     // preliminary ierr (0 or 3)
    int mr;       // continuation direction (+-1)
    real az;      // |z| = |w| (rotations and reflections keep the modulus bit for bit): shared by every pass
    int final_ierr, final_nz;  // CMB_FINAL: the result is 0 with these codes
};

// Conservative bound on |Re(-zeta1 + zeta2 -+ w)| of the leading uniform term for
// order gnu at w (x = |Re w|, az = |w|): if it stays below alim, uoik() returns 0 without touching y.
//   zeta2 = sqrt(gnu^2 + w^2):  Re zeta2 <= sqrt(gnu^2 + x^2) <= gnu + x   (if Re zeta2 were larger, |Im zeta2|
//           would exceed |Im w| and Re zeta2 Im zeta2 = x Im w could not hold);
//   zeta1 = gnu ln((gnu + zeta2)/w):  gnu/az <= |(gnu + zeta2)/w| <= 1 + 2 gnu/az, so with t = az/gnu
//           |Re zeta1| <= gnu ln(1 + 2/t + t);  pi gnu covers the branch bookkeeping of the Airy-type form;
// The classifier evaluates these tests for every point of a batch, in warps whose lanes hold unrelated points: they
// are written as straight-line code (every predicate computed, combined with & and |, regions chosen by selection)
// so that a warp does not walk the decision tree once per distinct path; only the logarithm of the full bound sits
// behind a branch, taken by the rare lanes that need it.  (Measured on config 3 before this form: 435 warp
// instructions per point, 24 of 32 lanes active.)
SPB_HD bool uoik_noop_cheap(real az, real gnu, real alim, real x) {
    // cheap sufficient tests (they cover the BASELINE configs without a division or a logarithm):
    // for 1/64 <= t = az/gnu <= 64, ln(1 + 2/t + t) <= ln(129.02) < 4.8601, so b < 9.0017 gnu + x;
    // for 2^-20 <= t <= 2^20, ln(1 + 2/t + t) < 15.3
    const bool a = (az >= gnu * real(0.015625)) & (az <= gnu * real(64.0)) & (gnu * real(9.0017) + x < alim);
    const bool b = (az >= gnu * real(9.5367431640625e-07)) & (az <= gnu * real(1048576.0)) &
                   (gnu * real(19.5) + x < alim);
    return a | b;
}
SPB_HD bool uoik_noop_full(real az, real gnu, real alim, real x) {
    real t = az / gnu;
    real b = gnu * (log(real(1.0) + real(2.0) / t + t) + real(SPB_PI) + real(1.0)) + x;
    return b < alim;
}
SPB_HD bool uoik_is_noop(real az, real gnu, real alim, real x, int kode) {
    // scaled variants test Re(zeta2 - zeta1 - w): x <= Re zeta2 <= x + gnu, so the x terms cancel
    if (kode != 1) x = real(0.0);
    bool ok = uoik_noop_cheap(az, gnu, alim, x);
    if (!ok) ok = uoik_noop_full(az, gnu, alim, x);
    return ok;
}

SPB_HD int kregion(cplx w, real fnu, real az_known = real(-1.0)) {
    const int inu = (int)(fnu + real(0.5));
    const real dnu = fnu - real(inu);
    const real az = az_known >= real(0.0) ? az_known : cabs_(w);
    const int r = (az > real(2.0)) ? KREG_MILLER : KREG_SERIES;
    return (rabs(dnu) == real(0.5)) ? KREG_HALF : r;
}

// binu_region() (spb_binu.h) for a single member, by selection: same comparisons, same result
SPB_HD int binu_region_sel(cplx z, real fnu, real rl, real fnul, real az) {
    const bool series = (az <= real(2.0)) | !(az * az * real(0.25) > fnu + real(1.0));
    const bool big = !(az < rl);
    const bool low_order = fnu <= real(1.0);
    const bool asy = big & (low_order | !(az + az < fnu * fnu));
    const bool direct = !big & low_order;
    const bool uni = (fnu > fnul) | (az > fnul);
    const bool form1 = !(rabs(z.im) > rabs(z.re) * real(1.7321));
    const int r_uni = form1 ? IREG_UNI1 : (wrsk_instead_of_uni2(az, fnu, fnul) ? IREG_WRSK : IREG_UNI2);
    const int r_rest = (az > rl) ? IREG_WRSK : IREG_MLRI;
    return series ? IREG_SERIES : asy ? IREG_ASYI : direct ? IREG_MLRI : uni ? r_uni : r_rest;
}

// Is the I evaluation at (w, fnu) simple, and in which region?  debye: carve the Debye-direct region out of the
// Miller / Wronskian / raised-order regions (only where the point is simple: both over/underflow pre-tests no-ops,
// which also keeps the exponent of the expansion on the middle scaling level).
SPB_HD int iregion_simple(cplx w, real fnu, real az, bool &simple, int kode, bool debye = false) {
    const real rl = SPB_RL, fnul = SPB_FNUL, alim = SPB_ALIM;
    const int r = binu_region_sel(w, fnu, rl, fnul, az);
    const bool iterative = (r >= IREG_MLRI) & (r <= IREG_UNI2);
    // these are entered through uoik(ikflg=1) unless dfnu <= 1 took the direct Miller route
    const bool direct = (r == IREG_MLRI) & (az < rl) & (fnu <= real(1.0));
    const bool need1 = iterative & !direct;
    const bool need2 = (r == IREG_WRSK);  // second pre-test on K_fnu, K_fnu+1
    const real x = (kode != 1) ? real(0.0) : rabs(w.re);
    const real gnu1 = rmax(fnu, real(1.0)), gnu2 = rmax(fnu + real(1.0), real(2.0));
    bool ok1 = uoik_noop_cheap(az, gnu1, alim, x), ok2 = uoik_noop_cheap(az, gnu2, alim, x);
    if (need1 & !ok1) ok1 = uoik_noop_full(az, gnu1, alim, x);
    if (need2 & !ok2) ok2 = uoik_noop_full(az, gnu2, alim, x);
    simple = (!need1 | ok1) & (!need2 | ok2);
    return (debye & simple & iterative & debye_ok(w, fnu)) ? IREG_DEBYE : r;
}

// Geometry of a family at z: the right-half-plane evaluation point w, the continuation direction and the way the
// pass results combine, assuming the point turns out simple.  Returns whether I_v(w) is needed.  Shared by the
// classifier (prepare) and by the region passes (prepare_pass), which skip the region tests.
SPB_HD bool prepare_geometry(int fam, cplx z, Prep &p) {
    if (fam == FAM_J || fam == FAM_I) {
        if (fam == FAM_J) {
            p.w = cplx(z.im, -z.re);
            if (z.im < real(0.0)) p.w = -p.w;
        } else if (z.re < real(0.0)) {
            p.w = -z;
        }
        p.combine = CMB_I_ROT;
        return true;
    }
    bool left;
    if (fam == FAM_K || fam == FAM_IK) {
        left = z.re < real(0.0);
        p.w = left ? -z : z;
        p.mr = (z.im < real(0.0)) ? -1 : 1;
        p.combine = left ? CMB_ACON : CMB_K_ROT;
        if (fam == FAM_IK) left = true;  // I_v(w) is an output everywhere
    } else if (fam == FAM_H1 || fam == FAM_H2) {
        int m = (fam == FAM_H1) ? 1 : 2;
        int mm = 3 - m - m;
        real fmm = real(mm);
        cplx zn(fmm * z.im, -fmm * z.re);
        left = (zn.re < real(0.0)) || (zn.re == real(0.0) && zn.im < real(0.0) && m == 2);
        p.w = left ? -zn : zn;
        p.mr = -mm;
        p.combine = left ? CMB_ACON : CMB_K_ROT;
    } else {  // FAM_Y, FAM_JY
        // H1 uses zn1 = -i z, H2 uses zn2 = +i z = -zn1
        cplx zn1(z.im, -z.re);
        if (z.im == real(0.0) && z.re > real(0.0)) {
            // both Hankel terms sit on the imaginary axis of the right half plane: conjugate points
            p.w = zn1;
            p.combine = CMB_Y_CONJ;
            left = false;
        } else {
            bool left1 = zn1.re < real(0.0);  // H1 continued, H2 direct
            p.w = left1 ? -zn1 : zn1;
            p.mr = left1 ? -1 : 1;  // mr = -mm of the continued term
            p.combine = CMB_Y_SHARED;
            left = true;
        }
    }
    if (fam == FAM_JY) left = true;  // J needs I(w) even where Y alone would not (positive real axis)
    return left;
}

// fuse: K-type points whose w lies in the large-|w| region of I get kreg = KREG_FUSED instead of a K bin: one
// routine evaluates I_v(w) and K_v(w) from the same series, sharing 1/w, 1/sqrt(w), the sincos of Im w and
// exp(+-Re w) (also where the family needs K only: the series is several times cheaper than the Miller algorithm).
SPB_FN Prep prepare(int fam, cplx z, real fnu, int kode, int fuse = 0) {
    const real fnul = SPB_FNUL, alim = SPB_ALIM, tol = SPB_TOL;
    Prep p;
    p.w = z;
    p.combine = CMB_GENERAL;
    p.ireg = -1;
    p.kreg = -1;
    p.ierr = 0;
    p.mr = 0;
    p.final_ierr = 0;
    p.final_nz = 0;
    // anything unusual (invalid input, NaN, huge or tiny arguments) goes to the driver
    real az = cabs_(z);
    p.az = az;
    const bool odd = !(fnu >= real(0.0)) | !(az == az) | !(az < real(32767.0)) | !(fnu < real(32767.0));
    if (fam == FAM_J || fam == FAM_I) {
        if (odd | (az == real(0.0))) return p;
        prepare_geometry(fam, z, p);
        p.combine = CMB_GENERAL;
        bool simple;
        p.ireg = iregion_simple(p.w, fnu, az, simple, kode, (fuse & FUSE_DEBYE) != 0);
        if (simple) p.combine = CMB_I_ROT;
        return p;
    }
    // K-type families
    if (odd | (az < real(1.0e-290)) | (fnu > fnul) | (!(fnu > real(2.0)) & (fnu > real(1.0)) & !(az > tol))) return p;
    const bool left = prepare_geometry(fam, z, p);
    // over/underflow pre-test of the K drivers (made at the rotated point) must be a no-op
    const real gnu1 = rmax(fnu, real(1.0));
    const bool final_fam = (kode == 1) & ((fam == FAM_K) | (fam == FAM_H1) | (fam == FAM_H2) | (fam == FAM_Y));
    // Far beyond the underflow limit the outcome below is decided without the logarithm: for 2^-20 <= |w|/gnu <= 2^20
    // ln(1 + 2/t + t) < 15.3, so |Re w| - gnu (15.3 + pi) > elim implies the provable-underflow test of the block
    // below (and |Re w| > elim > alim makes every form of the no-op test fail, so the block is entered).  On an
    // unscaled log-uniform batch one point in seven is of this kind.
    if (final_fam & (fnu > real(2.0)) & (az >= gnu1 * real(9.5367431640625e-07)) & (az <= gnu1 * real(1048576.0)) &
        (rabs(p.w.re) - gnu1 * real(18.4416) > real(SPB_ELIM))) {
        const bool overflow = (p.combine == CMB_ACON) || fam == FAM_Y;
        p.combine = CMB_FINAL;
        p.final_ierr = overflow ? 2 : 0;
        p.final_nz = overflow ? 0 : 1;
        return p;
    }
    if (fnu > real(2.0)) {
        const real x = (kode != 1) ? real(0.0) : rabs(p.w.re);
        if (!uoik_noop_cheap(az, gnu1, alim, x)) {
            // ONE logarithm serves the full bound of the no-op test (uoik_noop_full: same expression) and the test
            // for a provable underflow of K_v(w): Re(zeta2 - zeta1) >= |Re w| - gnu (ln(1 + 2/t + t) + pi) > elim makes
            // the pre-test zero the whole (single-member) sequence.  The K and Hankel drivers then return 0 with
            // nz = 1 in the right half plane and report an overflow (ierr = 2) in the left one, where the vanished K
            // would have been continued with the growing I.  (Unscaled functions only: the scaled pre-test has no
            // |Re w|.)  Y = (H1 - H2)/2i evaluates both Hankel functions at the same |Re w|: one of them vanishes,
            // the other one is the overflow, so Y reports ierr = 2.
            const real t = az / gnu1;
            const real lt = log(real(1.0) + real(2.0) / t + t);
            const bool noop = gnu1 * (lt + real(SPB_PI) + real(1.0)) + x < alim;
            if (!noop) {
                if (final_fam && rabs(p.w.re) - gnu1 * (lt + real(SPB_PI)) > real(SPB_ELIM)) {
                    const bool overflow = (p.combine == CMB_ACON) || fam == FAM_Y;
                    p.combine = CMB_FINAL;
                    p.final_ierr = overflow ? 2 : 0;
                    p.final_nz = overflow ? 0 : 1;
                    return p;
                }
                p.combine = CMB_GENERAL;
                return p;
            }
        }
    }
    const int kreg = kregion(p.w, fnu, az);
    const bool debye_on = (fuse & FUSE_DEBYE) != 0;
    bool simple;
    const int ri = iregion_simple(p.w, fnu, az, simple, kode, debye_on);
    if (left) {
        // I and K from one evaluation (large-|w| series / Debye expansion / Wronskian route): no K bin
        const bool fused = (((fuse & FUSE_ASYI) != 0) & (ri == IREG_ASYI)) | (ri == IREG_DEBYE) |
                           (((fuse & FUSE_WRSK) != 0) & (ri == IREG_WRSK));
        p.combine = simple ? p.combine : (int)CMB_GENERAL;
        p.ireg = simple ? ri : -1;
        p.kreg = simple ? (fused ? (int)KREG_FUSED : kreg) : -1;
    } else {
        // I_v(w) is not needed here, but the point lies in a region where one of the single-evaluation routines
        // gives K at a fraction of the cost of the Miller algorithm + forward recurrence: same bin, the combine
        // step ignores I
        const int r = binu_region_sel(p.w, fnu, SPB_RL, SPB_FNUL, az);
        const bool asy = ((fuse & FUSE_ASYI) != 0) & (r == IREG_ASYI);
        const bool dby = !asy & debye_on & (r >= IREG_MLRI) & (r <= IREG_UNI2) & debye_ok(p.w, fnu);
        p.ireg = asy ? (int)IREG_ASYI : dby ? (int)IREG_DEBYE : -1;
        p.kreg = (asy | dby) ? (int)KREG_FUSED : kreg;
    }
    return p;
}

// Fast accept of the classifier for the bulk of a large-|z| batch: decides from |z|^2 (no square root) and a
// handful of comparisons whether prepare() is certain to return a simple point whose I region (if I_v(w) is
// needed) is the large-|w| asymptotic one, and which K region it has.  Every comparison of prepare() that
// involves the modulus is replaced by a sufficient condition on m2 = x^2 + y^2 with a relative margin of 1e-9
// (rounding of m2 and of its square root is below 1e-15), so an accepted point is classified exactly as
// prepare() would; everything near a region boundary, every unusual argument and every other region is
// declined and goes through prepare().  tests/test_pipeline_cpu.py checks the equivalence on every distribution.
// The order-dependent half of the test (computed once per thread when the order is a broadcast scalar): the
// window of m2 inside which every modulus comparison of prepare() is decided, and the K region.
struct FastOrder {
    bool ok;          // 0 <= fnu <= fnul
    bool pretest;     // fnu > 2: the over/underflow pre-test applies
    real m2_lo, m2_hi;
    real c19;         // 19.5 fnu, the order term of the cheap sufficient form of uoik_is_noop()
    int kreg;         // KREG_HALF or KREG_MILLER (|z| > 2 inside the window)
};
SPB_HD FastOrder fast_order(real fnu) {
    const real fnul = SPB_FNUL;
    const real up = real(1.0) + real(1.0e-9), dn = real(1.0) - real(1.0e-9);
    FastOrder o;
    o.ok = (fnu >= real(0.0)) && (fnu <= fnul);
    o.pretest = fnu > real(2.0);
    // rl <= |z| < 32767
    real lo = real(SPB_RL * SPB_RL) * up, hi = real(32767.0 * 32767.0) * dn;
    // large-|w| region of I: not the power series (|z|^2/4 > fnu + 1), and fnu <= 1 or 2|z| >= fnu^2
    lo = rmax(lo, (fnu + real(1.0)) * real(4.0) * up);
    const real f2 = fnu * fnu;
    if (!(fnu <= real(1.0))) lo = rmax(lo, f2 * f2 * real(0.25) * up);
    if (o.pretest) {
        // fnu 2^-20 <= |z| <= fnu 2^20
        const real a = fnu * real(9.5367431640625e-07), b = fnu * real(1048576.0);
        lo = rmax(lo, a * a * up);
        hi = rmin(hi, b * b * dn);
    }
    o.m2_lo = lo;
    o.m2_hi = hi;
    o.c19 = fnu * real(19.5);
    const int inu = (int)(fnu + real(0.5));
    o.kreg = (rabs(fnu - real(inu)) == real(0.5)) ? KREG_HALF : KREG_MILLER;
    return o;
}
SPB_HD bool fast_classify(int fam, cplx z, const FastOrder &o, int kode, int fuse, int &ireg, int &kreg) {
    const real alim = SPB_ALIM;
    const real m2 = z.re * z.re + z.im * z.im;
    if (!o.ok || !(m2 >= o.m2_lo) || !(m2 <= o.m2_hi)) return false;  // (a NaN fails the comparisons)
    if (fam == FAM_J || fam == FAM_I) {
        ireg = IREG_ASYI;
        kreg = -1;
        return true;
    }
    // K-type families: which half plane the evaluation point lies in, |Re w| for the pre-test
    bool left;
    real x;
    if (fam == FAM_K || fam == FAM_IK) {
        left = (fam == FAM_IK) || z.re < real(0.0);
        x = rabs(z.re);
    } else if (fam == FAM_H1 || fam == FAM_H2) {
        const real fmm = (fam == FAM_H1) ? real(1.0) : real(-1.0);
        const real znr = fmm * z.im, zni = -fmm * z.re;
        left = (znr < real(0.0)) || (znr == real(0.0) && zni < real(0.0) && fam == FAM_H2);
        x = rabs(z.im);
    } else {
        left = (fam == FAM_JY) || !(z.im == real(0.0) && z.re > real(0.0));
        x = rabs(z.im);
    }
    if (o.pretest) {
        if (kode != 1) x = real(0.0);
        if (!(o.c19 + x < alim)) return false;
    }
    kreg = o.kreg;
    ireg = left ? IREG_ASYI : -1;
    if (fuse & FUSE_ASYI) {
        // the window is the large-|w| region of I: K comes from the same series whether or not I is needed
        ireg = IREG_ASYI;
        kreg = KREG_FUSED;
    }
    return true;
}
SPB_HD bool fast_classify(int fam, cplx z, real fnu, int kode, int fuse, int &ireg, int &kreg) {
    const FastOrder o = fast_order(fnu);
    return fast_classify(fam, z, o, kode, fuse, ireg, kreg);
}

// class byte of a point: bits 0-2 I bin (7 none), 3-4 K bin (3 none, also for KREG_FUSED), 5 general
SPB_HD unsigned int class_of(int ireg, int kreg) {
    const unsigned int bi = ireg >= 0 ? (unsigned int)ireg : 7u;
    const unsigned int bk = kreg >= 0 ? (unsigned int)kreg : 3u;
    return bi | (bk << 3);
}
// class byte of a prepared point (see class_byte); CMB_FINAL points carry no bin and no general flag
SPB_HD unsigned int class_of_prep(const Prep &p) {
    if (p.combine == CMB_FINAL) return 7u | (3u << 3);
    if (p.combine == CMB_GENERAL) return 7u | (3u << 3) | (1u << 5);
    return class_of(p.ireg, p.kreg);
}
SPB_HD unsigned int class_byte(int fam, cplx z, real fnu, int kode, int fuse, bool allow_fast = true) {
    int ireg, kreg;
    if (!(allow_fast && fast_classify(fam, z, fnu, kode, fuse, ireg, kreg))) {
        const Prep p = prepare(fam, z, fnu, kode, fuse);
        // (CMB_FINAL: decided by prepare() -- the classifier kernel stores the result itself, no bin at all)
        if (p.combine == CMB_FINAL) return 7u | (3u << 3);
        if (p.combine == CMB_GENERAL) return 7u | (3u << 3) | (1u << 5);
        ireg = p.ireg;
        kreg = p.kreg;
    }
    return class_of(ireg, kreg);
}

// prepare() for a point the classifier has already binned: the region tests are known to have passed, so only
// the geometry and the modulus are recomputed.  has_i = the class byte carries an I bin.
SPB_FN Prep prepare_pass(int fam, cplx z, bool with_modulus = true) {
    Prep p;
    p.w = z;
    p.combine = CMB_GENERAL;
    p.kreg = -1;
    p.ierr = 0;
    p.mr = 0;
    p.az = with_modulus ? cabs_(z) : real(-1.0);  // (the fused routine computes |w| with its reciprocal)
    // a binned point carries an I bin exactly when its geometry needs I(w) (prepare() sets ireg under `left`)
    p.ireg = prepare_geometry(fam, z, p) ? 0 : -1;  // the passes only ask whether I(w) exists
    return p;
}

// One region of the I pass.  Returns the AMOS nz (>= 0) or < 0 when the region
// wants to hand over (-> general path).
SPB_FN int ipass_region(int region, cplx w, real fnu, int kode, cplx &val, real az = real(-1.0),
                        const OrderPhase *oph = nullptr, const WShared *ws = nullptr, const real *gl = nullptr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL;
    cplx y[1];
    y[0] = cplx(0.0, 0.0);
    int nw;
    switch (region) {
        case IREG_SERIES: nw = seri(w, fnu, kode, 1, y, tol, elim, alim, az, gl); break;
        case IREG_ASYI: nw = asyi(w, fnu, kode, 1, y, rl, tol, elim, alim, az, oph, ws); break;
        case IREG_MLRI: nw = mlri(w, fnu, kode, 1, y, tol, az, gl); break;
        case IREG_WRSK: nw = wrsk<true>(w, fnu, kode, 1, y, tol, elim, alim, az); break;
        case IREG_UNI1:
        case IREG_UNI2: {
            // order raised to fnul, uniform expansion there, backward recurrence (binu()'s route for
            // |z| > fnul or fnu > fnul once the over/underflow pre-test is known to be a no-op)
            const real fnul = SPB_FNUL;
            int nui = (int)(fnul - fnu) + 1;
            if (nui < 0) nui = 0;
            int nlast = 0;
            nw = buni(w, fnu, kode, 1, y, nui, nlast, fnul, tol, elim, alim, region == IREG_UNI1 ? 1 : 2);
            if (nw != 0 || nlast != 0) nw = -1;  // the driver would continue with another method
            break;
        }
        default: nw = -9; break;
    }
    val = y[0];
    return nw;
}

SPB_FN int kpass_region(cplx w, real fnu, int kode, cplx &val, real az = real(-1.0), const WShared *ws = nullptr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    // binned points stay on the middle scaling level (their pre-test is a no-op): the lean form of the routine
    return bknu_simple(w, fnu, kode, val, tol, elim, alim, az, ws);
}

// I_v(w) and K_v(w) of a KREG_FUSED point (large-|w| region of I, |w| >= rl and 2|w| >= fnu^2 or fnu <= 1).  Both
// functions have the same Hankel coefficients a_k(fnu) there,
//     I_v(w) ~ e^w  / sqrt(2 pi w) sum (-1)^k a_k / w^k  (+ the second exponential),
//     K_v(w) ~ e^-w sqrt(pi / 2w)  sum        a_k / w^k,
// so the K series is the all-positive sum the I routine forms anyway for its second exponential: K costs one
// complex product on top of I instead of the Miller algorithm the published region map runs here (third
// departure from it, after the two of DESIGN.md section 2; |arg w| <= pi/2, so the remainder of the K series is
// bounded by twice the first neglected term, DLMF 10.40(iii), and the convergence test of the I series bounds that
// term by tol).  Shared quantities (1/w, 1/sqrt(w), sincos(Im w), exp(+-Re w)) are computed once.  Returns < 0
// when the point has to go to the general path.
SPB_FN int fused_ik_region(cplx w, real fnu, int kode, cplx &ival, int &inz, cplx &kval, int &knz, real az,
                           const OrderPhase *oph = nullptr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM, rl = SPB_RL;
    const real rthpi = 1.25331413731550025;  // sqrt(pi/2)
    // (az < 0: |w| not computed yet, see w_shared)
    const WShared ws = w_shared(w, az, kode == 1);
    cplx y[1], ksum(1.0, 0.0);
    y[0] = cplx(0.0, 0.0);
    inz = asyi(w, fnu, kode, 1, y, rl, tol, elim, alim, ws.az, oph, &ws, &ksum);
    ival = y[0];
    knz = 0;
    if (inz < 0) return inz;
    cplx coef = ws.rsq * rthpi;
    if (kode == 1) {
        if (w.re > alim) return -9;  // exp(-Re w) leaves the middle scaling level: like the lean K routine
        coef = coef * cplx(ws.em * ws.cs, -ws.em * ws.sn);
    }
    kval = coef * ksum;
    return 0;
}

// I_v(w) and K_v(w) of a K-type point whose I region is Miller + Wronskian (FUSE_WRSK): the Wronskian
// normalisation needs K_v(w) and K_v+1(w) anyway (same lean K routine, same arguments as the K pass would use), so
// K_v(w) is a by-product.  Returns < 0 when the point has to go to the general path.
SPB_FN int wrsk_ik_region(cplx w, real fnu, int kode, cplx &ival, int &inz, cplx &kval, int &knz, real az) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    cplx y[1];
    y[0] = cplx(0.0, 0.0);
    kval = cplx(0.0, 0.0);
    inz = 0;
    knz = 0;
    const int nw = wrsk<true>(w, fnu, kode, 1, y, tol, elim, alim, az, &kval);
    ival = y[0];
    return nw;
}

// both outputs of a KREG_FUSED point outside the Debye-direct region, by the routine of its I region
SPB_FN int fused_region(int ireg, cplx w, real fnu, int kode, cplx &ival, int &inz, cplx &kval, int &knz, real az,
                        const OrderPhase *oph = nullptr) {
    if (ireg == IREG_WRSK) return wrsk_ik_region(w, fnu, kode, ival, inz, kval, knz, az);
    return fused_ik_region(w, fnu, kode, ival, inz, kval, knz, az, oph);
}

// I_v(w) and / or K_v(w) of a point in the Debye-direct region (IREG_DEBYE): one evaluation of the uniform
// expansion at the order itself (spb_debye.h).  need_i: the family needs I_v(w) at this point.
SPB_FN int debye_region(int fam, cplx w, real fnu, int kode, bool need_i, cplx &ival, int &inz, cplx &kval, int &knz,
                        real az, const OrderPhase *oph = nullptr) {
    const bool need_k = !(fam == FAM_J || fam == FAM_I);
    inz = 0;
    knz = 0;
    ival = cplx(0.0, 0.0);
    kval = cplx(0.0, 0.0);
    if (az < real(0.0)) az = cabs_(w);
    return debye_ik(w, fnu, kode, need_i, need_k, ival, kval, az, oph);
}

// continuation K(w e^{i pi mr}) from (K(w), I(w)) for one member (acon with n = 1)
SPB_FN cplx acon_combine(cplx w, const OrderPhase &ph, int kode, int mr, cplx kval, cplx ival, int &nz) {
    const real tol = SPB_TOL, alim = SPB_ALIM, pi = SPB_PI;
    nz = 0;
    real sgn = (mr < 0) ? pi : -pi;
    cplx csgn(0.0, sgn);
    if (kode != 1) {
        real yy = -w.im;
        real sn, cs;
        sincos_(yy, sn, cs);
        csgn = csgn * cplx(cs, sn);
    }
    // cspn = exp(i sgn (fnu - inu)) (-1)^inu
    cplx cspn(ph.cp, (mr < 0) ? ph.sp : -ph.sp);
    if (ph.inu % 2 != 0) cspn = -cspn;
    cplx c1 = kval, c2 = ival;
    if (kode != 1) {
        int iuf = 0;
        real ascle = real(1.0e3) * SPB_D1MACH1 / tol;
        nz += s1s2(w, c1, c2, ascle, alim, iuf);
    }
    return cspn * c1 + csgn * c2;
}

// rotation applied by the Hankel driver after K (or its continuation):
// H(m) = -+ (2i/pi) exp(-+ i fnu pi/2) K
SPB_FN cplx hankel_rotate(int m, const OrderPhase &ph, cplx y) {
    const real tol = SPB_TOL;
    const real tpi = 0.63661977236758134308;  // 2/pi
    cplx csgn(-tpi * ph.sh, (m == 1) ? -tpi * ph.ch : tpi * ph.ch);
    if ((ph.inu / 2) % 2 == 1) csgn = -csgn;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * real(1.0e3) * rtol;
    return rot_guard(y, csgn, rtol, ascle, tol);
}

SPB_FN cplx y_from_hankels(cplx z, int kode, cplx h1, cplx h2, int nz1, int nz2, int &nz) {
    const real tol = SPB_TOL, elim = SPB_ELIM, hci = 0.5;
    nz = nz1 < nz2 ? nz1 : nz2;
    if (kode != 2) {
        cplx d = h2 - h1;
        return cplx(-d.im * hci, d.re * hci);
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
    cplx a = rot_guard(h2, c2, rtol, ascle, tol);
    cplx b = rot_guard(h1, c1, rtol, ascle, tol);
    cplx st = a - b;
    if (st.re == real(0.0) && st.im == real(0.0) && ey == real(0.0)) nz += 1;
    return cplx(-st.im * hci, st.re * hci);
}

// J from I(w): the rotation of the J driver (shared by finish() and finish_jy()).
SPB_HD cplx j_from_i(cplx z, const OrderPhase &ph, cplx ival) {
    const real tol = SPB_TOL;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * rtol * real(1.0e3);
    cplx csgn(ph.ch, ph.sh);
    if ((ph.inu / 2) % 2 == 1) csgn = -csgn;
    if (z.im < real(0.0)) csgn.im = -csgn.im;
    return rot_guard(ival, csgn, rtol, ascle, tol);
}

// Combine pass results into the family value.  inz/knz are the nz codes of the
// passes (>= 0).  Sets ierr/nz like the per-point drivers.
SPB_FN cplx finish(int fam, const Prep &p, cplx z, real fnu, int kode, cplx ival, int inz, cplx kval, int knz,
                   int &ierr, int &nz, const OrderPhase *oph = nullptr) {
    const real tol = SPB_TOL;
    ierr = p.ierr;
    nz = 0;
    real rtol = real(1.0) / tol;
    real ascle = SPB_D1MACH1 * rtol * real(1.0e3);
    if (p.combine == CMB_I_ROT) {
        nz = inz;
        if (nz != 0) return cplx(0.0, 0.0);
        if (fam == FAM_I && !(z.re < real(0.0))) return ival;
        const OrderPhase ph = oph ? *oph : order_phase(fnu);
        if (fam == FAM_I) {
            // csgn = exp(+-i pi fnu)
            cplx csgn(ph.cp, (z.im < real(0.0)) ? -ph.sp : ph.sp);
            if (ph.inu % 2 != 0) csgn = -csgn;
            return rot_guard(ival, csgn, rtol, ascle, tol);
        }
        // csgn = exp(+-i fnu pi/2)
        return j_from_i(z, ph, ival);
    }
    if (fam == FAM_K && p.combine == CMB_K_ROT) {
        nz = knz;
        return kval;
    }
    const OrderPhase ph = oph ? *oph : order_phase(fnu);
    if (fam == FAM_K) return acon_combine(p.w, ph, kode, p.mr, kval, ival, nz);
    if (fam == FAM_H1 || fam == FAM_H2) {
        int m = (fam == FAM_H1) ? 1 : 2;
        cplx y;
        if (p.combine == CMB_K_ROT) {
            nz = knz;
            y = kval;
        } else {
            y = acon_combine(p.w, ph, kode, p.mr, kval, ival, nz);
        }
        return hankel_rotate(m, ph, y);
    }
    // FAM_Y
    cplx h1, h2;
    int nz1 = 0, nz2 = 0;
    if (p.combine == CMB_Y_CONJ) {
        // H1 from K(w), w = (0,-x); H2 from K(conj w) = conj K(w)
        nz1 = nz2 = knz;
        h1 = hankel_rotate(1, ph, kval);
        h2 = hankel_rotate(2, ph, conj(kval));
    } else if (p.mr == -1) {
        // H1 continued (mr = -mm = -1), H2 direct
        cplx yc = acon_combine(p.w, ph, kode, -1, kval, ival, nz1);
        h1 = hankel_rotate(1, ph, yc);
        nz2 = knz;
        h2 = hankel_rotate(2, ph, kval);
    } else {
        nz1 = knz;
        h1 = hankel_rotate(1, ph, kval);
        cplx yc = acon_combine(p.w, ph, kode, 1, kval, ival, nz2);
        h2 = hankel_rotate(2, ph, yc);
    }
    return y_from_hankels(z, kode, h1, h2, nz1, nz2, nz);
}

// J and Y from one (I(w), K(w)) pair.  Same arithmetic as finish(FAM_J) and finish(FAM_Y).
SPB_FN void finish_jy(const Prep &p, cplx z, real fnu, int kode, cplx ival, int inz, cplx kval, int knz, cplx &jv,
                      int &jerr, int &jnz, cplx &yv, int &yerr, int &ynz, const OrderPhase *oph = nullptr) {
    const OrderPhase ph = oph ? *oph : order_phase(fnu);
    jerr = p.ierr;
    yerr = p.ierr;
    jnz = inz;
    jv = (inz != 0) ? cplx(0.0, 0.0) : j_from_i(z, ph, ival);
    cplx h1, h2;
    int nz1 = 0, nz2 = 0;
    if (p.combine == CMB_Y_CONJ) {
        nz1 = nz2 = knz;
        h1 = hankel_rotate(1, ph, kval);
        h2 = hankel_rotate(2, ph, conj(kval));
    } else if (p.mr == -1) {
        cplx yc = acon_combine(p.w, ph, kode, -1, kval, ival, nz1);
        h1 = hankel_rotate(1, ph, yc);
        nz2 = knz;
        h2 = hankel_rotate(2, ph, kval);
    } else {
        nz1 = knz;
        h1 = hankel_rotate(1, ph, kval);
        cplx yc = acon_combine(p.w, ph, kode, 1, kval, ival, nz2);
        h2 = hankel_rotate(2, ph, yc);
    }
    yv = y_from_hankels(z, kode, h1, h2, nz1, nz2, ynz);
}

// I and K from one (I(w), K(w)) pair.  Same arithmetic as finish(FAM_I) and finish(FAM_K).
SPB_FN void finish_ik(const Prep &p, cplx z, real fnu, int kode, cplx ival, int inz, cplx kval, int knz, cplx &iv,
                      int &ierr_i, int &nz_i, cplx &kv, int &ierr_k, int &nz_k, const OrderPhase *oph = nullptr) {
    Prep pi = p;
    pi.combine = CMB_I_ROT;
    iv = finish(FAM_I, pi, z, fnu, kode, ival, inz, cplx(0.0, 0.0), 0, ierr_i, nz_i, oph);
    kv = finish(FAM_K, p, z, fnu, kode, ival, inz, kval, knz, ierr_k, nz_k, oph);
}

// The general per-point path (same routine on CPU checker and GPU).
SPB_FN cplx general_eval(int fam, cplx z, real fnu, int kode, int &ierr, int &nz) {
    cplx cy[1];
    cy[0] = cplx(0.0, 0.0);
    if (!(fnu == fnu) || !(z.re == z.re) || !(z.im == z.im)) {
        // NaN in, NaN out (flagged as an input error)
        ierr = 1;
        nz = 0;
        real q = fnu + z.re + z.im;
        return cplx(q, q);
    }
    switch (fam) {
        case FAM_J: nz = besj(z, fnu, kode, 1, cy, ierr); break;
        case FAM_Y: nz = besy(z, fnu, kode, cy, ierr); break;
        case FAM_I: nz = besi(z, fnu, kode, 1, cy, ierr); break;
        case FAM_K: nz = besk(z, fnu, kode, 1, cy, ierr); break;
        case FAM_H1: nz = besh(z, fnu, kode, 1, 1, cy, ierr); break;
        case FAM_H2: nz = besh(z, fnu, kode, 2, 1, cy, ierr); break;
        default: ierr = 1; nz = 0; break;
    }
    return cy[0];
}

// Post-processing that reproduces the value conventions of scipy.special, the
// reference's declared test oracle (/root/reference/tests/special.rs:10-16),
// for orders >= 0: NaN on domain / overflow / no-result codes, signed infinity
// where the overflow direction is known.  ierr / nz are left untouched.
SPB_FN cplx apply_semantics(int fam, cplx z, real fnu, int kode, cplx val, int ierr, int nz) {
    (void)nz;
    if (ierr == 0 || ierr == 3) return val;
    const real inf = real(1.0) / real(0.0);
    const real nan = inf - inf;
    cplx out(nan, nan);
    if (!(fnu == fnu) || !(z.re == z.re) || !(z.im == z.im)) return out;
    if (fam == FAM_Y && kode == 1 && z.re == real(0.0) && z.im == real(0.0) && fnu >= real(0.0))
        return cplx(-inf, 0.0);
    if (ierr != 2) return out;
    const bool pos_real = z.re >= real(0.0) && z.im == real(0.0);
    if (fam == FAM_K) return pos_real ? cplx(inf, 0.0) : out;
    if (fam == FAM_Y) {
        // the oracle reports -inf for yv and +inf for yve on the positive real axis
        if (pos_real) return cplx(kode == 1 ? -inf : inf, 0.0);
        return out;
    }
    if (kode != 1) return out;
    if (fam == FAM_J) {
        int ie, nzz;
        cplx e = general_eval(FAM_J, z, fnu, 2, ie, nzz);
        if (ie != 0 && ie != 3) return out;
        return cplx(e.re * inf, e.im * inf);
    }
    if (fam == FAM_I) {
        if (z.im == real(0.0) && (z.re >= real(0.0) || fnu == floor(fnu))) {
            real half = fnu * real(0.5);
            if (z.re < real(0.0) && half != floor(half)) return cplx(-inf, 0.0);
            return cplx(inf, 0.0);
        }
        int ie, nzz;
        cplx e = general_eval(FAM_I, z, fnu, 2, ie, nzz);
        if (ie != 0 && ie != 3) return out;
        return cplx(e.re * inf, e.im * inf);
    }
    return out;
}

// ---- negative orders by reflection (the C ABI's default, SPB_RAW_NEG_ORDERS switches it off) --------------------
// The reference calls kv(-3.5, 1000) (/root/reference/src/special/kv.rs:53-54) and unwraps the result, so a negative
// order must produce a value where the raw TOMS-644 drivers report ierr = 1.  Formulae (DLMF 10.4, 10.27), av = -fnu:
//     J_-av = cos(pi av) J_av - sin(pi av) Y_av        Y_-av = sin(pi av) J_av + cos(pi av) Y_av
//     I_-av = I_av + (2/pi) sin(pi av) K_av            K_-av = K_av
//     H1_-av = e^{i pi av} H1_av                       H2_-av = e^{-i pi av} H2_av
// with the integer orders reduced to sign changes (sin(pi n) = 0 exactly).  The scaled variants carry the scale
// factor of the target function (J, Y: e^{-|Im z|} for both terms; I: K_av is rescaled from e^{z} to e^{-|Re z|}).

// sin(pi a), cos(pi a) for a >= 0, exact at the multiples of 1/2
SPB_FN void sincospi_(real a, real &s, real &c) {
    real q = floor(a + a + real(0.5));   // nearest half-integer index
    real r = a - real(0.5) * q;         // |r| <= 1/4
    real sr, cr;
    sincos_(real(SPB_PI) * r, sr, cr);
    if (r == real(0.0)) {
        sr = 0.0;
        cr = 1.0;
    }
    real qm = q - real(4.0) * floor(q * real(0.25));  // q mod 4
    int k = (int)qm;
    switch (k) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
}

// merge of the error codes of the two terms of a reflection formula: the first hard failure wins, else a 3 survives
SPB_HD int merge_ierr(int a, int b) {
    if (a != 0 && a != 3) return a;
    if (b != 0 && b != 3) return b;
    return (a == 3 || b == 3) ? 3 : 0;
}

SPB_FN void reflect_eval_jy(cplx z, real fnu, int kode, cplx &jv, int &jerr, int &jnz, cplx &yv, int &yerr, int &ynz,
                            bool sem) {
    const real av = -fnu;
    real s, c;
    sincospi_(av, s, c);
    int ej, ey, nj, ny;
    cplx j = general_eval(FAM_J, z, av, kode, ej, nj);
    cplx y = general_eval(FAM_Y, z, av, kode, ey, ny);
    const bool whole = (s == real(0.0));
    // integer order: J_-n = (-1)^n J_n needs no Y (which may fail where J does not, e.g. at z = 0)
    jerr = whole ? ej : merge_ierr(ej, ey);
    yerr = whole ? ey : merge_ierr(ey, ej);
    jnz = whole ? nj : 0;
    ynz = whole ? ny : 0;
    jv = whole ? j * c : j * c - y * s;
    yv = whole ? y * c : j * s + y * c;
    if (sem) {
        if (whole) {
            jv = apply_semantics(FAM_J, z, av, kode, j, ej, nj) * c;
            yv = apply_semantics(FAM_Y, z, av, kode, y, ey, ny) * c;
        } else {
            const real inf = real(1.0) / real(0.0);
            const real nan = inf - inf;
            if (jerr != 0 && jerr != 3) jv = cplx(nan, nan);
            if (yerr != 0 && yerr != 3) yv = cplx(nan, nan);
        }
    }
}

SPB_FN cplx reflect_eval(int fam, cplx z, real fnu, int kode, int &ierr, int &nz, bool sem) {
    const real av = -fnu;
    if (!(av == av) || !(z.re == z.re) || !(z.im == z.im)) return general_eval(fam, z, fnu, kode, ierr, nz);
    if (fam == FAM_K) {
        cplx k = general_eval(FAM_K, z, av, kode, ierr, nz);
        return sem ? apply_semantics(FAM_K, z, av, kode, k, ierr, nz) : k;
    }
    if (fam == FAM_J || fam == FAM_Y) {
        cplx jv, yv;
        int ej, ey, nj, ny;
        reflect_eval_jy(z, fnu, kode, jv, ej, nj, yv, ey, ny, sem);
        ierr = fam == FAM_J ? ej : ey;
        nz = fam == FAM_J ? nj : ny;
        return fam == FAM_J ? jv : yv;
    }
    real s, c;
    sincospi_(av, s, c);
    const real inf = real(1.0) / real(0.0);
    const real nan = inf - inf;
    if (fam == FAM_H1 || fam == FAM_H2) {
        cplx h = general_eval(fam, z, av, kode, ierr, nz);
        if (sem) h = apply_semantics(fam, z, av, kode, h, ierr, nz);
        return h * cplx(c, fam == FAM_H1 ? s : -s);
    }
    // FAM_I
    int ei, ni;
    cplx iv = general_eval(FAM_I, z, av, kode, ei, ni);
    if (s == real(0.0)) {  // I_-n = I_n
        ierr = ei;
        nz = ni;
        return sem ? apply_semantics(FAM_I, z, av, kode, iv, ei, ni) : iv;
    }
    int ek, nk;
    cplx kv = general_eval(FAM_K, z, av, kode, ek, nk);
    if (kode == 2) {
        // e^{z} K -> e^{-|Re z|} K: factor e^{-z - |Re z|} (modulus <= 1)
        real sn, cs;
        sincos_(-z.im, sn, cs);
        real m = z.re > real(0.0) ? exp_(real(-2.0) * z.re) : real(1.0);
        if (z.re > real(0.0) && real(2.0) * z.re > real(SPB_ELIM)) m = 0.0;
        kv = kv * cplx(m * cs, m * sn);
    }
    ierr = merge_ierr(ei, ek);
    nz = 0;
    cplx out = iv + kv * (real(0.63661977236758134308) * s);
    if (sem && ierr != 0 && ierr != 3) out = cplx(nan, nan);
    return out;
}

// Whole pipeline for one point, in the order the GPU runs it (prepare -> I pass
// -> K pass -> finish, general path on any hand-over).  `path` reports which
// route was taken: 0 binned, 1 general from prepare, 2 general after a pass.
SPB_FN cplx pipeline_eval(int fam, cplx z, real fnu, int kode, int &ierr, int &nz, int &path, int fuse = 0) {
    Prep p = prepare(fam, z, fnu, kode, fuse);
    path = 1;
    if (p.combine == CMB_GENERAL) return general_eval(fam, z, fnu, kode, ierr, nz);
    if (p.combine == CMB_FINAL) {
        // decided by the classifier (a provable underflow of K)
        path = 0;
        ierr = p.final_ierr;
        nz = p.final_nz;
        return cplx(0.0, 0.0);
    }
    bool geom_needs_i;
    {
        // the region passes see only the class (bins) and recompute the geometry: same thing here
        const int ireg = p.ireg, kreg = p.kreg;
        p = prepare_pass(fam, z);
        geom_needs_i = p.ireg >= 0;
        // (cannot happen except for fused points that do not need I; would show up as a path-1 mismatch)
        if ((p.ireg >= 0) != (ireg >= 0) && kreg != KREG_FUSED) p.combine = CMB_GENERAL;
        p.ireg = ireg;
        p.kreg = kreg;
    }
    cplx ival(0.0, 0.0), kval(0.0, 0.0);
    int inz = 0, knz = 0;
    path = 2;
    if (p.ireg == IREG_DEBYE) {
        if (debye_region(fam, p.w, fnu, kode, geom_needs_i, ival, inz, kval, knz, p.az) < 0)
            return general_eval(fam, z, fnu, kode, ierr, nz);
        p.kreg = -1;
    } else if (p.kreg == KREG_FUSED) {
        if (fused_region(p.ireg, p.w, fnu, kode, ival, inz, kval, knz, p.az) < 0)
            return general_eval(fam, z, fnu, kode, ierr, nz);
        if (knz != 0 && (p.combine == CMB_ACON || p.combine == CMB_Y_SHARED))
            return general_eval(fam, z, fnu, kode, ierr, nz);
    } else if (p.ireg >= 0) {
        inz = ipass_region(p.ireg, p.w, fnu, kode, ival, p.az);
        if (inz < 0) return general_eval(fam, z, fnu, kode, ierr, nz);
    }
    if (p.kreg >= 0) {
        knz = kpass_region(p.w, fnu, kode, kval, p.az);
        if (knz < 0) return general_eval(fam, z, fnu, kode, ierr, nz);
        if (knz != 0 && (p.combine == CMB_ACON || p.combine == CMB_Y_SHARED))
            return general_eval(fam, z, fnu, kode, ierr, nz);
    }
    path = 0;
    return finish(fam, p, z, fnu, kode, ival, inz, kval, knz, ierr, nz);
}

// J and Y together, in the order the GPU runs the FAM_JY pipeline.
SPB_FN void pipeline_eval_jy(cplx z, real fnu, int kode, cplx &jv, int &jerr, int &jnz, cplx &yv, int &yerr,
                             int &ynz, int &path, int fuse = 0) {
    Prep p = prepare(FAM_JY, z, fnu, kode, fuse);
    path = 1;
    cplx ival(0.0, 0.0), kval(0.0, 0.0);
    int inz = 0, knz = 0;
    bool general = (p.combine == CMB_GENERAL);
    if (!general) {
        const int ireg = p.ireg, kreg = p.kreg;
        p = prepare_pass(FAM_JY, z);
        p.ireg = ireg;
        p.kreg = kreg;
    }
    if (!general && p.ireg == IREG_DEBYE) {
        path = 2;
        if (debye_region(FAM_JY, p.w, fnu, kode, true, ival, inz, kval, knz, p.az) < 0) general = true;
    } else if (!general && p.kreg == KREG_FUSED) {
        path = 2;
        if (fused_region(p.ireg, p.w, fnu, kode, ival, inz, kval, knz, p.az) < 0) general = true;
        if (knz != 0 && p.combine == CMB_Y_SHARED) general = true;
    } else {
        if (!general) {
            path = 2;
            inz = ipass_region(p.ireg, p.w, fnu, kode, ival, p.az);
            if (inz < 0) general = true;
        }
        if (!general) {
            knz = kpass_region(p.w, fnu, kode, kval, p.az);
            if (knz < 0 || (knz != 0 && p.combine == CMB_Y_SHARED)) general = true;
        }
    }
    if (general) {
        jv = general_eval(FAM_J, z, fnu, kode, jerr, jnz);
        yv = general_eval(FAM_Y, z, fnu, kode, yerr, ynz);
        return;
    }
    path = 0;
    finish_jy(p, z, fnu, kode, ival, inz, kval, knz, jv, jerr, jnz, yv, yerr, ynz);
}

// I and K together, in the order the GPU runs the FAM_IK pipeline.
SPB_FN void pipeline_eval_ik(cplx z, real fnu, int kode, cplx &iv, int &ierr_i, int &nz_i, cplx &kv, int &ierr_k,
                             int &nz_k, int &path, int fuse = 0) {
    Prep p = prepare(FAM_IK, z, fnu, kode, fuse);
    path = 1;
    cplx ival(0.0, 0.0), kval(0.0, 0.0);
    int inz = 0, knz = 0;
    bool general = (p.combine == CMB_GENERAL || p.combine == CMB_FINAL);
    if (!general) {
        const int ireg = p.ireg, kreg = p.kreg;
        p = prepare_pass(FAM_IK, z);
        p.ireg = ireg;
        p.kreg = kreg;
        path = 2;
        if (p.ireg == IREG_DEBYE) {
            if (debye_region(FAM_IK, p.w, fnu, kode, true, ival, inz, kval, knz, p.az) < 0) general = true;
        } else if (p.kreg == KREG_FUSED) {
            if (fused_region(p.ireg, p.w, fnu, kode, ival, inz, kval, knz, p.az) < 0) general = true;
            if (knz != 0 && p.combine == CMB_ACON) general = true;
        } else {
            inz = ipass_region(p.ireg, p.w, fnu, kode, ival, p.az);
            if (inz < 0) general = true;
            if (!general) {
                knz = kpass_region(p.w, fnu, kode, kval, p.az);
                if (knz < 0 || (knz != 0 && p.combine == CMB_ACON)) general = true;
            }
        }
    }
    if (general) {
        iv = general_eval(FAM_I, z, fnu, kode, ierr_i, nz_i);
        kv = general_eval(FAM_K, z, fnu, kode, ierr_k, nz_k);
        return;
    }
    path = 0;
    finish_ik(p, z, fnu, kode, ival, inz, kval, knz, iv, ierr_i, nz_i, kv, ierr_k, nz_k);
}

}  // namespace spb
