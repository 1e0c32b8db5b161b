// spb_sweep.h — order sweeps  I_{fnu+k}(w)  /  J_{fnu+k}(z),  k = 0..n-1,  written ONCE per member.
//
// BASELINE config 5 (J_n(z), n = 0..255 per point, output [order][point]) is bound by its 16-byte stores, so
// the sequence form of the classic driver — which keeps the members in an array and re-reads the two it has
// just written at every step of its recurrences — is the wrong shape for a GPU.  This routine produces the
// same quantities from the same ingredients the published algorithm uses for its Miller / Wronskian route
// (ratios by backward recurrence started from an index found by a forward magnitude walk; normalisation
// through  I_v K_{v+1} + I_{v+1} K_v = 1/w  with K from bknu), arranged as two passes over the recurrence:
//
//   pass A  runs the backward recurrence from the start index down to order fnu in registers, rescaling by
//           2^-200 whenever the iterate grows past 2^200 and counting the rescalings; its last two iterates
//           give the ratio I_{fnu+1}/I_{fnu}, hence (Wronskian) I_fnu itself and the normalising factor;
//   pass B  repeats the identical recurrence (same seeds, same rescalings, so the same rounding) and stores
//           member k as  factor * iterate * 2^(-200 j)  the moment it appears, j = rescalings still to come.
//
// Leading (highest-order) members with both components below 2^-1011 = 4.6e-305 (about exp(-elim)) are flushed to
// zero and counted in nz, as the sequence drivers do for the leading members.  The routine DECLINES (returns < 0 before storing anything) for
// arguments near the exponent limits or outside its cost envelope; the caller then uses the classic sequence
// driver, so every input is still served.
#pragma once
#include "spb_common.h"
#include "spb_kregions.h"
#include "spb_drivers.h"

namespace spb {

#define SPB_SWEEP_AZ_MAX 1024.0
#define SPB_SWEEP_AZ_MIN 9.765625e-4 /* 2^-10: bounds the growth per recurrence step (see the rescaling test) */
#define SPB_SWEEP_MIN_ORDERS 16
#define SPB_SWEEP_BIG 1.6069380442589903e+60    /* 2^200 */
#define SPB_SWEEP_SMALL 6.2230152778611417e-61  /* 2^-200 */

// One step of the backward recurrence, a * c + b, with the fused multiply-adds written out in a fixed order:
// pass A and pass B must round identically (their rescaling counts are compared), whatever the compiler would
// otherwise contract, and the same holds between the CPU checker and the device.
SPB_HD cplx sweep_step(cplx a, cplx c, cplx b) {
    return cplx(fma_(a.re, c.re, fma_(-a.im, c.im, b.re)), fma_(a.re, c.im, fma_(a.im, c.re, b.im)));
}

// i^m * v (exact: a swap and sign changes, no floating-point operation)
SPB_HD cplx rot_i(cplx v, int m) {
    real x = (m & 1) ? neg_(v.im) : v.re;
    real y = (m & 1) ? v.re : v.im;
    if (m & 2) {
        x = neg_(x);
        y = neg_(y);
    }
    return cplx(x, y);
}

// |x| >= 2^200 for either component (also true for Inf / NaN), by the exponent fields
SPB_HD bool sweep_big(cplx a) {
    const int e = expo_(a.re) > expo_(a.im) ? expo_(a.re) : expo_(a.im);
    return e >= 1023 + 200;
}
// both components below 2^-1011 = 4.6e-305 (the flush threshold, about exp(-elim))
SPB_HD bool sweep_tiny(cplx a) { return expo_(a.re) < 12 && expo_(a.im) < 12; }

// four steps of the backward recurrence; on entry and on exit p1 is the newest iterate (order rap - 1 on entry of
// the first step), p2 the one before.  The two iterates swap roles from step to step.
SPB_HD void sweep_group(cplx &p1, cplx &p2, cplx rz, real &rap) {
    p2 = sweep_step(p1, rz * rap, p2);
    rap -= real(1.0);
    p1 = sweep_step(p2, rz * rap, p1);
    rap -= real(1.0);
    p2 = sweep_step(p1, rz * rap, p2);
    rap -= real(1.0);
    p1 = sweep_step(p2, rz * rap, p1);
    rap -= real(1.0);
}
// rescaling test after a group: both iterates times 2^-200 when the newest has grown past 2^200; returns 1 then
SPB_HD int sweep_rescale(cplx &a, cplx &b) {
    const bool big = sweep_big(a);
    const real sc = big ? real(SPB_SWEEP_SMALL) : real(1.0);
    a = a * sc;
    b = b * sc;
    return big ? 1 : 0;
}
// scale of a member with j rescalings still to come, as two exact powers of two around the multiplication by the
// normalising factor (|iterate| < 2^300, the factor may be as large as e^alim for unscaled values): neither stage can
// overflow, and a stage that underflows belongs to a member far below the flush threshold anyway.
// j = 0: 1;  j <= 5: s1 = 2^(-200 j);  5 < j <= 10: s1 = 2^-1000, s2 = 2^(-200 (j - 5));  beyond: zero
SPB_HD void sweep_scales(int j, real &s1, real &s2) {
    const int j1 = j < 5 ? j : 5;
    s1 = pow2_(-200 * j1);
    s2 = (j <= 5) ? real(1.0) : (j <= 10) ? pow2_(-200 * (j - 5)) : real(0.0);
}
// Where pass B puts its members: they come out top order first, one after the other, so the destination is a
// cursor that steps down by one member per store.  The generic form indexes the result view; a view with a cheaper
// way of stepping (the [order][point] device view: one pointer subtraction) specialises it.
template <class Y>
struct SweepCursor {
    Y y;
    int m;
    SPB_HD SweepCursor(Y y_, int top) : y(y_), m(top) {}
    SPB_HD void put(cplx v) {
        y[m] = v;
        --m;
    }
};

#if defined(__CUDA_ARCH__)
#define SPB_UNLIKELY(x) (__builtin_expect(!!(x), 0))
#else
#define SPB_UNLIKELY(x) (x)
#endif

// one member: value = iterate * scale * (rotated normalising factor), flushed to zero while it is a leading
// (highest-order) member below the threshold: underflow concerns the leading members only, once one member is on
// scale the lower orders are too (they grow towards the order |w| and oscillate at O(|w|^-1/2) below it)
template <class C>
SPB_HD void sweep_store(C &cur, cplx x, cplx qrot, int j, real s1, real s2, bool &leading, int &nz) {
    cplx val = (x * s1) * qrot;
    if (SPB_UNLIKELY(j > 5)) val = val * s2;
    if (leading) {
        if (sweep_tiny(val)) {
            val = cplx(0.0, 0.0);
            nz += 1;
        } else {
            leading = false;
        }
    }
    cur.put(val);
}

// one group of pass B that holds members (PARTIAL: the top group of a sequence whose length is not a multiple of
// four; its first steps belong to orders above the sequence)
template <bool PARTIAL, class C>
SPB_HD void sweep_store_group(C &cur, int it, int n, cplx &a, cplx &b, cplx rz, real &rap, cplx q0, cplx q1, cplx q2,
                              cplx q3, int &j, real &s1, real &s2, bool &leading, int &nz) {
    b = sweep_step(a, rz * rap, b);
    rap -= real(1.0);
    if (!PARTIAL || it + 3 < n) sweep_store(cur, b, q3, j, s1, s2, leading, nz);
    a = sweep_step(b, rz * rap, a);
    rap -= real(1.0);
    if (!PARTIAL || it + 2 < n) sweep_store(cur, a, q2, j, s1, s2, leading, nz);
    b = sweep_step(a, rz * rap, b);
    rap -= real(1.0);
    if (!PARTIAL || it + 1 < n) sweep_store(cur, b, q1, j, s1, s2, leading, nz);
    a = sweep_step(b, rz * rap, a);
    rap -= real(1.0);
    // (the test of this group comes before its last member is stored: the member is the rescaled iterate)
    if (sweep_rescale(a, b)) sweep_scales(--j, s1, s2);
    sweep_store(cur, a, q0, j, s1, s2, leading, nz);
}

// e^w K_fnu(w) and e^w K_fnu+1(w) for |w| >= rl, Re w >= 0, by the large-|w| (Hankel) expansion
//     e^w K_v(w) = sqrt(pi / 2w) sum_j a_j(v) / w^j,   a_j / a_j-1 = (4 v^2 - (2j-1)^2) / (8 j):
// the all-positive form of the series the large-|w| routine of I sums (spb_iregions.h, asyi), with its term ratio
// and its convergence test; |arg w| <= pi/2, so the remainder is bounded by twice the first neglected term
// (DLMF 10.40(iii)).  The binned pipeline takes K of this region from the same series (DESIGN.md section 2,
// departure 3).  Both orders advance in one loop until both have converged: ten real operations per term and
// order with nearly equal trip counts across a warp, where the Miller algorithm of bknu() runs three
// data-dependent loops.  Returns false when a series has not converged within the published term limit.
SPB_FN bool sweep_k_pair_large(cplx w, real fnu, real az, real tol, cplx &k0, cplx &k1) {
    static const double RJ[48] = {
        0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
        1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
        1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31,
        1.0 / 32, 1.0 / 33, 1.0 / 34, 1.0 / 35, 1.0 / 36, 1.0 / 37, 1.0 / 38, 1.0 / 39, 1.0 / 40, 1.0 / 41,
        1.0 / 42, 1.0 / 43, 1.0 / 44, 1.0 / 45, 1.0 / 46, 1.0 / 47};
    const real rthpi = 1.25331413731550025;  // sqrt(pi/2)
    const real raz = rcp_fast(az);
    const cplx rcz = w_recip(w, raz);
    const cplx coef = w_rsqrt(w, az, raz) * rthpi;
    const cplx rez = rcz * real(0.125);  // 1/(8w)
    const real raez = real(0.125) * raz;
    const real s = tol * raez;
    const int jl = (int)(real(SPB_RL) + real(SPB_RL)) + 2;
    const real d0 = fnu + fnu, d1 = d0 + real(2.0);
    real sq0 = d0 * d0 - real(1.0), sq1 = d1 * d1 - real(1.0);
    const real atol0 = s * rabs(sq0), atol1 = s * rabs(sq1);
    cplx c0(1.0, 0.0), c1(1.0, 0.0), sum0(1.0, 0.0), sum1(1.0, 0.0);
    real aa0 = 1.0, aa1 = 1.0, ak = 0.0;
    bool conv = false;
    for (int j = 1; j <= jl; ++j) {
        const real t0 = sq0 * real(RJ[j]), t1 = sq1 * real(RJ[j]);
        c0 = (c0 * rez) * t0;
        c1 = (c1 * rez) * t1;
        sum0 = sum0 + c0;
        sum1 = sum1 + c1;
        aa0 = aa0 * (rabs(t0) * raez);
        aa1 = aa1 * (rabs(t1) * raez);
        ak += real(8.0);
        sq0 -= ak;
        sq1 -= ak;
        if (aa0 <= atol0 && aa1 <= atol1) {
            conv = true;
            break;
        }
    }
    k0 = coef * sum0;
    k1 = coef * sum1;
    return conv;
}

// The right-half-plane point w at which the sweep of family J / I evaluates I_{fnu+k}(w) for the argument z
// (the rotation of besj_sweep / besi_sweep below; the K-pair pre-pass of the device kernels uses it too).
SPB_HD cplx sweep_point_w(bool jfam, cplx z) {
    if (jfam) {
        cplx zn(z.im, -z.re);
        return (z.im < real(0.0)) ? -zn : zn;
    }
    return (z.re < real(0.0)) ? -z : z;
}

// y[k] = csgn0 * i^(dm k) * I_{fnu+k}(w),  Re w >= 0,  az = |w|.  The order-independent factor csgn0 is folded into
// the normalising factor once; the per-order unit i^(dm k) is applied exactly (dm = +-1 for J, 2 for I continued
// to the left half plane, 0 for none).  Returns nz >= 0, or -1 when it declines.
// kpre (device kernels): where |w| < rl, the pair e^w K_fnu(w), e^w K_fnu+1(w) evaluated beforehand by a dense pass
// over the points that need the Miller algorithm (a few such lanes would hold their whole warp in bknu(): 19 % of the
// sweep kernel's instructions at 12 of 32 lanes); NaN there = that evaluation failed.  With kpre the routine
// takes K of |w| >= rl from the large-|w| series; without it (CPU drivers, classic behaviour) from bknu().
template <class Y>
SPB_FN int isweep(cplx w, real fnu, int kode, int n, Y y, real az, cplx csgn0, int dm, const cplx *kpre = nullptr) {
    const real tol = SPB_TOL, elim = SPB_ELIM, alim = SPB_ALIM;
    // short sequences stay with the classic drivers (cheap there, and a direct evaluation of so few members is
    // slightly more accurate than the ratio / Wronskian route)
    if (n < SPB_SWEEP_MIN_ORDERS || !(az >= real(SPB_SWEEP_AZ_MIN)) || !(az <= real(SPB_SWEEP_AZ_MAX))) return -1;
    if (!(fnu + real(n) < real(4096.0))) return -1;
    if (kode == 1 && !(w.re < alim)) return -1;
    // K_fnu(w), K_fnu+1(w) scaled by e^w (never underflow that way; failure or a set-to-zero member declines)
    cplx ck[2];
    ck[0] = cplx(0.0, 0.0);
    ck[1] = cplx(0.0, 0.0);
    int nw = 0;
    if (kpre) {
        if (az >= real(SPB_RL)) {
            if (!sweep_k_pair_large(w, fnu, az, tol, ck[0], ck[1])) return -1;  // (never observed: classic path)
        } else {
            ck[0] = kpre[0];
            ck[1] = kpre[1];
        }
    } else {
        nw = bknu(w, fnu, 2, 2, ck, tol, elim, alim, az);
    }
    // ---- start index: forward magnitude walk of the ratio algorithm (squared moduli)
    const int inu = (int)fnu;
    const int idnu = inu + n - 1;
    const int magz = (int)az;
    const real amagz = real(magz + 1);
    const real fdnu = real(idnu);
    const real fnup = rmax(amagz, fdnu);
    int id = idnu - magz - 1;
    if (id > 0) id = 0;
    const cplx rz = crecip(w) * real(2.0);
    int k = 1;
    {
        cplx t1 = rz * fnup;
        cplx p2 = -t1;
        cplx p1(1.0, 0.0);
        t1 = t1 + rz;
        real ap2s = p2.re * p2.re + p2.im * p2.im;
        real ap1s;
        const real test1s = real(2.0) * sqrt(ap2s) / tol;  // (sqrt(2 |p2| / tol))^2
        real tests = test1s;
        int itime = 1;
        for (;;) {
            k += 1;
            ap1s = ap2s;
            cplx pt = p2;
            p2 = p1 - t1 * pt;
            p1 = pt;
            t1 = t1 + rz;
            ap2s = p2.re * p2.re + p2.im * p2.im;
            if (ap1s <= tests && k < 100000) continue;
            if (itime == 2) break;
            real ak = cabs_(t1) * real(0.5);
            real flam = ak + sqrt(ak * ak - real(1.0));
            real rho = rmin(sqrt(ap2s / ap1s), flam);
            tests = test1s * (rho / (rho * rho - real(1.0)));
            itime = 2;
        }
    }
    const int kk = k + 1 - id;
    // The recurrence runs in GROUPS of four steps with the rescaling test after each group (one step grows the
    // iterate by at most 2 (order + steps) / |w| < 2^24, so it stays below 2^300 in between); the start index is
    // rounded up to a whole number of groups (a higher start only helps the ratio algorithm).  A group alternates
    // the roles of its two iterates, so no value is moved between registers.  Both passes count the groups by
    // g = groups still to come; step `it` = 4 g + 3 .. 4 g of a group produces the iterate of order fnu + it,
    // member it is stored when it < n.  On the device the lanes of a warp enter pass B staggered by (largest
    // group count of the warp - own group count) idle groups: g is then warp-uniform, so in pass B the 32 stores
    // of one order go out in the same instruction as one coalesced 512-byte request.  Timing only: the arithmetic
    // of a lane does not depend on its neighbours.  The rescaling itself is branch-free (a multiplication by 1
    // or by 2^-200).
    const int steps = (kk + n - 1 + 3) & ~3;
    const int groups = steps >> 2;
    // ---- pass A: recurrence in registers, count the rescalings
    int etot = 0;
    cplx p1(1.0, 0.0), p2(0.0, 0.0);
    {
        real rap = fnu + real(steps);
        for (int g = groups - 1; g >= 0; --g) {
            sweep_group(p1, p2, rz, rap);
            etot += sweep_rescale(p1, p2);
        }
    }
    if (p1.re == real(0.0) && p1.im == real(0.0)) p1 = cplx(tol, tol);
    // ---- I_fnu from the Wronskian:  I_fnu = e^{i Im w} / (w (r K_fnu + K_fnu+1)),  r = I_fnu+1 / I_fnu
    cplx q;
    {
        const cplx r = cdiv(p2, p1);
        const cplx ct = w * (r * ck[0] + ck[1]);
        const real act = cabs_(ct);
        const real ract = real(1.0) / act;
        real sn, cs;
        sincos_(w.im, sn, cs);
        cplx cinu(cs, sn);
        if (kode == 1) cinu = cinu * exp_(w.re);
        const cplx ifnu = (cinu * ract) * (conj(ct) * ract);
        // normalising factor q = I_fnu / (bottom iterate)
        const real bm = real(1.0) / cabs_(p1);
        q = ((ifnu * conj(p1 * bm)) * bm) * csgn0;
        // a failed K evaluation or non-finite factor: decline before anything is stored
        if (nw != 0 || !(q.re == q.re) || !(q.im == q.im) || !(rabs(q.re) <= real(SPB_D1MACH2)) ||
            !(rabs(q.im) <= real(SPB_D1MACH2)))
            return -1;
    }
    // ---- pass B: same recurrence, store each member once (top order first)
    int nz = 0;
    bool leading = true;
    {
        // member it carries the unit i^(dm it), which depends on it mod 4 only: four exactly rotated copies of the
        // normalising factor (swaps and sign changes), one per step of a group
        const cplx q0 = q, q1 = rot_i(q, dm & 3), q2 = rot_i(q, (2 * dm) & 3), q3 = rot_i(q, (3 * dm) & 3);
        int j = etot;  // rescalings still to come: the true size of an iterate is 2^(-200 j) of its value
        real s1, s2;
        sweep_scales(j, s1, s2);
        real rap = fnu + real(steps);
        cplx a(1.0, 0.0), b(0.0, 0.0);
        const int gn = (n + 3) >> 2;  // groups that hold members
        const int g_first = SPB_WARP_MAX(groups) - 1;
        SweepCursor<Y> cur(y, n - 1);
        SPB_NO_UNROLL
        for (int g = g_first; g >= 0; --g) {
            if (g >= groups) continue;  // (device: this lane's recurrence starts later)
            if (g >= gn) {
                sweep_group(a, b, rz, rap);
                if (sweep_rescale(a, b)) sweep_scales(--j, s1, s2);
                continue;
            }
            const int it = 4 * g;
            if (it + 3 < n)
                sweep_store_group<false>(cur, it, n, a, b, rz, rap, q0, q1, q2, q3, j, s1, s2, leading, nz);
            else
                sweep_store_group<true>(cur, it, n, a, b, rz, rap, q0, q1, q2, q3, j, s1, s2, leading, nz);
        }
    }
    return nz;
}

// J_{fnu+k}(z), k < n, through the sweep.  declined = true: nothing was stored, use besj().
template <class Y>
SPB_FN int besj_sweep(cplx z, real fnu, int kode, int n, Y cy, int &ierr, bool &declined, const cplx *kpre = nullptr) {
    ierr = 0;
    declined = true;
    if (!(fnu >= real(0.0)) || !(fnu < real(32000.0)) || n < 2 || n > 100000) return 0;
    if (!(z.re == z.re) || !(z.im == z.im)) return 0;
    const real az = cabs_(z);
    const OrderPhase ph = order_phase(fnu);
    cplx csgn(ph.ch, ph.sh);
    if ((ph.inu / 2) % 2 == 1) csgn = -csgn;
    cplx zn(z.im, -z.re);
    real cii = 1.0;
    if (z.im < real(0.0)) {
        zn = -zn;
        csgn.im = -csgn.im;
        cii = -cii;
    }
    const int nz = isweep(zn, fnu, kode, n, cy, az, csgn, cii > real(0.0) ? 1 : -1, kpre);
    if (nz < 0) return 0;
    declined = false;
    return nz;
}

// I_{fnu+k}(z), k < n, through the sweep.  declined = true: nothing was stored, use besi().
template <class Y>
SPB_FN int besi_sweep(cplx z, real fnu, int kode, int n, Y cy, int &ierr, bool &declined, const cplx *kpre = nullptr) {
    ierr = 0;
    declined = true;
    if (!(fnu >= real(0.0)) || !(fnu < real(32000.0)) || n < 2 || n > 100000) return 0;
    if (!(z.re == z.re) || !(z.im == z.im)) return 0;
    const real az = cabs_(z);
    cplx zn = z;
    cplx csgn(1.0, 0.0);
    int dm = 0;
    if (z.re < real(0.0)) {
        zn = -z;
        const OrderPhase ph = order_phase(fnu);
        csgn = cplx(ph.cp, (z.im < real(0.0)) ? -ph.sp : ph.sp);
        if (ph.inu % 2 != 0) csgn = -csgn;
        dm = 2;
    }
    const int nz = isweep(zn, fnu, kode, n, cy, az, csgn, dm, kpre);
    if (nz < 0) return 0;
    declined = false;
    return nz;
}

}  // namespace spb
