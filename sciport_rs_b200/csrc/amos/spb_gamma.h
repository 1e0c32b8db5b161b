// spb_gamma.h — real gamma / log-gamma companions (memory-bound kernels).
//   gamma_real : shift to x >= 12 by the recurrence, Stirling series there
//                (Bernoulli coefficients CF from tools/gen_tables.py), reflection
//                for x < 0; the power (x/e)^x is split to stay on scale up to 171.6.
//   lgamma_real: ln|Gamma(x)|; Maclaurin series of lgamma(1+t) on (0.5, 2.5] so
//                the zeros at 1 and 2 keep relative accuracy, Stirling above 12,
//                recurrence in between, reflection for x < 0.
// Value conventions (poles, overflow) follow scipy.special.gamma / gammaln, the
// reference's test oracle (/root/reference/tests/special.rs:10-16).
#pragma once
#include "spb_common.h"
#include "spb_gamln.h"

namespace spb {

#if defined(__CUDACC__)
// device copies of the constants of the real gamma / ln|gamma| kernels (constant-bank operands, see fm::K)
namespace gm {
enum { LN2_HI21 = 0, LN2_LO21, HL2PI, ATANH0, STIRLING0 = ATANH0 + 11, COUNT = STIRLING0 + 9 };
static __constant__ double K[COUNT] = {
    0.6931471803691238, 1.9082149292705877e-10,  // ln 2: head with 21 trailing zero bits, tail
    0.91893853320467274178,                       // ln(2 pi)/2
    // 2/(2k+1), k = 11 .. 1: series of 2 atanh(s)/s - 2 in s^2, highest power first
    0.08695652173913043, 0.09523809523809523, 0.10526315789473684, 0.11764705882352941, 0.13333333333333333,
    0.15384615384615385, 0.18181818181818182, 0.2222222222222222, 0.2857142857142857, 0.4, 0.6666666666666666,
    // B_2k / (2k (2k-1)), k = 1 .. 9
    8.33333333333333333333e-2, -2.77777777777777777778e-3, 7.93650793650793650794e-4, -5.95238095238095238095e-4,
    8.41750841750841750842e-4, -1.91752691752691752692e-3, 6.41025641025641025641e-3, -2.95506535947712418301e-2,
    1.79644372368830573165e-1};
}  // namespace gm
// the Maclaurin tables of ln Gamma(1 + t) once more at namespace scope, in constant memory
namespace gmt {
#pragma push_macro("SPB_TABLE")
#undef SPB_TABLE
#define SPB_TABLE(name, n) static __constant__ double name[n]
#include "spb_tables_gamma.inc"
#pragma pop_macro("SPB_TABLE")
}  // namespace gmt
#endif

// sum_{k>=1} B_2k / (2k(2k-1) x^(2k-1)),  x >= 12.  TERMS: 9 everywhere on the host; the device kernels sum what
// their result needs at x >= 12, where the k-th term is below 6.9e-3 / 144^(k-1): 7 terms leave 1.9e-18 (the
// exponent of Gamma), 6 terms 6e-17 (ln|Gamma| >= 17.5 there)
template <int TERMS = 9>
SPB_HD real stirling_tail(real x) {
    real r = rcp_fast(x);  // x >= 12
    real r2 = r * r;
#if defined(__CUDA_ARCH__)
    real s = gm::K[gm::STIRLING0 + TERMS - 1];
#pragma unroll
    for (int k = TERMS - 2; k >= 0; --k) s = fma(s, r2, gm::K[gm::STIRLING0 + k]);
    return s * r;
#else
#include "spb_tables_gln.inc"
    (void)GLN;
    real s = real(CF[8]);
    for (int k = 7; k >= 0; --k) s = s * r2 + real(CF[k]);
    return s * r;
#endif
}

SPB_HD real sinpi_(real x) {
#if defined(__CUDA_ARCH__)
    return sinpi(x);
#else
    // exact argument reduction on the host
    real r = fmod(x, real(2.0));
    if (r > real(1.0)) r -= real(2.0);
    if (r < real(-1.0)) r += real(2.0);
    if (r > real(0.5)) r = real(1.0) - r;
    if (r < real(-0.5)) r = real(-1.0) - r;
    return sin(real(SPB_PI) * r);
#endif
}

#if defined(__CUDA_ARCH__)
// ln y = hi + lo with an absolute error of a few 1e-18 (y > 0, finite, normal): the Stirling exponent
// (y - 1/2) ln y - y is as large as 709, so a plain double logarithm (error 1.1e-16 |ln y|) would cost 1e-13 in
// Gamma.  m = y 2^-e in [0.707, 1.414], s = (m - 1)/(m + 1) carried with its rounding residual,
// ln m = 2 s + 2 s^3/3 + ... (eleven terms), e ln 2 added in two parts.
__device__ __forceinline__ void log_dd(double y, double &hi, double &lo) {
    int hx = __double2hiint(y);
    const int lx = __double2loint(y);
    int e = (hx >> 20) - 1023;
    hx = (hx & 0x000fffff) | 0x3ff00000;
    double m = __hiloint2double(hx, lx);
    const bool upper = m > fm::K[fm::K_SQRT2];
    m = upper ? m * 0.5 : m;
    e = upper ? e + 1 : e;
    const double f = m - 1.0;            // exact
    const double dh = m + 1.0;           // may round: keep the residual
    const double dl = m - (dh - 1.0);
    const double r = rcp_onscale(dh);  // 1.7 <= dh <= 2.42; the residual below absorbs its rounding
    const double sh = f * r;
    const double sl = (fma(-sh, dh, f) - sh * dl) * r;
    const double s2 = sh * sh;
    double p = gm::K[gm::ATANH0];  // 2/23
#pragma unroll
    for (int k = 1; k < 11; ++k) p = fma(p, s2, gm::K[gm::ATANH0 + k]);
    const double mh = sh + sh;
    const double ml = fma(sh * s2, p, sl + sl);
    const double fe = (double)e;
    const double eh = fe * gm::K[gm::LN2_HI21];  // exact: 21 trailing zero bits
    const double t = eh + mh;                    // |eh| >= |mh| unless e == 0 (then t == mh exactly)
    const double err = mh - (t - eh);
    hi = t;
    lo = err + fma(fe, gm::K[gm::LN2_LO21], ml);
}
#endif

#if defined(__CUDA_ARCH__)
// Gamma(y) = exp((y - 1/2) ln y - y + ln sqrt(2 pi) + tail(y)) for 12 <= y <= 171.6, the exponent in double-double.
// Straight-line: the whole of Gamma(x) for x >= 12 (fast path of k_real_map_queued).
__device__ __forceinline__ double gamma_stirling_dev(double y) {
    double lh, ll;
    log_dd(y, lh, ll);
    const double a = y - 0.5;  // exact
    const double p = a * lh;
    const double pe = fma(a, lh, -p) + a * ll;
    const double t1 = p - y;  // |p| > y for y >= 12
    const double te = -y - (t1 - p);
    const double rest = ((te + pe) + stirling_tail<7>(y)) + gm::K[gm::HL2PI];
    const double th = t1 + rest;
    const double tl = rest - (th - t1);
    const double r = exp_(0.5 * th);
    return r * fma(r, tl, r);
}
#endif

SPB_FN real gamma_pos(real x) {
    // x > 0
    const real sq2pi = 2.50662827463100050242;
    if (x > real(171.6243769563027)) return real(1.0) / real(0.0);
    real prod = 1.0;
    real y = x;
    while (y < real(12.0)) {
        prod *= y;
        y += real(1.0);
    }
#if defined(__CUDA_ARCH__)
    return gamma_stirling_dev(y) * rcp_onscale(prod);  // 1 <= prod < 12^12
#endif
    real w = pow(y, real(0.5) * y - real(0.25));
    real g = sq2pi * exp(stirling_tail(y));
    g = (w * exp(-y)) * w * g;
    return g / prod;
}

SPB_FN real gamma_real(real x) {
    const real inf = real(1.0) / real(0.0);
    if (!(x == x)) return x;
    if (x == inf) return inf;
    if (x == -inf) return inf - inf;
    if (x == real(0.0)) return real(1.0) / x;  // signed infinity
    if (x > real(0.0)) return gamma_pos(x);
    if (x == floor(x)) return inf - inf;  // pole at a negative integer
    // reflection: Gamma(x) = pi / (sin(pi x) Gamma(1 - x))
    real s = sinpi_(x);
    real ax = -x;
    if (ax > real(170.6243769563027)) {
        // |Gamma| underflows; keep the sign like the oracle
        real p = floor(ax);
        bool odd = fmod(p, real(2.0)) != real(0.0);
        return odd ? real(0.0) : -real(0.0);
    }
    real g = gamma_pos(real(1.0) - x);
    return real(SPB_PI) / (s * g);
}

// ln Gamma(y) for y >= 12 by the Stirling series (straight-line: the fast path of k_real_map_queued)
SPB_HD real lgamma_stirling(real y) {
#if defined(__CUDA_ARCH__)
    const real ly = log_(y);
    return (y - real(0.5)) * ly - y + gm::K[gm::HL2PI] + (y < real(1.0e17) ? stirling_tail<6>(y) : real(0.0));
#else
    const real hl2pi = 0.91893853320467274178;  // ln(2 pi)/2
    const real ly = log_(y);
    return (y - real(0.5)) * ly - y + hl2pi + (y < real(1.0e17) ? stirling_tail(y) : real(0.0));
#endif
}

SPB_FN real lgamma_pos(real x) {
#if defined(__CUDA_ARCH__)
    using gmt::LG1;
#else
#include "spb_tables_gamma.inc"
#endif
    const real euler1 = 0.42278433509846713939;  // 1 - Euler's constant
    const real hl2pi = 0.91893853320467274178;   // ln(2 pi)/2
    if (x > real(2.5)) {
        // Stirling series at y = x + n >= 12, minus the log of the recurrence product x (x+1) ... (x+n-1) < 12^10:
        // lgamma >= 0.28 on this branch, so the absolute error of the two logarithms (a few 1e-15) stays far
        // below the tolerance; the zeros at 1 and 2 belong to the Maclaurin branch below
        if (x > real(1.0e305)) return real(1.0) / real(0.0);
        real prod = 1.0;
        real y = x;
        while (y < real(12.0)) {
            prod *= y;
            y += real(1.0);
        }
        real ly = log_(y);
        real r = (y - real(0.5)) * ly - y + hl2pi + (y < real(1.0e17) ? stirling_tail(y) : real(0.0));
        return x < real(12.0) ? r - log_(prod) : r;
    }
    // (0, 2.5]: lgamma(1+t), |t| <= 1/2, plus the log term of the recurrence
    real t, extra;
    if (x > real(1.5)) {
        t = x - real(2.0);
        extra = log1p_half_(t);  // lgamma(2+t) = lgamma(1+t) + ln(1+t)
    } else if (x > real(0.5)) {
        t = x - real(1.0);
        extra = 0.0;
    } else {
        t = x;
        extra = -log_(x);  // lgamma(t) = lgamma(1+t) - ln t
    }
    real s = real(LG1[31]);
#pragma unroll
    for (int k = 30; k >= 0; --k) s = s * t + real(LG1[k]);
    return extra + (t * euler1 - log1p_half_(t)) + s * t * t;
}

SPB_FN real lgamma_real(real x) {
    const real inf = real(1.0) / real(0.0);
    if (!(x == x)) return x;
    if (x == inf) return inf;
    if (x == -inf) return -inf;
    if (x == real(0.0)) return inf;
    if (x > real(0.0)) return lgamma_pos(x);
    if (x == floor(x)) return inf;
    // ln|Gamma(x)| = ln(pi / |sin(pi x)|) - lgamma(1 - x)
    real s = rabs(sinpi_(x));
    return log(real(SPB_PI) / s) - lgamma_pos(real(1.0) - x);
}

SPB_HD real sinc_real(real x) {
    // special::sinc, /root/reference/src/special/trig.rs:4-13: sin(pi a)/(pi a), 1 at a == 0
    if (x == real(0.0)) return real(1.0);
    real y = x * real(SPB_PI);
#if defined(__CUDA_ARCH__)
    // the reference's formula with the device sine (three-part reduction, <= 1 ulp; library form beyond 2^20)
    real sn, cs;
    sincos_(y, sn, cs);
    return sn / y;
#else
    return sin(y) / y;
#endif
}

}  // namespace spb

// ---------------------------------------------------------------------------
// Complex loggamma / gamma.  Region logic after Hare (1997), the algorithm the
// oracle's loggamma documents: Stirling series for Re z > 7 or |Im z| > 7,
// Taylor series around 1 and 2, reflection for Re z < 0.1, otherwise the upward
// recurrence with branch counting.  gamma(z) = exp(loggamma(z)).
namespace spb {

SPB_HD cplx csinpi_(cplx z) {
    real x = z.re, piy = real(SPB_PI) * z.im;
    real sx, cx;
#if defined(__CUDA_ARCH__)
    sincospi(x, &sx, &cx);
#else
    sx = sinpi_(x);
    // cos(pi x) with an exact +0 at the half-integers (the sign of that zero decides the branch of ln sin(pi z)
    // on the lines Re z = n + 1/2: like CUDA's cospi and the oracle's, never -0)
    real r = fmod(rabs(x), real(2.0));
    if (r == real(0.5) || r == real(1.5))
        cx = real(0.0);
    else if (r < real(1.0))
        cx = -sin(real(SPB_PI) * (r - real(0.5)));
    else
        cx = sin(real(SPB_PI) * (r - real(1.5)));
#endif
    return cplx(sx * cosh(piy), cx * sinh(piy));
}

SPB_FN cplx clgam_stirling(cplx z) {
#include "spb_tables_gln.inc"
    (void)GLN;
    const real hl2pi = 0.91893853320467274178;
    cplx rz = crecip(z);
    cplx rzz = rz * rz;
    cplx s(real(CF[7]), 0.0);
#pragma unroll
    for (int k = 6; k >= 0; --k) s = s * rzz + real(CF[k]);
    cplx lz = clog_(z);
    return (z - real(0.5)) * lz - z + hl2pi + s * rz;
}

SPB_FN cplx clgam_taylor(cplx t) {
#include "spb_tables_gamma.inc"
    (void)LG1;
    const real euler = 0.57721566490153286061;
    cplx s(real(LGZ[22]), 0.0);
#pragma unroll
    for (int k = 21; k >= 0; --k) s = s * t + real(LGZ[k]);
    return (s * t - euler) * t;
}

SPB_FN cplx clgam_recurrence(cplx z) {
    // Im z >= 0 here
    int signflips = 0, sb = 0;
    cplx shiftprod = z;
    z.re += real(1.0);
    while (z.re <= real(7.0)) {
        shiftprod = shiftprod * z;
        int nsb = shiftprod.im < real(0.0) ? 1 : 0;
        if (nsb != 0 && sb == 0) signflips += 1;
        sb = nsb;
        z.re += real(1.0);
    }
    cplx r = clgam_stirling(z) - clog_(shiftprod);
    r.im -= real(signflips) * real(2.0 * SPB_PI);
    return r;
}

SPB_FN cplx cloggamma(cplx z) {
    const real inf = real(1.0) / real(0.0);
    const real nan = inf - inf;
    const real logpi = 1.1447298858494001741434262;
    if (!(z.re == z.re) || !(z.im == z.im)) return cplx(nan, nan);
    if (z.re <= real(0.0) && z.im == real(0.0) && z.re == floor(z.re)) return cplx(nan, nan);
    if (z.re > real(7.0) || rabs(z.im) > real(7.0)) return clgam_stirling(z);
    cplx t1 = z - real(1.0);
    if (cabs_(t1) <= real(0.2)) return clgam_taylor(t1);
    cplx t2 = z - real(2.0);
    if (cabs_(t2) <= real(0.2)) return clog_(t1) + clgam_taylor(t2);
    if (z.re < real(0.1)) {
        // reflection: loggamma(z) = ln(pi) - ln(sin(pi z)) - loggamma(1 - z) with the branch tracked
        real tmp = floor(real(0.5) * z.re + real(0.25)) * real(2.0 * SPB_PI);
        if (z.im < real(0.0) || (z.im == real(0.0) && signbit(z.im))) tmp = -tmp;
        cplx one_minus(real(1.0) - z.re, -z.im);
        cplx rec;
        // 1 - z has Re > 0.9: Stirling / Taylor / recurrence, no second reflection
        if (one_minus.re > real(7.0) || rabs(one_minus.im) > real(7.0)) {
            rec = clgam_stirling(one_minus);
        } else {
            cplx u1 = one_minus - real(1.0), u2 = one_minus - real(2.0);
            if (cabs_(u1) <= real(0.2))
                rec = clgam_taylor(u1);
            else if (cabs_(u2) <= real(0.2))
                rec = clog_(u1) + clgam_taylor(u2);
            else if (!(one_minus.im < real(0.0)))
                rec = clgam_recurrence(one_minus);
            else
                rec = conj(clgam_recurrence(conj(one_minus)));
        }
        cplx ls = clog_(csinpi_(z));
        return cplx(logpi - ls.re - rec.re, tmp - ls.im - rec.im);
    }
    if (!(z.im < real(0.0))) return clgam_recurrence(z);
    return conj(clgam_recurrence(conj(z)));
}

SPB_FN cplx cgamma(cplx z) {
    const real inf = real(1.0) / real(0.0);
    const real nan = inf - inf;
    if (z.re <= real(0.0) && z.im == real(0.0) && z.re == floor(z.re)) return cplx(nan, nan);
    return cexp_(cloggamma(z));
}

}  // namespace spb
