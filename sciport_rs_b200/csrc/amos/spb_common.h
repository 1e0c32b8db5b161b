// spb_common.h — scalar/complex helpers and machine constants shared by every
// region routine of the batched special-function engine.
//
// The arithmetic the reference delegates to `complex-bessel-rs 1.2.0`
// (/root/reference/Cargo.toml:16; call sites src/special/mod.rs:10 and
// src/special/kv.rs:19) is D. E. Amos' Algorithm 644 (ACM TOMS 12, 1986).  The
// dependency's source is not vendored, so the routines below restate the
// published algorithm; constants follow IEEE-754 binary64.
//
// Everything here is `SPB_HD` (host + device) so the same region math can be
// driven (a) per-region by the CUDA kernels in ../kernels.cu and (b) point by
// point by the CPU checker built from oracle/.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SPB_HD __host__ __device__ __forceinline__
#define SPB_FN __host__ __device__ inline
#else
#define SPB_HD inline
#define SPB_FN inline
#endif

// Data-dependent loops of the region routines stay rolled on the device: unrolled by the compiler they make the
// pass kernels a quarter of a megabyte each, and warps sitting in different loops then miss the instruction cache
// (28 % of the stall samples of the K pass with per-element orders were "no instruction").
#if defined(__CUDACC__)
#define SPB_NO_UNROLL _Pragma("unroll 1")
#define SPB_UNROLL _Pragma("unroll")
#else
#define SPB_NO_UNROLL
#define SPB_UNROLL
#endif

// Rarely taken fallbacks (scaled modulus, library transcendentals outside the fast range) are kept out of line on
// the device: force-inlined at every call site they made up more than half of the code of the pass kernels.
#if defined(__CUDACC__)
#define SPB_COLD static __host__ __device__ __noinline__
#else
#define SPB_COLD static inline
#endif

// The scalar type can be swapped for an op-counting type on the host
// (oracle/flopcount.cpp) to tally algorithmic flops per evaluation.
#ifndef SPB_REAL
#define SPB_REAL double
#endif

namespace spb {

typedef SPB_REAL real;

// ---- machine constants (binary64) ------------------------------------------
// tol = eps; elim = 2.303*(1021*log10(2)-3); alim = elim - 2.303*52*log10(2)
// rl = 1.2*dig+3, fnul = 10+6*(dig-3), dig = 52*log10(2)
#define SPB_TOL 2.220446049250313e-16
#define SPB_R1M5 0.30102999566398120
#define SPB_ELIM 700.9217936944459         /* 2.303*(1021*R1M5-3) */
#define SPB_ALIM 664.8716455337102         /* elim - 2.303*52*R1M5 */
#define SPB_DIG 15.653559774527022         /* 52*R1M5 */
#define SPB_RL 21.784271729432426          /* 1.2*dig+3 */
#define SPB_FNUL 85.92135864716213         /* 10+6*(dig-3) */
#define SPB_D1MACH1 2.2250738585072014e-308
#define SPB_D1MACH2 1.7976931348623157e+308
#define SPB_PI 3.14159265358979323846
#define SPB_HPI 1.57079632679489661923

struct cplx {
    real re, im;
    SPB_HD cplx() {}
    SPB_HD cplx(real r, real i) : re(r), im(i) {}
};

SPB_HD cplx operator+(cplx a, cplx b) { return cplx(a.re + b.re, a.im + b.im); }
SPB_HD cplx operator-(cplx a, cplx b) { return cplx(a.re - b.re, a.im - b.im); }
SPB_HD cplx operator-(cplx a) { return cplx(-a.re, -a.im); }
SPB_HD cplx operator*(cplx a, cplx b) {
    return cplx(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re);
}
SPB_HD cplx operator*(cplx a, real s) { return cplx(a.re * s, a.im * s); }
SPB_HD cplx operator*(real s, cplx a) { return cplx(a.re * s, a.im * s); }
SPB_HD cplx operator+(cplx a, real s) { return cplx(a.re + s, a.im); }
SPB_HD cplx operator-(cplx a, real s) { return cplx(a.re - s, a.im); }
SPB_HD cplx conj(cplx a) { return cplx(a.re, -a.im); }
// multiplication by +i / -i (exact)
SPB_HD cplx mul_i(cplx a) { return cplx(-a.im, a.re); }
SPB_HD cplx mul_mi(cplx a) { return cplx(a.im, -a.re); }

SPB_HD real rabs(real x) { return fabs(x); }
SPB_HD real rmax(real a, real b) { return a > b ? a : b; }
SPB_HD real rmin(real a, real b) { return a < b ? a : b; }

// explicit fused multiply-add (never left to the compiler's contraction rules where two code paths must agree
// bit for bit); an op-counting scalar type supplies its own overload
SPB_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return fma(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

// 2^e for -1022 <= e <= 1023 (exact, built from the exponent field)
SPB_HD double pow2_(int e) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)(1023 + e) << 52);
#else
    union {
        unsigned long long u;
        double d;
    } x;
    x.u = (unsigned long long)(1023 + e) << 52;
    return x.d;
#endif
}

// Exponent field (biased, 0..2047) and sign flip of a double through its high word: integer-pipe operations where
// a comparison of moduli or a negation would otherwise occupy the FP64 pipe of a store-bound loop.
SPB_HD int expo_(double x) {
#if defined(__CUDA_ARCH__)
    return (__double2hiint(x) >> 20) & 0x7ff;
#else
    union {
        double d;
        unsigned long long u;
    } t;
    t.d = x;
    return (int)((t.u >> 52) & 0x7ffull);
#endif
}
SPB_HD double neg_(double x) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
#else
    return -x;
#endif
}

// largest value of x among the currently converged lanes of the warp (device); the value itself on the host
#if defined(__CUDA_ARCH__)
#define SPB_WARP_MAX(x) __reduce_max_sync(__activemask(), (x))
#else
#define SPB_WARP_MAX(x) (x)
#endif

// |z| without premature over/underflow (the ZABS of the algorithm): scaled form for components whose squares
// would leave the range
SPB_COLD real cabs_scaled(real u, real v) {
    real s = u + v;
    if (s == real(0.0)) return real(0.0);
    if (u > v) {
        real q = v / u;
        return u * sqrt(real(1.0) + q * q);
    }
    real q = u / v;
    return v * sqrt(real(1.0) + q * q);
}
SPB_HD real cabs_(cplx z) {
    real u = rabs(z.re), v = rabs(z.im);
    // both squares on scale: the plain form (one square root, no division) is as accurate as the scaled one
    real m = rmax(u, v);
    if (m > real(1.0e-140) && m < real(1.0e140)) return sqrt(u * u + v * v);
    return cabs_scaled(u, v);
}

// a / b via the modulus of b (keeps intermediate values on scale).
SPB_HD cplx cdiv(cplx a, cplx b) {
    real bm = real(1.0) / cabs_(b);
    real cc = b.re * bm, cd = b.im * bm;
    return cplx((a.re * cc + a.im * cd) * bm, (a.im * cc - a.re * cd) * bm);
}
// 1 / b
SPB_HD cplx crecip(cplx b) {
    real bm = real(1.0) / cabs_(b);
    real cc = b.re * bm, cd = b.im * bm;
    return cplx(cc * bm, -cd * bm);
}

// principal square root, branch cut on the negative real axis (algebraic
// form: one modulus, one real square root, one division)
SPB_HD cplx csqrt_(cplx a) {
    if (a.im == real(0.0)) {
        if (a.re >= real(0.0)) return cplx(sqrt(a.re), real(0.0));
        return cplx(real(0.0), sqrt(-a.re));
    }
    real d = cabs_(a);
    if (a.re > real(0.0)) {
        real r = sqrt(real(0.5) * (d + a.re));
        return cplx(r, a.im / (r + r));
    }
    real s = sqrt(real(0.5) * (d - a.re));
    real r = rabs(a.im) / (s + s);
    return cplx(r, a.im < real(0.0) ? -s : s);
}

// square root of a when its modulus d = |a| is already known
SPB_HD cplx csqrt_mod(cplx a, real d) {
    if (a.im == real(0.0)) {
        if (a.re >= real(0.0)) return cplx(sqrt(a.re), real(0.0));
        return cplx(real(0.0), sqrt(-a.re));
    }
    if (a.re > real(0.0)) {
        real r = sqrt(real(0.5) * (d + a.re));
        return cplx(r, a.im / (r + r));
    }
    real s = sqrt(real(0.5) * (d - a.re));
    real r = rabs(a.im) / (s + s);
    return cplx(r, a.im < real(0.0) ? -s : s);
}

// ---- device transcendentals --------------------------------------------------------------------------------
// exp and sincos account for 15-30 % of the instructions of the hot region kernels.  The library versions
// materialise every polynomial coefficient with two moves before the DFMA that uses it and carry slow-path
// calls for huge arguments; the versions below keep the coefficients in constant memory (a DFMA reads them as
// constant-bank operands), reduce with fused multiply-adds (three-part pi/2, two-part ln 2); sincos falls back
// to the library for |x| >= 2^20 (unless the caller guarantees the range), exp is branch-free (see exp_pair).  The
// coefficient tables are generated by tools/gen_tables.py (Chebyshev fits of (sin r / r - 1) / r^2 and
// (cos r - 1 + r^2/2) / r^4 on |r| <= pi/4, Taylor coefficients of exp); measured error <= 1.6 ulp.
#if defined(__CUDACC__)
namespace fm {
static __constant__ double SINC[7] = {-7.58549250782124013e-13, 1.60585109795353493e-10, -2.50521061009266746e-08,
                                      2.75573192189567318e-06,  -1.98412698412645693e-04, 8.33333333333333148e-03,
                                      -1.66666666666666657e-01};
static __constant__ double COSC[7] = {4.74520215570235880e-14,  -1.14704494124097104e-11, 2.08767557179572958e-09,
                                      -2.75573192211901978e-07, 2.48015873015843708e-05,  -1.38888888888888873e-03,
                                      4.16666666666666644e-02};
// Taylor coefficients of cosh r - 1 (1/12! .. 1/2!) and of sinh r / r - 1 (1/13! .. 1/3!), highest power first
static __constant__ double EXPE[6] = {2.08767569878681e-09, 2.755731922398589e-07, 2.48015873015873e-05,
                                      0.001388888888888889, 0.041666666666666664,  0.5};
static __constant__ double EXPO[6] = {1.6059043836821613e-10, 2.505210838544172e-08, 2.7557319223985893e-06,
                                      0.0001984126984126984,  0.008333333333333333,  0.16666666666666666};
static __constant__ double LOGC[7] = {1.46437581696042513e-01, 1.53294900617389485e-01, 1.81829569361832921e-01,
                                      2.22222101946705108e-01, 2.85714286317930388e-01, 3.99999999998865485e-01,
                                      6.66666666666666963e-01};
// Full-precision scalar constants of the routines below.  A 64-bit literal whose low word is not zero costs two
// moves into a register pair in front of every instruction that uses it (measured: a quarter of the instructions of
// the real gamma kernel, 6 % of the Debye-direct kernel); as an element of a __constant__ array it is a constant-bank
// operand of that instruction.
enum {
    K_LOG2E, K_NLN2_HI, K_NLN2_LO,          // exp: 1/ln 2, -ln 2 in two parts
    K_LN2_HEAD, K_LN2_TAIL, K_SQRT2,        // log: ln 2 as a 40-bit head + tail, the mantissa split point
    K_TWO_O_PI, K_NPIO2_A, K_NPIO2_B, K_NPIO2_C,  // sincos: 2/pi, -pi/2 in three parts
    K_COUNT
};
static __constant__ double K[K_COUNT] = {
    1.4426950408889634, -0.6931471805599453, -2.3190468138462996e-17,
    0.6931471805592082, 7.371002565167799e-13, 1.4142135623730951,
    0.6366197723675814, -1.5707963267948966, -6.123233995736766e-17, 1.4973849048591698e-33};
}  // namespace fm
#endif

#if defined(__CUDACC__)
static __device__ __noinline__ void sincos_library(double x, double *s, double *c) { sincos(x, s, c); }
#endif

#if defined(__CUDA_ARCH__)
// Branch-free reciprocal and reciprocal square root for arguments known to be on scale (the fused region has
// 21.78 <= |w| < 32767): hardware seed (about 20 bits) and one third-order correction, error below 1 ulp.  The
// library forms carry a slow-path branch each, which also keeps the compiler from scheduling the independent
// exp / sincos arithmetic across them.
__device__ __forceinline__ double rcp_onscale(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);  // y (1 + e + e^2)
}
__device__ __forceinline__ double rsqrt_onscale(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x * y, y, 1.0);              // 1 - x y^2
    return fma(y, e * fma(0.375, e, 0.5), y);          // y (1 + e/2 + 3 e^2/8)
}
#endif
// 1/x for an argument known to be far from the ends of the exponent range (loop counters, moduli inside a region):
// the branch-free form on the device, a plain division on the host
SPB_HD real rcp_fast(real x) {
#if defined(__CUDA_ARCH__)
    return rcp_onscale(x);
#else
    return real(1.0) / x;
#endif
}

// Natural logarithm for the region kernels: x = 2^e m with sqrt(1/2) < m <= sqrt(2), f = m - 1, s = f / (2 + f),
// ln(1 + f) = 2 atanh(s) = f - f^2/2 + s (f^2/2 + s^2 P(s^2)) with P a degree-6 fit of (2 atanh(s)/s - 2)/s^2 on
// s^2 <= 0.03 (tools/gen_tables.py --fastmath; fit error 4e-16, below 1 ulp overall), e ln 2 added in two parts.
// Zero, negative, subnormal, infinite and NaN arguments go to the library.
#if defined(__CUDACC__)
static __device__ __noinline__ double log_library(double x) { return log(x); }
#endif
#if defined(__CUDA_ARCH__)
// ln(2^e (1 + f)), sqrt(1/2) - 1 < f <= sqrt(2) - 1
__device__ __forceinline__ double log_core(double f, int e) {
    const double s = f * rcp_onscale(2.0 + f);
    const double u = s * s;
    double p = fm::LOGC[0];
#pragma unroll
    for (int k = 1; k < 7; ++k) p = fma(p, u, fm::LOGC[k]);
    // f - f^2/2 + s (f^2/2 + R), R = s^2 P(s^2): the rounded quotient s multiplies only second-order terms
    const double hfsq = 0.5 * f * f;
    const double l1 = f - (hfsq - s * fma(u, p, hfsq));
    const double fe = (double)e;
    return fma(fe, fm::K[fm::K_LN2_HEAD], fma(fe, fm::K[fm::K_LN2_TAIL], l1));  // ln 2 = 40-bit head + tail
}
#endif
SPB_HD real log_(real x) {
#if defined(__CUDA_ARCH__)
    const int hi = __double2hiint(x);
    if ((unsigned int)(hi - 0x00100000) >= 0x7fe00000u) return log_library(x);
    int e = (hi >> 20) - 1023;
    double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));  // [1, 2)
    const bool upper = m > fm::K[fm::K_SQRT2];
    m = upper ? m * 0.5 : m;
    e = upper ? e + 1 : e;
    return log_core(m - 1.0, e);
#else
    return log(x);
#endif
}
// ln(1 + t) for |t| <= 1/2 with the relative accuracy of log1p near t = 0 (t itself is the series argument there)
SPB_HD real log1p_half_(real t) {
#if defined(__CUDA_ARCH__)
    if (t >= -0.29 && t <= 0.41) return log_core(t, 0);
    return log_(1.0 + t);
#else
    return log1p(t);
#endif
}

// sin and cos of the same argument in one call.  BOUNDED: the caller guarantees |x| < 2^20 (no fallback branch)
template <bool BOUNDED = false>
SPB_HD void sincos_(real x, real &s, real &c) {
#if defined(__CUDA_ARCH__)
    if (!BOUNDED && !(fabs(x) < 1048576.0)) {
        sincos_library(x, &s, &c);
        return;
    }
    const int n = __double2int_rn(x * fm::K[fm::K_TWO_O_PI]);
    const double fn = (double)n;
    double r = fma(fn, fm::K[fm::K_NPIO2_A], x);
    r = fma(fn, fm::K[fm::K_NPIO2_B], r);
    r = fma(fn, fm::K[fm::K_NPIO2_C], r);
    const double u = r * r;
    double ps = fm::SINC[0], pc = fm::COSC[0];
#pragma unroll
    for (int k = 1; k < 7; ++k) {
        ps = fma(ps, u, fm::SINC[k]);
        pc = fma(pc, u, fm::COSC[k]);
    }
    const double sr = fma(r * u, ps, r);
    const double cr = fma(u * u, pc, fma(u, -0.5, 1.0));
    // quadrant: n mod 4 = 0: (s, c), 1: (c, -s), 2: (-s, -c), 3: (-c, s)
    double ss = (n & 1) ? cr : sr;
    double cc = (n & 1) ? sr : cr;
    if (n & 2) ss = -ss;
    if ((n + 1) & 2) cc = -cc;
    s = ss;
    c = cc;
#else
    s = sin(x);
    c = cos(x);
#endif
}

// e^x and e^-x from one argument reduction.  exp r = 1 + (cosh r - 1) + sinh r on |r| <= ln2/2: the even and
// the odd part are two independent Horner chains in r^2 (half the dependent depth of one chain in r), and
// e^-r is their difference, so the pair costs little more than one exponential.  Branch-free over the whole
// range: beyond the limits the argument is pinned where the result is already 0 or inf, and the power of two is
// applied in two exact halves (the scale never leaves the exponent range).  exp_(x) is the first member of the
// pair bit for bit (the unused half is dead code).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void exp_pair_dev(double x, double &ep, double &em) {
    x = x < -746.0 ? -746.0 : x;  // (comparisons, not fmin / fmax: a NaN passes through and comes out as NaN)
    x = x > 746.0 ? 746.0 : x;
    const int k = __double2int_rn(x * fm::K[fm::K_LOG2E]);
    const double fk = (double)k;
    double r = fma(fk, fm::K[fm::K_NLN2_HI], x);
    r = fma(fk, fm::K[fm::K_NLN2_LO], r);
    const double u = r * r;
    double pe = fm::EXPE[0], po = fm::EXPO[0];
#pragma unroll
    for (int j = 1; j < 6; ++j) {
        pe = fma(pe, u, fm::EXPE[j]);
        po = fma(po, u, fm::EXPO[j]);
    }
    const double c = pe * u;                // cosh r - 1
    const double sh = fma(r * u, po, r);    // sinh r
    const int k1 = k >> 1, k2 = k - k1;
    const double s1 = __longlong_as_double((long long)(1023 + k1) << 52);
    const double s2 = __longlong_as_double((long long)(1023 + k2) << 52);
    const double t1 = __longlong_as_double((long long)(1023 - k1) << 52);
    const double t2 = __longlong_as_double((long long)(1023 - k2) << 52);
    ep = fma(c + sh, s1, s1) * s2;
    em = fma(c - sh, t1, t1) * t2;
}
#endif
SPB_HD void exp_pair(real x, real &ep, real &em) {
#if defined(__CUDA_ARCH__)
    exp_pair_dev(x, ep, em);
#elif defined(SPB_EXP_PAIR_HOOK)
    ::exp_pair_hook(x, ep, em);  // op-counting build (oracle/flopcount.cpp)
#else
    ep = exp(x);
    em = exp(-x);
#endif
}

// e^x
SPB_HD real exp_(real x) {
#if defined(__CUDA_ARCH__)
    double ep, em;
    exp_pair_dev(x, ep, em);
    return ep;
#else
    return exp(x);
#endif
}

// (cos x, sin x)
SPB_HD cplx cis_(real x) {
    real s, c;
    sincos_(x, s, c);
    return cplx(c, s);
}

SPB_HD cplx cexp_(cplx a) {
    real zm = exp_(a.re);
    real s, c;
    sincos_(a.im, s, c);
    return cplx(zm * c, zm * s);
}

// principal logarithm; ierr=1 for a == 0
SPB_HD cplx clog_(cplx a, int &ierr) {
    ierr = 0;
    if (a.re == real(0.0) && a.im == real(0.0)) {
        ierr = 1;
        return cplx(real(0.0), real(0.0));
    }
    return cplx(log_(cabs_(a)), atan2(a.im, a.re));
}
SPB_HD cplx clog_(cplx a) {
    int ie;
    return clog_(a, ie);
}

// Underflow check of one scaled value: the smaller component is considered
// lost when it is below ascle and tol-times smaller than the larger one.
SPB_HD int uchk(cplx y, real ascle, real tol) {
    real wr = rabs(y.re), wi = rabs(y.im);
    real st = rmin(wr, wi);
    if (st > ascle) return 0;
    real ss = rmax(wr, wi);
    st = st / tol;
    return (ss < st) ? 1 : 0;
}

// Phase factors that depend on the order only: one sincos per point serves every rotation of the
// combine step (half-angle for the J / Hankel rotations, full angle by the double-angle formulas for
// the I / continuation factors).
struct OrderPhase {
    real ch, sh;  // cos, sin of (pi/2) (fnu - (inu - ir)),  ir = inu mod 2
    real cp, sp;  // cos, sin of pi (fnu - inu)
    int inu;
};

SPB_HD OrderPhase order_phase(real fnu) {
    OrderPhase ph;
    ph.inu = (int)fnu;
    int ir = ph.inu & 1;
    real a = (fnu - real(ph.inu - ir)) * real(SPB_HPI);
    sincos_<true>(a, ph.sh, ph.ch);  // 0 <= a < pi
    real c2 = ph.ch * ph.ch - ph.sh * ph.sh;
    real s2 = (ph.sh + ph.sh) * ph.ch;
    if (ir) {
        c2 = -c2;
        s2 = -s2;
    }
    ph.cp = c2;
    ph.sp = s2;
    return ph;
}

// Quantities of a right-half-plane point w that BOTH the large-|w| routine of I and the K routine use: where a
// family needs I_v(w) and K_v(w) at the same point (analytic continuation, Y, the J/Y pair) the fused region
// kernel computes them once and hands them to both routines; each routine computes the very same expressions
// itself when called alone, so the fused and the separate paths round identically.
struct WShared {
    real az;      // |w|
    real raz;     // 1/|w|
    cplx rcz;     // 1/w
    cplx rsq;     // 1/sqrt(w) = conj(sqrt w)/|w|
    real sn, cs;  // sin, cos of Im w
    real ep, em;  // exp(Re w), exp(-Re w) when have_em
    bool have_em;
};
SPB_HD cplx w_recip(cplx w, real raz) { return cplx(w.re * raz, -w.im * raz) * raz; }
SPB_HD cplx w_rsqrt(cplx w, real az, real raz) { return conj(csqrt_mod(w, az)) * raz; }
// az < 0: the modulus is not known yet (fused region kernel: |w| and 1/|w| come out of one reciprocal square root)
SPB_HD WShared w_shared(cplx w, real az, bool need_em) {
    WShared s;
#if defined(__CUDA_ARCH__)
    if (az < real(0.0)) {
        const double m2 = fma(w.re, w.re, w.im * w.im);
        const double mi = rsqrt_onscale(m2);
        double m = m2 * mi;
        m = fma(fma(-m, m, m2), 0.5 * mi, m);
        az = m;
        s.raz = fma(mi, fma(-m, mi, 1.0), mi);  // 1/|w| from the seed 1/sqrt(|w|^2), one correction
    } else {
        s.raz = rcp_onscale(az);
    }
    s.az = az;
    s.rcz = w_recip(w, s.raz);
    // 1/sqrt(w) = conj(sqrt w)/|w| with sqrt w = (r, Im w/(2r)), r = sqrt((|w| + Re w)/2), Re w >= 0: through
    // 1/r, so no division and no second square root
    const double t = 0.5 * (az + w.re);
    const double ri = rsqrt_onscale(t);
    double r = t * ri;
    r = fma(fma(-r, r, t), 0.5 * ri, r);
    s.rsq = cplx(r * s.raz, -(w.im * (0.5 * ri)) * s.raz);
    sincos_<true>(w.im, s.sn, s.cs);  // |Im w| <= |w| < 32767 for every binned point
#else
    if (az < real(0.0)) az = cabs_(w);
    s.az = az;
    s.raz = real(1.0) / az;
    s.rcz = w_recip(w, s.raz);
    s.rsq = w_rsqrt(w, az, s.raz);
    sincos_(w.im, s.sn, s.cs);
#endif
    s.have_em = need_em;
    s.ep = real(0.0);
    s.em = real(0.0);
    if (need_em) exp_pair(w.re, s.ep, s.em);
    return s;
}

}  // namespace spb

namespace spb {
// Strided view of a result sequence stored [order][point] in global memory: y[k] is member k of one point.
struct StridedSeq {
    cplx *base;
    long long stride;
    SPB_HD cplx &operator[](int k) const { return base[(long long)k * stride]; }
};
}  // namespace spb
