// oracle/flopcount.cpp — TEST / MEASUREMENT INFRASTRUCTURE ONLY.
//
// Compiles the region math with an op-counting scalar type to tally the ALGORITHMIC FP64 operations per
// point of every kernel of the binned pipeline (binding definition: add/sub/mul = 1, div = 1,
// sqrt = 1, each of exp/log/sin/cos/atan/... = 30).  tools/count_flops.py drives it over the parity sample
// of a BASELINE config and freezes the per-kernel means into profiles/flops_per_eval.json, which bench.py
// multiplies by the live launch durations to get roofline.achieved.
#include <cstdint>
#include <cmath>

struct counted;
static thread_local uint64_t g_ops = 0;

struct counted {
    double v;
    counted() : v(0) {}
    counted(double x) : v(x) {}
    counted(int x) : v(x) {}
    explicit operator int() const { return (int)v; }
    explicit operator double() const { return v; }
    explicit operator bool() const { return v != 0; }
};
#define BINOP(op)                                                                            \
    inline counted operator op(counted a, counted b) { g_ops += 1; return counted(a.v op b.v); } \
    inline counted operator op(counted a, double b) { g_ops += 1; return counted(a.v op b); }     \
    inline counted operator op(double a, counted b) { g_ops += 1; return counted(a op b.v); }
BINOP(+) BINOP(-) BINOP(*) BINOP(/)
inline counted operator-(counted a) { return counted(-a.v); }
inline counted &operator+=(counted &a, counted b) { g_ops += 1; a.v += b.v; return a; }
inline counted &operator-=(counted &a, counted b) { g_ops += 1; a.v -= b.v; return a; }
inline counted &operator*=(counted &a, counted b) { g_ops += 1; a.v *= b.v; return a; }
#define CMP(op)                                                         \
    inline bool operator op(counted a, counted b) { return a.v op b.v; } \
    inline bool operator op(counted a, double b) { return a.v op b; }    \
    inline bool operator op(double a, counted b) { return a op b.v; }
CMP(<) CMP(>) CMP(<=) CMP(>=) CMP(==) CMP(!=)
#define TRANS(name) \
    inline counted name(counted a) { g_ops += 30; return counted(std::name(a.v)); }
TRANS(exp) TRANS(log) TRANS(sin) TRANS(cos) TRANS(atan) TRANS(sinh) TRANS(cosh) TRANS(log1p)
inline counted sqrt(counted a) { g_ops += 1; return counted(std::sqrt(a.v)); }
inline counted fabs(counted a) { return counted(std::fabs(a.v)); }
inline counted floor(counted a) { return counted(std::floor(a.v)); }
inline counted fmod(counted a, counted b) { g_ops += 1; return counted(std::fmod(a.v, b.v)); }
inline counted atan2(counted a, counted b) { g_ops += 30; return counted(std::atan2(a.v, b.v)); }
inline counted pow(counted a, counted b) { g_ops += 60; return counted(std::pow(a.v, b.v)); }
inline bool signbit(counted a) { return std::signbit(a.v); }

inline counted fma_(counted a, counted b, counted c) { return a * b + c; }
inline int expo_(counted a) {
    union { double d; unsigned long long u; } t;
    t.d = a.v;
    return (int)((t.u >> 52) & 0x7ffull);
}
inline counted neg_(counted a) { return counted(-a.v); }
// e^x and e^-x from one argument reduction (spb::exp_pair): charged as ONE exponential plus the two combining
// additions, which is what the kernels spend on it
#define SPB_EXP_PAIR_HOOK 1
inline void exp_pair_hook(counted x, counted &ep, counted &em) {
    g_ops += 32;
    ep = counted(std::exp(x.v));
    em = counted(std::exp(-x.v));
}
#define SPB_REAL counted
#include "../sciport_rs_b200/csrc/amos/spb_family.h"

using namespace spb;

static inline int islot(int ireg) { return ireg < 4 ? 2 + ireg : 5 + ireg; }  // regions 4, 5, 6 -> ids 9, 10, 11
static const int NSLOT = 12;

extern "C" {

// ops[12]: classify(prepare), 7 ipass regions (region + its share of finish), kpass (+finish), general, and
// pts[11] the number of points charged to each, following the kernel ids of include/sciport_b200.h
// (0 classify, 2..5 ipass series/asyi/mlri/wrsk, 6 kpass, 7 general, 9 / 10 ipass uniform forms 1 / 2, 11 the
// Debye-direct region).  fuse is the bitmask of spb_family.h (FUSE_ASYI | FUSE_DEBYE).
// fuse = 1: points that need I and K with I in the large-|w| region are charged, as one unit (fused routine +
// finish), to the large-|w| region kernel, like the library does for a broadcast scalar order.
int spbo_count_ops2(int fam, int kode, int fuse, const double *v, int64_t v_stride, const double *z, int64_t n,
                    double *ops, int64_t *pts) {
    for (int k = 0; k < NSLOT; ++k) {
        ops[k] = 0;
        pts[k] = 0;
    }
    for (int64_t i = 0; i < n; ++i) {
        cplx zz(counted(z[2 * i]), counted(z[2 * i + 1]));
        counted fnu(v[i * v_stride]);
        g_ops = 0;
        Prep p = prepare(fam, zz, fnu, kode, fuse);
        ops[0] += g_ops;
        pts[0] += 1;
        // a region kernel recomputes the geometry of its point only (prepare_pass: rotation and modulus), not the
        // region tests of the classifier: that is what it is charged with
        g_ops = 0;
        (void)prepare_pass(fam, zz);
        const uint64_t prep_ops = g_ops;
        int ierr, nz;
        if (fam == FAM_JY || fam == FAM_IK) {
            const bool ik = (fam == FAM_IK);
            // pair family: I pass charged to its region kernel, K pass + both finishes to kpass
            bool general = (p.combine == CMB_GENERAL);
            cplx iv(0.0, 0.0), kv(0.0, 0.0), jv, yv;
            int inz = 0, knz = 0, e2, n2;
            if (!general && p.ireg == IREG_DEBYE) {
                g_ops = prep_ops;
                if (debye_region(fam, p.w, fnu, kode, true, iv, inz, kv, knz, p.az) < 0)
                    general = true;
                else
                    { if (ik) finish_ik(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); else finish_jy(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); }
                ops[islot(IREG_DEBYE)] += g_ops;
                pts[islot(IREG_DEBYE)] += 1;
            } else if (!general && p.kreg == KREG_FUSED) {
                g_ops = prep_ops;
                if (fused_region(p.ireg, p.w, fnu, kode, iv, inz, kv, knz, p.az) < 0 ||
                    (knz != 0 && (p.combine == CMB_Y_SHARED || p.combine == CMB_ACON)))
                    general = true;
                else
                    { if (ik) finish_ik(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); else finish_jy(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); }
                ops[islot(p.ireg)] += g_ops;
                pts[islot(p.ireg)] += 1;
            } else if (!general) {
                g_ops = prep_ops;
                inz = ipass_region(p.ireg, p.w, fnu, kode, iv);
                ops[islot(p.ireg)] += g_ops;
                pts[islot(p.ireg)] += 1;
                if (inz < 0) general = true;
            }
            if (!general && p.kreg != KREG_FUSED) {
                g_ops = prep_ops;
                knz = kpass_region(p.w, fnu, kode, kv);
                if (knz < 0 || (knz != 0 && (p.combine == CMB_Y_SHARED || p.combine == CMB_ACON))) general = true;
                else { if (ik) finish_ik(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); else finish_jy(p, zz, fnu, kode, iv, inz, kv, knz, jv, ierr, nz, yv, e2, n2); }
                ops[6] += g_ops;
                pts[6] += 1;
            }
            if (general) {
                g_ops = 0;
                general_eval(ik ? FAM_I : FAM_J, zz, fnu, kode, ierr, nz);
                general_eval(ik ? FAM_K : FAM_Y, zz, fnu, kode, ierr, nz);
                ops[7] += g_ops;
                pts[7] += 1;
            }
            continue;
        }
        if (p.combine == CMB_FINAL) continue;  // decided (and stored) by the classifier
        if (p.combine == CMB_GENERAL) {
            g_ops = 0;
            general_eval(fam, zz, fnu, kode, ierr, nz);
            ops[7] += g_ops;
            pts[7] += 1;
            continue;
        }
        cplx ival(0.0, 0.0), kval(0.0, 0.0);
        int inz = 0, knz = 0;
        bool handed = false;
        if (p.ireg == IREG_DEBYE) {
            g_ops = prep_ops;
            Prep q = p;
            const bool need_i = (fam == FAM_J || fam == FAM_I) || prepare_geometry(fam, zz, q);
            if (debye_region(fam, p.w, fnu, kode, need_i, ival, inz, kval, knz, p.az) < 0)
                handed = true;
            else
                finish(fam, p, zz, fnu, kode, ival, inz, kval, knz, ierr, nz);
            ops[islot(IREG_DEBYE)] += g_ops;
            pts[islot(IREG_DEBYE)] += 1;
        } else if (p.kreg == KREG_FUSED) {
            g_ops = prep_ops;
            if (fused_region(p.ireg, p.w, fnu, kode, ival, inz, kval, knz, p.az) < 0 ||
                (knz != 0 && (p.combine == CMB_ACON || p.combine == CMB_Y_SHARED)))
                handed = true;
            else
                finish(fam, p, zz, fnu, kode, ival, inz, kval, knz, ierr, nz);
            ops[islot(p.ireg)] += g_ops;
            pts[islot(p.ireg)] += 1;
        } else if (p.ireg >= 0) {
            g_ops = prep_ops;  // the region kernel recomputes the geometry
            inz = ipass_region(p.ireg, p.w, fnu, kode, ival);
            if (inz < 0) handed = true;
            if (!handed && p.combine == CMB_I_ROT) finish(fam, p, zz, fnu, kode, ival, inz, kval, 0, ierr, nz);
            ops[islot(p.ireg)] += g_ops;
            pts[islot(p.ireg)] += 1;
        }
        if (!handed && p.kreg >= 0) {
            g_ops = prep_ops;
            knz = kpass_region(p.w, fnu, kode, kval);
            if (knz < 0 || (knz != 0 && (p.combine == CMB_ACON || p.combine == CMB_Y_SHARED)))
                handed = true;
            else
                finish(fam, p, zz, fnu, kode, ival, inz, kval, knz, ierr, nz);
            ops[6] += g_ops;
            pts[6] += 1;
        }
        if (handed) {
            g_ops = 0;
            general_eval(fam, zz, fnu, kode, ierr, nz);
            ops[7] += g_ops;
            pts[7] += 1;
        }
    }
    return 0;
}

// Airy: mean operation count per point of Ai / Ai' / Bi / Bi' (which = 0..3) through the complete routine
double spbo_count_ops_airy(int which, int kode, const double *z, int64_t n) {
    g_ops = 0;
    for (int64_t i = 0; i < n; ++i) {
        cplx zz(counted(z[2 * i]), counted(z[2 * i + 1]));
        int ierr = 0, nz = 0;
        if (which >= 2)
            biry(zz, which & 1, kode, ierr);
        else
            airy(zz, which & 1, kode, nz, ierr);
    }
    return n > 0 ? (double)g_ops / (double)n : 0.0;
}

int spbo_count_ops(int fam, int kode, const double *v, int64_t v_stride, const double *z, int64_t n, double *ops,
                   int64_t *pts) {
    return spbo_count_ops2(fam, kode, 0, v, v_stride, z, n, ops, pts);
}

}  // extern "C"
