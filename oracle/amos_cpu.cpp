// oracle/amos_cpu.cpp — CPU checker for the batched special-function path.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit compiles the scalar,
// point-by-point TOMS-644 drivers (sciport_rs_b200/csrc/amos/spb_drivers.h)
// for the host and exposes them through a small C ABI so that tests/ and
// bench.py's cpu_baseline leg can (a) pin the restatement against the golden
// vectors generated from scipy.special (the reference's own test oracle,
// /root/reference/tests/special.rs:10-16) and (b) time a multi-threaded CPU
// run of the same batch (the stand-in for the reference's
// `x.mapv(|v| bessel_i(0.0, v.into()))`, /root/reference/src/special/mod.rs:10,
// parallelised over contiguous chunks).  Nothing in the product path links it.
#include <cstdint>
#include <thread>
#include <limits>
#include <vector>
#include <algorithm>
#include <cmath>

#include "../sciport_rs_b200/csrc/amos/spb_family.h"
#include "../sciport_rs_b200/csrc/amos/spb_gamma.h"
#include "../sciport_rs_b200/csrc/amos/spb_airy_lean.h"

using namespace spb;

namespace {

enum Fn { FN_J = 0, FN_Y = 1, FN_I = 2, FN_K = 3, FN_H1 = 4, FN_H2 = 5 };

inline void eval_one(int fn, int kode, double v, cplx z, cplx &out, int &ierr, int &nz) {
    out = general_eval(fn, z, v, kode, ierr, nz);
}

// negative orders like the C ABI's default (include/sciport_b200.h, SPB_RAW_NEG_ORDERS clear): reflection formulae
inline void eval_one_reflect(int fn, int kode, int sem, double v, cplx z, cplx &out, int &ierr, int &nz) {
    if (v < 0.0) {
        out = reflect_eval(fn, z, v, kode, ierr, nz, sem != 0);
        return;
    }
    out = general_eval(fn, z, v, kode, ierr, nz);
    if (sem) out = apply_semantics(fn, z, v, kode, out, ierr, nz);
}

template <class F>
void parallel_for(int64_t n, int threads, F f) {
    if (threads <= 1 || n < 4096) {
        f(0, n);
        return;
    }
    std::vector<std::thread> pool;
    int64_t chunk = (n + threads - 1) / threads;
    for (int t = 0; t < threads; ++t) {
        int64_t lo = t * chunk, hi = std::min<int64_t>(n, lo + chunk);
        if (lo >= hi) break;
        pool.emplace_back([=] { f(lo, hi); });
    }
    for (auto &th : pool) th.join();
}

}  // namespace

extern "C" {

int spbo_bessel_sem(int fn, int kode, int sem, const double *v, int64_t v_stride, const double *z, double *out,
                    int32_t *ierr, int32_t *nz, int64_t n, int threads);

// Raw AMOS semantics: value as computed, ierr and nz per element.
int spbo_bessel(int fn, int kode, const double *v, int64_t v_stride, const double *z /*re,im pairs*/,
                double *out /*re,im pairs*/, int32_t *ierr, int32_t *nz, int64_t n, int threads) {
    return spbo_bessel_sem(fn, kode, 0, v, v_stride, z, out, ierr, nz, n, threads);
}

// sem = 1 additionally applies the scipy value conventions (NaN / signed infinity on failure).
int spbo_bessel_sem(int fn, int kode, int sem, const double *v, int64_t v_stride, const double *z, double *out,
                    int32_t *ierr, int32_t *nz, int64_t n, int threads) {
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            cplx o;
            int ie, nzz;
            cplx zz(z[2 * i], z[2 * i + 1]);
            // sem bit 0: scipy value conventions; bit 1: raw behaviour for negative orders (ierr = 1)
            if (sem & 2) {
                eval_one(fn, kode, v[i * v_stride], zz, o, ie, nzz);
                if (sem & 1) o = apply_semantics(fn, zz, v[i * v_stride], kode, o, ie, nzz);
            } else {
                eval_one_reflect(fn, kode, sem & 1, v[i * v_stride], zz, o, ie, nzz);
            }
            out[2 * i] = o.re;
            out[2 * i + 1] = o.im;
            if (ierr) ierr[i] = ie;
            if (nz) nz[i] = nzz;
        }
    });
    return 0;
}

// Same batch through the restructured flow the GPU kernels use (prepare / I pass /
// K pass / finish); path[i] = 0 binned, 1 general from prepare, 2 general after a pass.
int spbo_bessel_pipeline(int fn, int kode, const double *v, int64_t v_stride, const double *z, double *out,
                         int32_t *ierr, int32_t *nz, int32_t *path, int64_t n, int threads_and_mode) {
    // bit 16 of the last argument: route points that need I and K in the large-|w| region of I through the
    // fused routine (the flow of the fused region kernel)
    // bit 17: the Debye-direct region (spb_debye.h) is enabled as well
    const int threads = threads_and_mode & 0xffff;
    const int fuse = ((threads_and_mode & 0x10000) ? FUSE_ASYI : 0) | ((threads_and_mode & 0x20000) ? FUSE_DEBYE : 0) | ((threads_and_mode & 0x40000) ? FUSE_WRSK : 0);
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            int ie, nzz, pth;
            cplx o = pipeline_eval(fn, cplx(z[2 * i], z[2 * i + 1]), v[i * v_stride], kode, ie, nzz, pth, fuse);
            out[2 * i] = o.re;
            out[2 * i + 1] = o.im;
            if (ierr) ierr[i] = ie;
            if (nz) nz[i] = nzz;
            if (path) path[i] = pth;
        }
    });
    return 0;
}

// Class bytes of the classifier: fast[i] = what the fast accept alone decides (-1 where it declines),
// full[i] = what prepare() decides.  The fast accept is only allowed to agree or to decline.
int spbo_classify(int fn, int kode, int fuse, const double *v, int64_t v_stride, const double *z, int8_t *fast,
                  int8_t *full, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        const cplx zz(z[2 * i], z[2 * i + 1]);
        const double fnu = v[i * v_stride];
        int ireg, kreg;
        fast[i] = -1;
        if (fast_classify(fn, zz, fnu, kode, fuse, ireg, kreg))
            fast[i] = (int8_t)class_byte(fn, zz, fnu, kode, fuse, true);
        full[i] = (int8_t)class_byte(fn, zz, fnu, kode, fuse, false);
    }
    return 0;
}

// I and K together through the FAM_IK flow (out_i, out_k as re,im pairs); codes [4][n]: ierr_i, ierr_k, nz_i, nz_k
int spbo_bessel_ik_pipeline(int kode, const double *v, int64_t v_stride, const double *z, double *out_i, double *out_k,
                            int32_t *codes, int64_t n, int threads_and_mode) {
    const int threads = threads_and_mode & 0xffff;
    const int fuse = ((threads_and_mode & 0x10000) ? FUSE_ASYI : 0) | ((threads_and_mode & 0x20000) ? FUSE_DEBYE : 0) | ((threads_and_mode & 0x40000) ? FUSE_WRSK : 0);
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            cplx a, b;
            int ea, eb, na, nb, pth;
            pipeline_eval_ik(cplx(z[2 * i], z[2 * i + 1]), v[i * v_stride], kode, a, ea, na, b, eb, nb, pth, fuse);
            out_i[2 * i] = a.re;
            out_i[2 * i + 1] = a.im;
            out_k[2 * i] = b.re;
            out_k[2 * i + 1] = b.im;
            codes[i] = ea;
            codes[n + i] = eb;
            codes[2 * n + i] = na;
            codes[3 * n + i] = nb;
        }
    });
    return 0;
}

// J and Y together through the FAM_JY flow (out_j, out_y as re,im pairs)
int spbo_bessel_jy_pipeline(int kode, const double *v, int64_t v_stride, const double *z, double *out_j, double *out_y,
                            int32_t *codes /* [4][n]: ierr_j, ierr_y, nz_j, nz_y */, int64_t n,
                            int threads_and_mode) {
    const int threads = threads_and_mode & 0xffff;
    const int fuse = ((threads_and_mode & 0x10000) ? FUSE_ASYI : 0) | ((threads_and_mode & 0x20000) ? FUSE_DEBYE : 0) | ((threads_and_mode & 0x40000) ? FUSE_WRSK : 0);
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            cplx jv, yv;
            int je, jn, ye, yn, pth;
            pipeline_eval_jy(cplx(z[2 * i], z[2 * i + 1]), v[i * v_stride], kode, jv, je, jn, yv, ye, yn, pth,
                             fuse);
            out_j[2 * i] = jv.re;
            out_j[2 * i + 1] = jv.im;
            out_y[2 * i] = yv.re;
            out_y[2 * i + 1] = yv.im;
            codes[i] = je;
            codes[n + i] = ye;
            codes[2 * n + i] = jn;
            codes[3 * n + i] = yn;
        }
    });
    return 0;
}

// Order sequences: out[k*n + i] = F_{v_i + k}(z_i), k < n_orders (J, I, K, H1, H2; Y through its Hankel pair).
int spbo_bessel_seq(int fn, int kode, const double *v, int64_t v_stride, const double *z, double *out, int32_t *ierr,
                    int32_t *nz, int64_t n, int n_orders, int threads_and_mode) {
    // bit 16 of the last argument switches the sweep off (classic sequence drivers only: cross-check); bit 17 runs
    // the sweep the way the CUDA kernels do (k_sequence.cu): the K pair of the Wronskian normalisation from the
    // large-|w| series for |w| >= rl, from a bknu() evaluation handed in for the other points
    const int threads = threads_and_mode & 0xffff;
    const bool sweep = (threads_and_mode & 0x10000) == 0;
    const bool device_flow = (threads_and_mode & 0x20000) != 0;
    auto kpair = [](bool jfam, cplx zz, double fnu, cplx *kpre) {
        kpre[0] = cplx(0.0, 0.0);
        kpre[1] = cplx(0.0, 0.0);
        const double az = cabs_(zz);
        if (!(az >= SPB_RL)) {
            const cplx w = sweep_point_w(jfam, zz);
            if (bknu(w, fnu, 2, 2, kpre, SPB_TOL, SPB_ELIM, SPB_ALIM, az) != 0) {
                const double nan = std::numeric_limits<double>::quiet_NaN();
                kpre[0] = cplx(nan, nan);
                kpre[1] = cplx(nan, nan);
            }
        }
    };
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        std::vector<cplx> cy(n_orders), cw(n_orders);
        for (int64_t i = lo; i < hi; ++i) {
            cplx zz(z[2 * i], z[2 * i + 1]);
            double fnu = v[i * v_stride];
            int ie = 0, nzz = 0;
            for (int k = 0; k < n_orders; ++k) cy[k] = cplx(0.0, 0.0);
            switch (fn) {
                // J and I sweeps: the write-once two-pass recurrence (spb_sweep.h) where it applies, the classic
                // sequence driver where it declines — the same rule the CUDA sequence kernels follow
                case FN_J: {
                    bool declined = true;
                    cplx kpre[2];
                    if (sweep && device_flow) kpair(true, zz, fnu, kpre);
                    if (sweep) nzz = besj_sweep(zz, fnu, kode, n_orders, cy.data(), ie, declined, device_flow ? kpre : nullptr);
                    if (declined) nzz = besj(zz, fnu, kode, n_orders, cy.data(), ie);
                    break;
                }
                case FN_I: {
                    bool declined = true;
                    cplx kpre[2];
                    if (sweep && device_flow) kpair(false, zz, fnu, kpre);
                    if (sweep) nzz = besi_sweep(zz, fnu, kode, n_orders, cy.data(), ie, declined, device_flow ? kpre : nullptr);
                    if (declined) nzz = besi(zz, fnu, kode, n_orders, cy.data(), ie);
                    break;
                }
                case FN_K: nzz = besk(zz, fnu, kode, n_orders, cy.data(), ie); break;
                case FN_H1: nzz = besh(zz, fnu, kode, 1, n_orders, cy.data(), ie); break;
                case FN_H2: nzz = besh(zz, fnu, kode, 2, n_orders, cy.data(), ie); break;
                case FN_Y:
                    // member by member through the Hankel pair (no sequence recurrence of its own)
                    for (int k = 0; k < n_orders; ++k) {
                        cplx one[1];
                        one[0] = cplx(0.0, 0.0);
                        int iek = 0;
                        nzz += besy(zz, fnu + (double)k, kode, one, iek);
                        cy[k] = one[0];
                        if (iek != 0 && (ie == 0 || ie == 3)) ie = iek;
                    }
                    break;
                default: ie = 1;
            }
            for (int k = 0; k < n_orders; ++k) {
                out[2 * (k * n + i)] = cy[k].re;
                out[2 * (k * n + i) + 1] = cy[k].im;
            }
            if (ierr) ierr[i] = ie;
            if (nz) nz[i] = nzz;
        }
    });
    return 0;
}

// signal::windows::kaiser (/root/reference/src/signal/fir_filter_design/windows.rs:540-558), restated with the
// checker's I_0: l = beta sqrt(1 - ((n - alpha)/alpha)^2); w = |I0(l) / I0(beta)|; periodic form = first m points of
// the symmetric window of length m + 1; lengths <= 1 give ones.
int spbo_kaiser(int64_t m, double beta, int sym, double *out) {
    if (m <= 1) {
        for (int64_t i = 0; i < m; ++i) out[i] = 1.0;
        return 0;
    }
    const int64_t m_ext = sym ? m : m + 1;
    const double alpha = 0.5 * ((double)m_ext - 1.0);
    int ie, nz;
    const cplx r = general_eval(FN_I, cplx(beta, 0.0), 0.0, 1, ie, nz);
    for (int64_t n = 0; n < m; ++n) {
        const double t = ((double)n - alpha) / alpha;
        const double l = beta * std::sqrt(1.0 - t * t);
        const cplx q = cdiv(general_eval(FN_I, cplx(l, 0.0), 0.0, 1, ie, nz), r);
        out[n] = cabs_(q);
    }
    return 0;
}

// signal::windows::lanczos (windows.rs:509-538)
int spbo_lanczos(int64_t m, int sym, double *out) {
    if (m <= 1) {
        for (int64_t i = 0; i < m; ++i) out[i] = 1.0;
        return 0;
    }
    const int64_t m_ext = sym ? m : m + 1;
    const int64_t half = (m_ext % 2 == 0) ? m_ext / 2 : (m_ext + 1) / 2;
    for (int64_t n = 0; n < m; ++n) {
        if (m_ext % 2 == 1 && n == (m_ext - 1) / 2) {
            out[n] = 1.0;
            continue;
        }
        const int64_t j = n >= half ? n : (m_ext - 1 - n);
        out[n] = sinc_real(2.0 * (double)j / ((double)m_ext - 1.0) - 1.0);
    }
    return 0;
}

// which: 0 Ai, 1 Ai', 2 Bi, 3 Bi'
int spbo_airy(int which, int kode, const double *z, double *out, int32_t *ierr, int32_t *nz, int64_t n,
              int threads) {
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            cplx zz(z[2 * i], z[2 * i + 1]);
            cplx o;
            int ie = 0, nzz = 0;
            if (which < 2)
                o = airy(zz, which, kode, nzz, ie);
            else
                o = biry(zz, which - 2, kode, ie);
            out[2 * i] = o.re;
            out[2 * i + 1] = o.im;
            if (ierr) ierr[i] = ie;
            if (nz) nz[i] = nzz;
        }
    });
    return 0;
}

// The flow of the Airy kernels: |zeta| = (2/3)|z|^(3/2) >= rl through the lean large-|zeta| routines
// (spb_airy_lean.h; a declined point falls back to the complete routine like on the GPU), the rest through the
// complete routines.  path[i] = 1 where the lean routine produced the value.
int spbo_airy_flow(int which, int kode, const double *z, double *out, int32_t *ierr, int32_t *nz, int32_t *path,
                   int64_t n, int threads) {
    parallel_for(n, threads, [=](int64_t lo, int64_t hi) {
        for (int64_t i = lo; i < hi; ++i) {
            cplx zz(z[2 * i], z[2 * i + 1]);
            cplx o(0.0, 0.0);
            int ie = 0, nzz = 0;
            const double az = cabs_(zz);
            const double azeta = (2.0 / 3.0) * az * sqrt(az);
            bool lean = az > 1.0 && azeta >= SPB_RL * (1.0 + 1.0e-9) && az < 1.0e300;
            if (lean) {
                if (which < 2)
                    lean = airy_large(zz, which, kode, o, nzz, ie);
                else
                    o = biry_large(zz, which - 2, kode, ie);
            } else if (az > 1.0 && azeta < SPB_RL) {
                lean = (which < 2) ? airy_mid(zz, which, kode, o, nzz, ie) : biry_mid(zz, which - 2, kode, o, ie);
            }
            if (!lean) {
                nzz = 0;
                if (which < 2)
                    o = airy(zz, which, kode, nzz, ie);
                else
                    o = biry(zz, which - 2, kode, ie);
            }
            out[2 * i] = o.re;
            out[2 * i + 1] = o.im;
            if (ierr) ierr[i] = ie;
            if (nz) nz[i] = nzz;
            if (path) path[i] = lean ? 1 : 0;
        }
    });
    return 0;
}

// which: 0 gamma, 1 ln|gamma|, 2 sinc
int spbo_gamma(int which, const double *x, double *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i)
        out[i] = which == 0 ? gamma_real(x[i]) : which == 1 ? lgamma_real(x[i]) : sinc_real(x[i]);
    return 0;
}

// which: 0 gamma(z), 1 loggamma(z)
int spbo_cgamma(int which, const double *z, double *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        cplx r = which == 0 ? cgamma(cplx(z[2 * i], z[2 * i + 1])) : cloggamma(cplx(z[2 * i], z[2 * i + 1]));
        out[2 * i] = r.re;
        out[2 * i + 1] = r.im;
    }
    return 0;
}

int spbo_gamln(const double *x, double *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = gamln(x[i]);
    return 0;
}

int spbo_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
