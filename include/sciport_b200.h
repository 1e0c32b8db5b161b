/* sciport_b200.h — C ABI of the B200 batched special-function engine.
 *
 * Drop-in boundary for sciport-rs's `special` module.  The
 * reference reaches its arithmetic through two scalar calls into the external
 * crate complex-bessel-rs 1.2.0:
 *
 *   complex_bessel_rs::bessel_i::bessel_i(order: f64, z: Complex<f64>) -> Result<Complex<f64>, i32>
 *       call site /root/reference/src/special/mod.rs:10   (special::i0, batched over Array1<f64>)
 *   complex_bessel_rs::bessel_k::bessel_k(order: f64, z: Complex<f64>) -> Result<Complex<f64>, E>
 *       call site /root/reference/src/special/kv.rs:19    (special::kv / special::kve)
 *
 * The entry points below are what a `build.rs` + `extern "C"` block in the
 * Rust crate binds instead (INTEGRATION.md shows the stub): batched forms with
 * the reference's argument order (order first, then z: kv.rs:11), the
 * unscaled / exponentially scaled pair selected by `kode` (kv vs kve,
 * kv.rs:11,39), complex values as `#[repr(C)] {re, im}` pairs (= CUDA
 * double2), and the per-element error code the reference's `Result<_, i32>`
 * carries (mod.rs:8).
 *
 * All functions return 0 on success, a negative SPB_E_* code on failure
 * (spb_last_error() gives the text).  There is NO CPU fallback: without a
 * CUDA device every compute entry point fails with SPB_E_NO_DEVICE.
 */
#ifndef SCIPORT_B200_H
#define SCIPORT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct spb_c128 {
    double re, im;
} spb_c128; /* == num::Complex<f64> (#[repr(C)]) == double2 */

/* function selector of spb_bessel (order-then-z like kv(v, z), kv.rs:11) */
enum spb_fn {
    SPB_FN_J = 0,  /* J_v(z)                                  */
    SPB_FN_Y = 1,  /* Y_v(z)                                  */
    SPB_FN_I = 2,  /* I_v(z)   — replaces bessel_i, mod.rs:10 */
    SPB_FN_K = 3,  /* K_v(z)   — replaces bessel_k, kv.rs:19  */
    SPB_FN_H1 = 4, /* H^(1)_v(z)                              */
    SPB_FN_H2 = 5, /* H^(2)_v(z)                              */
    SPB_FN_JY = 6, /* J_v and Y_v together (spb_bessel_jy only): both need I_v at the same rotated point,
                      so the pair costs one I pass + one K pass instead of two + one */
    SPB_FN_IK = 7  /* I_v and K_v together (spb_bessel_ik only): K in the left half plane is continued with I
                      at the same point, and the large-|z| / Debye-direct regions give both from one series */
};

/* kode: 1 = unscaled (kv), 2 = exponentially scaled (kve; I,J,Y: e^{-|Re z|} resp. e^{-|Im z|};
 * K: e^{z}; H1: e^{-iz}; H2: e^{iz}) */
enum spb_kode { SPB_KODE_UNSCALED = 1, SPB_KODE_SCALED = 2 };

/* which of spb_airy */
enum spb_airy_which { SPB_AIRY_AI = 0, SPB_AIRY_AIP = 1, SPB_AIRY_BI = 2, SPB_AIRY_BIP = 3 };

/* per-element error code (AMOS IERR, the i32 of Result<_, i32>, mod.rs:8) */
enum spb_ierr {
    SPB_IERR_OK = 0,
    SPB_IERR_INPUT = 1,    /* invalid input / singular point (e.g. K at z = 0) */
    SPB_IERR_OVERFLOW = 2, /* result too large                                  */
    SPB_IERR_HALF = 3,     /* argument so large that half the digits are lost   */
    SPB_IERR_LOST = 4,     /* argument so large that no digits remain           */
    SPB_IERR_NOCONV = 5    /* an internal iteration failed to converge          */
};

/* error returns of the library calls themselves */
enum spb_status {
    SPB_OK = 0,
    SPB_E_ARG = -1,       /* bad argument (null pointer, unknown selector, n < 0 …) */
    SPB_E_NO_DEVICE = -2, /* no CUDA device: this library has no CPU path          */
    SPB_E_CUDA = -3,      /* a CUDA runtime call failed                            */
    SPB_E_ALLOC = -4      /* device or pinned-host allocation failed               */
};

/* exec.flags */
enum spb_flags {
    SPB_SEM_AMOS = 0,   /* value exactly as the algorithm leaves it (0 on failure) + ierr/nz  */
    SPB_SEM_SCIPY = 1,  /* value post-processed like scipy.special, the reference's test oracle
                           (tests/special.rs:10-16): NaN on ierr 1/2/4/5, +-inf on overflow
                           where the sign is known                                          */
    SPB_RAW_NEG_ORDERS = 2, /* orders v < 0 are reported as ierr = 1 like the raw TOMS-644 drivers.
                           DEFAULT (flag clear): negative orders are evaluated by the reflection
                           formulae on the device -- K_-v = K_v (the reference calls kv(-3.5, 1000),
                           kv.rs:53-54, and expects a value, not an Err), J_-v = cos(pi v) J_v -
                           sin(pi v) Y_v, Y_-v = sin(pi v) J_v + cos(pi v) Y_v, I_-v = I_v +
                           (2/pi) sin(pi v) K_v, H1_-v = e^{i pi v} H1_v, H2_-v = e^{-i pi v} H2_v,
                           integer orders by sign changes; order sequences (n_orders > 1) keep
                           the raw behaviour                                                 */
    SPB_PATH_BINNED = 0,     /* classify -> stable compaction by region -> per-region kernels */
    SPB_PATH_MONOLITHIC = 16, /* one thread runs the whole driver per point (debug / checker) */
    SPB_PROFILE = 32,         /* device pointers only: CUDA events between the kernel launches;
                                 read the records back with spb_last_profile()                 */
    SPB_PATH_NO_DIRECT = 64   /* broadcast-order jobs: every point through the classifier (the direct
                                 large-|z| kernel that runs ahead of it is switched off; checker) */
};

/* kernel ids reported by spb_last_profile */
enum spb_kernel_id {
    SPB_K_CLASSIFY = 0,      /* region classification + per-tile histogram          */
    SPB_K_SCAN_SCATTER = 1,  /* offsets scan + warp-aggregated stable compaction     */
    SPB_K_IPASS_SERIES = 2,  /* I pass: power series                                 */
    SPB_K_IPASS_ASYI = 3,    /* I pass: large-|z| asymptotic expansion               */
    SPB_K_IPASS_MLRI = 4,    /* I pass: Miller recurrence, Neumann normalisation     */
    SPB_K_IPASS_WRSK = 5,    /* I pass: Miller ratios, Wronskian normalisation       */
    SPB_K_KPASS = 6,         /* K pass + family combine                              */
    SPB_K_GENERAL = 7,       /* general per-point path                               */
    SPB_K_MONOLITHIC = 8,    /* whole batch through the general path                 */
    SPB_K_IPASS_UNI1 = 9,    /* I pass: uniform (Debye-type) expansion at raised order */
    SPB_K_IPASS_UNI2 = 10,   /* I pass: uniform (Airy-type) expansion at raised order  */
    SPB_K_IPASS_DEBYE = 11,  /* I and/or K: Debye-type expansion at the order itself   */
    SPB_K_DIRECT = 12        /* broadcast order: large-|z| points in input order, ahead of the classifier */
};

/* where the buffers live and which GPUs to use */
typedef struct spb_exec {
    int n_devices;      /* 0 or 1: one device; >1: slice [i*n/G,(i+1)*n/G) per device (host pointers only), one
                           host thread (bound to the CPUs local to its GPU) and two streams per device; every
                           batched entry point below honours it (Bessel, Airy, gamma family, sinc)          */
    const int *devices; /* device ordinals, NULL = 0..n_devices-1 (or the current device)                  */
    int device_ptrs;    /* 0: all buffers are host memory (pinned or pageable); 1: device memory on devices[0] */
    void *stream;       /* cudaStream_t to enqueue on when device_ptrs = 1 (NULL = default stream); such calls
                           only ENQUEUE work and return: synchronise the stream before reading results      */
    int flags;          /* spb_flags, OR-ed                                                                */
} spb_exec;

/* ---- Bessel / Hankel ---------------------------------------------------------------------
 * out[k*n + i] = F_{v_i + k}(z_i),  k = 0..n_orders-1,  i = 0..n-1
 * v[i*v_stride] is the (starting) order of point i; v_stride = 0 broadcasts one scalar order.
 * ierr / nz may be NULL.  nz = number of members set to zero by underflow. */
int spb_bessel(int fn, int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out,
               int32_t *ierr, int32_t *nz, int64_t n, int n_orders, const spb_exec *ex);

/* J_v(z_i) and Y_v(z_i) in one pass (n_orders = 1).  Results equal spb_bessel(SPB_FN_J) / spb_bessel(SPB_FN_Y). */
int spb_bessel_jy(int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out_j, spb_c128 *out_y,
                  int32_t *ierr_j, int32_t *ierr_y, int32_t *nz_j, int32_t *nz_y, int64_t n, const spb_exec *ex);

/* I_v(z_i) and K_v(z_i) in one pass (kode 2: e^{-|Re z|} I and e^{z} K, the ive / kve pair of BASELINE config 3).
 * Results equal spb_bessel(SPB_FN_I) / spb_bessel(SPB_FN_K) to the parity tolerance (a point may take a different
 * region routine in the pair than in the single calls). */
int spb_bessel_ik(int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out_i, spb_c128 *out_k,
                  int32_t *ierr_i, int32_t *ierr_k, int32_t *nz_i, int32_t *nz_k, int64_t n, const spb_exec *ex);

/* same with real arguments promoted to complex, the shape of special::i0 (mod.rs:8-13: `v.into()`) */
int spb_bessel_real(int fn, int kode, const double *v, int64_t v_stride, const double *x, spb_c128 *out,
                    int32_t *ierr, int32_t *nz, int64_t n, int n_orders, const spb_exec *ex);

/* ---- Airy -------------------------------------------------------------------------------- */
int spb_airy(int which, int kode, const spb_c128 *z, spb_c128 *out, int32_t *ierr, int32_t *nz, int64_t n,
             const spb_exec *ex);

/* ---- gamma companions (memory-bound) ------------------------------------------------------ */
int spb_gamma(const double *x, double *out, int64_t n, const spb_exec *ex);    /* Gamma(x), real    */
int spb_lgamma(const double *x, double *out, int64_t n, const spb_exec *ex);   /* ln|Gamma(x)|      */
int spb_cgamma(const spb_c128 *z, spb_c128 *out, int64_t n, const spb_exec *ex);   /* Gamma(z)      */
int spb_clgamma(const spb_c128 *z, spb_c128 *out, int64_t n, const spb_exec *ex);  /* loggamma(z)   */

/* ---- sinc, special::sinc (trig.rs:4-13): sin(pi x)/(pi x), 1 at x = 0 ---------------------- */
int spb_sinc(const double *x, double *out, int64_t n, const spb_exec *ex);

/* ---- the device transcendentals of the region kernels, exported for accuracy tests only (no reference
 * counterpart): which = 0 exp(x), 1 exp(-x) (second member of the exp pair), 2 sin(x), 3 cos(x), 4 log(x) -- */
int spb_devmath(int which, const double *x, double *out, int64_t n, const spb_exec *ex);

/* ---- windows built on the path (the batched in-tree consumers) -------------------------------------
 * spb_kaiser: signal::windows::kaiser(m, beta, sym) (fir_filter_design/windows.rs:540-558): argument generation,
 * I_0 through the complex routine, division by I_0(beta) and modulus fused in one pass; out has m doubles.
 * spb_lanczos: signal::windows::lanczos(m, sym) (windows.rs:509-538) on the fused sinc.
 * sym = 0 gives the periodic (fftbins) form: first m points of the symmetric window of length m + 1. */
int spb_kaiser(int64_t m, double beta, int sym, double *out, const spb_exec *ex);
int spb_lanczos(int64_t m, int sym, double *out, const spb_exec *ex);

/* ---- every window of signal::windows as ONE fused kernel per kind: the selector of
 * get_window(WindowType, nx, fftbins) (windows.rs:116-152).  Index -> argument -> value -> store, nothing is read.
 * params (n_params doubles):
 *   GENERAL_COSINE  the coefficients a_k (<= 16); blackman / hamming / hann / flattop / blackmanharris / nuttall /
 *                   general_hamming are general_cosine with fixed coefficients (windows.rs:220-250, 272-285, 342-356)
 *   EXPONENTIAL     center (NaN = (m - 1)/2), tau           TUKEY  alpha (alpha <= 0: ones, alpha >= 1: hann)
 *   TAYLOR          nbar (<= 17), sll, norm (0 / 1; the reference accepts `norm` and ignores it, windows.rs:466)
 *   GAUSSIAN        std                                      GENERAL_GAUSSIAN  power, width
 *   KAISER, KBD     beta                                     the others take none
 * KBD (kaiser_bessel_derived, windows.rs:560-589) needs an even m; its running sum is a device scan. */
enum spb_window_kind {
    SPB_WIN_BOXCAR = 0, SPB_WIN_TRIANG = 1, SPB_WIN_GENERAL_COSINE = 2, SPB_WIN_BARTLETT = 3, SPB_WIN_PARZEN = 4,
    SPB_WIN_BOHMAN = 5, SPB_WIN_BARTHANN = 6, SPB_WIN_COSINE = 7, SPB_WIN_EXPONENTIAL = 8, SPB_WIN_TUKEY = 9,
    SPB_WIN_TAYLOR = 10, SPB_WIN_GAUSSIAN = 11, SPB_WIN_GENERAL_GAUSSIAN = 12, SPB_WIN_KAISER = 13,
    SPB_WIN_LANCZOS = 14, SPB_WIN_KBD = 15
};
int spb_window(int kind, int64_t m, int sym, const double *params, int n_params, double *out, const spb_exec *ex);

/* ---- first error in index order = `.collect::<Result<_, i32>>()` (mod.rs:11-12) ------------
 * *idx = -1, *code = 0 when every element is OK.  ierr lives where ex says. */
int spb_first_error(const int32_t *ierr, int64_t n, int64_t *idx, int32_t *code, const spb_exec *ex);

/* Same answer for the LAST spb_bessel* / spb_airy call of this thread without any per-element array: every kernel
 * keeps the smallest (index, ierr) of its failing elements in one device word, so callers that only need the
 * reference's `Result<Array, i32>` may pass ierr = nz = NULL and save 8 bytes per element of device-to-host
 * traffic.  which = 0: primary output, 1: second output of spb_bessel_jy.  Waits for a device-pointer call to
 * finish.  *idx = -1, *code = 0 when no element failed. */
int spb_last_first_error(int which, int64_t *idx, int32_t *code);

/* ---- synthetic inputs: index-addressable splitmix64 stream -----------------
 * Fills v[offset..offset+n) / z[...] for BASELINE config `cfg` (2,3,4,5) directly where ex says. */
int spb_synth(int cfg, uint64_t seed, int64_t offset, int64_t n, int64_t n_total, double *v, spb_c128 *z,
              const spb_exec *ex);

/* ---- measurement helpers ------------------------------------------------------------------ */
/* sustained DFMA throughput of `device` in TFLOP/s (register-resident, 8 independent chains/thread) */
int spb_fp64_peak(int device, double seconds, double *tflops, double *sm_mhz);
/* device time (ms) spent in kernels of the last spb_* compute call on this thread, measured with
 * CUDA events on the launching stream (asking for kernel_ms waits for that call to finish; pass NULL
 * to only read the launch count), and how many kernels it launched */
int spb_last_timing(double *kernel_ms, int *launches);
/* per-kernel records of the last spb_bessel call made with SPB_PROFILE on this thread: returns the number of
 * records written (<= max_records); points = elements the launch processed */
int spb_last_profile(int max_records, int *kernel_ids, double *ms, int64_t *points);

/* where device `device` hangs in the host: writes "bdf=<pci address> numa=<node or -1> cpus=<local_cpulist>" (what
 * the per-device worker threads of a multi-device call bind to) into buf; returns the length written or < 0 */
int spb_device_topology(int device, char *buf, int buflen);

int spb_device_count(void);
int spb_version(void);
const char *spb_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SCIPORT_B200_H */
