// k_window.cu — the batched consumers of the hot path inside the reference's window module:
//
//   kaiser   signal::windows::kaiser (/root/reference/src/signal/fir_filter_design/windows.rs:540-558):
//              l_n = beta sqrt(1 - ((n - alpha)/alpha)^2),  w_n = | I0(l_n) / I0(beta) |,  alpha = (M - 1)/2
//            fused into one pass: argument generation, I_0 through the complex routine (x + 0i, order 0.0, exactly
//            the promotion special::i0 performs, mod.rs:10), the quotient by I0(beta) (one thread evaluates it
//            first) and the modulus.  Neighbouring n have neighbouring arguments, so warps stay inside one region
//            of the algorithm without any sort.  8 bytes stored per window point, nothing read.
//   lanczos  signal::windows::lanczos (windows.rs:509-538): sinc(2 n / (M - 1) - 1), mirrored about the centre.
#include "spb_dev.cuh"
#include "amos/spb_gamma.h"
#include "../../include/sciport_b200.h"

namespace spbk {

__device__ __forceinline__ cplx i0_complex(double x) {
    int ierr = 0, nz = 0;
    return spb::general_eval(spb::FAM_I, cplx(x, 0.0), 0.0, 1, ierr, nz);
}

__global__ void k_kaiser_denominator(double beta, double2 *den) {
    cplx r = i0_complex(beta);
    *den = make_double2(r.re, r.im);
}

// m_ext = window length before the periodic truncation; the first n_out points are stored
__global__ void __launch_bounds__(128) k_kaiser(long long m_ext, long long n_out, double beta, const double2 *den,
                                                double *out) {
    const double alpha = 0.5 * ((double)m_ext - 1.0);
    const double2 d = *den;
    const cplx r(d.x, d.y);
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n_out; n += stride) {
        const double t = ((double)n - alpha) / alpha;
        const double l = beta * sqrt(1.0 - t * t);
        const cplx q = spb::cdiv(i0_complex(l), r);  // (l / r) as complex numbers, then .norm() (windows.rs:552-553)
        out[n] = spb::cabs_(q);
    }
}

// lanczos: w_n = sinc(2 n / (M - 1) - 1) evaluated on the right half and mirrored (windows.rs:513-535)
__global__ void __launch_bounds__(256) k_lanczos(long long m_ext, long long n_out, double *out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double mm = (double)m_ext;
    const long long half = (m_ext % 2 == 0) ? m_ext / 2 : (m_ext + 1) / 2;  // first index of the right half
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < n_out; n += stride) {
        double w;
        if (m_ext % 2 == 1 && n == (m_ext - 1) / 2) {
            w = 1.0;
        } else {
            // left half mirrors the right half: index j in the right half for this n
            long long j = n >= half ? n : (m_ext - 1 - n);
            w = spb::sinc_real(2.0 * (double)j / (mm - 1.0) - 1.0);
        }
        out[n] = w;
    }
}

// ---- the elementwise windows of the reference's window module -----------------------------------
// One fused kernel per window kind (template parameter): index -> argument -> value -> 8-byte store; nothing is read.
// Formulas follow /root/reference/src/signal/fir_filter_design/windows.rs (line ranges per kind below; they are the
// scipy.signal.windows definitions the reference's tests compare against at 1e-14).  m_ext = length of the symmetric
// window before the periodic truncation (windows.rs:628-647); the first n_out points are stored.
template <int KIND>
__device__ __forceinline__ double window_value(long long i, long long m_ext, const WinParams &p, const double *fm) {
    const double n = (double)i, m = (double)m_ext;
    switch (KIND) {
        case SPB_WIN_BOXCAR: return 1.0;  // windows.rs:154-165
        case SPB_WIN_TRIANG: {            // windows.rs:167-196
            const long long j = (2 * i < m_ext) ? i : m_ext - 1 - i;  // mirrored about the centre
            const double k = (double)(j + 1);
            return (m_ext % 2 == 0) ? (2.0 * k - 1.0) / m : 2.0 * k / (m + 1.0);
        }
        case SPB_WIN_GENERAL_COSINE: {    // windows.rs:198-218 (blackman, hamming, hann, flattop, ... map here)
            // fac = linspace(-pi, pi, m): start + i * step, the last point exactly pi
            const double step = (2.0 * SPB_PI) / (m - 1.0);
            const double fac = (i == m_ext - 1) ? SPB_PI : -SPB_PI + n * step;
            double w = 0.0;
            for (int k = 0; k < p.n; ++k) w += p.a[k] * cos((double)k * fac);
            return w;
        }
        case SPB_WIN_BARTLETT:            // windows.rs:252-270
            return (n <= (m - 1.0) / 2.0) ? 2.0 * n / (m - 1.0) : 2.0 - 2.0 * n / (m - 1.0);
        case SPB_WIN_PARZEN: {            // windows.rs:301-320
            const double x = fabs(n - (m - 1.0) / 2.0);
            const double r = x / (m / 2.0);
            if (x <= (m - 1.0) / 4.0) return 1.0 - 6.0 * (r * r) + 6.0 * (r * r * r);
            const double q = 1.0 - r;
            return 2.0 * (q * q * q);
        }
        case SPB_WIN_BOHMAN: {            // windows.rs:322-340
            if (i == 0 || i == m_ext - 1) return 0.0;
            const double step = 2.0 / (m - 1.0);
            const double fac = fabs(-1.0 + n * step);
            return (1.0 - fac) * cos(SPB_PI * fac) + (1.0 / SPB_PI) * sin(SPB_PI * fac);
        }
        case SPB_WIN_BARTHANN: {          // windows.rs:358-373
            const double fac = fabs(n / (m - 1.0) - 0.5);
            return 0.62 - 0.48 * fac + 0.38 * cos(2.0 * SPB_PI * fac);
        }
        case SPB_WIN_COSINE:              // windows.rs:375-387
            return sin(SPB_PI / m * (n + 0.5));
        case SPB_WIN_EXPONENTIAL: {       // windows.rs:389-417: a[0] = center (NaN: (m - 1)/2), a[1] = tau
            const double c = (p.a[0] == p.a[0]) ? p.a[0] : (m - 1.0) / 2.0;
            return exp(-fabs(n - c) / p.a[1]);
        }
        case SPB_WIN_TUKEY: {             // windows.rs:419-455: a[0] = alpha in (0, 1)
            const double alpha = p.a[0];
            const long long width = (long long)floor(alpha * (m - 1.0) / 2.0);
            if (i <= width) return 0.5 * (1.0 + cos(SPB_PI * (-1.0 + 2.0 * n / alpha / (m - 1.0))));
            if (i >= m_ext - width - 1) return 0.5 * (1.0 + cos(SPB_PI * (-2.0 / alpha + 1.0 + 2.0 * n / alpha / (m - 1.0))));
            return 1.0;
        }
        case SPB_WIN_TAYLOR: {            // windows.rs:457-508: 1 + 2 sum_k F_k cos(2 pi k (n - m/2 + 1/2) / m)
            const int nb = (int)p.a[0] - 1;
            const double x = (n - m / 2.0 + 0.5) / m;
            double acc = 0.0;
            for (int k = 0; k < nb; ++k) acc += fm[k] * cos(2.0 * SPB_PI * (double)(k + 1) * x);
            return 1.0 + 2.0 * acc;
        }
        case SPB_WIN_GAUSSIAN: {          // windows.rs:591-605: a[0] = std
            const double x = n - (m - 1.0) / 2.0;
            return exp(-(x * x) / (2.0 * p.a[0] * p.a[0]));
        }
        case SPB_WIN_GENERAL_GAUSSIAN: {  // windows.rs:607-622: a[0] = power, a[1] = width
            const double x = n - (m - 1.0) / 2.0;
            return exp(-0.5 * pow(fabs(x / p.a[1]), 2.0 * p.a[0]));
        }
        default: return 0.0;
    }
}

// Taylor coefficients F_k, k = 1..nbar-1 (windows.rs:465-486), one thread each into shared memory
__device__ __forceinline__ void taylor_coefficients(const WinParams &p, double *fm) {
    const int nbar = (int)p.a[0];
    const double sll = p.a[1];
    const int mi = threadIdx.x;
    if (mi < nbar - 1) {
        const double b = pow(10.0, sll / 20.0);
        const double a = acosh(b) / SPB_PI;
        const double fnbar = (double)nbar;
        const double s2 = fnbar * fnbar / (a * a + (fnbar - 0.5) * (fnbar - 0.5));
        const double m2 = (double)(mi + 1) * (double)(mi + 1);
        double numer = (mi % 2 == 0) ? 1.0 : -1.0, denom = 2.0;
        for (int j = 0; j < nbar - 1; ++j) {
            const double mj = (double)(j + 1);
            numer *= 1.0 - m2 / s2 / (a * a + (mj - 0.5) * (mj - 0.5));
            if (j != mi) denom *= 1.0 - m2 / (mj * mj);
        }
        fm[mi] = numer / denom;
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) k_window(long long m_ext, long long n_out, WinParams p, double *out) {
    __shared__ double fm[SPB_WIN_MAX_PARAMS];
    __shared__ double s_norm;
    if (KIND == SPB_WIN_TAYLOR) {
        taylor_coefficients(p, fm);
        __syncthreads();
        if (threadIdx.x == 0) {
            // norm (a[2] != 0): divide by the value at the centre (m - 1)/2 (scipy's `norm=True`)
            double nv = 1.0;
            if (p.a[2] != 0.0) {
                const int nb = (int)p.a[0] - 1;
                const double m = (double)m_ext;
                const double x = ((m - 1.0) / 2.0 - m / 2.0 + 0.5) / m;
                double acc = 0.0;
                for (int k = 0; k < nb; ++k) acc += fm[k] * cos(2.0 * SPB_PI * (double)(k + 1) * x);
                nv = 1.0 + 2.0 * acc;
            }
            s_norm = nv;
        }
        __syncthreads();
    }
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_out; i += stride) {
        double w = window_value<KIND>(i, m_ext, p, fm);
        if (KIND == SPB_WIN_TAYLOR) w = w / s_norm;
        out[i] = w;
    }
}

// Kaiser-Bessel derived window (windows.rs:560-589) from the Kaiser window k[0..h] of length h + 1 = m/2 + 1 that
// k_kaiser has left in `k`: out[i] = out[m-1-i] = sqrt(cumsum(k)[i] / cumsum(k)[h]), i < h.  One CTA: a block
// reduction for the total, then a chunked block scan (1024 values per round, running carry in a register).
__global__ void __launch_bounds__(1024) k_kbd_finish(const double *k, long long h, double *out) {
    __shared__ double warp_part[32];
    __shared__ double s_total, s_round;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double part = 0.0;
    for (long long i = threadIdx.x; i <= h; i += 1024) part += k[i];
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) warp_part[wid] = part;
    __syncthreads();
    if (wid == 0) {
        double t = warp_part[lane];
        for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) s_total = t;
    }
    __syncthreads();
    const double total = s_total;
    double carry = 0.0;
    for (long long base = 0; base < h; base += 1024) {
        const long long i = base + threadIdx.x;
        const double x = (i < h) ? k[i] : 0.0;
        double incl = x;
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        __syncthreads();  // warp_part / s_round of the previous round have been read by everyone
        if (lane == 31) warp_part[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            const double w = warp_part[lane];
            double wi = w;
            for (int d = 1; d < 32; d <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_part[lane] = wi - w;
            if (lane == 31) s_round = wi;
        }
        __syncthreads();
        if (i < h) {
            const double c = carry + warp_part[wid] + incl;
            const double w = sqrt(c / total);
            out[i] = w;
            out[2 * h - 1 - i] = w;
        }
        carry += s_round;
    }
}

template <int KIND>
static void launch_window_kind(long long m_ext, long long n_out, const WinParams &p, double *out, cudaStream_t st) {
    long long blocks = (n_out + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_window<KIND><<<(int)blocks, 256, 0, st>>>(m_ext, n_out, p, out);
}

void launch_window(int kind, long long m_ext, long long n_out, const WinParams &p, double *out, cudaStream_t st) {
    switch (kind) {
#define SPB_WIN_CASE(K) case K: launch_window_kind<K>(m_ext, n_out, p, out, st); break;
        SPB_WIN_CASE(SPB_WIN_BOXCAR)
        SPB_WIN_CASE(SPB_WIN_TRIANG)
        SPB_WIN_CASE(SPB_WIN_GENERAL_COSINE)
        SPB_WIN_CASE(SPB_WIN_BARTLETT)
        SPB_WIN_CASE(SPB_WIN_PARZEN)
        SPB_WIN_CASE(SPB_WIN_BOHMAN)
        SPB_WIN_CASE(SPB_WIN_BARTHANN)
        SPB_WIN_CASE(SPB_WIN_COSINE)
        SPB_WIN_CASE(SPB_WIN_EXPONENTIAL)
        SPB_WIN_CASE(SPB_WIN_TUKEY)
        SPB_WIN_CASE(SPB_WIN_TAYLOR)
        SPB_WIN_CASE(SPB_WIN_GAUSSIAN)
        SPB_WIN_CASE(SPB_WIN_GENERAL_GAUSSIAN)
#undef SPB_WIN_CASE
        default: break;
    }
}

void launch_kbd_finish(const double *k, long long h, double *out, cudaStream_t st) {
    k_kbd_finish<<<1, 1024, 0, st>>>(k, h, out);
}

void launch_kaiser(long long m_ext, long long n_out, double beta, double2 *den, double *out, int sms, cudaStream_t st) {
    k_kaiser_denominator<<<1, 1, 0, st>>>(beta, den);
    long long blocks = (n_out + 127) / 128;
    long long cap = (long long)resident_grid(k_kaiser, 128, sms);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_kaiser<<<(int)blocks, 128, 0, st>>>(m_ext, n_out, beta, den, out);
}

void launch_lanczos(long long m_ext, long long n_out, double *out, cudaStream_t st) {
    long long blocks = (n_out + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_lanczos<<<(int)blocks, 256, 0, st>>>(m_ext, n_out, out);
}

}  // namespace spbk
