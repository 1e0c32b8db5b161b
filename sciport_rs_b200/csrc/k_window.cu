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
