// k_cgamma.cu — complex gamma / loggamma over a batch (16 B in, 16 B out per point).
#include "spb_dev.cuh"
#include "amos/spb_gamma.h"

namespace spbk {

template <int WHICH>
__global__ void __launch_bounds__(256) k_cgamma(const double2 *z, double2 *out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 t = __ldg(z + i);
        cplx r = WHICH == 0 ? spb::cgamma(cplx(t.x, t.y)) : spb::cloggamma(cplx(t.x, t.y));
        out[i] = make_double2(r.re, r.im);
    }
}

void launch_cgamma(int which, const double2 *z, double2 *out, long long n, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    if (which == 0)
        k_cgamma<0><<<(int)blocks, 256, 0, st>>>(z, out, n);
    else
        k_cgamma<1><<<(int)blocks, 256, 0, st>>>(z, out, n);
}

}  // namespace spbk
