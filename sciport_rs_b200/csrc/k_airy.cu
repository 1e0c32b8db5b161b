// k_airy.cu — Airy Ai, Ai', Bi, Bi' over a batch (one thread per point).
#include "spb_dev.cuh"
#include "amos/spb_airy.h"

namespace spbk {

__global__ void __launch_bounds__(128) k_airy(int which, int kode, int sem, const double2 *z, double2 *out, int *ierr,
                                              int *nz, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 t = __ldg(z + i);
        cplx zz(t.x, t.y);
        int ie = 0, nzz = 0;
        cplx val;
        if (which < 2)
            val = spb::airy(zz, which, kode, nzz, ie);
        else
            val = spb::biry(zz, which - 2, kode, ie);
        if (sem) val = spb::airy_semantics(zz, val, ie);
        out[i] = make_double2(val.re, val.im);
        if (ierr) ierr[i] = ie;
        if (nz) nz[i] = nzz;
    }
}

void launch_airy(int which, int kode, int sem, const double2 *z, double2 *out, int *ierr, int *nz, long long n,
                 cudaStream_t st) {
    long long blocks = (n + 127) / 128;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    k_airy<<<(int)blocks, 128, 0, st>>>(which, kode, sem, z, out, ierr, nz, n);
}

}  // namespace spbk
