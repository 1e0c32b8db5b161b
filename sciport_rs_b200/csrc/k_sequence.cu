// k_sequence.cu — order sequences: out[k*stride + i] = F_{v_i + k}(z_i), k = 0..n_orders-1 (BASELINE config 5:
// J_n(z), n = 0..255 per point).  One thread per point runs the sequence form of the driver (series /
// asymptotic / Miller / uniform expansions for the top members, then the backward recurrence in the order),
// writing every member straight into its [order][point] slot through a strided view, so that at each order
// the 32 lanes of a warp store 32 consecutive double2 values (coalesced) and no per-thread array exists.
// The recurrences re-read the two members just written (L1/L2 hits).
#include "spb_dev.cuh"

namespace spbk {

__global__ void __launch_bounds__(128) k_sequence(DevJob job, int n_orders, long long out_stride) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += stride) {
        cplx z;
        double v;
        load_point(job, i, z, v);
        spb::StridedSeq y{reinterpret_cast<cplx *>(job.out + i), out_stride};
        int ierr = 0, nz = 0;
        if (!(v == v) || !(z.re == z.re) || !(z.im == z.im)) {
            ierr = 1;
        } else {
            switch (job.fam) {
                case spb::FAM_J: nz = spb::besj(z, v, job.kode, n_orders, y, ierr); break;
                case spb::FAM_I: nz = spb::besi(z, v, job.kode, n_orders, y, ierr); break;
                case spb::FAM_K: nz = spb::besk(z, v, job.kode, n_orders, y, ierr); break;
                case spb::FAM_H1: nz = spb::besh(z, v, job.kode, 1, n_orders, y, ierr); break;
                case spb::FAM_H2: nz = spb::besh(z, v, job.kode, 2, n_orders, y, ierr); break;
                case spb::FAM_Y: {
                    // Y has no sequence recurrence of its own here: member by member through its Hankel pair
                    for (int k = 0; k < n_orders; ++k) {
                        cplx cy[1];
                        cy[0] = cplx(0.0, 0.0);
                        int ie = 0;
                        int nzk = spb::besy(z, v + (double)k, job.kode, cy, ie);
                        y[k] = cy[0];
                        if (ie != 0 && (ierr == 0 || ierr == 3)) ierr = ie;
                        nz += nzk;
                    }
                    break;
                }
                default: ierr = 1; break;
            }
        }
        if (ierr != 0 && ierr != 3 && job.fam != spb::FAM_Y) {
            for (int k = 0; k < n_orders; ++k) y[k] = cplx(0.0, 0.0);  // failed points carry zeros like the single-order path
        }
        if (job.ierr) job.ierr[i] = ierr;
        if (job.nz) job.nz[i] = nz;
    }
}

void launch_sequence(const DevJob &job, int n_orders, long long out_stride, int sms, cudaStream_t st) {
    k_sequence<<<resident_grid(k_sequence, 128, sms), 128, 0, st>>>(job, n_orders, out_stride);
}

}  // namespace spbk
