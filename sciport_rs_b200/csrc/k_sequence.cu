// k_sequence.cu — order sequences: out[k*stride + i] = F_{v_i + k}(z_i), k = 0..n_orders-1 (BASELINE config 5:
// J_n(z), n = 0..255 per point; 16 bytes stored per eval, so the path is bound by its stores).
//
//   k_sweep     J and I families: the write-once two-pass recurrence of amos/spb_sweep.h.  One thread per
//               point; every member is produced in registers and stored exactly once, as one 16-byte store
//               into its [order][point] slot, so at each order the lanes of a warp write consecutive double2
//               values.  Points the sweep declines (arguments near the exponent limits, very large |z|, short
//               sequences) are appended to a list.
//   k_sequence  the classic sequence form of the drivers through a strided view of the result array (members
//               are re-read by the recurrences): K, H1, H2, Y sequences, and the points k_sweep declined.
#include "spb_dev.cuh"

namespace spbk {

// store-only view of one point's column of the [order][point] result: y[k] = value is one 128-bit store
struct SeqStore {
    double2 *base;
    long long stride;
    struct Slot {
        double2 *p;
        __device__ __forceinline__ void operator=(const cplx &v) const { *p = make_double2(v.re, v.im); }
    };
    __device__ __forceinline__ Slot operator[](int k) const { return Slot{base + (long long)k * stride}; }
};

}  // namespace spbk

// pass B of the sweep steps down the [order][point] column of its point: one pointer subtraction per member
template <>
struct spb::SweepCursor<spbk::SeqStore> {
    double2 *p;
    long long stride;
    __device__ __forceinline__ SweepCursor(spbk::SeqStore y, int top) : p(y.base + (long long)top * y.stride), stride(y.stride) {}
    __device__ __forceinline__ void put(spb::cplx v) {
        *p = make_double2(v.re, v.im);
        p -= stride;
    }
};

namespace spbk {

// ---- K-pair pre-pass of the sweep --------------------------------------------------------------------------
// The sweep normalises through the Wronskian with e^w K_fnu(w), e^w K_fnu+1(w).  For |w| >= rl (nine points in
// ten of config 5) both come from the large-|w| series inside k_sweep; below, they need the Miller algorithm of
// bknu(), whose three data-dependent loops would hold a whole warp for the three or four lanes that need it.
// Those points are listed (k_sweep_mark) and their pairs evaluated densely (k_sweep_kpair), parked in rows 0 and 1
// of the point's own result column -- the two rows pass B of k_sweep writes last -- where k_sweep picks them up.
__global__ void __launch_bounds__(256) k_sweep_mark(DevJob job, unsigned int *defer, unsigned int *defer_len) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += stride) {
        cplx z;
        double v;
        load_point(job, i, z, v);
        if (!(spb::cabs_(z) >= SPB_RL)) defer[atomicAdd(defer_len, 1u)] = (unsigned int)i;  // (also NaN: declined later)
    }
}

__global__ void __launch_bounds__(128) k_sweep_kpair(DevJob job, long long out_stride, const unsigned int *defer,
                                                     const unsigned int *defer_len) {
    const unsigned int count = *defer_len;
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const long long i = defer[t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        const cplx w = spb::sweep_point_w(job.fam == spb::FAM_J, z);
        cplx ck[2];
        ck[0] = cplx(0.0, 0.0);
        ck[1] = cplx(0.0, 0.0);
        const int nw = spb::bknu(w, v, 2, 2, ck, SPB_TOL, SPB_ELIM, SPB_ALIM, spb::cabs_(z));
        if (nw != 0) {  // failed or set a member to zero: the sweep declines the point (NaN normalising factor)
            const double nan = __longlong_as_double(0x7ff8000000000000ll);
            ck[0] = cplx(nan, nan);
            ck[1] = cplx(nan, nan);
        }
        job.out[i] = make_double2(ck[0].re, ck[0].im);
        job.out[out_stride + i] = make_double2(ck[1].re, ck[1].im);
    }
}

__global__ void __launch_bounds__(128) k_sweep(DevJob job, int n_orders, long long out_stride, unsigned int *list,
                                               unsigned int *list_len) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += stride) {
        cplx z;
        double v;
        load_point(job, i, z, v);
        SeqStore y{job.out + i, out_stride};
        // the pair parked by k_sweep_kpair (same test as k_sweep_mark)
        cplx kpre[2];
        kpre[0] = cplx(0.0, 0.0);
        kpre[1] = cplx(0.0, 0.0);
        if (!(spb::cabs_(z) >= SPB_RL)) {
            const double2 a = job.out[i], b = job.out[out_stride + i];
            kpre[0] = cplx(a.x, a.y);
            kpre[1] = cplx(b.x, b.y);
        }
        int ierr = 0, nz = 0;
        bool declined = true;
        if (job.fam == spb::FAM_J)
            nz = spb::besj_sweep(z, v, job.kode, n_orders, y, ierr, declined, kpre);
        else
            nz = spb::besi_sweep(z, v, job.kode, n_orders, y, ierr, declined, kpre);
        if (declined) {
            list[atomicAdd(list_len, 1u)] = (unsigned int)i;
            continue;
        }
        if (job.ierr) job.ierr[i] = ierr;
        if (job.nz) job.nz[i] = nz;
        note_error(job.first_err, job.index_base + i, ierr);
    }
}

// classic sequence drivers; list == nullptr: every point of the chunk, else the listed points
__global__ void __launch_bounds__(128) k_sequence(DevJob job, int n_orders, long long out_stride,
                                                  const unsigned int *list, const unsigned int *list_len) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long count = list ? (long long)*list_len : job.n;
#pragma unroll 1
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const long long i = list ? (long long)list[t] : t;
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
        note_error(job.first_err, job.index_base + i, ierr);
    }
}

int launch_sequence(const DevJob &job, int n_orders, long long out_stride, const Scratch &s, int sms,
                    cudaStream_t st) {
    if (job.fam == spb::FAM_J || job.fam == spb::FAM_I) {
        unsigned int *len = s.bins + GLIST_LEN_SLOT, *dlen = s.bins + SLOT_G2;
        cudaMemsetAsync(len, 0, 3 * sizeof(unsigned int), st);  // GLIST_LEN_SLOT, SLOT_DONE (stays 0), SLOT_G2
        // (perm_i is free on this path: the list of the pre-pass)
        long long mark_blocks = (job.n + 255) / 256;
        if (mark_blocks > (long long)sms * 8) mark_blocks = (long long)sms * 8;
        if (n_orders >= SPB_SWEEP_MIN_ORDERS) {  // (shorter sequences: the sweep declines every point anyway)
            k_sweep_mark<<<(int)mark_blocks, 256, 0, st>>>(job, s.perm_i, dlen);
            k_sweep_kpair<<<resident_grid(k_sweep_kpair, 128, sms), 128, 0, st>>>(job, out_stride, s.perm_i, dlen);
        }
        k_sweep<<<resident_grid(k_sweep, 128, sms), 128, 0, st>>>(job, n_orders, out_stride, s.glist, len);
        k_sequence<<<resident_grid(k_sequence, 128, sms), 128, 0, st>>>(job, n_orders, out_stride, s.glist, len);
        return 4;
    }
    k_sequence<<<resident_grid(k_sequence, 128, sms), 128, 0, st>>>(job, n_orders, out_stride, nullptr, nullptr);
    return 1;
}

}  // namespace spbk
