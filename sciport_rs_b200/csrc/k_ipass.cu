// k_ipass.cu — I pass of the binned pipeline: one kernel per I region
// (power series / large-|z| asymptotic / Miller + Neumann / Miller + Wronskian),
// each a persistent grid that walks its bin segment of perm_i.  For the J and I
// families the kernel also applies the final rotation and stores the result;
// for K-type families it parks I(w) in itemp for the K pass.
#include "spb_dev.cuh"

namespace spbk {

template <int REGION>
__global__ void __launch_bounds__(128) k_ipass(DevJob job, Scratch s) {
    const unsigned int start = s.bins[REGION] - s.bins[0];
    const unsigned int count = s.bins[REGION + 1] - s.bins[REGION];
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const long long i = s.perm_i[start + t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        spb::Prep p = spb::prepare(job.fam, z, v, job.kode);
        cplx ival;
        int inz = spb::ipass_region(REGION, p.w, v, job.kode, ival);
        if (inz < 0) {
            // the region wants to hand over: finish this point on the general path
            unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
            s.glist[g] = (unsigned int)i;
            s.icode[i] = (signed char)-1;
            continue;
        }
        if (p.combine == spb::CMB_I_ROT) {
            int ierr, nz;
            cplx val = spb::finish(job.fam, p, z, v, job.kode, ival, inz, cplx(0.0, 0.0), 0, ierr, nz);
            store_point(job, i, val, ierr, nz);
        } else {
            s.itemp[i] = make_double2(ival.re, ival.im);
            s.icode[i] = (signed char)inz;
        }
    }
}

void launch_ipass(const DevJob &job, const Scratch &s, int region, int sms, cudaStream_t st) {
    // persistent grid: SM count x resident CTAs per SM, 128-thread CTAs
    switch (region) {
        case spb::IREG_SERIES:
            k_ipass<spb::IREG_SERIES><<<resident_grid(k_ipass<spb::IREG_SERIES>, 128, sms), 128, 0, st>>>(job, s);
            break;
        case spb::IREG_ASYI:
            k_ipass<spb::IREG_ASYI><<<resident_grid(k_ipass<spb::IREG_ASYI>, 128, sms), 128, 0, st>>>(job, s);
            break;
        case spb::IREG_MLRI:
            k_ipass<spb::IREG_MLRI><<<resident_grid(k_ipass<spb::IREG_MLRI>, 128, sms), 128, 0, st>>>(job, s);
            break;
        case spb::IREG_WRSK:
            k_ipass<spb::IREG_WRSK><<<resident_grid(k_ipass<spb::IREG_WRSK>, 128, sms), 128, 0, st>>>(job, s);
            break;
        default: break;
    }
}

}  // namespace spbk
