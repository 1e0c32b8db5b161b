// k_ipass.cu — I pass of the binned pipeline: one kernel per I region
// (power series / large-|z| asymptotic / Miller + Neumann / Miller + Wronskian /
// uniform expansion of Debye type / uniform expansion of Airy type),
// each a persistent grid that walks its bin segment of perm_i.  For the J and I
// families the kernel also applies the final rotation and stores the result;
// for K-type families it parks I(w) in itemp for the K pass.
//
// The walk is software-pipelined two deep: while point t is being evaluated the
// 16-byte load of point t+stride is in flight and the permutation index of point
// t+2*stride is being fetched, so the dependent pair of global loads
// (perm_i -> z) never sits on the critical path of a warp.
#include "k_ipass.cuh"

namespace spbk {

template <int REGION>
static void launch_ipass_region(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    if (job.v_stride == 0)
        launch_pdl(job.pdl != 0, k_ipass<REGION, true>, resident_grid(k_ipass<REGION, true>, 128, sms), 128, st, job, s);
    else
        launch_pdl(job.pdl != 0, k_ipass<REGION, false>, resident_grid(k_ipass<REGION, false>, 128, sms), 128, st, job, s);
}

void launch_ipass(const DevJob &job, const Scratch &s, int region, int sms, cudaStream_t st) {
    // persistent grid: SM count x resident CTAs per SM, 128-thread CTAs
    switch (region) {
        case spb::IREG_SERIES: launch_ipass_region<spb::IREG_SERIES>(job, s, sms, st); break;
        case spb::IREG_ASYI:
            if (job.fuse & spb::FUSE_ASYI)
                launch_fused(job, s, sms, st, spb::IREG_ASYI);  // I and K of these points in one kernel (k_fused.cu)
            else
                launch_ipass_region<spb::IREG_ASYI>(job, s, sms, st);
            break;
        case spb::IREG_MLRI: launch_ipass_region<spb::IREG_MLRI>(job, s, sms, st); break;
        case spb::IREG_WRSK:
            if ((job.fuse & spb::FUSE_WRSK) && !(job.fam == spb::FAM_J || job.fam == spb::FAM_I))
                launch_fused(job, s, sms, st, spb::IREG_WRSK);  // K_v(w) from the Wronskian normalisation: no K pass
            else if (job.v_stride != 0 && !job.cellsort)
                launch_pdl(job.pdl != 0, k_ipass_sorted<spb::IREG_WRSK>, resident_grid(k_ipass_sorted<spb::IREG_WRSK>, 128, sms), 128, st,
                           job, s);
            else
                launch_ipass_region<spb::IREG_WRSK>(job, s, sms, st);
            break;
        case spb::IREG_UNI1: launch_ipass_region<spb::IREG_UNI1>(job, s, sms, st); break;
        case spb::IREG_UNI2: launch_ipass_region<spb::IREG_UNI2>(job, s, sms, st); break;
        case spb::IREG_DEBYE: launch_debye(job, s, sms, st); break;
        default: break;
    }
}

}  // namespace spbk
