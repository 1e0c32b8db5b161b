// k_debye.cu — region kernel of the Debye-direct region (amos/spb_debye.h): one evaluation of the uniform expansion
// at the order itself gives I_v(w), K_v(w) or both for a point, followed by the family's combine step and the store.
//
// The classifier carves this region out of the Miller / Wronskian / raised-order regions (classic cost 2500-2700
// flops per point, data-dependent trip counts) wherever |s| = |sqrt(v^2 + w^2)| and |s|^3 / v^2 are large enough;
// the routine here is straight-line code of a few hundred flops (14 Horner chains with compile-time coefficients),
// so the lanes of a warp stay together up to the convergence exit of the series.  K-type points of the region carry
// no K bin (spb::KREG_FUSED): like the fused large-|w| kernel this one finishes them completely.
//
// Same two-deep software pipeline over the sorted index list as the other region kernels (k_ipass.cuh).
#include "spb_dev.cuh"

namespace spbk {

#ifndef SPB_DEBYE_MINB
#define SPB_DEBYE_MINB 4
#endif

template <int FAM>
__device__ __forceinline__ void debye_point(const DevJob &job, const Scratch &s, long long i, cplx z, double v,
                                            const spb::OrderPhase *ph) {
    // the classifier has already established region and simplicity of this point: geometry only
    spb::Prep p = spb::prepare_pass(job.fam, z);
    const bool need_i = (FAM == spb::FAM_J || FAM == spb::FAM_I) ? true : (p.ireg >= 0);
    cplx ival, kval;
    int inz = 0, knz = 0;
    const int rc = spb::debye_region(job.fam, p.w, v, job.kode, need_i, ival, inz, kval, knz, p.az, ph);
    if (rc < 0) {
        // off the middle scaling level after all, or the series did not converge: general path
        unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
        s.glist[g] = (unsigned int)i;
        return;
    }
    finish_store(job, p, i, z, v, ival, inz, kval, knz, ph);
}

template <bool SCALAR, int FAM, int KODE>
__global__ void __launch_bounds__(128, SPB_DEBYE_MINB) k_debye(DevJob job, Scratch s) {
    pdl_enter();
    job.fam = FAM;
    job.kode = KODE;
    __shared__ spb::OrderPhase s_ph;
    if (SCALAR) {
        if (threadIdx.x == 0) s_ph = spb::order_phase(scalar_order(job));
        __syncthreads();
    }
    const unsigned int start = s.bins[spb::IREG_DEBYE] - s.bins[0];
    const unsigned int count = s.bins[spb::IREG_DEBYE + 1] - s.bins[spb::IREG_DEBYE];
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int *perm = s.perm_i + start;
    unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool have1 = t < count, have2 = (t + stride) < count;
    unsigned int i1 = have1 ? __ldg(perm + t) : 0u;
    unsigned int i2 = have2 ? __ldg(perm + t + stride) : 0u;
    cplx z1(0.0, 0.0);
    double v1 = 0.0;
    if (have1) load_point(job, i1, z1, v1);
#pragma unroll 1
    while (have1) {
        const long long i = i1;
        const cplx z = z1;
        const double v = v1;
        // advance the pipeline before the arithmetic so that both loads overlap it
        t += stride;
        have1 = have2;
        i1 = i2;
        have2 = (t + stride) < count;
        if (have2) i2 = __ldg(perm + t + stride);
        if (have1) load_point(job, i1, z1, v1);
        if (SCALAR) {
            debye_point<FAM>(job, s, i, z, v, &s_ph);
        } else {
            const spb::OrderPhase ph = spb::order_phase(v);
            debye_point<FAM>(job, s, i, z, v, &ph);
        }
    }
}

template <int FAM, int KODE>
struct DebyeLaunch {
    static void run(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
        if (job.v_stride == 0)
            launch_pdl(job.pdl != 0, k_debye<true, FAM, KODE>, resident_grid(k_debye<true, FAM, KODE>, 128, sms), 128, st,
                       job, s);
        else
            launch_pdl(job.pdl != 0, k_debye<false, FAM, KODE>, resident_grid(k_debye<false, FAM, KODE>, 128, sms), 128,
                       st, job, s);
    }
};

void launch_debye(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    dispatch_fam_kode<DebyeLaunch>(job.fam, job.kode, job, s, sms, st);
}

}  // namespace spbk
