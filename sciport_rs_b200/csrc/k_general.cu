// k_general.cu — the general per-point path on the GPU: every point the binned
// pipeline does not treat as "simple" (uniform-asymptotic orders, arguments near
// the exponent limits, invalid input, region hand-overs) runs the complete
// driver here.  k_monolithic runs the same driver over a whole batch and is the
// in-GPU cross-check of the binned path (SPB_PATH_MONOLITHIC).
#include "spb_dev.cuh"

namespace spbk {

// list / len: the general list of the classifier, or the second-stage list k_general_shortcuts left behind
__global__ void __launch_bounds__(128) k_general(DevJob job, const unsigned int *list, const unsigned int *len) {
    pdl_enter();
    const unsigned int count = *len;
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const long long i = list[t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        eval_general_store(job, i, z, v);
    }
}

// First stage for the unscaled K / H1 / H2 / Y families: most of their non-simple points sit at huge |Re w|, where
// prepare() has already PROVED that K underflows (CMB_FINAL: value 0 with known codes).  Those are stored here;
// the few that really need the drivers are appended to a second list, so that the expensive kernel runs them in
// full warps instead of one or two lanes per warp.
__global__ void __launch_bounds__(256) k_general_shortcuts(DevJob job, Scratch s) {
    pdl_enter();
    const unsigned int count = s.bins[GLIST_LEN_SLOT];
    const unsigned int stride = gridDim.x * blockDim.x;
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const unsigned int i = s.glist[t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        const spb::Prep p = spb::prepare(job.fam, z, v, job.kode);
        if (p.combine == spb::CMB_FINAL) {
            cplx val(0.0, 0.0);
            if (job.sem && p.final_ierr == 2) {
                // scipy conventions for an overflow of K / H: +inf for K on the positive real axis, NaN otherwise
                const double inf = __longlong_as_double(0x7ff0000000000000ll);
                const bool pos_real = z.re >= 0.0 && z.im == 0.0;
                val = (job.fam == spb::FAM_K && pos_real) ? cplx(inf, 0.0) : cplx(inf - inf, inf - inf);
            }
            store_point(job, i, val, p.final_ierr, p.final_nz);
        } else {
            s.perm_k[atomicAdd(&s.bins[SLOT_G2], 1u)] = i;  // perm_k is free once the K pass has run
        }
    }
}

__global__ void __launch_bounds__(128) k_monolithic(DevJob job) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += stride) {
        cplx z;
        double v;
        load_point(job, i, z, v);
        eval_general_store(job, i, z, v);
    }
}

int launch_general(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    const bool shortcuts = job.kode == 1 && (job.fam == spb::FAM_K || job.fam == spb::FAM_H1 ||
                                             job.fam == spb::FAM_H2 || job.fam == spb::FAM_Y);
    if (shortcuts) {
        launch_pdl(job.pdl != 0, k_general_shortcuts, sms * 4, 256, st, job, s);
        launch_pdl(job.pdl != 0, k_general, resident_grid(k_general, 128, sms), 128, st, job, (const unsigned int *)s.perm_k,
                   (const unsigned int *)(s.bins + SLOT_G2));
        return 2;
    }
    launch_pdl(job.pdl != 0, k_general, resident_grid(k_general, 128, sms), 128, st, job, (const unsigned int *)s.glist,
               (const unsigned int *)(s.bins + GLIST_LEN_SLOT));
    return 1;
}

void launch_monolithic(const DevJob &job, cudaStream_t st) {
    k_monolithic<<<resident_grid(k_monolithic, 128, 148), 128, 0, st>>>(job);
}

}  // namespace spbk
