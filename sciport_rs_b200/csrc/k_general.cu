// k_general.cu — the general per-point path on the GPU: every point the binned
// pipeline does not treat as "simple" (uniform-asymptotic orders, arguments near
// the exponent limits, invalid input, region hand-overs) runs the complete
// driver here.  k_monolithic runs the same driver over a whole batch and is the
// in-GPU cross-check of the binned path (SPB_PATH_MONOLITHIC).
#include "spb_dev.cuh"

namespace spbk {

// list / len: the general list of the classifier, or the second-stage list k_general_shortcuts left behind
// Entry t of the list goes to warp (t mod W), lane (t div W) mod 32, W = warps of the grid: a short list (config 3
// leaves about ninety such points per 2^24-point chunk) is spread one point per warp over the whole machine instead
// of filling three warps of one SM, whose lanes then walked up to 32 different branches of the driver one after the
// other (measured: 89 us per chunk for 91 points); a long list fills all 32 lanes of every warp as before.
__global__ void __launch_bounds__(128, 4) k_general(DevJob job, const unsigned int *list, const unsigned int *len) {
    pdl_enter();
    const unsigned int count = *len;
    const unsigned int nwarps = (gridDim.x * blockDim.x) >> 5;
    // (warps numbered CTA-minor: the first gridDim.x entries land on different CTAs, i.e. on different SMs)
    const unsigned int warp = (threadIdx.x >> 5) * gridDim.x + blockIdx.x, lane = threadIdx.x & 31;
    const unsigned long long stride = 32ull * nwarps;
#pragma unroll 1
    for (unsigned long long t = (unsigned long long)lane * nwarps + warp; t < count; t += stride) {
        const long long i = list[t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        eval_general_store(job, i, z, v);
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
    // (points whose result prepare() has already proved -- CMB_FINAL -- were stored by the classifier)
    launch_pdl(job.pdl != 0, k_general, resident_grid(k_general, 128, sms), 128, st, job, (const unsigned int *)s.glist,
               (const unsigned int *)(s.bins + GLIST_LEN_SLOT));
    return 1;
}

void launch_monolithic(const DevJob &job, cudaStream_t st) {
    k_monolithic<<<resident_grid(k_monolithic, 128, 148), 128, 0, st>>>(job);
}

}  // namespace spbk
