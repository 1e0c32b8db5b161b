// k_direct.cu — large-|w| points of a broadcast-order job evaluated straight from the input order, BEFORE the
// classification pipeline.
//
// On a gridded batch (config 2: one order, a 4096 x 4096 grid of arguments) 98 % of the points sit deep inside the
// large-|w| region, and neighbours in memory are neighbours in the plane.  The binned pipeline still spent a third
// of such a step on them outside the arithmetic: the classifier read every point (16 B) to write a class byte, the
// compaction turned the bytes into an index list (4 B written, 4 B read back), and the region kernel gathered the
// points through that list.  Here a warp takes 32 CONSECUTIVE points (one coalesced 512-byte load), runs the fast
// accept of the classifier on them (spb::fast_classify: |z|^2 against a window, a sufficient condition that can only
// agree with prepare() or decline) and, if all 32 lanes accept, evaluates them on the spot with the very routine the
// region kernel would run (ipass_point: same code, same roundings) and stores the results.  One flag byte per group
// (Scratch::done) tells the classifier and the compaction which groups are finished: a tile whose 64 groups are all
// done costs them one 64-byte read; everything else — declined groups, the ragged last group, a group in which the
// routine reported a point it cannot finish — goes through the unchanged pipeline.
//
// A batch of unrelated points (config 4) never has 32 accepting lanes at once: a warp that has declined its first
// DIRECT_BAIL groups without a single accept stops (the flags are cleared before the launch and only ever set), and
// when no warp has finished anything the classifier runs exactly as it does without this kernel.  The launcher side
// (capi.cu) keeps the kernel away from where it cannot pay: chunks below 2^22 points never get it, and a multi-chunk
// call reads the outcome of its first chunk back without blocking (DirectProbe) and drops it from the later chunks
// if it finished nothing.
#include "k_ipass.cuh"

namespace spbk {

constexpr int DIRECT_BAIL = 8;

template <int FAM, int KODE>
__global__ void __launch_bounds__(128, (FAM == spb::FAM_J || FAM == spb::FAM_I) ? SPB_IPASS_MINB_ASYI : SPB_FUSED_MINB)
    k_direct(DevJob job, Scratch s) {
    pdl_enter();
    job.fam = FAM;
    job.kode = KODE;
    constexpr bool KTYPE = !(FAM == spb::FAM_J || FAM == spb::FAM_I);
    __shared__ spb::OrderPhase s_ph;
    const double v = scalar_order(job);
    if (threadIdx.x == 0) s_ph = spb::order_phase(v);
    __syncthreads();
    const spb::FastOrder fo = spb::fast_order(v);
    const int lane = threadIdx.x & 31;
    const unsigned int nwarps = (gridDim.x * blockDim.x) >> 5;
    const unsigned int ngroups = (unsigned int)s.nblocks * (TILE / 32);  // flags exist for every group of every tile
    unsigned int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned int accepted = 0, declined = 0;
    // one group ahead: the load of the next group is in flight while this one is evaluated
    cplx z1(0.0, 0.0);
    double vdummy;
    bool in1 = g < ngroups && (long long)g * 32 + lane < job.n;
    if (in1) load_point(job, (long long)g * 32 + lane, z1, vdummy);
#pragma unroll 1
    while (g < ngroups) {
        const long long i = (long long)g * 32 + lane;
        const cplx z = z1;
        const bool in = in1;
        const unsigned int gn = g + nwarps;
        in1 = gn < ngroups && (long long)gn * 32 + lane < job.n;
        if (in1) load_point(job, (long long)gn * 32 + lane, z1, vdummy);
        int ireg, kreg;
        const bool ok = in && spb::fast_classify(FAM, z, fo, KODE, job.fuse, ireg, kreg);
        // (K-type families: the launcher has made sure that job.fuse carries FUSE_ASYI, so an accepted point is a
        // KREG_FUSED point of the large-|w| region)
        bool fin = false;
        if (__all_sync(0xffffffffu, ok)) {
            const bool good = ipass_point<spb::IREG_ASYI, KTYPE, true>(job, s, i, z, v, &s_ph);
            fin = __all_sync(0xffffffffu, good);
        }
        // the flags were cleared before the launch: only finished groups are marked
        if (fin) {
            if (lane == 0) s.done[g] = 1;
            ++accepted;
        } else if (++declined >= (unsigned int)DIRECT_BAIL && accepted == 0) {
            break;  // a batch of unrelated points: leave it to the classifier
        }
        g = gn;
    }
    // (the classifier and the compaction look at the flags only if this counter is non-zero)
    if (lane == 0 && accepted) atomicAdd(&s.bins[SLOT_DIRECT], accepted * 32u);
}

template <int FAM, int KODE>
struct DirectLaunch {
    static void run(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
        cudaMemsetAsync(s.done, 0, (size_t)s.nblocks * (TILE / 32), st);
        launch_pdl(job.pdl != 0, k_direct<FAM, KODE>, resident_grid(k_direct<FAM, KODE>, 128, sms), 128, st, job, s);
    }
};

void launch_direct(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    dispatch_fam_kode<DirectLaunch>(job.fam, job.kode, job, s, sms, st);
}

}  // namespace spbk
