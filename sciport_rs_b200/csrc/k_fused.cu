// k_fused.cu — fused I + K form of the large-|w| region kernel.
//
// Families that need I_v(w) AND K_v(w) at the same right-half-plane point (the J/Y pair and Y everywhere, K and
// the Hankel functions where they are continued into the left half plane) spend most of a large-|z| batch in
// the asymptotic region of I.  There both routines start from the same quantities — 1/w, 1/sqrt(w), the sincos
// of Im w, exp(-Re w) — so one kernel evaluates I, K, the family combine and the store for such a point
// (spb::fused_ik_region): no parked I(w), no second walk over the point, and the shared quantities once.
// The classifier gives these points no K bin (spb::KREG_FUSED), so the K pass never sees them.
#include "k_ipass.cuh"

namespace spbk {

template <int REGION, int FAM, int KODE>
static void launch_fused_region(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    if (job.v_stride == 0)
        launch_pdl(job.pdl != 0, k_ipass<REGION, true, FAM, KODE>, resident_grid(k_ipass<REGION, true, FAM, KODE>, 128, sms),
                   128, st, job, s);
    else
        launch_pdl(job.pdl != 0, k_ipass<REGION, false, FAM, KODE>,
                   resident_grid(k_ipass<REGION, false, FAM, KODE>, 128, sms), 128, st, job, s);
}

// region: IREG_ASYI (I and K from one large-|w| series) or IREG_WRSK (K_v(w) from the Wronskian normalisation of I)
template <int FAM, int KODE>
struct FusedLaunch {
    static void run(const DevJob &job, const Scratch &s, int sms, cudaStream_t st, int region) {
        if (region == spb::IREG_WRSK)
            launch_fused_region<spb::IREG_WRSK, FAM, KODE>(job, s, sms, st);
        else
            launch_fused_region<spb::IREG_ASYI, FAM, KODE>(job, s, sms, st);
    }
};
// J and I never need K
template <int KODE>
struct FusedLaunch<spb::FAM_J, KODE> {
    static void run(const DevJob &, const Scratch &, int, cudaStream_t, int) {}
};
template <int KODE>
struct FusedLaunch<spb::FAM_I, KODE> {
    static void run(const DevJob &, const Scratch &, int, cudaStream_t, int) {}
};

void launch_fused(const DevJob &job, const Scratch &s, int sms, cudaStream_t st, int region) {
    dispatch_fam_kode<FusedLaunch>(job.fam, job.kode, job, s, sms, st, region);
}

}  // namespace spbk
