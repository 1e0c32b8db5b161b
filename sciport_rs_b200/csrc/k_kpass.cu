// k_kpass.cu — K pass of the binned pipeline: K_v(w) for the right-half-plane
// point of every K-type point (families K, H1, H2, Y), walking perm_k (sorted by
// K region: Temme series / Miller / half-odd closed form), then the family's
// combine step (rotation, analytic continuation with the parked I(w), Hankel
// difference for Y) and the final store.
//
// Same two-deep software pipeline as the I pass: the point (z, v, parked I(w) and
// its code) of iteration t+stride and the permutation index of t+2*stride are
// loaded while iteration t computes.
#include "spb_dev.cuh"

namespace spbk {

#ifndef SPB_KPASS_MINB
#define SPB_KPASS_MINB 4
#endif

// one point of the K pass: K_v(w), combine with the parked I(w), store
__device__ __forceinline__ void kpass_point(const DevJob &job, const Scratch &s, long long i, cplx z, double v,
                                            double2 itv, int icv, const spb::OrderPhase *ph) {
    // the classifier has already established regions and simplicity of this point: geometry only
    spb::Prep p = spb::prepare_pass(job.fam, z);
    cplx ival(0.0, 0.0);
    int inz = 0;
    if (p.ireg >= 0) {
        inz = icv;
        if (inz < 0) return;  // already queued on the general path by the I pass
        ival = cplx(itv.x, itv.y);
    }
    cplx kval;
    int knz = spb::kpass_region(p.w, v, job.kode, kval, p.az);
    if (knz < 0 || (knz != 0 && (p.combine == spb::CMB_ACON || p.combine == spb::CMB_Y_SHARED))) {
        unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
        s.glist[g] = (unsigned int)i;
        return;
    }
    finish_store(job, p, i, z, v, ival, inz, kval, knz, ph);
}

// broadcast scalar order (v_stride = 0): the order-dependent phase factors are computed once per CTA and live in
// shared memory, not in registers of the persistent loop
// SCALAR = false: per-element orders of a cell-sorted job (perm_k2: neighbours in the list are neighbours in |z| and
// in the order), phase factors per point
template <int FAM, int KODE, bool SCALAR = true>
__global__ void __launch_bounds__(128, SPB_KPASS_MINB) k_kpass(DevJob job, Scratch s) {
    pdl_enter();
    job.fam = FAM;  // compile-time family and scaling flag (dispatch_fam_kode)
    job.kode = KODE;
    __shared__ spb::OrderPhase s_ph;
    if (SCALAR) {
        if (threadIdx.x == 0) s_ph = spb::order_phase(scalar_order(job));
        __syncthreads();
    }
    const unsigned int count = s.bins[BIN_GENERAL] - s.bins[BIN_K0];
    const unsigned int stride = gridDim.x * blockDim.x;
    const unsigned int *perm = job.cellsort ? s.perm_k2 : s.perm_k;
    unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool have1 = t < count, have2 = (t + stride) < count;
    unsigned int i1 = have1 ? __ldg(perm + t) : 0u;
    unsigned int i2 = have2 ? __ldg(perm + t + stride) : 0u;
    cplx z1(0.0, 0.0);
    double v1 = 0.0;
    double2 it1 = make_double2(0.0, 0.0);
    int ic1 = 0;
    if (have1) {
        load_point(job, i1, z1, v1);
        it1 = s.itemp[i1];
        ic1 = s.icode[i1];
    }
#pragma unroll 1
    while (have1) {
        const long long i = i1;
        const cplx z = z1;
        const double v = v1;
        const double2 itv = it1;
        const int icv = ic1;
        t += stride;
        have1 = have2;
        i1 = i2;
        have2 = (t + stride) < count;
        if (have2) i2 = __ldg(perm + t + stride);
        if (have1) {
            load_point(job, i1, z1, v1);
            it1 = s.itemp[i1];  // only meaningful where the point went through the I pass
            ic1 = s.icode[i1];
        }
        if (SCALAR) {
            kpass_point(job, s, i, z, v, itv, icv, &s_ph);
        } else {
            const spb::OrderPhase ph = spb::order_phase(v);
            kpass_point(job, s, i, z, v, itv, icv, &ph);
        }
    }
}

// Per-element orders: batches of SORT_B points staged in shared memory and grouped by predicted work
// (spb_dev.cuh) instead of the register pipeline above.
template <int FAM, int KODE>
__global__ void __launch_bounds__(128, SPB_KPASS_MINB) k_kpass_sorted(DevJob job, Scratch s) {
    pdl_enter();
    job.fam = FAM;
    job.kode = KODE;
    __shared__ SortSmem sm;
    __shared__ unsigned int s_idx[SORT_B];
    __shared__ double2 s_z[SORT_B], s_it[SORT_B];
    __shared__ double s_v[SORT_B];
    __shared__ signed char s_ic[SORT_B];
    const unsigned int count = s.bins[BIN_GENERAL] - s.bins[BIN_K0];
    const unsigned int *perm = s.perm_k;
#pragma unroll 1
    for (unsigned int t0 = blockIdx.x * SORT_B; t0 < count; t0 += gridDim.x * SORT_B) {
        const unsigned int m = min((unsigned int)SORT_B, count - t0);
        unsigned int bucket[SORT_R];
#pragma unroll
        for (int r = 0; r < SORT_R; ++r) {
            const unsigned int slot = threadIdx.x + 128 * r;
            bucket[r] = 0xffu;
            if (slot < m) {
                const unsigned int i = __ldg(perm + t0 + slot);
                cplx z;
                double v;
                load_point(job, i, z, v);
                s_idx[slot] = i;
                s_z[slot] = make_double2(z.re, z.im);
                s_v[slot] = v;
                s_it[slot] = s.itemp[i];
                s_ic[slot] = s.icode[i];
                bucket[r] = bucket_of(kpass_work(spb::cabs_(z), v), 0.25);
            }
        }
        batch_sort(sm, bucket);
#pragma unroll 1
        for (int r = 0; r < SORT_R; ++r) {
            const unsigned int pos = threadIdx.x + 128 * r;
            if (pos < m) {
                const unsigned int slot = sm.order[pos];
                const double2 zz = s_z[slot];
                const double vv = s_v[slot];
                const spb::OrderPhase ph = spb::order_phase(vv);  // per-element orders: one phase per point
                kpass_point(job, s, s_idx[slot], cplx(zz.x, zz.y), vv, s_it[slot], s_ic[slot], &ph);
            }
        }
        __syncthreads();  // the next batch overwrites the staging arrays
    }
}

template <int FAM, int KODE>
struct KpassLaunch {
    static void run(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
        if (job.cellsort && job.v_stride != 0)
            launch_pdl(job.pdl != 0, k_kpass<FAM, KODE, false>, resident_grid(k_kpass<FAM, KODE, false>, 128, sms), 128, st,
                       job, s);
        else if (job.v_stride != 0 || job.n < 2)
            launch_pdl(job.pdl != 0, k_kpass_sorted<FAM, KODE>, resident_grid(k_kpass_sorted<FAM, KODE>, 128, sms), 128, st, job, s);
        else
            launch_pdl(job.pdl != 0, k_kpass<FAM, KODE>, resident_grid(k_kpass<FAM, KODE>, 128, sms), 128, st, job, s);
    }
};
// J and I never reach the K pass: their instantiations stay empty
template <int KODE>
struct KpassLaunch<spb::FAM_J, KODE> {
    static void run(const DevJob &, const Scratch &, int, cudaStream_t) {}
};
template <int KODE>
struct KpassLaunch<spb::FAM_I, KODE> {
    static void run(const DevJob &, const Scratch &, int, cudaStream_t) {}
};

void launch_kpass(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    dispatch_fam_kode<KpassLaunch>(job.fam, job.kode, job, s, sms, st);
}

}  // namespace spbk
