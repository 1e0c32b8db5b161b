// k_ipass.cuh — kernel templates of the I pass, shared by k_ipass.cu (one kernel per I region) and k_fused.cu
// (the fused I + K form of the large-|w| region kernel, one instantiation per family and scaling flag).
#pragma once
#include "spb_dev.cuh"

namespace spbk {

// minimum resident CTAs per SM the register allocation must allow (occupancy hides the FP64 dependency
// latency of the region routines; measured on B200)
#ifndef SPB_IPASS_MINB_ASYI
#define SPB_IPASS_MINB_ASYI 8
#endif
// (Debye-type uniform kernel: 3 CTAs need <= 168 registers, which is what the routine takes when left alone)
#define SPB_IPASS_MINB(R) ((R) == 1 ? SPB_IPASS_MINB_ASYI : (R) == 4 ? 3 : 1)
#ifndef SPB_FUSED_MINB
#define SPB_FUSED_MINB 5
#endif
// (fused Miller + Wronskian kernel: 96 registers without spills, 5 CTAs per SM; left alone it takes 104 = 4 CTAs)
#ifndef SPB_WRSKF_MINB
#define SPB_WRSKF_MINB 5
#endif

// one point of the I pass: I_v(w) by the region routine; J / I families finish here, the others park I(w)
// FUSED (large-|w| region, families that need K as well): I_v(w) and K_v(w) by the fused routine, combine, store
// DIRECT (k_direct.cu: the point has no class byte and no list entry yet): a point the routine cannot finish is
// reported to the caller (false) instead of being appended to the general list, and nothing is parked
template <int REGION, bool FUSED = false, bool DIRECT = false>
__device__ __forceinline__ bool ipass_point(const DevJob &job, const Scratch &s, long long i, cplx z, double v,
                                            const spb::OrderPhase *ph, const double *gl = nullptr) {
    // the classifier has already established region and simplicity of this point: geometry only
    spb::Prep p = spb::prepare_pass(job.fam, z, !FUSED || REGION != spb::IREG_ASYI);
    if (FUSED) {
        cplx ival, kval;
        int inz = 0, knz = 0;
        const int rc = (REGION == spb::IREG_WRSK)
                           ? spb::wrsk_ik_region(p.w, v, job.kode, ival, inz, kval, knz, p.az)
                           : spb::fused_ik_region(p.w, v, job.kode, ival, inz, kval, knz, p.az, ph);
        if (rc < 0 || (knz != 0 && (p.combine == spb::CMB_ACON || p.combine == spb::CMB_Y_SHARED))) {
            if (DIRECT) return false;
            unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
            s.glist[g] = (unsigned int)i;
            return false;
        }
        finish_store(job, p, i, z, v, ival, inz, kval, knz, ph);
        return true;
    }
    cplx ival;
    int inz = spb::ipass_region(REGION, p.w, v, job.kode, ival, p.az, ph, nullptr, gl);
    if (inz < 0) {
        if (DIRECT) return false;
        // the region wants to hand over: finish this point on the general path
        unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
        s.glist[g] = (unsigned int)i;
        s.icode[i] = (signed char)-1;
        return false;
    }
    if (DIRECT || p.combine == spb::CMB_I_ROT) {  // (DIRECT without FUSED: the J and I families only)
        int ierr, nz;
        cplx val = spb::finish(job.fam, p, z, v, job.kode, ival, inz, cplx(0.0, 0.0), 0, ierr, nz, ph);
        store_point(job, i, val, ierr, nz);
    } else {
        s.itemp[i] = make_double2(ival.re, ival.im);
        s.icode[i] = (signed char)inz;
    }
    return true;
}

// SCALAR = the order is one broadcast value (v_stride = 0): its phase factors are computed once per CTA into shared
// memory; otherwise they are computed per point
// FAM >= 0: the fused form of the kernel with the family and the scaling flag as compile-time constants
template <int REGION, bool SCALAR, int FAM = -1, int KODE = 0>
__global__ void __launch_bounds__(128, FAM >= 0 ? (REGION == spb::IREG_ASYI ? SPB_FUSED_MINB : SPB_WRSKF_MINB) : SPB_IPASS_MINB(REGION))
    k_ipass(DevJob job, Scratch s) {
    pdl_enter();
    if (FAM >= 0) {
        job.fam = FAM;
        job.kode = KODE;
    }
    __shared__ spb::OrderPhase s_ph;
    // broadcast order, power series / Miller routine: the order-only values of ln Gamma once per CTA (order_gln; the
    // two of the Miller routine on two warps, side by side) instead of once per point — a quarter of the
    // instructions of the series kernel on config 4
    constexpr bool GLN = SCALAR && FAM < 0 && (REGION == spb::IREG_SERIES || REGION == spb::IREG_MLRI);
    __shared__ double s_gl[3];
    if (SCALAR) {
        if (threadIdx.x == 0) s_ph = spb::order_phase(scalar_order(job));
        if (GLN) {
            const double vs = scalar_order(job);
            if (REGION == spb::IREG_SERIES && threadIdx.x == 0) s_gl[0] = spb::order_gln(vs, 0);
            if (REGION == spb::IREG_MLRI && (threadIdx.x == 32 || threadIdx.x == 64))
                s_gl[threadIdx.x >> 5] = spb::order_gln(vs, (int)(threadIdx.x >> 5));
        }
        __syncthreads();
    }
    const unsigned int start = s.bins[REGION] - s.bins[0];
    const unsigned int count = s.bins[REGION + 1] - s.bins[REGION];
    const unsigned int stride = gridDim.x * blockDim.x;
    // iterative regions of a cell-sorted job walk the re-sorted index array (k_pipeline.cu)
    const unsigned int *perm = (job.cellsort ? s.perm_i2 : s.perm_i) + start;
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
            ipass_point<REGION, (FAM >= 0)>(job, s, i, z, v, &s_ph, GLN ? s_gl : nullptr);
        } else {
            const spb::OrderPhase ph = spb::order_phase(v);
            ipass_point<REGION, (FAM >= 0)>(job, s, i, z, v, &ph);
        }
    }
}

// Per-element orders, Miller / Wronskian region: batches grouped by predicted work (spb_dev.cuh)
template <int REGION>
__global__ void __launch_bounds__(128) k_ipass_sorted(DevJob job, Scratch s) {
    pdl_enter();
    __shared__ SortSmem sm;
    __shared__ unsigned int s_idx[SORT_B];
    __shared__ double2 s_z[SORT_B];
    __shared__ double s_v[SORT_B];
    const unsigned int start = s.bins[REGION] - s.bins[0];
    const unsigned int count = s.bins[REGION + 1] - s.bins[REGION];
    const unsigned int *perm = s.perm_i + start;
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
                bucket[r] = bucket_of(wrsk_work(spb::cabs_(z), v), 0.1);
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
                const spb::OrderPhase ph = spb::order_phase(vv);
                ipass_point<REGION>(job, s, s_idx[slot], cplx(zz.x, zz.y), vv, &ph);
            }
        }
        __syncthreads();
    }
}

}  // namespace spbk
