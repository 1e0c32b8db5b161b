// spb_dev.cuh — device-side helpers shared by the kernel translation units:
// coalesced 16-byte point loads / result stores and the persistent-grid loop.
#pragma once
#include "spb_kernels.cuh"
#include "amos/spb_family.h"

namespace spbk {

using spb::cplx;

__device__ __forceinline__ void load_point(const DevJob &j, long long i, cplx &z, double &v) {
    if (j.z) {
        double2 t = __ldg(j.z + i);  // one 128-bit load per point
        z = cplx(t.x, t.y);
    } else {
        z = cplx(__ldg(j.x + i), 0.0);
    }
    v = __ldg(j.v + i * j.v_stride);
    // K_{-v} = K_v: fold the order where it is loaded, so negative orders of K take the binned path like any other
    if (j.reflect && j.fam == spb::FAM_K) v = fabs(v);
}

// the broadcast order of a scalar-order job (v_stride = 0), folded like load_point() folds it
__device__ __forceinline__ double scalar_order(const DevJob &j) {
    double v = __ldg(j.v);
    if (j.reflect && j.fam == spb::FAM_K) v = fabs(v);
    return v;
}

// record a failing element in the call-wide "first error" word (smallest (index, ierr) of the call).  Failing
// elements need not be rare (a quarter of an unscaled log-uniform batch over- or underflows), and a million atomics
// on one address serialise (measured: 0.7 ms per 4M points): the word is read first -- an L2 load; a stale value only
// costs a redundant atomic, never a wrong result -- and the atomic is issued only by a candidate that would lower it.
__device__ __forceinline__ void note_error(unsigned long long *slot, long long index, int ierr) {
    if (ierr != 0 && slot) {
        const unsigned long long cand = ((unsigned long long)index << 8) | (unsigned long long)(ierr & 0xff);
        if (cand < __ldcg(slot)) atomicMin(slot, cand);
    }
}

__device__ __forceinline__ void store_point(const DevJob &j, long long i, cplx val, int ierr, int nz) {
    j.out[i] = make_double2(val.re, val.im);  // one 128-bit store per point
    if (j.ierr) j.ierr[i] = ierr;
    if (j.nz) j.nz[i] = nz;
    note_error(j.first_err, j.index_base + i, ierr);
}

__device__ __forceinline__ void store_point2(const DevJob &j, long long i, cplx val, int ierr, int nz) {
    j.out2[i] = make_double2(val.re, val.im);
    if (j.ierr2) j.ierr2[i] = ierr;
    if (j.nz2) j.nz2[i] = nz;
    note_error(j.first_err2, j.index_base + i, ierr);
}

// family combine (rotation, continuation with I(w), Hankel difference; for the JY pair both results) and store
__device__ __forceinline__ void finish_store(const DevJob &job, const spb::Prep &p, long long i, cplx z, double v,
                                             cplx ival, int inz, cplx kval, int knz, const spb::OrderPhase *ph) {
    int ierr, nz;
    if (job.fam == spb::FAM_JY) {
        cplx jv, yv;
        int yerr, ynz;
        spb::finish_jy(p, z, v, job.kode, ival, inz, kval, knz, jv, ierr, nz, yv, yerr, ynz, ph);
        store_point(job, i, jv, ierr, nz);
        store_point2(job, i, yv, yerr, ynz);
        return;
    }
    if (job.fam == spb::FAM_IK) {
        cplx iv, kv;
        int kerr, kz;
        spb::finish_ik(p, z, v, job.kode, ival, inz, kval, knz, iv, ierr, nz, kv, kerr, kz, ph);
        store_point(job, i, iv, ierr, nz);
        store_point2(job, i, kv, kerr, kz);
        return;
    }
    cplx val = spb::finish(job.fam, p, z, v, job.kode, ival, inz, kval, knz, ierr, nz, ph);
    store_point(job, i, val, ierr, nz);
}

// Per-thread cache of the order-dependent phase factors: one sincos per DISTINCT order a thread meets.  With a
// broadcast scalar order (v_stride = 0) it is filled once per thread; with per-element orders it is refilled
// per point (one compare extra).  The cached value is exactly spb::order_phase(v), so results do not depend on
// whether the cache hit.
struct PhaseCache {
    double v;
    spb::OrderPhase ph;
    __device__ __forceinline__ PhaseCache() : v(__builtin_nan("")) {}
    __device__ __forceinline__ const spb::OrderPhase *get(double fnu) {
        if (!(fnu == v)) {
            ph = spb::order_phase(fnu);
            v = fnu;
        }
        return &ph;
    }
};

// ---- block-level grouping by predicted work ----------------------------------------------------------------
// With per-element orders the trip counts of the region routines (forward recurrence over the order, Miller
// start index, ratio walk) differ from lane to lane even inside one region bin (measured: 14 of 32 lanes active
// in the K pass of config 3).  The passes then work in batches of SORT_B points per CTA: the batch is staged
// in shared memory, a counting sort by a cheap work estimate (32 buckets) orders it, and thread t takes the
// sorted positions t, t + 128, ... so that the lanes of a warp run points of similar length in every round.
constexpr int SORT_R = 4;              // points per thread per batch
constexpr int SORT_B = 128 * SORT_R;   // points per batch
constexpr int SORT_BUCKETS = 32;

struct SortSmem {
    unsigned int cnt[SORT_BUCKETS];
    unsigned short order[SORT_B];
};

// bucket[r] is the bucket of slot threadIdx.x + 128 r, or 0xff for an empty slot.  On return sm.order[0..m) holds
// the slots grouped by ascending bucket (m = number of non-empty slots).
__device__ __forceinline__ void batch_sort(SortSmem &sm, const unsigned int (&bucket)[SORT_R]) {
    if (threadIdx.x < SORT_BUCKETS) sm.cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned int rank[SORT_R];
#pragma unroll
    for (int r = 0; r < SORT_R; ++r)
        if (bucket[r] != 0xffu) rank[r] = atomicAdd(&sm.cnt[bucket[r]], 1u);
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned int c = sm.cnt[threadIdx.x];
        unsigned int incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)threadIdx.x >= d) incl += y;
        }
        sm.cnt[threadIdx.x] = incl - c;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_R; ++r)
        if (bucket[r] != 0xffu) sm.order[sm.cnt[bucket[r]] + rank[r]] = (unsigned short)(threadIdx.x + 128 * r);
    __syncthreads();
}

// Which side of the family's continuation a point lies on: the half of the z plane where the combine step takes the
// analytic continuation (K, the I/K pair and the Hankel functions: K_v(w) plus the I_v(w) term, an exponential and
// an underflow test) or, for Y and the J/Y pair, which of the two Hankel terms is the continued one.  A grouping
// hint only (top bit of the cell key; bit 6 of the class byte, which orders the stable bins inside a tile): warps
// that are uniform in it do not walk both forms of the combine step (measured on config 3: the combine step ran with
// 16 of 32 lanes and took a quarter of the issue slots of the Debye-direct kernel).
__device__ __forceinline__ unsigned int side_of(int fam, cplx z) {
    switch (fam) {
        case spb::FAM_H1: return z.im < 0.0 ? 1u : 0u;
        case spb::FAM_H2: return z.im > 0.0 ? 1u : 0u;
        case spb::FAM_Y:
        case spb::FAM_JY: return z.im < 0.0 ? 1u : 0u;
        default: return z.re < 0.0 ? 1u : 0u;  // I, K, the I/K pair (J: no branch depends on it)
    }
}

// Work estimates (arbitrary units, only their order matters) -> bucket 0..31
__device__ __forceinline__ unsigned int bucket_of(double est, double scale) {
    double b = est * scale;
    return b >= 31.0 ? 31u : (unsigned int)b;
}

// K_v(w): Miller start index ~ 9 + 175/|w| for 2 < |w| < 28.7 (about 12 beyond), each step two divisions;
// then int(v + 1/2) forward-recurrence steps
__device__ __forceinline__ double kpass_work(double az, double v) {
    double k = az <= 2.0 ? 20.0 : (az < 28.7 ? 9.0 + 175.0 / az : 12.0);
    return 3.0 * k + v;
}

// I_v(w) by ratios + Wronskian: walk of ~40 steps plus max(0, |w| - v) backward steps, plus K_v, K_v+1
__device__ __forceinline__ double wrsk_work(double az, double v) {
    double d = az - v;
    return 40.0 + (d > 0.0 ? d : 0.0) + 0.25 * kpass_work(az, v);
}

// general per-point evaluation + optional oracle-compatible post-processing
__device__ __forceinline__ void eval_general_store(const DevJob &j, long long i, cplx z, double v) {
    int ierr, nz;
    if (j.reflect && v < 0.0) {
        // negative order: reflection formulae composed from the drivers' results at |v| (spb_family.h)
        if (j.fam == spb::FAM_JY) {
            cplx a, b;
            int ierr2, nz2;
            spb::reflect_eval_jy(z, v, j.kode, a, ierr, nz, b, ierr2, nz2, j.sem != 0);
            store_point(j, i, a, ierr, nz);
            store_point2(j, i, b, ierr2, nz2);
            return;
        }
        if (j.fam == spb::FAM_IK) {
            cplx a = spb::reflect_eval(spb::FAM_I, z, v, j.kode, ierr, nz, j.sem != 0);
            store_point(j, i, a, ierr, nz);
            cplx b = spb::reflect_eval(spb::FAM_K, z, v, j.kode, ierr, nz, j.sem != 0);
            store_point2(j, i, b, ierr, nz);
            return;
        }
        cplx val = spb::reflect_eval(j.fam, z, v, j.kode, ierr, nz, j.sem != 0);
        store_point(j, i, val, ierr, nz);
        return;
    }
    if (j.fam == spb::FAM_IK) {
        cplx a = spb::general_eval(spb::FAM_I, z, v, j.kode, ierr, nz);
        if (j.sem) a = spb::apply_semantics(spb::FAM_I, z, v, j.kode, a, ierr, nz);
        store_point(j, i, a, ierr, nz);
        cplx b = spb::general_eval(spb::FAM_K, z, v, j.kode, ierr, nz);
        if (j.sem) b = spb::apply_semantics(spb::FAM_K, z, v, j.kode, b, ierr, nz);
        store_point2(j, i, b, ierr, nz);
        return;
    }
    if (j.fam == spb::FAM_JY) {
        cplx a = spb::general_eval(spb::FAM_J, z, v, j.kode, ierr, nz);
        if (j.sem) a = spb::apply_semantics(spb::FAM_J, z, v, j.kode, a, ierr, nz);
        store_point(j, i, a, ierr, nz);
        cplx b = spb::general_eval(spb::FAM_Y, z, v, j.kode, ierr, nz);
        if (j.sem) b = spb::apply_semantics(spb::FAM_Y, z, v, j.kode, b, ierr, nz);
        store_point2(j, i, b, ierr, nz);
        return;
    }
    cplx val = spb::general_eval(j.fam, z, v, j.kode, ierr, nz);
    if (j.sem) val = spb::apply_semantics(j.fam, z, v, j.kode, val, ierr, nz);
    store_point(j, i, val, ierr, nz);
}

}  // namespace spbk
