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
}

// record a failing element in the call-wide "first error" word (rare path: one atomic per failing element)
__device__ __forceinline__ void note_error(unsigned long long *slot, long long index, int ierr) {
    if (ierr != 0 && slot) atomicMin(slot, ((unsigned long long)index << 8) | (unsigned long long)(ierr & 0xff));
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

// general per-point evaluation + optional oracle-compatible post-processing
__device__ __forceinline__ void eval_general_store(const DevJob &j, long long i, cplx z, double v) {
    int ierr, nz;
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
