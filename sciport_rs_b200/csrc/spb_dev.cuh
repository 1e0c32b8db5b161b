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

__device__ __forceinline__ void store_point(const DevJob &j, long long i, cplx val, int ierr, int nz) {
    j.out[i] = make_double2(val.re, val.im);  // one 128-bit store per point
    if (j.ierr) j.ierr[i] = ierr;
    if (j.nz) j.nz[i] = nz;
}

// general per-point evaluation + optional oracle-compatible post-processing
__device__ __forceinline__ void eval_general_store(const DevJob &j, long long i, cplx z, double v) {
    int ierr, nz;
    cplx val = spb::general_eval(j.fam, z, v, j.kode, ierr, nz);
    if (j.sem) val = spb::apply_semantics(j.fam, z, v, j.kode, val, ierr, nz);
    store_point(j, i, val, ierr, nz);
}

}  // namespace spbk
