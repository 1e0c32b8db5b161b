// k_kpass.cu — K pass of the binned pipeline: K_v(w) for the right-half-plane
// point of every K-type point (families K, H1, H2, Y), walking perm_k (sorted by
// K region: Temme series / Miller / half-odd closed form), then the family's
// combine step (rotation, analytic continuation with the parked I(w), Hankel
// difference for Y) and the final store.
#include "spb_dev.cuh"

namespace spbk {

__global__ void __launch_bounds__(128) k_kpass(DevJob job, Scratch s) {
    const unsigned int count = s.bins[BIN_GENERAL] - s.bins[4];
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const long long i = s.perm_k[t];
        cplx z;
        double v;
        load_point(job, i, z, v);
        spb::Prep p = spb::prepare(job.fam, z, v, job.kode);
        cplx ival(0.0, 0.0);
        int inz = 0;
        if (p.ireg >= 0) {
            inz = s.icode[i];
            if (inz < 0) continue;  // already queued on the general path by the I pass
            double2 t2 = s.itemp[i];
            ival = cplx(t2.x, t2.y);
        }
        cplx kval;
        int knz = spb::kpass_region(p.w, v, job.kode, kval);
        if (knz < 0 || (knz != 0 && (p.combine == spb::CMB_ACON || p.combine == spb::CMB_Y_SHARED))) {
            unsigned int g = atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u);
            s.glist[g] = (unsigned int)i;
            continue;
        }
        int ierr, nz;
        if (job.fam == spb::FAM_JY) {
            cplx jv, yv;
            int yerr, ynz;
            spb::finish_jy(p, z, v, job.kode, ival, inz, kval, knz, jv, ierr, nz, yv, yerr, ynz);
            store_point(job, i, jv, ierr, nz);
            store_point2(job, i, yv, yerr, ynz);
            continue;
        }
        cplx val = spb::finish(job.fam, p, z, v, job.kode, ival, inz, kval, knz, ierr, nz);
        store_point(job, i, val, ierr, nz);
    }
}

void launch_kpass(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    k_kpass<<<resident_grid(k_kpass, 128, sms), 128, 0, st>>>(job, s);
}

}  // namespace spbk
