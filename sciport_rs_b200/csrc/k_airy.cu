// k_airy.cu — Airy Ai, Ai', Bi, Bi' over a batch.
//
// Same shape as the Bessel pipeline: a classification pass writes a region key per point (Maclaurin series,
// K_{1/3} by Temme series / Miller algorithm, analytic continuation through I by power series / Miller /
// large-|zeta| expansion), the shared scan + stable compaction sorts the point indices by key, and one
// persistent kernel walks the sorted list so that the 32 lanes of a warp run the same region of the
// algorithm.  The key only groups work: every point still goes through the complete routine, so a point whose
// key is off by one region (the key uses |zeta| = (2/3)|z|^(3/2) and arg z without forming zeta) is merely
// evaluated in a mixed warp.
#include "spb_dev.cuh"
#include "amos/spb_airy.h"
#include "amos/spb_airy_lean.h"

namespace spbk {

// bins 0..5 of the I-bin field: 0 Maclaurin; 1 K series; 2 K Miller; 3 I series; 4 I Miller; 5 large |zeta| (>= rl:
// K and I from the large-|zeta| routines only, lean kernel); non-finite arguments go to the complete routine
__device__ __forceinline__ unsigned int airy_key(bool bi, cplx z, double &azeta) {
    const double az = spb::cabs_(z);
    azeta = 0.0;
    if (az <= 1.0) return 0u;
    azeta = (2.0 / 3.0) * az * sqrt(az);
    if (azeta >= SPB_RL * (1.0 + 1.0e-9) && az < 1.0e300) return 5u;  // (margin: the routine forms zeta itself)
    if (!(azeta < SPB_RL)) return 4u;  // at the border, NaN, Inf: complete routine
    // Ai: K_{1/3}(zeta) directly when |arg z| <= pi/3 (Re zeta >= 0, Re z > 0); otherwise, and always for Bi,
    // through I_{+-1/3}(zeta)
    if (!bi && z.re >= 0.5 * az) return azeta > 2.0 ? 2u : 1u;
    if (azeta <= 2.0 || azeta * azeta * 0.25 <= 4.0 / 3.0) return 3u;
    return 4u;
}

// Points whose result the range / overflow / underflow tests decide (amos/spb_airy_lean.h: airy_pretest) are stored
// here and get no bin.
__global__ void __launch_bounds__(256) k_classify_airy(int which, int kode, int sem, const double2 *z, double2 *out,
                                                       int *ierr, int *nz, long long n, Scratch s,
                                                       unsigned long long *first_err, long long index_base) {
    __shared__ unsigned int hist[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * TILE;
    unsigned int cnt[NB_I];
#pragma unroll
    for (int b = 0; b < NB_I; ++b) cnt[b] = 0;
    for (int k = threadIdx.x; k < TILE; k += 256) {
        long long i = base + k;
        if (i < n) {
            double2 t = __ldg(z + i);
            const cplx zc(t.x, t.y);
            double azeta;
            unsigned int key = airy_key(which >= 2, zc, azeta);
            // (the pre-test can only decide a point whose |Re zeta| reaches alim; |zeta| bounds it)
            if (key == 5u && kode == 1 && azeta >= SPB_ALIM * (1.0 - 1.0e-9)) {
                spb::AiryPre pre;
                if (spb::airy_pretest(which >= 2, zc, kode, pre) == spb::AIRY_DONE) {
                    cplx val(0.0, 0.0);
                    if (sem) val = spb::airy_semantics(zc, val, pre.ierr);
                    out[i] = make_double2(val.re, val.im);
                    if (ierr) ierr[i] = pre.ierr;
                    if (nz) nz[i] = pre.nz;
                    note_error(first_err, index_base + i, pre.ierr);
                    key = 7u;  // no bin
                }
            }
            s.cls[i] = (unsigned char)(key | (3u << 3));  // no K bin, not general
#pragma unroll
            for (int b = 0; b < NB_I; ++b) cnt[b] += (key == (unsigned int)b);
        }
    }
#pragma unroll
    for (int b = 0; b < NB_I; ++b) {
        unsigned int c = cnt[b];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
        if ((threadIdx.x & 31) == 0 && c) atomicAdd(&hist[b], c);
    }
    __syncthreads();
    if (threadIdx.x < NBINS) s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x] = hist[threadIdx.x];
}

template <bool BI>
__device__ __forceinline__ void airy_point(int id, int kode, int sem, cplx zz, long long i, double2 *out, int *ierr,
                                           int *nz, unsigned long long *first_err, long long index_base) {
    int ie = 0, nzz = 0;
    cplx val = BI ? spb::biry(zz, id, kode, ie) : spb::airy(zz, id, kode, nzz, ie);
    if (sem) val = spb::airy_semantics(zz, val, ie);
    out[i] = make_double2(val.re, val.im);
    if (ierr) ierr[i] = ie;
    if (nz) nz[i] = nzz;
    note_error(first_err, index_base + i, ie);
}

// rough length of the evaluation at |zeta| inside one region key (arbitrary units; only the order matters)
__device__ __forceinline__ double airy_work(cplx z) {
    const double az = spb::cabs_(z);
    if (!(az > 1.0)) return 10.0 + 20.0 * az;  // Maclaurin terms grow with |z|
    const double azeta = (2.0 / 3.0) * az * sqrt(az);
    // Miller start indices fall like 1/|zeta| (K) or grow like |zeta| (I below rl); asymptotic series shorten
    double k = azeta <= 2.0 ? 25.0 : (azeta < 28.7 ? 9.0 + 175.0 / azeta : 12.0);
    double i = azeta < SPB_RL ? 30.0 + 2.0 * azeta : 12.0 + 600.0 / azeta;
    return 2.0 * k + i;
}

// Key 0 (|z| <= 1, the Maclaurin series: 43 % of a log-uniform batch) in its own lean kernel: a few dozen
// registers at full occupancy instead of a share of the heavy kernel below (255 registers, two CTAs per SM)
template <bool BI>
__global__ void __launch_bounds__(256) k_airy_small(int id, int kode, int sem, const double2 *z, double2 *out, int *ierr,
                                                    int *nz, Scratch s, unsigned long long *first_err,
                                                    long long index_base) {
    const unsigned int count = s.bins[1] - s.bins[0];
    const unsigned int *perm = s.perm_i;
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const unsigned int i = __ldg(perm + t);
        const double2 zz = __ldg(z + i);
        const cplx zc(zz.x, zz.y);
        const double az = spb::cabs_(zc);  // (the classifier's test, so az <= 1 here)
        cplx val = BI ? spb::biry_small(zc, az, id, kode) : spb::airy_small(zc, az, id, kode);
        if (sem) val = spb::airy_semantics(zc, val, 0);
        out[i] = make_double2(val.re, val.im);
        if (ierr) ierr[i] = 0;
        if (nz) nz[i] = 0;
    }
    (void)first_err;
    (void)index_base;
}

// keys 1..4 (1 < |z|, |zeta| < rl): persistent walk over the key-sorted index list, batches of SORT_B points staged in
// shared memory and grouped by predicted work (spb_dev.cuh), evaluated by the lean mid-range routines
// LARGE: key 5 instead (|zeta| >= rl: the large-|zeta| routines; the length of their series falls with |zeta|)
template <bool BI, bool LARGE = false>
__global__ void __launch_bounds__(128, LARGE ? 4 : 6) k_airy_sorted(int id, int kode, int sem, const double2 *z,
                                                                   double2 *out, int *ierr, int *nz, long long n,
                                                                   Scratch s, unsigned long long *first_err,
                                                                   long long index_base) {
    __shared__ SortSmem sm;
    __shared__ unsigned int s_idx[SORT_B];
    __shared__ double2 s_z[SORT_B];
    // keys 1..4 (key 0 is evaluated by k_airy_small, key 5 by k_airy_large)
    const unsigned int start = s.bins[LARGE ? 5 : 1] - s.bins[0];
    const unsigned int count = LARGE ? s.bins[6] - s.bins[5] : s.bins[5] - s.bins[1];
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
                const double2 zz = __ldg(z + i);
                s_idx[slot] = i;
                s_z[slot] = zz;
                unsigned int bk = bucket_of(airy_work(cplx(zz.x, zz.y)), 0.2);
                if (LARGE && !BI) {
                    // large-|zeta| Ai: half of an all-quadrant batch continues K through the cut (K and I combined,
                    // an underflow test); the work estimate spans few buckets there, the sector takes the top bit
                    const double az = spb::cabs_(cplx(zz.x, zz.y));
                    bk = (bk > 15u ? 15u : bk) + (zz.x >= 0.5 * az ? 0u : 16u);
                }
                bucket[r] = bk;
            }
        }
        batch_sort(sm, bucket);
#pragma unroll 1
        for (int r = 0; r < SORT_R; ++r) {
            const unsigned int pos = threadIdx.x + 128 * r;
            if (pos < m) {
                const unsigned int slot = sm.order[pos];
                const double2 zz = s_z[slot];
                const cplx zc(zz.x, zz.y);
                const unsigned int i = s_idx[slot];
                // 1 < |z|, |zeta| < rl: the lean routines (amos/spb_airy_lean.h); the border of the region,
                // non-finite arguments and declined points go to the complete routine (k_airy_list)
                const double az = spb::cabs_(zc);
                const double azeta = (2.0 / 3.0) * az * sqrt(az);
                int ie = 0, nzz = 0;
                cplx val;
                bool done;
                if (LARGE) {
                    done = true;  // (the classifier's key: |zeta| >= rl with a margin, finite)
                    if (BI)
                        val = spb::biry_large(zc, id, kode, ie);
                    else
                        done = spb::airy_large(zc, id, kode, val, nzz, ie);
                } else {
                    done = az > 1.0 && azeta < SPB_RL;
                    if (done) done = BI ? spb::biry_mid(zc, id, kode, val, ie) : spb::airy_mid(zc, id, kode, val, nzz, ie);
                }
                if (!done) {
                    s.glist[atomicAdd(&s.bins[GLIST_LEN_SLOT], 1u)] = i;
                } else {
                    if (sem) val = spb::airy_semantics(zc, val, ie);
                    out[i] = make_double2(val.re, val.im);
                    if (ierr) ierr[i] = ie;
                    if (nz) nz[i] = nzz;
                    note_error(first_err, index_base + i, ie);
                }
            }
        }
        __syncthreads();
    }
}

// the points the lean kernels declined, through the complete routine
template <bool BI>
__global__ void __launch_bounds__(128) k_airy_list(int id, int kode, int sem, const double2 *z, double2 *out, int *ierr,
                                                   int *nz, Scratch s, unsigned long long *first_err,
                                                   long long index_base) {
    const unsigned int count = s.bins[GLIST_LEN_SLOT];
    const unsigned int stride = gridDim.x * blockDim.x;
#pragma unroll 1
    for (unsigned int t = blockIdx.x * blockDim.x + threadIdx.x; t < count; t += stride) {
        const unsigned int i = s.glist[t];
        const double2 zz = __ldg(z + i);
        airy_point<BI>(id, kode, sem, cplx(zz.x, zz.y), i, out, ierr, nz, first_err, index_base);
    }
}

// small batches: one thread per point in index order (no sort)
template <bool BI>
__global__ void __launch_bounds__(128) k_airy(int id, int kode, int sem, const double2 *z, double2 *out, int *ierr,
                                              int *nz, long long n, unsigned long long *first_err,
                                              long long index_base) {
    const long long stride = (long long)gridDim.x * blockDim.x;
#pragma unroll 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 t = __ldg(z + i);
        airy_point<BI>(id, kode, sem, cplx(t.x, t.y), i, out, ierr, nz, first_err, index_base);
    }
}

int launch_airy(int which, int kode, int sem, const double2 *z, double2 *out, int *ierr, int *nz, long long n,
                const Scratch *s, int sms, cudaStream_t st, unsigned long long *first_err, long long index_base) {
    const bool bi = which >= 2;
    const int id = which & 1;
    if (!s || n < 8192) {
        long long blocks = (n + 127) / 128;
        if (blocks > 148 * 32) blocks = 148 * 32;
        if (blocks < 1) blocks = 1;
        if (bi)
            k_airy<true><<<(int)blocks, 128, 0, st>>>(id, kode, sem, z, out, ierr, nz, n, first_err, index_base);
        else
            k_airy<false><<<(int)blocks, 128, 0, st>>>(id, kode, sem, z, out, ierr, nz, n, first_err, index_base);
        return 1;
    }
    k_classify_airy<<<s->nblocks, 256, 0, st>>>(which, kode, sem, z, out, ierr, nz, n, *s, first_err, index_base);
    DevJob job{};
    job.n = n;
    launch_scan_scatter(job, *s, st);
    if (bi) {
        k_airy_small<true><<<resident_grid(k_airy_small<true>, 256, sms), 256, 0, st>>>(id, kode, sem, z, out, ierr, nz, *s,
                                                                                       first_err, index_base);
        k_airy_sorted<true, true><<<resident_grid(k_airy_sorted<true, true>, 128, sms), 128, 0, st>>>(
            id, kode, sem, z, out, ierr, nz, n, *s, first_err, index_base);
        k_airy_sorted<true><<<resident_grid(k_airy_sorted<true>, 128, sms), 128, 0, st>>>(id, kode, sem, z, out, ierr, nz,
                                                                                         n, *s, first_err, index_base);
        k_airy_list<true><<<148, 128, 0, st>>>(id, kode, sem, z, out, ierr, nz, *s, first_err, index_base);
    } else {
        k_airy_small<false><<<resident_grid(k_airy_small<false>, 256, sms), 256, 0, st>>>(id, kode, sem, z, out, ierr, nz,
                                                                                         *s, first_err, index_base);
        k_airy_sorted<false, true><<<resident_grid(k_airy_sorted<false, true>, 128, sms), 128, 0, st>>>(
            id, kode, sem, z, out, ierr, nz, n, *s, first_err, index_base);
        k_airy_sorted<false><<<resident_grid(k_airy_sorted<false>, 128, sms), 128, 0, st>>>(id, kode, sem, z, out, ierr,
                                                                                           nz, n, *s, first_err, index_base);
        k_airy_list<false><<<148, 128, 0, st>>>(id, kode, sem, z, out, ierr, nz, *s, first_err, index_base);
    }
    return 7;
}

}  // namespace spbk
