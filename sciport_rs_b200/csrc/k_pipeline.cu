// k_pipeline.cu — region classification and STABLE compaction by algorithm region.
//
//   k_classify : one thread per point (coalesced 16-byte loads); computes
//                spb::prepare() (family rotation, I region, K region,
//                simple/general), writes a 1-byte class and a per-tile histogram
//                of the 8 bins.  Counts are kept as eight 16-bit fields packed in
//                two 64-bit registers and reduced with warp shuffles, so the
//                histogram costs 10 shuffles + 16 shared atomics per warp.
//   k_scan     : exclusive scan of the [bin][tile] histogram (bin-major) so that
//                every (bin, tile) pair knows where its points start.
//   k_scatter  : each thread re-reads the class bytes of 8 CONSECUTIVE points with
//                one 8-byte load, a block-wide exclusive scan of the packed counts
//                gives every thread its start offset in each bin, and the point
//                indices are written in ascending order (stable => region kernels
//                read z and write results in ascending, mostly contiguous order).
#include "spb_dev.cuh"

namespace spbk {

__device__ __forceinline__ void point_bins(unsigned int c, int &bi, int &bk, bool &gen) {
    gen = (c >> 5) & 1;
    bi = c & 7;         // 0..3, 7 = none
    bk = (c >> 3) & 3;  // 0..2, 3 = none
}

// packed 16-bit counters: lo holds bins 0..3 (I regions), hi holds bins 4..7 (K regions, general)
__device__ __forceinline__ void packed_add(unsigned int c, unsigned long long &lo, unsigned long long &hi) {
    int bi, bk;
    bool gen;
    point_bins(c, bi, bk, gen);
    if (bi != 7) lo += 1ull << (16 * bi);
    if (bk != 3) hi += 1ull << (16 * bk);
    if (gen) hi += 1ull << 48;
}

__global__ void __launch_bounds__(256) k_classify(DevJob job, Scratch s) {
    __shared__ unsigned int hist[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * TILE;
    unsigned long long lo = 0, hi = 0;
#pragma unroll 2
    for (int k = threadIdx.x; k < TILE; k += 256) {
        long long i = base + k;
        if (i < job.n) {
            cplx z;
            double v;
            load_point(job, i, z, v);
            spb::Prep p = spb::prepare(job.fam, z, v, job.kode);
            unsigned int c;
            if (p.combine == spb::CMB_GENERAL) {
                c = 7u | (3u << 3) | (1u << 5);
            } else {
                unsigned int bi = p.ireg >= 0 ? (unsigned int)p.ireg : 7u;
                unsigned int bk = p.kreg >= 0 ? (unsigned int)p.kreg : 3u;
                c = bi | (bk << 3);
            }
            s.cls[i] = (unsigned char)c;
            packed_add(c, lo, hi);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        lo += __shfl_xor_sync(0xffffffffu, lo, d);
        hi += __shfl_xor_sync(0xffffffffu, hi, d);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            unsigned int a = (unsigned int)((lo >> (16 * b)) & 0xffffu), h = (unsigned int)((hi >> (16 * b)) & 0xffffu);
            if (a) atomicAdd(&hist[b], a);
            if (h) atomicAdd(&hist[4 + b], h);
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS) s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Exclusive scan of the [bin][tile] histogram: one CTA per bin scans that bin's tile counters (8 consecutive
// counters per thread, coalesced), so blockcnt[b][t] becomes the offset of tile t inside bin b.  The last CTA to
// finish turns the bin totals into the bin start offsets (bins[0..8]) and the general-list length.
__global__ void __launch_bounds__(1024) k_scan(Scratch s) {
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int carry_s;
    const int b = blockIdx.x;
    unsigned int *cnt = s.blockcnt + (size_t)b * s.nblocks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < s.nblocks; base += 8192) {
        const int i0 = base + (int)threadIdx.x * 8;
        unsigned int x[8];
        unsigned int sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            x[k] = (i0 + k < s.nblocks) ? cnt[i0 + k] : 0u;
            sum += x[k];
        }
        unsigned int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        const unsigned int carry = carry_s;
        if (wid == 0) {
            unsigned int w = warp_tot[lane];
            unsigned int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned int y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
            if (lane == 31) carry_s = carry + wi;
        }
        __syncthreads();
        unsigned int running = carry + warp_tot[wid] + incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k < s.nblocks) cnt[i0 + k] = running;
            running += x[k];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s.bins[18 + b] = carry_s;  // total of this bin
        __threadfence();
        unsigned int done = atomicAdd(&s.bins[17], 1u);
        if (done == NBINS - 1) {
            // absolute starts in the concatenated order: I bins 0..3 | K bins 4..6 | general 7
            unsigned int acc = 0;
            for (int k = 0; k < NBINS; ++k) {
                s.bins[k] = acc;
                acc += ((volatile unsigned int *)s.bins)[18 + k];
            }
            s.bins[NBINS] = acc;
            s.bins[GLIST_LEN_SLOT] = ((volatile unsigned int *)s.bins)[18 + BIN_GENERAL];
            s.bins[17] = 0;  // re-arm for the next chunk
        }
    }
}

__global__ void __launch_bounds__(256) k_scatter(DevJob job, Scratch s) {
    __shared__ unsigned long long warp_lo[8], warp_hi[8];
    __shared__ unsigned int tile_base[NBINS];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * TILE + (long long)threadIdx.x * 8;
    if (threadIdx.x < NBINS) {
        // where this tile's points of bin b start inside the bin's perm array
        unsigned int seg = threadIdx.x < 4 ? s.bins[0] : threadIdx.x < 7 ? s.bins[4] : s.bins[BIN_GENERAL];
        tile_base[threadIdx.x] =
            s.bins[threadIdx.x] - seg + s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x];
    }
    // class bytes of 8 consecutive points in one load (tiles are 2048-aligned, so 8-byte aligned)
    unsigned long long cls8 = 0;
    int valid = 0;
    if (base + 8 <= job.n) {
        cls8 = *reinterpret_cast<const unsigned long long *>(s.cls + base);
        valid = 8;
    } else if (base < job.n) {
        valid = (int)(job.n - base);
        for (int k = 0; k < valid; ++k) cls8 |= (unsigned long long)s.cls[base + k] << (8 * k);
    }
    unsigned long long lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < valid) packed_add((unsigned int)((cls8 >> (8 * k)) & 0xffu), lo, hi);
    // block-wide exclusive scan of the packed counters (fields never exceed 2048)
    unsigned long long ilo = lo, ihi = hi;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long a = __shfl_up_sync(0xffffffffu, ilo, d);
        unsigned long long b = __shfl_up_sync(0xffffffffu, ihi, d);
        if (lane >= d) {
            ilo += a;
            ihi += b;
        }
    }
    if (lane == 31) {
        warp_lo[wid] = ilo;
        warp_hi[wid] = ihi;
    }
    __syncthreads();
    unsigned long long olo = ilo - lo, ohi = ihi - hi;
    for (int w = 0; w < wid; ++w) {
        olo += warp_lo[w];
        ohi += warp_hi[w];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= valid) break;
        int bi, bk;
        bool gen;
        point_bins((unsigned int)((cls8 >> (8 * k)) & 0xffu), bi, bk, gen);
        const unsigned int idx = (unsigned int)(base + k);
        if (bi != 7) {
            unsigned int off = (unsigned int)((olo >> (16 * bi)) & 0xffffu);
            s.perm_i[tile_base[bi] + off] = idx;
            olo += 1ull << (16 * bi);
        }
        if (bk != 3) {
            unsigned int off = (unsigned int)((ohi >> (16 * bk)) & 0xffffu);
            s.perm_k[tile_base[4 + bk] + off] = idx;
            ohi += 1ull << (16 * bk);
        }
        if (gen) {
            unsigned int off = (unsigned int)((ohi >> 48) & 0xffffu);
            s.glist[tile_base[BIN_GENERAL] + off] = idx;
            ohi += 1ull << 48;
        }
    }
}

void launch_classify(const DevJob &job, const Scratch &s, cudaStream_t st) {
    k_classify<<<s.nblocks, 256, 0, st>>>(job, s);
}

void launch_scan_scatter(const DevJob &job, const Scratch &s, cudaStream_t st) {
    k_scan<<<NBINS, 1024, 0, st>>>(s);
    k_scatter<<<s.nblocks, 256, 0, st>>>(job, s);
}

}  // namespace spbk
