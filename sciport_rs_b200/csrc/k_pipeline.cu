// k_pipeline.cu — region classification and warp-aggregated STABLE compaction.
//
//   k_classify : one thread per point; computes spb::prepare() (family rotation,
//                I region, K region, simple/general), writes a 1-byte class and a
//                per-tile histogram of the 8 bins.
//   k_scan     : exclusive scan of the [bin][tile] histogram (bin-major) so that
//                every (bin, tile) pair knows where its points start.
//   k_scatter  : re-reads the class bytes of a tile and writes each point's index
//                to its bin segment; the rank inside the tile comes from warp
//                ballots + a per-warp prefix in shared memory, so indices stay in
//                ascending order inside a bin (stable => region kernels read and
//                write z / out in ascending, mostly contiguous order).
#include "spb_dev.cuh"

namespace spbk {

__device__ __forceinline__ void point_bins(unsigned char c, int &bi, int &bk, bool &gen) {
    gen = (c >> 5) & 1;
    bi = c & 7;          // 0..3, 7 = none
    bk = (c >> 3) & 3;   // 0..2, 3 = none
}

__global__ void __launch_bounds__(256) k_classify(DevJob job, Scratch s) {
    __shared__ unsigned int hist[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    long long base = (long long)blockIdx.x * TILE;
#pragma unroll 1
    for (int k = threadIdx.x; k < TILE; k += blockDim.x) {
        long long i = base + k;
        if (i >= job.n) break;
        cplx z;
        double v;
        load_point(job, i, z, v);
        spb::Prep p = spb::prepare(job.fam, z, v, job.kode);
        unsigned char c;
        if (p.combine == spb::CMB_GENERAL) {
            c = (unsigned char)(7 | (3 << 3) | (1 << 5));
        } else {
            int bi = p.ireg >= 0 ? p.ireg : 7;
            int bk = p.kreg >= 0 ? p.kreg : 3;
            c = (unsigned char)(bi | (bk << 3));
        }
        s.cls[i] = c;
        // warp-aggregated histogram: one shared atomic per (warp, bin present) instead of one per thread
        const unsigned int act = __activemask();
        int bi, bk;
        bool gen;
        point_bins(c, bi, bk, gen);
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            bool in = (b < 4) ? (bi == b) : (b < 7) ? (bk == b - 4) : gen;
            unsigned int m = __ballot_sync(act, in);
            if (m != 0 && (threadIdx.x & 31) == (__ffs(act) - 1)) atomicAdd(&hist[b], (unsigned int)__popc(m));
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS) s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x] = hist[threadIdx.x];
}

// single-block exclusive scan over NBINS*nblocks counters (bin-major); also
// records where each bin starts inside its own perm array.
__global__ void __launch_bounds__(1024) k_scan(Scratch s) {
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int carry;
    __shared__ unsigned int bin_start[NBINS + 1];
    const int total = NBINS * s.nblocks;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < total; base += 1024) {
        int idx = base + threadIdx.x;
        unsigned int x = idx < total ? s.blockcnt[idx] : 0u;
        unsigned int incl = x;
        int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned int y = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += y;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned int w = warp_tot[lane];
            unsigned int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned int y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
        }
        __syncthreads();
        unsigned int excl = carry + warp_tot[wid] + incl - x;
        if (idx < total) {
            s.blockcnt[idx] = excl;
            if (idx % s.nblocks == 0) bin_start[idx / s.nblocks] = excl;
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bin_start[NBINS] = carry;
        // absolute starts in the concatenated order: I bins 0..3 | K bins 4..6 | general 7
        for (int b = 0; b <= NBINS; ++b) s.bins[b] = bin_start[b];
        s.bins[GLIST_LEN_SLOT] = bin_start[NBINS] - bin_start[BIN_GENERAL];
    }
}

__global__ void __launch_bounds__(256) k_scatter(DevJob job, Scratch s) {
    // ranks[b]: running count of bin b inside this tile
    __shared__ unsigned int running[NBINS];
    __shared__ unsigned int warp_cnt[8][NBINS];
    if (threadIdx.x < NBINS) running[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * TILE;
    const unsigned int i_base = s.bins[0];              // = 0
    const unsigned int k_base = s.bins[4];              // start of K bins in scan order
    const unsigned int g_base = s.bins[BIN_GENERAL];    // start of the general bin in scan order
#pragma unroll 1
    for (int k0 = 0; k0 < TILE; k0 += 256) {
        long long i = base + k0 + threadIdx.x;
        bool valid = i < job.n;
        int bi = 7, bk = 3;
        bool gen = false;
        if (valid) point_bins(s.cls[i], bi, bk, gen);
        // per-warp rank for every bin via ballots
        unsigned int my_rank_i = 0, my_rank_k = 0, my_rank_g = 0;
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            bool in = valid && ((b < 4) ? (bi == b) : (b < 7) ? (bk == b - 4) : gen);
            unsigned int m = __ballot_sync(0xffffffffu, in);
            unsigned int r = __popc(m & ((1u << lane) - 1u));
            if (in) {
                if (b < 4) my_rank_i = r;
                else if (b < 7) my_rank_k = r;
                else my_rank_g = r;
            }
            if (lane == 0) warp_cnt[wid][b] = __popc(m);
        }
        __syncthreads();
        // exclusive prefix over the 8 warps (tiny: done redundantly by every thread that needs it)
        if (valid) {
            if (bi != 7) {
                unsigned int off = running[bi];
                for (int w = 0; w < wid; ++w) off += warp_cnt[w][bi];
                unsigned int dst = s.blockcnt[(size_t)bi * s.nblocks + blockIdx.x] - i_base + off + my_rank_i;
                s.perm_i[dst] = (unsigned int)i;
            }
            if (bk != 3) {
                int b = 4 + bk;
                unsigned int off = running[b];
                for (int w = 0; w < wid; ++w) off += warp_cnt[w][b];
                unsigned int dst = s.blockcnt[(size_t)b * s.nblocks + blockIdx.x] - k_base + off + my_rank_k;
                s.perm_k[dst] = (unsigned int)i;
            }
            if (gen) {
                unsigned int off = running[BIN_GENERAL];
                for (int w = 0; w < wid; ++w) off += warp_cnt[w][BIN_GENERAL];
                unsigned int dst =
                    s.blockcnt[(size_t)BIN_GENERAL * s.nblocks + blockIdx.x] - g_base + off + my_rank_g;
                s.glist[dst] = (unsigned int)i;
            }
        }
        __syncthreads();
        if (threadIdx.x < NBINS) {
            unsigned int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_cnt[w][threadIdx.x];
            running[threadIdx.x] += t;
        }
        __syncthreads();
    }
}

void launch_classify(const DevJob &job, const Scratch &s, cudaStream_t st) {
    k_classify<<<s.nblocks, 256, 0, st>>>(job, s);
}

void launch_scan_scatter(const DevJob &job, const Scratch &s, cudaStream_t st) {
    k_scan<<<1, 1024, 0, st>>>(s);
    k_scatter<<<s.nblocks, 256, 0, st>>>(job, s);
}

}  // namespace spbk
