// k_pipeline.cu — region classification and STABLE compaction by algorithm region.
//
//   k_classify : one thread per point (coalesced 16-byte loads); computes
//                spb::prepare() (family rotation, I region, K region,
//                simple/general), writes a 1-byte class and a per-tile histogram
//                of the NBINS bins.  Counts are kept as 12-bit fields packed five
//                to a 64-bit register and reduced with warp shuffles, so the
//                histogram costs 5 shuffles per word + <= NBINS shared atomics per warp.
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
    bi = c & 7;         // 0..NB_I-1, 7 = none
    bk = (c >> 3) & 3;  // 0..NB_K-1, 3 = none
}

// packed counters, one 12-bit field per bin (a tile has 2048 points; a count of exactly 2048 still fits):
// field f lives in word f / 5 at bit 12 * (f % 5)
struct Packed {
    unsigned long long w[NWORDS];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) w[k] = 0;
    }
    __device__ __forceinline__ void bump(int f) {
        const int word = f / FIELDS_PER_WORD, sh = FIELD_BITS * (f % FIELDS_PER_WORD);
#pragma unroll
        for (int k = 0; k < NWORDS; ++k)
            if (word == k) w[k] += 1ull << sh;
    }
    __device__ __forceinline__ unsigned int field(int f) const {
        const int word = f / FIELDS_PER_WORD, sh = FIELD_BITS * (f % FIELDS_PER_WORD);
        unsigned long long x = 0;
#pragma unroll
        for (int k = 0; k < NWORDS; ++k)
            if (word == k) x = w[k];
        return (unsigned int)((x >> sh) & ((1u << FIELD_BITS) - 1u));
    }
};

// The three keys of a class byte land in fixed places of the three words (NB_I = 7, five fields per word):
// I bins 0..4 -> word 0, field bi; I bins 5, 6 -> word 1, fields 0, 1; K bins -> word 1, fields 2..4; general ->
// word 2, field 0 (= field number bin of the generic Packed::field()).  Written out so that the only
// variable-dependent work is one multiply and one shift per key.
static_assert(NB_I == 7 && NB_K == 3 && FIELDS_PER_WORD == 5 && NWORDS == 3, "packed layout below assumes 7 + 3 + 1 bins");

__device__ __forceinline__ void packed_add(unsigned int c, Packed &p) {
    int bi, bk;
    bool gen;
    point_bins(c, bi, bk, gen);
    if (bi < 5)
        p.w[0] += 1ull << (FIELD_BITS * bi);
    else if (bi != 7)
        p.w[1] += 1ull << (FIELD_BITS * (bi - 5));
    if (bk != 3) p.w[1] += 1ull << (FIELD_BITS * (bk + 2));
    if (gen) p.w[2] += 1ull;
}

// Compaction only: an entry of a stable bin (large-|w|: bin 1, Debye-direct: bin 6) that carries the side bit is
// counted in a field of its own (word 2, fields 1 and 2), so that inside a tile the bin's entries come out as
// [side 0 in index order | side 1 in index order].  The bin totals of a tile are unchanged.
__device__ __forceinline__ int side_field(int bi) { return bi == 1 ? 1 : 2; }
__device__ __forceinline__ void packed_add_sided(unsigned int c, Packed &p) {
    if (!(c & 0x40u)) {
        packed_add(c, p);
        return;
    }
    int bi, bk;
    bool gen;
    point_bins(c, bi, bk, gen);
    p.w[2] += 1ull << (FIELD_BITS * side_field(bi));
    if (bk != 3) p.w[1] += 1ull << (FIELD_BITS * (bk + 2));
    if (gen) p.w[2] += 1ull;
}
__device__ __forceinline__ unsigned int field_side(const Packed &p, int bi) {
    return (unsigned int)(p.w[2] >> (FIELD_BITS * side_field(bi))) & ((1u << FIELD_BITS) - 1u);
}

__device__ __forceinline__ unsigned int field_i(const Packed &p, int bi) {
    unsigned long long x = bi < 5 ? (p.w[0] >> (FIELD_BITS * bi)) : (p.w[1] >> (FIELD_BITS * (bi - 5)));
    return (unsigned int)x & ((1u << FIELD_BITS) - 1u);
}
__device__ __forceinline__ unsigned int field_k(const Packed &p, int bk) {
    return (unsigned int)(p.w[1] >> (FIELD_BITS * (bk + 2))) & ((1u << FIELD_BITS) - 1u);
}
__device__ __forceinline__ unsigned int field_g(const Packed &p) {
    return (unsigned int)p.w[2] & ((1u << FIELD_BITS) - 1u);
}

// The increments of the three packed words for every class byte, as a shared-memory table: the if / else form above
// ran with 15 of 32 lanes and was 45 % of the instructions of k_scatter (the bytes of a warp's points differ; in the
// classifier, a tenth of the instructions, the table measured no gain); a branch-free arithmetic form paid as much in 64-bit variable shifts.  One table per
// CTA, filled by its first 128 threads from the very functions above; a point then costs three 8-byte shared loads
// (a handful of distinct addresses per warp) and three adds.  sided: bit 6 counts in the side fields (compaction).
__device__ __forceinline__ void packed_table_init(unsigned long long (*tab)[NWORDS], bool sided) {
    for (unsigned int c = threadIdx.x; c < 128u; c += blockDim.x) {
        Packed p;
        p.clear();
        if (sided)
            packed_add_sided(c, p);
        else
            packed_add(c, p);
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) tab[c][k] = p.w[k];
    }
}
__device__ __forceinline__ void packed_add_tab(const unsigned long long (*tab)[NWORDS], unsigned int c, Packed &p) {
#pragma unroll
    for (int k = 0; k < NWORDS; ++k) p.w[k] += tab[c & 0x7fu][k];
}

// (side, |z|, order) cell of a point for the second-level sort (see k_cell_hist below): 2^CELL_ZBITS levels of |z|
// over [1, 256] (8 per octave with 6 bits, 4 with 5), 2^CELL_VBITS levels of the order over [0, 64), 2 sides
__device__ __forceinline__ unsigned int cell_of(int fam, cplx z, double v, bool with_side) {
    const double m2 = z.re * z.re + z.im * z.im;
    const int hi = __double2hiint(m2);
    // levels of |z|^2 per octave = half of the levels of |z| per octave: the exponent and the top mantissa bits
    constexpr int MB = CELL_ZBITS - 4;  // mantissa bits of |z|^2 used: 16 octaves of |z|^2 fill 4 bits
    int q = (((hi >> 20) - 1023) << MB) + ((hi >> (20 - MB)) & ((1 << MB) - 1));
    q = q < 0 ? 0 : (q > (1 << CELL_ZBITS) - 1 ? (1 << CELL_ZBITS) - 1 : q);
    int b = (int)(fabs(v) * (double)(1 << CELL_VBITS) * (1.0 / 64.0));
    b = b > (1 << CELL_VBITS) - 1 ? (1 << CELL_VBITS) - 1 : b;
    return (unsigned int)(q | (b << CELL_ZBITS)) | (with_side ? (side_of(fam, z) << (CELL_ZBITS + CELL_VBITS)) : 0u);
}

// Direct jobs (k_direct.cu ran first): is every group of 32 points of this tile already finished?  (one flag byte
// per group, 64 per tile; the barrier makes the answer CTA-uniform)
constexpr int GROUPS = TILE / 32;
__device__ __forceinline__ bool tile_done(const Scratch &s) {
    const bool mine = threadIdx.x < GROUPS ? s.done[(size_t)blockIdx.x * GROUPS + threadIdx.x] != 0 : true;
    return __syncthreads_and(mine) != 0;
}

// FAM / KODE: the job's family and scaling flag as compile-time constants (dispatch_fam_kode)
template <int FAM, int KODE>
__global__ void __launch_bounds__(256, 4) k_classify(DevJob job, Scratch s) {
    pdl_enter();
    job.fam = FAM;
    job.kode = KODE;
    // (k_direct finished nothing -- a batch of unrelated points: classify as if it had not run)
    if (job.direct && s.bins[SLOT_DIRECT] == 0) job.direct = 0;
    if (job.direct && tile_done(s)) {
        // nothing of this tile enters a bin (k_scatter skips the tile on the same test, so no class bytes either)
        if (threadIdx.x < NBINS) s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x] = 0;
        return;
    }
    __shared__ unsigned int hist[NBINS];
    if (threadIdx.x < NBINS) hist[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * TILE;
    Packed cnt;
    cnt.clear();
    constexpr int PER_THREAD = TILE / 256;
    // Phase 1: the fast accept (|z|^2 and a few comparisons) for the thread's 8 points at once: all loads are
    // issued before the first test, so the phase runs at memory speed on batches that live in the large-|z|
    // region.  A warp keeps a fast result only if all its lanes accepted that point.
    unsigned int cls4[PER_THREAD / 4];  // class bytes, 0xff = not decided yet
#pragma unroll
    for (int q = 0; q < PER_THREAD / 4; ++q) cls4[q] = 0xffffffffu;
    // (per-element orders: a warp of 32 unrelated (order, argument) pairs practically never sits in the large-|z|
    // region as a whole, so the fast accept would only add its cost to every point)
    if (job.direct) {
        // the fast accept has been tried by k_direct: its finished groups enter no bin (class byte of a decided
        // point), everything else goes through prepare().  Thread t holds points t + 256 k: group (t / 32) + 8 k.
        static_assert(GROUPS == 8 * PER_THREAD, "one group per warp and round");
#pragma unroll
        for (int q = 0; q < PER_THREAD / 4; ++q) cls4[q] = 0;
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const bool dn = s.done[(size_t)blockIdx.x * GROUPS + (threadIdx.x >> 5) + 8 * k] != 0;
            cls4[k / 4] |= (dn ? (7u | (3u << 3)) : 0xffu) << (8 * (k % 4));
        }
    } else if (job.v_stride == 0) {
        double2 zk[PER_THREAD];
        double vk[PER_THREAD];
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            const long long i = base + threadIdx.x + 256 * k;
            zk[k] = make_double2(0.0, 0.0);
            vk[k] = 0.0;
            if (i < job.n) {
                cplx z;
                load_point(job, i, z, vk[k]);
                zk[k] = make_double2(z.re, z.im);
            }
        }
#pragma unroll
        for (int q = 0; q < PER_THREAD / 4; ++q) cls4[q] = 0;
        const spb::FastOrder fo = spb::fast_order(vk[0]);  // a broadcast order: its half of the test once per thread
        // A warp whose first point is not accepted as a whole (a batch of unrelated points, not a grid) skips the
        // fast test for its other seven points: on such batches the test practically never succeeds for 32 lanes
        // at once and would only add its cost to every point (config 4: a quarter of the classifier's time).
        bool try_fast = true;
#pragma unroll
        for (int k = 0; k < PER_THREAD; ++k) {
            unsigned int c = 0xffu;
            if (try_fast) {
                int ireg, kreg;
                const bool ok = spb::fast_classify(job.fam, cplx(zk[k].x, zk[k].y), fo, job.kode, job.fuse, ireg, kreg);
                const bool all = __all_sync(0xffffffffu, ok);
                if (all) c = spb::class_of(ireg, kreg);
                if (k == 0 && !all) try_fast = false;
            }
            cls4[k / 4] |= c << (8 * (k % 4));
        }
    }
    // Phase 2: prepare() for the points the fast accept left open (warp-uniform per point), then store and count
#pragma unroll 2
    for (int k = 0; k < PER_THREAD; ++k) {
        const long long i = base + threadIdx.x + 256 * k;
        unsigned int c = (cls4[k / 4] >> (8 * (k % 4))) & 0xffu;
        if (i < job.n) {
            if (c == 0xffu) {
                cplx z;
                double v;
                load_point(job, i, z, v);
                const spb::Prep p = spb::prepare(job.fam, z, v, job.kode, job.fuse);
                c = spb::class_of_prep(p);
                if (p.combine == spb::CMB_FINAL) {
                    // decided here (a provable underflow of K / overflow of its continuation, unscaled K, H, Y far from
                    // the real axis): stored at once, the point enters no bin and is never visited again
                    cplx val(0.0, 0.0);
                    if (job.sem && p.final_ierr == 2) {
                        // scipy conventions for an overflow: +inf for K on the positive real axis, NaN otherwise
                        const double inf = __longlong_as_double(0x7ff0000000000000ll);
                        const bool pos_real = z.re >= 0.0 && z.im == 0.0;
                        val = (job.fam == spb::FAM_K && pos_real) ? cplx(inf, 0.0) : cplx(inf - inf, inf - inf);
                    }
                    store_point(job, i, val, p.final_ierr, p.final_nz);
                }
                if (job.cellsort) s.cellkey[i] = (unsigned short)cell_of(FAM, z, v, (job.cellsort & 2) != 0);
                // stable bins: ordering hint for the compaction (the histogram below ignores it)
                if (job.group && ((c & 7u) == (unsigned int)spb::IREG_ASYI || (c & 7u) == (unsigned int)spb::IREG_DEBYE))
                    c |= side_of(FAM, z) << 6;
            } else if (job.cellsort) {
                cplx z;
                double v;
                load_point(job, i, z, v);
                s.cellkey[i] = (unsigned short)cell_of(FAM, z, v, (job.cellsort & 2) != 0);
            }
            s.cls[i] = (unsigned char)c;
            packed_add(c, cnt);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) cnt.w[k] += __shfl_xor_sync(0xffffffffu, cnt.w[k], d);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int b = 0; b < NBINS; ++b) {
            unsigned int a = cnt.field(b);
            if (a) atomicAdd(&hist[b], a);
        }
    }
    __syncthreads();
    if (threadIdx.x < NBINS) s.blockcnt[(size_t)threadIdx.x * s.nblocks + blockIdx.x] = hist[threadIdx.x];
}

// Exclusive scan of the [bin][tile] histogram: one CTA per bin scans that bin's tile counters (8 consecutive
// counters per thread, coalesced), so blockcnt[b][t] becomes the offset of tile t inside bin b.  The last CTA to
// finish turns the bin totals into the bin start offsets (bins[0..8]) and the general-list length.
__global__ void __launch_bounds__(1024) k_scan(Scratch s) {
    pdl_enter();
    __shared__ unsigned int warp_tot[32];
    __shared__ unsigned int round_total;
    const int b = blockIdx.x;
    unsigned int *cnt = s.blockcnt + (size_t)b * s.nblocks;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // The running total of the rounds already scanned lives in a REGISTER that every thread updates identically
    // from round_total after the second barrier of a round; the trailing barrier keeps round_total (and warp_tot)
    // of this round intact until every thread has read them.  (No thread ever reads a shared word that another
    // thread may be overwriting between the same pair of barriers.)
    unsigned int carry = 0;
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
        if (wid == 0) {
            unsigned int w = warp_tot[lane];
            unsigned int wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned int y = __shfl_up_sync(0xffffffffu, wi, d);
                if (lane >= d) wi += y;
            }
            warp_tot[lane] = wi - w;  // exclusive prefix of warp totals
            if (lane == 31) round_total = wi;
        }
        __syncthreads();
        unsigned int running = carry + warp_tot[wid] + incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (i0 + k < s.nblocks) cnt[i0 + k] = running;
            running += x[k];
        }
        carry += round_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        s.bins[SLOT_TOTALS + b] = carry;  // total of this bin
        __threadfence();
        unsigned int done = atomicAdd(&s.bins[SLOT_DONE], 1u);
        if (done == NBINS - 1) {
            // absolute starts in the concatenated order: I bins | K bins | general
            unsigned int acc = 0;
            for (int k = 0; k < NBINS; ++k) {
                s.bins[k] = acc;
                acc += ((volatile unsigned int *)s.bins)[SLOT_TOTALS + k];
            }
            s.bins[NBINS] = acc;
            s.bins[GLIST_LEN_SLOT] = ((volatile unsigned int *)s.bins)[SLOT_TOTALS + BIN_GENERAL];
            s.bins[SLOT_DONE] = 0;  // re-arm for the next chunk
            s.bins[SLOT_G2] = 0;
            s.bins[SLOT_DIRECT_LAST] = ((volatile unsigned int *)s.bins)[SLOT_DIRECT];
            s.bins[SLOT_DIRECT] = 0;
        }
    }
}

__global__ void __launch_bounds__(256) k_scatter(DevJob job, Scratch s) {
    pdl_enter();
    // counters of the second-level sort of this chunk (k_cell_hist runs after this kernel, in stream order)
    if (blockIdx.x == 0)
        for (int c = threadIdx.x; c < CELL_SLOTS * NCELLS; c += 256) s.cellcnt[c] = 0;
    // (k_classify wrote no class bytes for such a tile; k_scan has moved the direct counter on by now)
    if (job.direct && s.bins[SLOT_DIRECT_LAST] != 0 && tile_done(s)) return;
    __shared__ unsigned long long warp_tot[8][NWORDS];
    __shared__ unsigned int tile_base[NBINS], lstart[NBINS], seg_total[3];
    __shared__ unsigned int stage[3][TILE];
    __shared__ unsigned long long ptab[128][NWORDS];
    packed_table_init(ptab, true);  // (published by the barrier of the homogeneous-tile test below)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long base = (long long)blockIdx.x * TILE + (long long)threadIdx.x * 8;
    if (threadIdx.x < NBINS) {
        // where this tile's points of bin b start inside the bin's index array (perm_i / perm_k / glist)
        const int b = threadIdx.x;
        unsigned int seg = b < BIN_K0 ? s.bins[0] : b < BIN_GENERAL ? s.bins[BIN_K0] : s.bins[BIN_GENERAL];
        tile_base[b] = s.bins[b] - seg + s.blockcnt[(size_t)b * s.nblocks + blockIdx.x];
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
    // Homogeneous tile (every point in the same class: the rule on gridded inputs, where 2048 consecutive points
    // are neighbours): the sorted order is the index order, so the indices are written straight out.
    {
        const long long tile0 = (long long)blockIdx.x * TILE;
        const unsigned int c0 = s.cls[tile0];
        const unsigned long long rep = 0x0101010101010101ull * c0;
        const bool same = valid == 8 ? (cls8 == rep) : (((cls8 ^ rep) & ((1ull << (8 * valid)) - 1ull)) == 0ull);
        if (__syncthreads_and(same)) {  // (also orders the tile_base[] writes before their use)
            int bi, bk;
            bool gen;
            point_bins(c0, bi, bk, gen);
            const unsigned int m = (unsigned int)min((long long)TILE, job.n - tile0);
            for (unsigned int j = threadIdx.x; j < m; j += 256) {
                const unsigned int idx = (unsigned int)(tile0 + j);
                if (bi != 7) s.perm_i[tile_base[bi] + j] = idx;
                if (bk != 3) s.perm_k[tile_base[BIN_K0 + bk] + j] = idx;
                if (gen) s.glist[tile_base[BIN_GENERAL] + j] = idx;
            }
            return;
        }
    }
    Packed mine;
    mine.clear();
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k < valid) packed_add_tab(ptab, (unsigned int)((cls8 >> (8 * k)) & 0xffu), mine);
    // block-wide exclusive scan of the packed counters (fields never exceed 2048)
    Packed incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) {
            unsigned long long a = __shfl_up_sync(0xffffffffu, incl.w[k], d);
            if (lane >= d) incl.w[k] += a;
        }
    }
    if (lane == 31) {
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) warp_tot[wid][k] = incl.w[k];
    }
    __syncthreads();
    Packed off, total;
#pragma unroll
    for (int k = 0; k < NWORDS; ++k) {
        off.w[k] = incl.w[k] - mine.w[k];
        total.w[k] = 0;
    }
    for (int w = 0; w < 8; ++w) {
#pragma unroll
        for (int k = 0; k < NWORDS; ++k) {
            if (w < wid) off.w[k] += warp_tot[w][k];
            total.w[k] += warp_tot[w][k];
        }
    }
    // The indices are first laid out in shared memory in their final (bin-major, stable) order and then
    // written out by consecutive threads: a thread's 8 consecutive points would otherwise produce stride-8
    // 4-byte stores (eight L2 write requests per warp instruction instead of one).
    if (threadIdx.x < NBINS) {
        const int b = threadIdx.x;
        const int first = b < BIN_K0 ? 0 : b < BIN_GENERAL ? BIN_K0 : BIN_GENERAL;
        unsigned int st = 0;
        for (int k = first; k < b; ++k) st += total.field(k);
        if (b < BIN_K0) {  // side-1 entries of the stable bins in front of bin b
            if (b > 1) st += field_side(total, 1);
            if (b > 6) st += field_side(total, 6);
        }
        lstart[b] = st;
    }
    if (threadIdx.x == 0) {
        unsigned int ti = field_side(total, 1) + field_side(total, 6), tk = 0;
        for (int k = 0; k < BIN_K0; ++k) ti += total.field(k);
        for (int k = BIN_K0; k < BIN_GENERAL; ++k) tk += total.field(k);
        seg_total[0] = ti;
        seg_total[1] = tk;
        seg_total[2] = total.field(BIN_GENERAL);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= valid) break;
        int bi, bk;
        bool gen;
        point_bins((unsigned int)((cls8 >> (8 * k)) & 0xffu), bi, bk, gen);
        const unsigned int idx = (unsigned int)(base + k);
        const unsigned int ck = (unsigned int)((cls8 >> (8 * k)) & 0xffu);
        if (bi != 7)
            stage[0][lstart[bi] + ((ck & 0x40u) ? field_i(total, bi) + field_side(off, bi) : field_i(off, bi))] = idx;
        if (bk != 3) stage[1][lstart[BIN_K0 + bk] + field_k(off, bk)] = idx;
        if (gen) stage[2][field_g(off)] = idx;
        packed_add_tab(ptab, ck, off);
    }
    __syncthreads();
    // coalesced write-out, one segment at a time; position j of a segment belongs to the last bin whose local
    // start is <= j
    for (unsigned int j = threadIdx.x; j < seg_total[0]; j += 256) {
        int b = 0;
#pragma unroll
        for (int k = 1; k < NB_I; ++k)
            if (j >= lstart[k]) b = k;
        s.perm_i[tile_base[b] + (j - lstart[b])] = stage[0][j];
    }
    for (unsigned int j = threadIdx.x; j < seg_total[1]; j += 256) {
        int b = BIN_K0;
#pragma unroll
        for (int k = BIN_K0 + 1; k < BIN_GENERAL; ++k)
            if (j >= lstart[k]) b = k;
        s.perm_k[tile_base[b] + (j - lstart[b])] = stage[1][j];
    }
    for (unsigned int j = threadIdx.x; j < seg_total[2]; j += 256) s.glist[tile_base[BIN_GENERAL] + j] = stage[2][j];
}

// ---- second-level sort inside the iterative bins: by (|z|, order) cell ------------------------------------------
// The region routines of these bins are loops whose trip counts are smooth functions of |z| and the order (series
// length, Miller start index, forward recurrence over the order, ratio walk).  Inside one region bin the stable
// compaction leaves the points in index order, i.e. in random (|z|, order) order for unrelated inputs: the lanes of
// a warp then run loops of very different lengths (measured: 20-21 of 32 lanes active in the Miller / Wronskian /
// K kernels of config 3 even with a per-CTA batch sort of 512 points).  A counting sort of each bin segment by a
// (CELL_ZBITS + CELL_VBITS + 1)-bit cell key -- 32 levels of |z|, four per octave, times 16 levels of the order, times
// the side of the family's continuation (side_of); 64 x 32 x 2 cells measured slower: the sort pays more than the
// kernels gain (config 3 pair 9.67 -> 10.27 ms per 2^26 points) -- makes 32 consecutive
// entries near neighbours in both variables, over the whole chunk instead of a 512-point batch:
// The classifier of a cell-sorted job stores the key of every point (2 bytes, coalesced) next to its class byte.
//   k_cell_hist    histogram of the cell keys per bin (shared-memory histograms, one global add per cell and CTA)
//   k_cell_scan    exclusive scan of the NCELLS cell counters of every bin -> cell cursors
//   k_cell_scatter batches of 1024 entries: shared-memory counts, ONE global atomic per (batch, cell) reserves the
//                  range, entries are written to perm_i2 / perm_k2 (order inside a cell is arbitrary: results are
//                  per point, only the memory order depends on it)
// Bin boundaries do not move, so the region kernels only switch the index array they walk.
// slot -> (bin, index array): I bins 0..5 (series, asyi, mlri, wrsk, uniform forms 1 and 2), then the three K bins
__device__ __forceinline__ void cell_slot(const Scratch &s, int slot, const unsigned int *&src, unsigned int *&dst,
                                          unsigned int &count) {
    const int bin = slot < 6 ? slot : BIN_K0 + (slot - 6);
    const bool kbin = slot >= 6;
    const unsigned int start = s.bins[bin] - (kbin ? s.bins[BIN_K0] : s.bins[0]);
    count = s.bins[bin + 1] - s.bins[bin];
    src = (kbin ? s.perm_k : s.perm_i) + start;
    dst = (kbin ? s.perm_k2 : s.perm_i2) + start;
}

__global__ void __launch_bounds__(256) k_cell_hist(DevJob job, Scratch s) {
    pdl_enter();
    __shared__ unsigned int hist[NCELLS];
    const int slot = blockIdx.y;
    if (slot == 1) return;  // (large-|z| bin: see k_cell_scatter)
    const unsigned int *src;
    unsigned int *dst, count;
    cell_slot(s, slot, src, dst, count);
    if ((unsigned long long)blockIdx.x * 256ull >= count) return;
    for (int c = threadIdx.x; c < NCELLS; c += 256) hist[c] = 0;
    __syncthreads();
    for (unsigned int t = blockIdx.x * 256 + threadIdx.x; t < count; t += gridDim.x * 256) {
        const unsigned int i = __ldg(src + t);
        atomicAdd(&hist[s.cellkey[i]], 1u);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < NCELLS; c += 256)
        if (hist[c]) atomicAdd(&s.cellcnt[slot * NCELLS + c], hist[c]);
}

constexpr int CSCAN_T = NCELLS < 1024 ? NCELLS : 1024;  // threads of k_cell_scan
__global__ void __launch_bounds__(CSCAN_T) k_cell_scan(Scratch s) {
    pdl_enter();
    constexpr int PER = NCELLS / CSCAN_T;  // consecutive cells per thread
    __shared__ unsigned int wsum[CSCAN_T / 32];
    const int slot = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int *cnt = s.cellcnt + slot * NCELLS + threadIdx.x * PER;
    unsigned int x[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        x[k] = cnt[k];
        sum += x[k];
    }
    unsigned int incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int y = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += y;
    }
    if (lane == 31) wsum[wid] = incl;
    __syncthreads();
    unsigned int off = 0;
    for (int w = 0; w < wid; ++w) off += wsum[w];
    unsigned int running = off + incl - sum;  // exclusive: where the thread's first cell starts inside its bin segment
#pragma unroll
    for (int k = 0; k < PER; ++k) {
        cnt[k] = running;
        running += x[k];
    }
}

__global__ void __launch_bounds__(256) k_cell_scatter(DevJob job, Scratch s) {
    pdl_enter();
    constexpr int R = 4;  // entries per thread and batch
    __shared__ unsigned int cnt[NCELLS], base[NCELLS];
    const int slot = blockIdx.y;
    const unsigned int *src;
    unsigned int *dst, count;
    cell_slot(s, slot, src, dst, count);
    if (slot == 1) {
        // The large-|z| bin keeps its stable (ascending index) order: its routine costs 0.06 ns per point, and at that
        // rate the kernel is bound by the memory system once neighbouring lanes stop reading neighbouring points
        // (measured on config 3: 27.5 instead of 23.5 active lanes, but 8.9 instead of 4.4 ms).  Copied, so that the
        // region kernels of a cell-sorted job read one index array.
        for (unsigned int t = blockIdx.x * 256 + threadIdx.x; t < count; t += gridDim.x * 256) dst[t] = src[t];
        return;
    }
#pragma unroll 1
    for (unsigned int t0 = blockIdx.x * (256 * R); t0 < count; t0 += gridDim.x * (256 * R)) {
        for (int c = threadIdx.x; c < NCELLS; c += 256) cnt[c] = 0;
        __syncthreads();
        unsigned int idx[R], key[R], rank[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned int t = t0 + threadIdx.x + 256 * r;
            key[r] = 0xffffffffu;
            if (t < count) {
                idx[r] = __ldg(src + t);
                key[r] = s.cellkey[idx[r]];
                rank[r] = atomicAdd(&cnt[key[r]], 1u);
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < NCELLS; c += 256)
            if (cnt[c]) base[c] = atomicAdd(&s.cellcnt[slot * NCELLS + c], cnt[c]);
        __syncthreads();
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (key[r] != 0xffffffffu) dst[base[key[r]] + rank[r]] = idx[r];
        __syncthreads();
    }
}

int launch_cellsort(const DevJob &job, const Scratch &s, int sms, cudaStream_t st) {
    // (the cell counters were zeroed by k_scatter of this chunk)
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = job.pdl ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.stream = st;
    cfg.dynamicSmemBytes = 0;
    cfg.blockDim = dim3(256, 1, 1);
    cfg.gridDim = dim3((unsigned int)(sms * 2), CELL_SLOTS, 1);
    cudaLaunchKernelEx(&cfg, k_cell_hist, job, s);
    launch_pdl(job.pdl != 0, k_cell_scan, CELL_SLOTS, CSCAN_T, st, s);
    cudaLaunchKernelEx(&cfg, k_cell_scatter, job, s);
    return 3;
}

template <int FAM, int KODE>
struct ClassifyLaunch {
    static void run(const DevJob &job, const Scratch &s, cudaStream_t st) {
        launch_pdl(job.pdl != 0, k_classify<FAM, KODE>, s.nblocks, 256, st, job, s);
    }
};

void launch_classify(const DevJob &job, const Scratch &s, cudaStream_t st) {
    dispatch_fam_kode<ClassifyLaunch>(job.fam, job.kode, job, s, st);
}

void launch_scan_scatter(const DevJob &job, const Scratch &s, cudaStream_t st) {
    launch_pdl(job.pdl != 0, k_scan, NBINS, 1024, st, s);
    launch_pdl(job.pdl != 0, k_scatter, s.nblocks, 256, st, job, s);
}

}  // namespace spbk
