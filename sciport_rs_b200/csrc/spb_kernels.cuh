// spb_kernels.cuh — declarations shared by the kernel translation units and the
// C-ABI launcher layer (capi.cu).  One `DevJob` describes one chunk of one
// spb_bessel() call on one device; all pointers are device pointers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace spbk {

struct DevJob {
    int fam;   // spb::Family
    int kode;  // 1 | 2
    int sem;   // 0 raw AMOS, 1 scipy-compatible post-processing
    int fuse;  // spb::FUSE_ASYI (1): points needing I and K with I in the large-|w| region are evaluated by the fused
               // region kernel; spb::FUSE_DEBYE (2): the Debye-direct region (spb_debye.h) is carved out
    int pdl;   // 1: the kernels of this chunk are launched as programmatic dependents of each other (launch_pdl)
    int cellsort;  // bit 0: inside the iterative region bins the point indices are sorted once more, by (|z|, order) cell
                   // (k_pipeline.cu): the bins' kernels then read perm_i2 / perm_k2; bit 1: the cell key also carries
                   // the side of the family's continuation
    int group;    // 1: inside a tile the compaction orders the entries of the stable bins (large-|w|, Debye-direct) by the
                  // side of the family's continuation (side_of, spb_dev.cuh; bit 6 of the class byte), then by index
    int reflect;  // 1: orders v < 0 by the reflection formulae (K folds v -> |v| at load; the others on the general path)
    int direct;   // 1: k_direct (k_direct.cu) ran over this chunk first: groups of 32 consecutive points that it accepted
                  // and finished are flagged in Scratch::done, the classifier and the compaction skip them
    const double *v;
    long long v_stride;  // 0 = broadcast
    const double2 *z;    // complex input, or
    const double *x;     // real input promoted to complex (z == nullptr)
    double2 *out;
    int *ierr;  // nullable
    int *nz;    // nullable
    double2 *out2;  // second result of a pair family (FAM_JY: out = J, out2 = Y; FAM_IK: out = I, out2 = K)
    int *ierr2;     // nullable
    int *nz2;       // nullable
    long long n;  // points in this chunk
    // first failing element of the whole call, packed (global index << 8 | ierr), kept with atomicMin so that a
    // caller who only wants `collect::<Result<_, i32>>` semantics never has to read per-element codes back
    unsigned long long *first_err;   // nullable; primary output
    unsigned long long *first_err2;  // nullable; second output of a pair family
    long long index_base;            // global index of element 0 of this chunk
};

// Scratch of one chunk (device memory, owned by the per-device cache in capi.cu).
struct Scratch {
    unsigned char *cls;     // [n] class byte: bits 0-2 I bin (7 none), 3-4 K bin (3 none), 5 general, 6 side (ordering hint)
    unsigned int *perm_i;   // [n] point indices sorted by I bin (stable)
    unsigned int *perm_k;   // [n] point indices sorted by K bin (stable)
    unsigned int *perm_i2;  // [n] perm_i with the segments of the iterative bins re-sorted by (|z|, order) cell
    unsigned int *perm_k2;  // [n] perm_k re-sorted likewise
    unsigned int *cellcnt;  // [CELL_SLOTS][NCELLS] cell histogram of a chunk, then (k_cell_scan) the cell cursors
    unsigned short *cellkey;  // [n] (|z|, order) cell of every point, written by the classifier of a cell-sorted job
    unsigned int *glist;    // [n] general-path list
    double2 *itemp;         // [n] I(w) for families that also need K
    signed char *icode;     // [n] nz of the I pass (<0: handed to the general path)
    unsigned int *blockcnt; // [NBINS][nblocks] per-tile histogram, then exclusive offsets
    unsigned int *bins;     // [NBINS+1] start offsets of bins (concatenated order); [SLOT_GLEN] = glist length
    unsigned char *done;    // [64 * nblocks] one flag per group of 32 consecutive points: finished by k_direct
    int nblocks;            // tiles in this chunk
};

// bins: 0..6 I regions (series, asyi, mlri, wrsk, uniform form 1, uniform form 2, Debye-direct), 7..9 K regions
// (series, miller, half), 10 general.  The I bins index perm_i, the K bins perm_k, the general bin glist.
constexpr int NB_I = 7;
constexpr int NB_K = 3;
constexpr int BIN_K0 = NB_I;
constexpr int BIN_GENERAL = NB_I + NB_K;
constexpr int NBINS = BIN_GENERAL + 1;
// packed per-tile counters: a tile holds 2048 points, so 12 bits per field and five fields per 64-bit word
constexpr int FIELD_BITS = 12;
constexpr int FIELDS_PER_WORD = 5;
constexpr int NWORDS = (NBINS + FIELDS_PER_WORD - 1) / FIELDS_PER_WORD;
constexpr int TILE = 2048;               // points per classification tile
// second-level sort inside a bin: 512 cells = 32 levels of |z| (four per octave, 1 <= |z| < 221) x 16 levels of the
// order (steps of 4); slots = the bins that are sorted this way (every I bin but the Debye-direct one, whose routine
// has no data-dependent loop; the three K bins)
// cell key of the second-level sort: CELL_ZBITS bits of |z| level, CELL_VBITS bits of order level, 1 bit of side
#ifndef SPB_CELL_ZBITS
#define SPB_CELL_ZBITS 5
#endif
#ifndef SPB_CELL_VBITS
#define SPB_CELL_VBITS 4
#endif
constexpr int CELL_ZBITS = SPB_CELL_ZBITS, CELL_VBITS = SPB_CELL_VBITS;
constexpr int NCELLS = 1 << (CELL_ZBITS + CELL_VBITS + 1);
constexpr int CELL_SLOTS = 9;
// slots of Scratch::bins beyond the NBINS+1 start offsets
constexpr int GLIST_LEN_SLOT = 24;  // current length of the general list (grows when a pass hands a point over)
constexpr int SLOT_DONE = 25;       // CTAs of k_scan that have finished
constexpr int SLOT_G2 = 26;         // length of the second-stage general list (points left after the shortcuts)
constexpr int SLOT_DIRECT = 27;     // points finished by k_direct in the chunk being classified (k_scan moves it on)
constexpr int SLOT_DIRECT_LAST = 28;  // ... in the chunk whose pipeline ran last (read by the profile)
constexpr int SLOT_TOTALS = 32;     // [NBINS] bin totals
constexpr int BINS_WORDS = 64;

// Persistent grids are sized to exactly the number of CTAs that are resident at once
// (SM count x occupancy), so every CTA starts in the first wave and the grid-stride
// loops finish together.
// Programmatic dependent launch for the kernels of one pipeline: every pipeline kernel begins with
// pdl_enter() (griddepcontrol.launch_dependents, then griddepcontrol.wait), so the NEXT kernel's CTAs may be
// scheduled as soon as all CTAs of this one have started, and they block in their own wait until this kernel has
// completed and its writes are visible.  Results are those of plain stream order; what overlaps is the launch
// latency and the ramp-up of each kernel with the tail of its predecessor (a dozen launches per chunk, several of
// them over empty bins).  Used for jobs with a broadcast scalar order (DevJob::pdl): with per-element orders the
// kernels run for milliseconds and the early dependents measured 2-3 % slower.  SPB_PDL=0 in the environment
// falls back to plain launches everywhere.
bool pdl_enabled();
template <class... KArgs, class... Args>
inline void launch_pdl(bool on, void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned int)grid, 1, 1);
    cfg.blockDim = dim3((unsigned int)block, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = on ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// The occupancy of every kernel is looked up once per (kernel, device) and cached (k_misc.cu).
int resident_ctas_per_sm(const void *kernel, int threads);
template <class Kernel>
inline int resident_grid(Kernel kernel, int threads, int sms) {
    return resident_ctas_per_sm((const void *)kernel, threads) * sms;
}

// Family and scaling flag as compile-time constants of the hot kernels: classify and the K pass branch on them in
// every routine they inline (geometry, combine, scaled/unscaled forms); as template parameters the dead forms
// vanish (K pass 10.4 k -> 4.8 k SASS instructions, 104 -> 94 registers; measured 0.45 -> 0.38 ms on config 2).
// F<FAM, KODE>::run(args...) is called with the job's family and kode.
template <template <int, int> class F, class... A>
inline void dispatch_fam_kode(int fam, int kode, A &&...a) {
#define SPB_FAM_CASE(FAM)               \
    case FAM:                           \
        if (kode == 1)                  \
            F<FAM, 1>::run(a...);       \
        else                            \
            F<FAM, 2>::run(a...);       \
        break;
    switch (fam) {
        SPB_FAM_CASE(0)
        SPB_FAM_CASE(1)
        SPB_FAM_CASE(2)
        SPB_FAM_CASE(3)
        SPB_FAM_CASE(4)
        SPB_FAM_CASE(5)
        SPB_FAM_CASE(6)
        SPB_FAM_CASE(7)
        default: break;  // families are validated at the C ABI
    }
#undef SPB_FAM_CASE
}

// kernels.cu
void launch_classify(const DevJob &job, const Scratch &s, cudaStream_t st);
void launch_scan_scatter(const DevJob &job, const Scratch &s, cudaStream_t st);
// large-|w| points of a broadcast-order job straight from the input order, before classification (k_direct.cu)
void launch_direct(const DevJob &job, const Scratch &s, int sms, cudaStream_t st);
int launch_cellsort(const DevJob &job, const Scratch &s, int sms, cudaStream_t st);  // returns the launches made
void launch_ipass(const DevJob &job, const Scratch &s, int region, int sms, cudaStream_t st);
// job.fuse: I + K of the K-type points of one I region (large-|w| series / Miller + Wronskian) in one kernel
void launch_fused(const DevJob &job, const Scratch &s, int sms, cudaStream_t st, int region);
void launch_debye(const DevJob &job, const Scratch &s, int sms, cudaStream_t st);  // Debye-direct region (k_debye.cu)
void launch_kpass(const DevJob &job, const Scratch &s, int sms, cudaStream_t st);
int launch_general(const DevJob &job, const Scratch &s, int sms, cudaStream_t st);
void launch_monolithic(const DevJob &job, cudaStream_t st);
int launch_sequence(const DevJob &job, int n_orders, long long out_stride, const Scratch &s, int sms,
                    cudaStream_t st);
// returns the number of kernels launched; with scratch (sized for n) the points are sorted by region first
int launch_airy(int which, int kode, int sem, const double2 *z, double2 *out, int *ierr, int *nz, long long n,
                const Scratch *s, int sms, cudaStream_t st, unsigned long long *first_err, long long index_base);
void launch_synth(int cfg, unsigned long long seed, long long offset, long long n, long long n_total, double *v,
                  double2 *z, cudaStream_t st);
void launch_first_error(const int *ierr, long long n, unsigned long long *packed, cudaStream_t st);
void launch_gamma(int which, const double *x, double *out, long long n, cudaStream_t st);
void launch_cgamma(int which, const double2 *z, double2 *out, long long n, cudaStream_t st);
void launch_sinc(const double *x, double *out, long long n, cudaStream_t st);
void launch_devmath(int which, const double *x, double *out, long long n, cudaStream_t st);
void launch_kaiser(long long m_ext, long long n_out, double beta, double2 *den, double *out, int sms, cudaStream_t st);
void launch_lanczos(long long m_ext, long long n_out, double *out, cudaStream_t st);
// elementwise windows (k_window.cu): parameters by value, at most SPB_WIN_MAX_PARAMS doubles
constexpr int SPB_WIN_MAX_PARAMS = 16;
struct WinParams {
    double a[SPB_WIN_MAX_PARAMS];
    int n;
};
void launch_window(int kind, long long m_ext, long long n_out, const WinParams &p, double *out, cudaStream_t st);
void launch_kbd_finish(const double *k, long long h, double *out, cudaStream_t st);
double run_fp64_peak(int device, double seconds, double *sm_mhz);

}  // namespace spbk
