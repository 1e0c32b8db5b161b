// k_misc.cu — small kernels around the hot path: synthetic inputs, first-error
// reduction, the memory-bound gamma / lgamma / sinc companions, and the DFMA
// microbenchmark that defines "FP64 peak" for the roofline.
#include "spb_dev.cuh"
#include "amos/spb_gamma.h"
#include <chrono>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

namespace spbk {

bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("SPB_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

// Resident CTAs per SM of one kernel on the current device (cudaOccupancyMaxActiveBlocksPerMultiprocessor),
// cached per (kernel, device, CTA size): persistent grids are sized from it on every launch.
int resident_ctas_per_sm(const void *kernel, int threads) {
    static std::mutex mu;
    static std::map<std::pair<const void *, long long>, int> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    const std::pair<const void *, long long> key(kernel, ((long long)dev << 16) | threads);
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, threads, 0) != cudaSuccess || nb < 1) nb = 1;
    cache[key] = nb;
    return nb;
}

// ---------------------------------------------------------------- synthetic inputs
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}
__device__ __forceinline__ double u01(unsigned long long seed, unsigned long long i, int k) {
    unsigned long long h = mix64(seed * 0x9E3779B97F4A7C15ull + 4ull * i + (unsigned long long)k);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}

// cfg 1: x = 100 u0 (real axis), v = 0      cfg 2: 4096 x 4096 grid, |z| <= 200, v = 2.5
// cfg 3: v = 50 u0, z = 200 u1 e^{i pi (2 u2 - 1)}      cfg 4: z = 10^(-3+7 u0) e^{i pi (2 u1 - 1)}, v = 2.5
// cfg 5: z as cfg 3, v = 0
__global__ void k_synth(int cfg, unsigned long long seed, long long offset, long long n, long long n_total,
                        double *v, double2 *z) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) {
        unsigned long long i = (unsigned long long)(offset + t);
        double vv = 0.0;
        double2 zz;
        if (cfg == 2) {
            long long side = 4096;
            if (n_total != side * side) {
                side = (long long)llrint(sqrt((double)n_total));
            }
            long long j = (long long)(i % (unsigned long long)side), k = (long long)(i / (unsigned long long)side);
            const double a = 141.42135623730951;  // 200/sqrt(2)
            zz.x = -a + (2.0 * a) * ((double)j + 0.5) / (double)side;
            zz.y = -a + (2.0 * a) * ((double)k + 0.5) / (double)side;
            vv = 2.5;
        } else if (cfg == 3 || cfg == 5) {
            vv = (cfg == 3) ? 50.0 * u01(seed, i, 0) : 0.0;
            double r = 200.0 * u01(seed, i, 1);
            double s, c;
            sincospi(2.0 * u01(seed, i, 2) - 1.0, &s, &c);
            zz.x = r * c;
            zz.y = r * s;
        } else if (cfg == 4) {
            vv = 2.5;
            double r = exp10(-3.0 + 7.0 * u01(seed, i, 0));
            double s, c;
            sincospi(2.0 * u01(seed, i, 1) - 1.0, &s, &c);
            zz.x = r * c;
            zz.y = r * s;
        } else {  // cfg 1
            zz.x = 100.0 * u01(seed, i, 0);
            zz.y = 0.0;
        }
        if (v) v[t] = vv;
        if (z) z[t] = zz;
    }
}

void launch_synth(int cfg, unsigned long long seed, long long offset, long long n, long long n_total, double *v,
                  double2 *z, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    k_synth<<<(int)blocks, 256, 0, st>>>(cfg, seed, offset, n, n_total, v, z);
}

// ---------------------------------------------------------------- first error in index order
// packed = min over i with ierr != 0 of (i << 8 | ierr); ~0 when none
__global__ void k_first_error(const int *ierr, long long n, unsigned long long *packed) {
    unsigned long long best = ~0ull;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int e = ierr[i];
        if (e != 0) {
            unsigned long long p = ((unsigned long long)i << 8) | (unsigned long long)(e & 0xff);
            if (p < best) best = p;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        unsigned long long o = __shfl_down_sync(0xffffffffu, best, d);
        if (o < best) best = o;
    }
    if ((threadIdx.x & 31) == 0 && best != ~0ull) atomicMin(packed, best);
}

void launch_first_error(const int *ierr, long long n, unsigned long long *packed, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    k_first_error<<<(int)blocks, 256, 0, st>>>(ierr, n, packed);
}

// ---------------------------------------------------------------- gamma / lgamma / sinc (HBM-bound)
// two points per thread through one 16-byte load / store
// WHICH 0..2: the companions; 3..7: the device transcendentals of the region kernels (exp, exp(-x) as the second
// member of the pair, sin, cos, log), exported through spb_devmath so that their accuracy is testable on its own
template <int WHICH>
__device__ __forceinline__ double gfun(double x) {
    if (WHICH == 0) return spb::gamma_real(x);
    if (WHICH == 1) return spb::lgamma_real(x);
    if (WHICH == 2) return spb::sinc_real(x);
    if (WHICH == 3) return spb::exp_(x);
    if (WHICH == 4) {
        double ep, em;
        spb::exp_pair(x, ep, em);
        return em;
    }
    if (WHICH == 7) return spb::log_(x);
    double sn, cs;
    spb::sincos_(x, sn, cs);
    return WHICH == 5 ? sn : cs;
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_real_map(const double *x, double *out, long long n) {
    const long long n2 = n >> 1;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
        double2 a = __ldg(x2 + i);
        double2 r;
        r.x = gfun<WHICH>(a.x);
        r.y = gfun<WHICH>(a.y);
        o2[i] = r;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = gfun<WHICH>(x[n - 1]);
}

// gamma / ln|gamma|: fast path + in-kernel queue.
// Both functions are a straight-line Stirling evaluation for x >= 12 and something longer and data-dependent below
// (recurrence product of up to 12 factors, Maclaurin series around 1 and 2, reflection for x < 0).  On unrelated
// inputs every warp holds a few such lanes, and the whole warp then walks every branch: ln|gamma| ran with 14.4 of 32
// lanes active (ncu, 2^27 points in [0, 100]).  Here a CTA takes tiles of 2048 points: every thread evaluates the
// fast path of its points at once and pushes the others (value + index) into a shared-memory queue; after a barrier
// the queue is processed densely, 32 slow points per warp.  A point takes the fast or the slow routine by its own
// value alone, so results do not depend on the neighbours.  One pass over memory: 8 B in, 8 B out per point.
template <int WHICH>
__device__ __forceinline__ bool gfast(double x, double &r) {
#if defined(__CUDA_ARCH__)
    if (WHICH == 0) {
        if (!(x >= 12.0 && x <= 171.6243769563027)) return false;
        r = spb::gamma_stirling_dev(x);
        return true;
    }
    if (!(x >= 12.0 && x < 1.0e305)) return false;
    r = spb::lgamma_stirling(x);
    return true;
#else
    (void)x;
    (void)r;
    return false;  // (host pass of nvcc: never called)
#endif
}

// Second class of points: below the Stirling range but with a straight-line evaluation of their own, the Stirling
// form after a FIXED upward shift (no data-dependent product loop): Gamma(x) = Gamma(x + 12) / (x (x+1) ... (x+11))
// for 1 <= x < 12, ln Gamma(x) = ln Gamma(x + 10) - ln(x (x+1) ... (x+9)) for 2.5 < x < 12 (ln Gamma >= 0.28 there:
// the zeros at 1 and 2 belong to the Maclaurin branch of the general routine).
template <int WHICH>
__device__ __forceinline__ bool gmid_ok(double x) {
    // (Gamma keeps its points below 12 in the general class: its product loop stops at the first y >= 12, about half
    // the factors of the fixed shift -- measured 1.17 vs 1.24 ms per 2^27 points)
    return WHICH == 0 ? false : (x > 2.5 && x < 12.0);
}
template <int WHICH>
__device__ __forceinline__ double gmid(double x) {
#if defined(__CUDA_ARCH__)
    constexpr int SHIFT = WHICH == 0 ? 12 : 10;
    double prod = x;
#pragma unroll
    for (int k = 1; k < SHIFT; ++k) prod *= x + (double)k;
    const double y = x + (double)SHIFT;
    if (WHICH == 0) return spb::gamma_stirling_dev(y) * spb::rcp_onscale(prod);  // 4.8e8 <= prod < 2e16
    return spb::lgamma_stirling(y) - spb::log_(prod);
#else
    return x;
#endif
}

// Three classes per tile: the fast path is evaluated in place; the two others are pushed (value + index) into one
// shared-memory queue each -- mid-range points from the front, general points from the back of the same arrays -- and
// processed densely after a barrier, every warp running ONE routine.
#ifndef SPB_GAMMA_NT
#define SPB_GAMMA_NT 256
#endif
template <int WHICH>
__global__ void __launch_bounds__(SPB_GAMMA_NT) k_real_map_queued(const double *x, double *out, long long n) {
    constexpr int NT = SPB_GAMMA_NT, IT = 4, TILE_PTS = NT * 2 * IT;
    __shared__ double qx[TILE_PTS];
    __shared__ unsigned int qi[TILE_PTS];
    __shared__ unsigned int qn_mid, qn_gen;
    const double2 *x2 = reinterpret_cast<const double2 *>(x);
    double2 *o2 = reinterpret_cast<double2 *>(out);
    const long long n2 = n >> 1;
    const long long tiles = (n2 + NT * IT - 1) / (NT * IT);
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (threadIdx.x == 0) {
            qn_mid = 0;
            qn_gen = 0;
        }
        __syncthreads();
        const long long base = tile * (NT * IT);
        double2 a[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const long long i = base + threadIdx.x + NT * k;
            a[k] = (i < n2) ? __ldg(x2 + i) : make_double2(20.0, 20.0);
        }
        // fast path for all eight points of the thread; what it declines is recorded in two bit masks
        unsigned int need_mid = 0, need_gen = 0;
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const long long i = base + threadIdx.x + NT * k;
            double2 r;
            const bool ok0 = gfast<WHICH>(a[k].x, r.x), ok1 = gfast<WHICH>(a[k].y, r.y);
            if (i < n2) {
                if (!ok0) (gmid_ok<WHICH>(a[k].x) ? need_mid : need_gen) |= 1u << (2 * k);
                if (!ok1) (gmid_ok<WHICH>(a[k].y) ? need_mid : need_gen) |= 1u << (2 * k + 1);
                o2[i] = r;  // (queued positions are overwritten below, after the barrier)
            }
        }
        // queue slots: the thread's counts, an exclusive scan over the warp, ONE shared atomic per warp and class
        // (an atomic per queued point -- or the compiler's aggregated form of it inside the divergent branches --
        // was 15 % of the kernel's instructions at two lanes active)
        {
            const unsigned int lane = threadIdx.x & 31u;
            const unsigned int cm = __popc(need_mid), cg = __popc(need_gen);
            unsigned int packed = cm | (cg << 16), incl = packed;  // (at most 8 x 32 per class and warp)
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= (unsigned int)d) incl += y;
            }
            const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int base_mid = 0, base_gen = 0;
            if (lane == 0) {
                if (total & 0xffffu) base_mid = atomicAdd(&qn_mid, total & 0xffffu);
                if (total >> 16) base_gen = atomicAdd(&qn_gen, total >> 16);
            }
            base_mid = __shfl_sync(0xffffffffu, base_mid, 0);
            base_gen = __shfl_sync(0xffffffffu, base_gen, 0);
            unsigned int qm = base_mid + ((incl - packed) & 0xffffu);
            unsigned int qg = base_gen + ((incl - packed) >> 16);
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                const unsigned int local = (unsigned int)(threadIdx.x + NT * k) * 2u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned int bit = 1u << (2 * k + h);
                    const double xv = h ? a[k].y : a[k].x;
                    if (need_mid & bit) {
                        qx[qm] = xv;
                        qi[qm] = local + h;
                        ++qm;
                    } else if (need_gen & bit) {
                        const unsigned int q = (unsigned int)TILE_PTS - 1u - qg;
                        qx[q] = xv;
                        qi[q] = local + h;
                        ++qg;
                    }
                }
            }
        }
        __syncthreads();
        const unsigned int m_mid = qn_mid, m_gen = qn_gen;
        for (unsigned int q = threadIdx.x; q < m_mid; q += NT) out[2 * base + qi[q]] = gmid<WHICH>(qx[q]);
        for (unsigned int q = threadIdx.x; q < m_gen; q += NT) {
            const unsigned int s = (unsigned int)TILE_PTS - 1u - q;
            out[2 * base + qi[s]] = gfun<WHICH>(qx[s]);
        }
        __syncthreads();
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = gfun<WHICH>(x[n - 1]);
}

template <int WHICH>
__global__ void __launch_bounds__(256) k_real_map_unaligned(const double *x, double *out, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = gfun<WHICH>(__ldg(x + i));
}

template <int WHICH>
static void launch_real_map(const double *x, double *out, long long n, cudaStream_t st) {
    long long blocks = (n / 2 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (aligned && (WHICH == 0 || WHICH == 1) && n >= 4096) {
        // gamma / ln|gamma|: fast path + in-kernel queue, persistent grid of resident CTAs
        const int grid = resident_grid(k_real_map_queued<WHICH>, SPB_GAMMA_NT, 148);
        k_real_map_queued<WHICH><<<grid, SPB_GAMMA_NT, 0, st>>>(x, out, n);
    } else if (aligned)
        k_real_map<WHICH><<<(int)blocks, 256, 0, st>>>(x, out, n);
    else
        k_real_map_unaligned<WHICH><<<(int)blocks, 256, 0, st>>>(x, out, n);
}

void launch_gamma(int which, const double *x, double *out, long long n, cudaStream_t st) {
    if (which == 0)
        launch_real_map<0>(x, out, n, st);
    else
        launch_real_map<1>(x, out, n, st);
}
void launch_sinc(const double *x, double *out, long long n, cudaStream_t st) { launch_real_map<2>(x, out, n, st); }
void launch_devmath(int which, const double *x, double *out, long long n, cudaStream_t st) {
    switch (which) {
        case 0: launch_real_map<3>(x, out, n, st); break;
        case 1: launch_real_map<4>(x, out, n, st); break;
        case 2: launch_real_map<5>(x, out, n, st); break;
        case 3: launch_real_map<6>(x, out, n, st); break;
        default: launch_real_map<7>(x, out, n, st); break;
    }
}

// ---------------------------------------------------------------- FP64 peak microbenchmark
// 8 independent DFMA chains per thread, register resident; 2 flops per DFMA.
__global__ void __launch_bounds__(256) k_dfma(double *sink, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3;
    double x4 = x0 + 4e-3, x5 = x0 + 5e-3, x6 = x0 + 6e-3, x7 = x0 + 7e-3;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 12345.678) sink[0] = s;  // never true; keeps the chains alive
}

double run_fp64_peak(int device, double seconds, double *sm_mhz) {
    cudaSetDevice(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    double *sink = nullptr;
    cudaMalloc(&sink, 8);
    const int blocks = prop.multiProcessorCount * 8;
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_dfma<<<blocks, 256>>>(sink, iters, 0.999999, 1e-7);  // warm-up
    cudaDeviceSynchronize();
    double best = 0.0;
    auto t0 = std::chrono::steady_clock::now();
    do {
        cudaEventRecord(e0);
        for (int r = 0; r < 8; ++r) k_dfma<<<blocks, 256>>>(sink, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        double flops = 8.0 * (double)blocks * 256.0 * (double)iters * 16.0 * 8.0 * 2.0;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    } while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds);
    if (sm_mhz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
        *sm_mhz = khz / 1000.0;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace spbk
