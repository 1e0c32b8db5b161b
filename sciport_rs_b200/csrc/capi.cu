// capi.cu — the extern "C" launcher layer declared in include/sciport_b200.h.
//
// Host side of the drop-in boundary: argument checking, per-device scratch
// cache, chunking, the host-buffer pipeline (pinned async copies, two streams
// per device, one host thread per device) and the device-pointer fast path
// used when the caller already keeps the batch in HBM.  No CPU fallback: every
// compute entry point fails with SPB_E_NO_DEVICE when no CUDA device exists.
#include "../../include/sciport_b200.h"
#include "spb_kernels.cuh"
#include "amos/spb_family.h"


#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace spbk;

namespace {

thread_local std::string g_err;
thread_local double g_last_ms = 0.0;
thread_local int g_last_launches = 0;
// device-pointer calls are asynchronous: their events are resolved when the timing is asked for
thread_local cudaEvent_t g_pending_ev0 = nullptr, g_pending_ev1 = nullptr;

int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CK(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(SPB_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

// points per pipeline pass when data is resident (2^24) and per staged chunk for host buffers (2^22); the
// environment knobs SPB_CHUNK_LOG2 / SPB_HOST_CHUNK_LOG2 (read once) exist for the tests: small chunks exercise the
// multi-chunk paths on small batches, 2^25 and more makes k_scan take several rounds per bin
long long env_log2(const char *name, int dflt, int lo, int hi) {
    const char *e = getenv(name);
    int l = e ? atoi(e) : dflt;
    if (l < lo) l = lo;
    if (l > hi) l = hi;
    return 1ll << l;
}
const long long CHUNK_DEVICE = env_log2("SPB_CHUNK_LOG2", 24, 12, 27);
const long long CHUNK_HOST = env_log2("SPB_HOST_CHUNK_LOG2", 22, 12, 26);

constexpr int NSTREAM = 3;

struct ScratchSet {
    Scratch s{};
    long long cap = 0;
    void *arena = nullptr;
};

struct StageSet {  // device staging for host-pointer mode
    double2 *z = nullptr;
    double *x = nullptr;
    double *v = nullptr;
    double2 *out = nullptr;
    int *ierr = nullptr;
    int *nz = nullptr;
    double2 *out2 = nullptr;
    int *ierr2 = nullptr;
    int *nz2 = nullptr;
    double2 *seq = nullptr;  // [n_orders][m] block of an order-sequence chunk
    long long cap = 0;
    long long cap_out = 0;   // capacity of seq in values
};

struct DeviceCtx {
    int dev = -1;
    int sms = 148;
    std::mutex mu;
    // host-buffer calls stage their chunks round-robin over NSTREAM streams, each with its own staging set and
    // pipeline scratch: with two, the device-to-host engine idles whenever both streams are in their upload / compute
    // phase; three keep both copy engines busy (config 3 pair, host buffers: 189 -> see DESIGN.md section 6)
    ScratchSet scratch[NSTREAM];
    StageSet stage[NSTREAM];
    cudaStream_t streams[NSTREAM] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t fork = nullptr, join[NSTREAM] = {};  // chunk interleaving of device-pointer calls (run_device)
    unsigned long long *d_packed = nullptr;
    unsigned long long *d_first = nullptr;         // [2] first failing element (packed) of the call in flight
    unsigned long long h_first[2] = {~0ull, ~0ull};  // host copy, valid once the call has completed
    double2 *d_scalar = nullptr;                   // one complex scalar (I0(beta) of the Kaiser window)
    // outcome of k_direct on the first chunk of a multi-chunk device-pointer call, read back without blocking
    cudaEvent_t probe_ev = nullptr;
    unsigned int *h_probe = nullptr;  // pinned
    // The scratch of a device is shared by all calls on it.  Device-pointer calls return while their kernels are
    // still queued, so a following call on ANOTHER stream first waits (on the device) for the event that ends the
    // previous one; calls on the same stream are ordered by the stream itself.
    cudaEvent_t busy = nullptr;
    cudaStream_t busy_stream = nullptr;
    bool busy_valid = false;
};

static void order_after_previous(DeviceCtx *c, cudaStream_t st) {
    if (c->busy_valid && c->busy_stream != st) cudaStreamWaitEvent(st, c->busy, 0);
}
static void mark_busy(DeviceCtx *c, cudaStream_t st) {
    cudaEventRecord(c->busy, st);
    c->busy_stream = st;
    c->busy_valid = true;
}

// which devices the last compute call of this thread used, for spb_last_first_error()
thread_local std::vector<int> g_first_devs;
thread_local cudaStream_t g_first_stream = nullptr;
thread_local bool g_first_pending = false;  // device-pointer call: result still to be fetched from d_first

std::mutex g_ctx_mu;
std::vector<DeviceCtx *> g_ctx;

int device_count() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

DeviceCtx *get_ctx(int dev) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if ((int)g_ctx.size() <= dev) g_ctx.resize(dev + 1, nullptr);
    if (!g_ctx[dev]) {
        DeviceCtx *c = new DeviceCtx();
        c->dev = dev;
        cudaSetDevice(dev);
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) == cudaSuccess) c->sms = p.multiProcessorCount;
        for (int b = 0; b < NSTREAM; ++b) cudaStreamCreateWithFlags(&c->streams[b], cudaStreamNonBlocking);
        cudaEventCreate(&c->ev0);
        cudaEventCreate(&c->ev1);
        cudaEventCreateWithFlags(&c->busy, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
        for (int b = 0; b < NSTREAM; ++b) cudaEventCreateWithFlags(&c->join[b], cudaEventDisableTiming);
        cudaMalloc(&c->d_packed, sizeof(unsigned long long));
        cudaMalloc(&c->d_first, 2 * sizeof(unsigned long long));
        cudaMalloc(&c->d_scalar, sizeof(double2));
        cudaEventCreateWithFlags(&c->probe_ev, cudaEventDisableTiming);
        cudaMallocHost(&c->h_probe, sizeof(unsigned int));
        g_ctx[dev] = c;
    }
    return g_ctx[dev];
}


// ---- host topology of a device (sysfs): PCI address, NUMA node, CPUs local to its PCIe root ---------------------
struct DevTopo {
    std::string bdf, cpulist;
    int numa = -1;
    std::vector<int> cpus;
};

static std::string read_line(const std::string &path) {
    std::string out;
    if (FILE *f = fopen(path.c_str(), "r")) {
        char buf[4096];
        if (fgets(buf, sizeof buf, f)) out = buf;
        fclose(f);
    }
    while (!out.empty() && (out.back() == '\n' || out.back() == ' ')) out.pop_back();
    return out;
}

static std::vector<int> parse_cpulist(const std::string &spec) {
    std::vector<int> cpus;
    size_t i = 0;
    while (i < spec.size()) {
        size_t j = spec.find(',', i);
        if (j == std::string::npos) j = spec.size();
        const std::string part = spec.substr(i, j - i);
        const size_t dash = part.find('-');
        const int lo = atoi(part.c_str());
        const int hi = dash == std::string::npos ? lo : atoi(part.c_str() + dash + 1);
        for (int c = lo; c <= hi && c - lo < 4096; ++c) cpus.push_back(c);
        i = j + 1;
    }
    return cpus;
}

static DevTopo device_topology(int dev) {
    DevTopo t;
    char id[32] = {0};
    if (cudaDeviceGetPCIBusId(id, sizeof id, dev) != cudaSuccess) {
        cudaGetLastError();
        return t;
    }
    t.bdf = id;
    for (auto &ch : t.bdf) ch = (char)tolower(ch);
    const std::string base = "/sys/bus/pci/devices/" + t.bdf + "/";
    const std::string node = read_line(base + "numa_node");
    if (!node.empty()) t.numa = atoi(node.c_str());
    // local_cpulist names the CPUs behind the device's PCIe root even where numa_node reads -1 (virtualised hosts)
    t.cpulist = read_line(base + "local_cpulist");
    if (t.cpulist.empty() && t.numa >= 0)
        t.cpulist = read_line("/sys/devices/system/node/node" + std::to_string(t.numa) + "/cpulist");
    t.cpus = parse_cpulist(t.cpulist);
    return t;
}

// Run the calling thread on the CPUs local to `dev` (best effort; a no-op when the topology cannot be read or the
// list covers every CPU anyway).  Used by the per-device worker threads of a multi-device call, so that the staging
// copies of each slice are driven from the socket its GPU hangs off.
static void bind_thread_near_device(int dev) {
    const DevTopo t = device_topology(dev);
    if (t.cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    int k = 0;
    for (int c : t.cpus)
        if (c >= 0 && c < CPU_SETSIZE) {
            CPU_SET(c, &set);
            ++k;
        }
    if (k > 0) pthread_setaffinity_np(pthread_self(), sizeof set, &set);
}

// Contiguous slice [g n / G, (g + 1) n / G) of an n-element batch per device: f(g, lo, hi, &launches) runs on one
// host thread per device (the calling thread itself when there is one device).  No collective: shards never talk.
template <class F>
int slice_over_devices(const std::vector<int> &devs, long long n, int *launches_total, F f) {
    const int G = (int)devs.size();
    if (G == 1) {
        int l = 0;
        int rc = f(0, 0ll, n, &l);
        if (launches_total) *launches_total += l;
        return rc;
    }
    std::vector<int> rcs(G, SPB_OK), ls(G, 0);
    std::vector<std::string> errs(G);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        th.emplace_back([&, g] {
            bind_thread_near_device(devs[g]);
            const long long lo = n * g / G, hi = n * (g + 1) / G;
            if (hi > lo) rcs[g] = f(g, lo, hi, &ls[g]);
            errs[g] = g_err;
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g) {
        if (launches_total) *launches_total += ls[g];
        if (rcs[g] != SPB_OK) return fail(rcs[g], errs[g]);
    }
    return SPB_OK;
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int ensure_scratch(ScratchSet &ss, long long n, cudaStream_t st) {
    if (n <= ss.cap) {
        ss.s.nblocks = (int)((n + TILE - 1) / TILE);
        return SPB_OK;
    }
    if (ss.arena) cudaFree(ss.arena);
    ss.arena = nullptr;
    ss.cap = 0;
    long long cap = std::max<long long>(n, 1 << 16);
    int nblocks_cap = (int)((cap + TILE - 1) / TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    size_t o_cls = take(cap), o_pi = take(cap * 4), o_pk = take(cap * 4), o_gl = take(cap * 4),
           o_it = take(cap * 16), o_ic = take(cap), o_bc = take((size_t)NBINS * nblocks_cap * 4), o_bins = take(BINS_WORDS * 4),
           o_pi2 = take(cap * 4), o_pk2 = take(cap * 4), o_cc = take((size_t)CELL_SLOTS * NCELLS * 4), o_ck = take(cap * 2),
           o_dn = take((size_t)nblocks_cap * (TILE / 32));
    char *base = nullptr;
    if (cudaMalloc(&base, off) != cudaSuccess) {
        cudaGetLastError();
        return fail(SPB_E_ALLOC, "cudaMalloc of pipeline scratch failed");
    }
    ss.arena = base;
    ss.cap = cap;
    ss.s.cls = (unsigned char *)(base + o_cls);
    ss.s.perm_i = (unsigned int *)(base + o_pi);
    ss.s.perm_k = (unsigned int *)(base + o_pk);
    ss.s.glist = (unsigned int *)(base + o_gl);
    ss.s.perm_i2 = (unsigned int *)(base + o_pi2);
    ss.s.perm_k2 = (unsigned int *)(base + o_pk2);
    ss.s.cellcnt = (unsigned int *)(base + o_cc);
    ss.s.cellkey = (unsigned short *)(base + o_ck);
    ss.s.itemp = (double2 *)(base + o_it);
    ss.s.icode = (signed char *)(base + o_ic);
    ss.s.blockcnt = (unsigned int *)(base + o_bc);
    ss.s.bins = (unsigned int *)(base + o_bins);
    ss.s.done = (unsigned char *)(base + o_dn);
    ss.s.nblocks = (int)((n + TILE - 1) / TILE);
    // counters of the scan's last-CTA hand-off start at zero: on the stream the pipeline kernels will run on (the
    // pipeline streams are non-blocking, so a memset on the legacy stream would not be ordered before them)
    cudaMemsetAsync(ss.s.bins, 0, BINS_WORDS * 4, st);
    return SPB_OK;
}

// per-kernel profile of the last call made with SPB_PROFILE (events between launches)
struct ProfRec {
    int kernel;  // spb_kernel_id
    double ms;
    long long points;
};
thread_local std::vector<ProfRec> g_prof;

struct Prof {
    bool on = false;
    std::vector<cudaEvent_t> ev;
    std::vector<int> ids;
    void mark(int id, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
        ids.push_back(id);
    }
};

// enqueue the whole pipeline for one chunk; returns number of kernel launches
// SPB_FUSE (environment, read once): 0 never fuse, 1 fuse for a broadcast scalar order only, 2 (default) always
static int fuse_mode() {
    static const int mode = [] {
        const char *e = getenv("SPB_FUSE");
        return e ? atoi(e) : 2;
    }();
    return mode;
}

static int cellsort_mode() {
    static const int mode = [] {
        const char *e = getenv("SPB_CELLSORT");
        return e ? atoi(e) : 1;
    }();
    return mode;
}

// SPB_INTERLEAVE (environment, read once): number of internal streams (each with its own pipeline scratch) the
// chunks of a multi-chunk device-pointer call rotate over, forked from and joined back into the caller's stream;
// default 2, at most NSTREAM; 0 or 1: every chunk on the caller's stream.
static int interleave_lanes() {
    static const int lanes = [] {
        const char *e = getenv("SPB_INTERLEAVE");
        const int v = e ? atoi(e) : 2;
        return v < 1 ? 1 : (v > NSTREAM ? NSTREAM : v);  // 0 or 1: off
    }();
    return lanes;
}

static bool cellside_enabled() {
    static const bool on = [] {
        const char *e = getenv("SPB_CELLSIDE");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

static bool group_enabled() {
    static const bool on = [] {
        const char *e = getenv("SPB_GROUP");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// SPB_DIRECT (environment, read once): 0 switches off the direct large-|z| kernel of broadcast-order jobs
static bool direct_enabled() {
    static const bool on = [] {
        const char *e = getenv("SPB_DIRECT");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// smallest chunk the direct kernel is launched for (SPB_DIRECT_MIN_LOG2 in the environment, read once; the tests
// lower it to exercise the kernel on small batches)
static long long direct_min_points() {
    static const long long n = [] {
        const char *e = getenv("SPB_DIRECT_MIN_LOG2");
        const int l = e ? atoi(e) : 22;
        return 1ll << (l < 0 ? 0 : (l > 40 ? 40 : l));
    }();
    return n;
}

static bool debye_enabled() {
    static const bool on = [] {
        const char *e = getenv("SPB_DEBYE");
        return !(e && atoi(e) == 0);
    }();
    return on;
}

// Does k_direct pay on this call?  On a batch of unrelated points it finishes nothing, and although its warps give up
// after a few groups, one more persistent grid at the head of every chunk costs the interleaved chunks of a long call
// 0.1 ms each (config 4: 4.84 -> 5.29 ms per 2^26 points).  The first chunk of a call always runs it; the number of
// points it finished is copied to pinned host memory right behind the kernel, and the chunks enqueued after that
// copy has landed skip the kernel if the answer was zero.  Never blocks: while the answer is pending the kernel runs.
struct DirectProbe {
    cudaEvent_t ev;
    unsigned int *h;
    int state = 0;  // 0 nothing enqueued yet, 1 answer pending, 2 keep the kernel, 3 drop it
    bool wanted() {
        if (state == 1) {
            const cudaError_t e = cudaEventQuery(ev);
            if (e == cudaSuccess)
                state = (*h != 0) ? 2 : 3;
            else if (e == cudaErrorNotReady)
                cudaGetLastError();  // "not ready" is an answer, not an error: keep it out of the call's error check
        }
        return state != 3;
    }
};

int enqueue_chunk(const DevJob &job_in, ScratchSet &ss, int sms, int flags, cudaStream_t st, Prof *pf = nullptr,
                  DirectProbe *probe = nullptr) {
    Prof dummy;
    Prof &p = pf ? *pf : dummy;
    DevJob job = job_in;
    {
        // families that need K: points in the large-|w| region of I take the fused I + K kernel (K from the series
        // the I routine sums anyway)
        const bool ktype = !(job.fam == SPB_FN_J || job.fam == SPB_FN_I);
        const int mode = fuse_mode();
        job.fuse = (ktype && !(flags & SPB_PATH_MONOLITHIC) && (mode >= 2 || (mode == 1 && job.v_stride == 0)))
                       ? spb::FUSE_ASYI : 0;
        // Debye-direct region (every family; SPB_DEBYE=0 in the environment switches it off)
        if (debye_enabled() && !(flags & SPB_PATH_MONOLITHIC)) job.fuse |= spb::FUSE_DEBYE;
        // K of a Miller + Wronskian point from the Wronskian routine itself (with the fused large-|w| kernel)
        if (job.fuse & spb::FUSE_ASYI) job.fuse |= spb::FUSE_WRSK;
        // second-level (|z|, order) sort inside the iterative bins: SPB_CELLSORT = 0 never, 1 (default) for
        // per-element orders, 2 always
        job.cellsort = (cellsort_mode() == 2 || (cellsort_mode() == 1 && job.v_stride != 0)) ? 1 : 0;
        // bit 1: the cell key carries the side of the family's continuation (SPB_CELLSIDE=0 switches it off)
        if (job.cellsort && cellside_enabled()) job.cellsort |= 2;
        // the stable-order bins (Debye-direct, fused large-|w|) take the same hint inside CTA batches (SPB_GROUP=0: off)
        job.group = (ktype && job.v_stride != 0 && group_enabled()) ? 1 : 0;  // (I and J: no continuation in the combine step)
        job.pdl = (pdl_enabled() && job.v_stride == 0) ? 1 : 0;
        // broadcast order: groups of 32 consecutive points that all pass the classifier's fast accept are evaluated
        // in input order by k_direct before anything is classified (K-type families: by the fused large-|w| routine)
        // (not below 2^22 points per chunk: there the kernel can win a few microseconds at most, and on a batch of
        // unrelated points its launch, although its warps give up at once, is 4 - 9 % of a 2^20-point call: config 1)
        job.direct = (job.v_stride == 0 && direct_enabled() && !(flags & (SPB_PATH_MONOLITHIC | SPB_PATH_NO_DIRECT)) &&
                      (!ktype || (job.fuse & spb::FUSE_ASYI)) && job.n >= direct_min_points()) ? 1 : 0;
        if (job.direct && probe && !probe->wanted()) job.direct = 0;
    }
    if (flags & SPB_PATH_MONOLITHIC) {
        p.mark(SPB_K_MONOLITHIC, st);
        launch_monolithic(job, st);
        p.mark(-1, st);
        return 1;
    }
    int launches = 0;
    if (job.direct) {
        p.mark(SPB_K_DIRECT, st);
        launch_direct(job, ss.s, sms, st);
        ++launches;
        if (probe && probe->state == 0) {
            cudaMemcpyAsync(probe->h, ss.s.bins + SLOT_DIRECT, sizeof(unsigned int), cudaMemcpyDeviceToHost, st);
            cudaEventRecord(probe->ev, st);
            probe->state = 1;
        }
    }
    p.mark(SPB_K_CLASSIFY, st);
    launch_classify(job, ss.s, st);
    p.mark(SPB_K_SCAN_SCATTER, st);
    launch_scan_scatter(job, ss.s, st);
    launches += 3;
    if (job.cellsort) launches += launch_cellsort(job, ss.s, sms, st);
    const bool ktype = !(job.fam == SPB_FN_J || job.fam == SPB_FN_I);  // K, H1, H2, Y and the JY pair
    static const int ipass_id[NB_I] = {SPB_K_IPASS_SERIES, SPB_K_IPASS_ASYI, SPB_K_IPASS_MLRI, SPB_K_IPASS_WRSK,
                                       SPB_K_IPASS_UNI1,   SPB_K_IPASS_UNI2, SPB_K_IPASS_DEBYE};
    for (int r = 0; r < NB_I; ++r) {
        p.mark(ipass_id[r], st);
        launch_ipass(job, ss.s, r, sms, st);
        ++launches;
    }
    if (ktype) {
        p.mark(SPB_K_KPASS, st);
        launch_kpass(job, ss.s, sms, st);
        ++launches;
    }
    p.mark(SPB_K_GENERAL, st);
    launches += launch_general(job, ss.s, sms, st);
    p.mark(-1, st);
    return launches;
}

// turn the recorded events into (kernel, ms, points) records; call after the stream is idle
void collect_profile(Prof &p, const ScratchSet &ss, long long n) {
    if (!p.on) return;
    unsigned int bins[SLOT_DIRECT_LAST + 1] = {0};
    cudaMemcpy(bins, ss.s.bins, sizeof(bins), cudaMemcpyDeviceToHost);
    for (size_t i = 0; i + 1 < p.ev.size(); ++i) {
        if (p.ids[i] < 0) continue;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p.ev[i], p.ev[i + 1]);
        long long pts = n;
        int id = p.ids[i];
        if ((id >= SPB_K_IPASS_SERIES && id <= SPB_K_IPASS_WRSK) || id == SPB_K_IPASS_UNI1 || id == SPB_K_IPASS_UNI2 ||
            id == SPB_K_IPASS_DEBYE) {
            int r = id == SPB_K_IPASS_UNI1 ? 4 : id == SPB_K_IPASS_UNI2 ? 5 : id == SPB_K_IPASS_DEBYE ? 6
                                                                                : id - SPB_K_IPASS_SERIES;
            pts = (long long)bins[r + 1] - bins[r];
        } else if (id == SPB_K_KPASS) {
            pts = (long long)bins[BIN_GENERAL] - bins[BIN_K0];
        } else if (id == SPB_K_GENERAL) {
            pts = bins[GLIST_LEN_SLOT];
        } else if (id == SPB_K_DIRECT) {
            pts = bins[SLOT_DIRECT_LAST];
        }
        g_prof.push_back({id, ms, pts});
    }
    for (auto e : p.ev) cudaEventDestroy(e);
    p.ev.clear();
    p.ids.clear();
}

int check_exec(const spb_exec *ex, std::vector<int> &devs, bool &device_ptrs, cudaStream_t &stream, int &flags) {
    int ndev = device_count();
    if (ndev <= 0) return fail(SPB_E_NO_DEVICE, "no CUDA device visible: libsciport_b200 has no CPU path");
    device_ptrs = ex ? ex->device_ptrs != 0 : false;
    stream = ex ? (cudaStream_t)ex->stream : nullptr;
    flags = ex ? ex->flags : 0;
    devs.clear();
    int want = ex ? std::max(ex->n_devices, 1) : 1;
    if (ex && ex->devices) {
        for (int i = 0; i < want; ++i) devs.push_back(ex->devices[i]);
    } else if (want == 1) {
        int cur = 0;
        cudaGetDevice(&cur);
        devs.push_back(cur);
    } else {
        for (int i = 0; i < want; ++i) devs.push_back(i);
    }
    for (int d : devs)
        if (d < 0 || d >= ndev) return fail(SPB_E_ARG, "device ordinal out of range");
    if (device_ptrs && devs.size() != 1) return fail(SPB_E_ARG, "device pointers require exactly one device");
    return SPB_OK;
}

struct BesselArgs {
    int fn, kode;
    const double *v;
    long long v_stride;
    const double2 *z;
    const double *x;
    double2 *out;
    int *ierr;
    int *nz;
    long long n;
    int flags;
    double2 *out2 = nullptr;  // pair families
    int *ierr2 = nullptr;
    int *nz2 = nullptr;
    long long index_base = 0; // global index of element 0 of this slice (first-error bookkeeping)
    int n_orders = 1;         // > 1: order sequence, out is [n_orders][out_stride]
    long long out_stride = 0; // points per order row of the caller's result array
};

// Order sequences (n_orders > 1) with the batch resident on the device: one launch per chunk, every member
// written straight into its [order][point] slot of the caller's array.
int run_device_seq(DeviceCtx *c, const BesselArgs &a, cudaStream_t st, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    int launches = 0;
    order_after_previous(c, st);
    CK(cudaMemsetAsync(c->d_first, 0xff, 16, st));
    CK(cudaEventRecord(c->ev0, st));
    for (long long off = 0; off < a.n; off += CHUNK_DEVICE) {
        long long m = std::min(CHUNK_DEVICE, a.n - off);
        DevJob job{};
        job.first_err = c->d_first;
        job.index_base = a.index_base + off;
        job.fam = a.fn;
        job.kode = a.kode;
        job.sem = (a.flags & SPB_SEM_SCIPY) ? 1 : 0;
        job.reflect = 0;  // order sequences keep the raw behaviour for negative start orders
        job.v = a.v + off * a.v_stride;
        job.v_stride = a.v_stride;
        job.z = a.z ? a.z + off : nullptr;
        job.x = a.x ? a.x + off : nullptr;
        job.out = a.out + off;
        job.ierr = a.ierr ? a.ierr + off : nullptr;
        job.nz = a.nz ? a.nz + off : nullptr;
        job.n = m;
        int rc = ensure_scratch(c->scratch[0], m, st);
        if (rc != SPB_OK) return rc;
        launches += launch_sequence(job, a.n_orders, a.out_stride, c->scratch[0].s, c->sms, st);
    }
    CK(cudaEventRecord(c->ev1, st));
    mark_busy(c, st);
    CK(cudaGetLastError());
    g_pending_ev0 = c->ev0;
    g_pending_ev1 = c->ev1;
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

// data resident on the device: chunk over the batch on the caller's stream
int run_device(DeviceCtx *c, const BesselArgs &a, cudaStream_t st, double *ms_out, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    int launches = 0;
    Prof prof;
    prof.on = (a.flags & SPB_PROFILE) != 0;
    if (prof.on) g_prof.clear();
    order_after_previous(c, st);
    CK(cudaMemsetAsync(c->d_first, 0xff, 16, st));
    CK(cudaEventRecord(c->ev0, st));
    // Chunk interleaving.  A chunk is a dozen dependent launches; several of them have next to no parallelism (the
    // scans, the general path over a few hundred points, empty region bins) and every persistent region kernel ends
    // in a tail.  Consecutive chunks are independent, so they alternate between two internal streams with a scratch
    // set each: whatever one chunk leaves idle, the kernels of the next one fill.  Results are those of the
    // one-stream order (every point is written by its own chunk; the first-error word is an atomic minimum).
    const int lanes = (!prof.on && a.n > CHUNK_DEVICE) ? interleave_lanes() : 1;
    const bool interleave = lanes > 1;
    if (interleave) {
        CK(cudaEventRecord(c->fork, st));
        for (int b = 0; b < lanes; ++b) CK(cudaStreamWaitEvent(c->streams[b], c->fork, 0));
    }
    DirectProbe probe{c->probe_ev, c->h_probe};
    int chunk_no = 0;
    for (long long off = 0; off < a.n; off += CHUNK_DEVICE, ++chunk_no) {
        long long m = std::min(CHUNK_DEVICE, a.n - off);
        const int lane = interleave ? (chunk_no % lanes) : 0;
        cudaStream_t cst = interleave ? c->streams[lane] : st;
        int rc = ensure_scratch(c->scratch[lane], m, cst);
        if (rc != SPB_OK) return rc;
        DevJob job{};
        job.fam = a.fn;
        job.kode = a.kode;
        job.sem = (a.flags & SPB_SEM_SCIPY) ? 1 : 0;
        job.reflect = (a.flags & SPB_RAW_NEG_ORDERS) ? 0 : 1;
        job.v = a.v + off * a.v_stride;
        job.v_stride = a.v_stride;
        job.z = a.z ? a.z + off : nullptr;
        job.x = a.x ? a.x + off : nullptr;
        job.out = a.out + off;
        job.ierr = a.ierr ? a.ierr + off : nullptr;
        job.nz = a.nz ? a.nz + off : nullptr;
        job.out2 = a.out2 ? a.out2 + off : nullptr;
        job.ierr2 = a.ierr2 ? a.ierr2 + off : nullptr;
        job.nz2 = a.nz2 ? a.nz2 + off : nullptr;
        job.n = m;
        job.first_err = c->d_first;
        job.first_err2 = c->d_first + 1;
        job.index_base = a.index_base + off;
        launches += enqueue_chunk(job, c->scratch[lane], c->sms, a.flags, cst, &prof,
                                  (a.n > CHUNK_DEVICE && !prof.on && c->h_probe) ? &probe : nullptr);
        if (prof.on) {
            // profiling serialises the chunks so that the bin counters can be read back per chunk
            CK(cudaStreamSynchronize(st));
            if (!(a.flags & SPB_PATH_MONOLITHIC)) collect_profile(prof, c->scratch[0], m);
            else {
                unsigned int z17[1];
                (void)z17;
                for (size_t i = 0; i + 1 < prof.ev.size(); ++i) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, prof.ev[i], prof.ev[i + 1]);
                    if (prof.ids[i] >= 0) g_prof.push_back({prof.ids[i], ms, m});
                }
                for (auto e : prof.ev) cudaEventDestroy(e);
                prof.ev.clear();
                prof.ids.clear();
            }
        }
    }
    if (interleave) {
        for (int b = 0; b < lanes; ++b) {
            CK(cudaEventRecord(c->join[b], c->streams[b]));
            CK(cudaStreamWaitEvent(st, c->join[b], 0));
        }
    }
    CK(cudaEventRecord(c->ev1, st));
    mark_busy(c, st);
    CK(cudaGetLastError());
    // no host synchronisation: the call returns as soon as the launches are queued on the caller's stream
    g_pending_ev0 = c->ev0;
    g_pending_ev1 = c->ev1;
    if (ms_out) *ms_out = 0.0;
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

int ensure_stage(StageSet &s, long long n, bool real_in, bool per_elem_v, bool want_ierr, bool want_nz) {
    if (n > s.cap) {
        cudaFree(s.z);
        cudaFree(s.x);
        cudaFree(s.v);
        cudaFree(s.out);
        cudaFree(s.ierr);
        cudaFree(s.nz);
        cudaFree(s.out2);
        cudaFree(s.ierr2);
        cudaFree(s.nz2);
        cudaFree(s.seq);
        s = StageSet();
        if (cudaMalloc(&s.z, n * 16) != cudaSuccess || cudaMalloc(&s.x, n * 8) != cudaSuccess ||
            cudaMalloc(&s.v, n * 8) != cudaSuccess || cudaMalloc(&s.out, n * 16) != cudaSuccess ||
            cudaMalloc(&s.ierr, n * 4) != cudaSuccess || cudaMalloc(&s.nz, n * 4) != cudaSuccess ||
            cudaMalloc(&s.out2, n * 16) != cudaSuccess || cudaMalloc(&s.ierr2, n * 4) != cudaSuccess ||
            cudaMalloc(&s.nz2, n * 4) != cudaSuccess) {
            cudaGetLastError();
            return fail(SPB_E_ALLOC, "cudaMalloc of staging buffers failed");
        }
        s.cap = n;
    }
    (void)real_in;
    (void)per_elem_v;
    (void)want_ierr;
    (void)want_nz;
    return SPB_OK;
}

// Pin a caller buffer for the duration of the call unless it already is.
struct Pin {
    void *p = nullptr;
    bool registered = false;
    void pin(const void *ptr, size_t bytes) {
        if (!ptr || bytes == 0) return;
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) return;
        cudaGetLastError();
        if (cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterPortable) == cudaSuccess) {
            p = const_cast<void *>(ptr);
            registered = true;
        } else {
            cudaGetLastError();  // pageable copies still work, just slower
        }
    }
    ~Pin() {
        if (registered) cudaHostUnregister(p);
    }
};

// host buffers, one device: staged chunks, two streams, copies overlapped with compute
int run_host_slice(DeviceCtx *c, const BesselArgs &a, double *ms_out, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    int launches = 0;
    int k = 0;
    double v0 = a.v[0];
    {
        const unsigned long long ones[2] = {~0ull, ~0ull};
        if (c->busy_valid) cudaEventSynchronize(c->busy);  // a queued device-pointer call still owns the scratch
        CK(cudaMemcpy(c->d_first, ones, 16, cudaMemcpyHostToDevice));  // host-buffer calls are synchronous
    }
    for (long long off = 0; off < a.n; off += CHUNK_HOST, ++k) {
        long long m = std::min(CHUNK_HOST, a.n - off);
        int b = k % NSTREAM;
        cudaStream_t st = c->streams[b];
        int rc = ensure_stage(c->stage[b], std::min(CHUNK_HOST, a.n), a.x != nullptr, a.v_stride != 0, a.ierr, a.nz);
        if (rc != SPB_OK) return rc;
        rc = ensure_scratch(c->scratch[b], m, st);
        if (rc != SPB_OK) return rc;
        StageSet &s = c->stage[b];
        if (a.z) CK(cudaMemcpyAsync(s.z, a.z + off, m * 16, cudaMemcpyHostToDevice, st));
        if (a.x) CK(cudaMemcpyAsync(s.x, a.x + off, m * 8, cudaMemcpyHostToDevice, st));
        if (a.v_stride == 1) {
            CK(cudaMemcpyAsync(s.v, a.v + off, m * 8, cudaMemcpyHostToDevice, st));
        } else if (a.v_stride == 0) {
            CK(cudaMemcpyAsync(s.v, &v0, 8, cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMemcpy2DAsync(s.v, 8, a.v + off * a.v_stride, a.v_stride * 8, 8, m, cudaMemcpyHostToDevice, st));
        }
        DevJob job{};
        job.fam = a.fn;
        job.kode = a.kode;
        job.sem = (a.flags & SPB_SEM_SCIPY) ? 1 : 0;
        job.reflect = (a.flags & SPB_RAW_NEG_ORDERS) ? 0 : 1;
        job.v = s.v;
        job.v_stride = a.v_stride == 0 ? 0 : 1;
        job.z = a.z ? s.z : nullptr;
        job.x = a.x ? s.x : nullptr;
        job.out = s.out;
        job.ierr = a.ierr ? s.ierr : nullptr;
        job.nz = a.nz ? s.nz : nullptr;
        job.out2 = a.out2 ? s.out2 : nullptr;
        job.ierr2 = a.ierr2 ? s.ierr2 : nullptr;
        job.nz2 = a.nz2 ? s.nz2 : nullptr;
        job.n = m;
        job.first_err = c->d_first;
        job.first_err2 = c->d_first + 1;
        job.index_base = a.index_base + off;
        launches += enqueue_chunk(job, c->scratch[b], c->sms, a.flags, st);
        CK(cudaMemcpyAsync(a.out + off, s.out, m * 16, cudaMemcpyDeviceToHost, st));
        if (a.ierr) CK(cudaMemcpyAsync(a.ierr + off, s.ierr, m * 4, cudaMemcpyDeviceToHost, st));
        if (a.nz) CK(cudaMemcpyAsync(a.nz + off, s.nz, m * 4, cudaMemcpyDeviceToHost, st));
        if (a.out2) CK(cudaMemcpyAsync(a.out2 + off, s.out2, m * 16, cudaMemcpyDeviceToHost, st));
        if (a.ierr2) CK(cudaMemcpyAsync(a.ierr2 + off, s.ierr2, m * 4, cudaMemcpyDeviceToHost, st));
        if (a.nz2) CK(cudaMemcpyAsync(a.nz2 + off, s.nz2, m * 4, cudaMemcpyDeviceToHost, st));
        // the staging set of this stream is reused two chunks later: stream order protects it
    }
    for (int b = 0; b < NSTREAM; ++b) CK(cudaStreamSynchronize(c->streams[b]));
    CK(cudaGetLastError());
    CK(cudaMemcpy(c->h_first, c->d_first, 16, cudaMemcpyDeviceToHost));
    if (ms_out) *ms_out = 0.0;
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

// Order sequences with host buffers: chunks sized so that one staged [n_orders][m] block is <= 256 MiB; the block
// goes back with one strided copy (row k of the block -> row k of the caller's [n_orders][n] array).
int run_host_slice_seq(DeviceCtx *c, const BesselArgs &a, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    long long chunk = std::max<long long>(1024, ((1ll << 24) / a.n_orders) & ~1023ll);
    chunk = std::min(chunk, CHUNK_HOST);
    int launches = 0, k = 0;
    double v0 = a.v[0];
    {
        const unsigned long long ones[2] = {~0ull, ~0ull};
        if (c->busy_valid) cudaEventSynchronize(c->busy);  // a queued device-pointer call still owns the scratch
        CK(cudaMemcpy(c->d_first, ones, 16, cudaMemcpyHostToDevice));
    }
    for (long long off = 0; off < a.n; off += chunk, ++k) {
        long long m = std::min(chunk, a.n - off);
        int b = k % NSTREAM;
        cudaStream_t st = c->streams[b];
        StageSet &s = c->stage[b];
        int rc = ensure_stage(s, std::min(chunk, a.n), a.x != nullptr, a.v_stride != 0, a.ierr, a.nz);
        if (rc != SPB_OK) return rc;
        long long need = std::min(chunk, a.n) * a.n_orders;
        if (need > s.cap_out) {
            cudaFree(s.seq);
            s.seq = nullptr;
            s.cap_out = 0;
            if (cudaMalloc(&s.seq, need * 16) != cudaSuccess) {
                cudaGetLastError();
                return fail(SPB_E_ALLOC, "cudaMalloc of the order-sequence staging block failed");
            }
            s.cap_out = need;
        }
        if (a.z) CK(cudaMemcpyAsync(s.z, a.z + off, m * 16, cudaMemcpyHostToDevice, st));
        if (a.x) CK(cudaMemcpyAsync(s.x, a.x + off, m * 8, cudaMemcpyHostToDevice, st));
        if (a.v_stride == 1) {
            CK(cudaMemcpyAsync(s.v, a.v + off, m * 8, cudaMemcpyHostToDevice, st));
        } else if (a.v_stride == 0) {
            CK(cudaMemcpyAsync(s.v, &v0, 8, cudaMemcpyHostToDevice, st));
        } else {
            CK(cudaMemcpy2DAsync(s.v, 8, a.v + off * a.v_stride, a.v_stride * 8, 8, m, cudaMemcpyHostToDevice, st));
        }
        DevJob job{};
        job.fam = a.fn;
        job.kode = a.kode;
        job.sem = (a.flags & SPB_SEM_SCIPY) ? 1 : 0;
        job.reflect = 0;  // order sequences keep the raw behaviour for negative start orders
        job.v = s.v;
        job.v_stride = a.v_stride == 0 ? 0 : 1;
        job.z = a.z ? s.z : nullptr;
        job.x = a.x ? s.x : nullptr;
        job.out = s.seq;
        job.ierr = a.ierr ? s.ierr : nullptr;
        job.nz = a.nz ? s.nz : nullptr;
        job.n = m;
        job.first_err = c->d_first;
        job.index_base = a.index_base + off;
        rc = ensure_scratch(c->scratch[b], m, st);
        if (rc != SPB_OK) return rc;
        launches += launch_sequence(job, a.n_orders, m, c->scratch[b].s, c->sms, st);
        CK(cudaMemcpy2DAsync(a.out + off, a.out_stride * 16, s.seq, m * 16, m * 16, a.n_orders, cudaMemcpyDeviceToHost,
                             st));
        if (a.ierr) CK(cudaMemcpyAsync(a.ierr + off, s.ierr, m * 4, cudaMemcpyDeviceToHost, st));
        if (a.nz) CK(cudaMemcpyAsync(a.nz + off, s.nz, m * 4, cudaMemcpyDeviceToHost, st));
    }
    for (int b = 0; b < NSTREAM; ++b) CK(cudaStreamSynchronize(c->streams[b]));
    CK(cudaGetLastError());
    CK(cudaMemcpy(c->h_first, c->d_first, 16, cudaMemcpyDeviceToHost));
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

int bessel_impl(int fn, int kode, const double *v, int64_t v_stride, const spb_c128 *z, const double *x,
                spb_c128 *out, int32_t *ierr, int32_t *nz, int64_t n, int n_orders, const spb_exec *ex,
                spb_c128 *out2 = nullptr, int32_t *ierr2 = nullptr, int32_t *nz2 = nullptr) {
    if (fn < SPB_FN_J || fn > SPB_FN_IK) return fail(SPB_E_ARG, "unknown function selector");
    const bool pair = (fn == SPB_FN_JY || fn == SPB_FN_IK);
    if (pair != (out2 != nullptr) && n > 0) return fail(SPB_E_ARG, "pair family needs two result buffers");
    if (kode != 1 && kode != 2) return fail(SPB_E_ARG, "kode must be 1 or 2");
    if (n < 0 || v_stride < 0) return fail(SPB_E_ARG, "negative n or v_stride");
    if (n_orders < 1 || n_orders > 4096) return fail(SPB_E_ARG, "n_orders must be 1..4096");
    if (n_orders > 1 && pair) return fail(SPB_E_ARG, "the pair families have no order-sequence form");
    if (n > 0 && (!v || (!z && !x) || !out)) return fail(SPB_E_ARG, "null buffer");
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    g_last_ms = 0.0;
    g_last_launches = 0;
    g_pending_ev0 = g_pending_ev1 = nullptr;
    if (n == 0) return SPB_OK;
    BesselArgs a{fn, kode, v, v_stride, (const double2 *)z, x, (double2 *)out, ierr, nz, n, flags};
    a.out2 = (double2 *)out2;
    a.ierr2 = ierr2;
    a.nz2 = nz2;
    a.n_orders = n_orders;
    a.out_stride = n;
    const bool seq = n_orders > 1;
    g_first_devs = devs;
    g_first_stream = st;
    g_first_pending = dptr;
    if (dptr)
        return seq ? run_device_seq(get_ctx(devs[0]), a, st, &g_last_launches)
                   : run_device(get_ctx(devs[0]), a, st, &g_last_ms, &g_last_launches);
    // Pin the caller's buffers once for the whole call (a copy may not straddle separately registered ranges, and
    // the slices of a multi-device call share pages at their boundaries), unless they already are pinned.
    Pin pz, px, pv, po, pe, pn, po2, pe2, pn2;
    if (a.z) pz.pin(a.z, n * 16);
    if (a.x) px.pin(a.x, n * 8);
    if (v_stride != 0) pv.pin(v, ((n - 1) * v_stride + 1) * 8);
    po.pin(out, (size_t)n * n_orders * 16);
    if (ierr) pe.pin(ierr, n * 4);
    if (nz) pn.pin(nz, n * 4);
    if (out2) po2.pin(out2, n * 16);
    if (ierr2) pe2.pin(ierr2, n * 4);
    if (nz2) pn2.pin(nz2, n * 4);
    // contiguous slice g of G per device, one host thread per device (no collective: shards never talk)
    return slice_over_devices(devs, n, &g_last_launches, [&](int g, long long lo, long long hi, int *launches) {
        BesselArgs sl = a;
        sl.index_base = lo;
        sl.v = a.v + lo * a.v_stride;
        sl.z = a.z ? a.z + lo : nullptr;
        sl.x = a.x ? a.x + lo : nullptr;
        sl.out = a.out + lo;
        sl.ierr = a.ierr ? a.ierr + lo : nullptr;
        sl.nz = a.nz ? a.nz + lo : nullptr;
        sl.out2 = a.out2 ? a.out2 + lo : nullptr;
        sl.ierr2 = a.ierr2 ? a.ierr2 + lo : nullptr;
        sl.nz2 = a.nz2 ? a.nz2 + lo : nullptr;
        sl.n = hi - lo;
        return seq ? run_host_slice_seq(get_ctx(devs[g]), sl, launches)
                   : run_host_slice(get_ctx(devs[g]), sl, nullptr, launches);
    });
}

// generic "map over a batch" helper for the elementwise kernels (gamma, ln|gamma|, complex gamma, sinc, ...):
// device pointers -> one launch on the caller's stream; host pointers -> contiguous slice per device, each slice
// staged in chunks over two streams (H2D, kernel, D2H of consecutive chunks overlap), buffers from the cached
// staging set (no allocation in steady state)
template <class F>
int map_host_slice(DeviceCtx *c, const char *in, char *out, long long n, size_t in_bytes_per, size_t out_bytes_per,
                   F launch, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    if (c->busy_valid) cudaEventSynchronize(c->busy);
    int k = 0, launches = 0;
    for (long long off = 0; off < n; off += CHUNK_HOST, ++k) {
        const long long m = std::min(CHUNK_HOST, n - off);
        const int b = k % NSTREAM;
        cudaStream_t st = c->streams[b];
        int rc = ensure_stage(c->stage[b], std::min(CHUNK_HOST, n), false, false, false, false);
        if (rc != SPB_OK) return rc;
        StageSet &sg = c->stage[b];  // z: 16 B per element of input, out: 16 B per element of output
        CK(cudaMemcpyAsync(sg.z, in + (size_t)off * in_bytes_per, (size_t)m * in_bytes_per, cudaMemcpyHostToDevice, st));
        launch((const void *)sg.z, (void *)sg.out, m, st);
        ++launches;
        CK(cudaMemcpyAsync(out + (size_t)off * out_bytes_per, sg.out, (size_t)m * out_bytes_per, cudaMemcpyDeviceToHost,
                           st));
    }
    for (int b = 0; b < NSTREAM; ++b) CK(cudaStreamSynchronize(c->streams[b]));
    CK(cudaGetLastError());
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

template <class F>
int run_map(const spb_exec *ex, long long n, size_t in_bytes_per, size_t out_bytes_per, const void *in, void *out,
            F launch) {
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    if (n < 0) return fail(SPB_E_ARG, "negative n");
    g_last_ms = 0.0;
    g_last_launches = 0;
    g_pending_ev0 = g_pending_ev1 = nullptr;
    if (n == 0) return SPB_OK;
    if (!in || !out) return fail(SPB_E_ARG, "null buffer");
    if (in_bytes_per > 16 || out_bytes_per > 16) return fail(SPB_E_ARG, "element wider than the staging buffers");
    if (dptr) {
        DeviceCtx *c = get_ctx(devs[0]);
        std::lock_guard<std::mutex> lk(c->mu);
        CK(cudaSetDevice(c->dev));
        order_after_previous(c, st);
        CK(cudaEventRecord(c->ev0, st));
        launch(in, out, n, st);
        CK(cudaEventRecord(c->ev1, st));
        mark_busy(c, st);
        CK(cudaGetLastError());
        // only enqueued, like every device-pointer call: spb_last_timing() resolves the events on demand
        g_pending_ev0 = c->ev0;
        g_pending_ev1 = c->ev1;
        g_last_launches = 1;
        return SPB_OK;
    }
    Pin pi, po;
    pi.pin(in, (size_t)n * in_bytes_per);
    po.pin(out, (size_t)n * out_bytes_per);
    return slice_over_devices(devs, n, &g_last_launches, [&](int g, long long lo, long long hi, int *launches) {
        return map_host_slice(get_ctx(devs[g]), (const char *)in + (size_t)lo * in_bytes_per,
                              (char *)out + (size_t)lo * out_bytes_per, hi - lo, in_bytes_per, out_bytes_per, launch,
                              launches);
    });
}

// Airy with host buffers, one device: staged chunks over two streams like run_host_slice
int airy_host_slice(DeviceCtx *c, int which, int kode, int sem, const spb_c128 *z, spb_c128 *out, int32_t *ierr,
                    int32_t *nz, long long n, long long index_base, int *launches_out) {
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    int launches = 0, k = 0;
    {
        const unsigned long long ones[2] = {~0ull, ~0ull};
        if (c->busy_valid) cudaEventSynchronize(c->busy);  // a queued device-pointer call still owns the scratch
        CK(cudaMemcpy(c->d_first, ones, 16, cudaMemcpyHostToDevice));
    }
    for (long long off = 0; off < n; off += CHUNK_HOST, ++k) {
        const long long m = std::min<long long>(CHUNK_HOST, n - off);
        const int b = k % NSTREAM;
        cudaStream_t s0 = c->streams[b];
        int rc = ensure_stage(c->stage[b], std::min<long long>(CHUNK_HOST, n), false, false, true, true);
        if (rc != SPB_OK) return rc;
        StageSet &sg = c->stage[b];
        CK(cudaMemcpyAsync(sg.z, z + off, m * 16, cudaMemcpyHostToDevice, s0));
        rc = ensure_scratch(c->scratch[b], m, s0);
        if (rc != SPB_OK) return rc;
        launches += launch_airy(which, kode, sem, sg.z, sg.out, ierr ? sg.ierr : nullptr, nz ? sg.nz : nullptr, m,
                                &c->scratch[b].s, c->sms, s0, c->d_first, index_base + off);
        CK(cudaMemcpyAsync(out + off, sg.out, m * 16, cudaMemcpyDeviceToHost, s0));
        if (ierr) CK(cudaMemcpyAsync(ierr + off, sg.ierr, m * 4, cudaMemcpyDeviceToHost, s0));
        if (nz) CK(cudaMemcpyAsync(nz + off, sg.nz, m * 4, cudaMemcpyDeviceToHost, s0));
    }
    for (int b = 0; b < NSTREAM; ++b) CK(cudaStreamSynchronize(c->streams[b]));
    CK(cudaGetLastError());
    CK(cudaMemcpy(c->h_first, c->d_first, 16, cudaMemcpyDeviceToHost));
    if (launches_out) *launches_out = launches;
    return SPB_OK;
}

}  // namespace

extern "C" {

int spb_bessel(int fn, int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out, int32_t *ierr,
               int32_t *nz, int64_t n, int n_orders, const spb_exec *ex) {
    return bessel_impl(fn, kode, v, v_stride, z, nullptr, out, ierr, nz, n, n_orders, ex);
}

int spb_bessel_jy(int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out_j, spb_c128 *out_y,
                  int32_t *ierr_j, int32_t *ierr_y, int32_t *nz_j, int32_t *nz_y, int64_t n, const spb_exec *ex) {
    if (n > 0 && !out_y) return fail(SPB_E_ARG, "null buffer");
    return bessel_impl(SPB_FN_JY, kode, v, v_stride, z, nullptr, out_j, ierr_j, nz_j, n, 1, ex, out_y, ierr_y, nz_y);
}

int spb_bessel_ik(int kode, const double *v, int64_t v_stride, const spb_c128 *z, spb_c128 *out_i, spb_c128 *out_k,
                  int32_t *ierr_i, int32_t *ierr_k, int32_t *nz_i, int32_t *nz_k, int64_t n, const spb_exec *ex) {
    if (n > 0 && !out_k) return fail(SPB_E_ARG, "null buffer");
    return bessel_impl(SPB_FN_IK, kode, v, v_stride, z, nullptr, out_i, ierr_i, nz_i, n, 1, ex, out_k, ierr_k, nz_k);
}

int spb_bessel_real(int fn, int kode, const double *v, int64_t v_stride, const double *x, spb_c128 *out,
                    int32_t *ierr, int32_t *nz, int64_t n, int n_orders, const spb_exec *ex) {
    return bessel_impl(fn, kode, v, v_stride, nullptr, x, out, ierr, nz, n, n_orders, ex);
}

int spb_airy(int which, int kode, const spb_c128 *z, spb_c128 *out, int32_t *ierr, int32_t *nz, int64_t n,
             const spb_exec *ex) {
    if (which < 0 || which > 3) return fail(SPB_E_ARG, "which must be 0..3");
    if (kode != 1 && kode != 2) return fail(SPB_E_ARG, "kode must be 1 or 2");
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    if (n < 0) return fail(SPB_E_ARG, "negative n");
    if (n == 0) return SPB_OK;
    if (!z || !out) return fail(SPB_E_ARG, "null buffer");
    const int sem = (flags & SPB_SEM_SCIPY) ? 1 : 0;
    g_last_ms = 0.0;
    g_last_launches = 0;
    g_pending_ev0 = g_pending_ev1 = nullptr;
    g_first_devs = devs;
    g_first_stream = st;
    g_first_pending = dptr;
    if (dptr) {
        DeviceCtx *c = get_ctx(devs[0]);
        std::lock_guard<std::mutex> lk(c->mu);
        CK(cudaSetDevice(c->dev));
        order_after_previous(c, st);
        CK(cudaMemsetAsync(c->d_first, 0xff, 16, st));
        CK(cudaEventRecord(c->ev0, st));
        for (long long off = 0; off < n; off += CHUNK_DEVICE) {
            long long m = std::min<long long>(CHUNK_DEVICE, n - off);
            rc = ensure_scratch(c->scratch[0], m, st);
            if (rc != SPB_OK) return rc;
            g_last_launches += launch_airy(which, kode, sem, (const double2 *)z + off, (double2 *)out + off,
                                           ierr ? ierr + off : nullptr, nz ? nz + off : nullptr, m, &c->scratch[0].s,
                                           c->sms, st, c->d_first, off);
        }
        CK(cudaEventRecord(c->ev1, st));
        mark_busy(c, st);
        CK(cudaGetLastError());
        // only enqueued, like the Bessel entry points: synchronise the stream before reading the results
        g_pending_ev0 = c->ev0;
        g_pending_ev1 = c->ev1;
        return SPB_OK;
    }
    Pin pz, po, pe, pn;
    pz.pin(z, (size_t)n * 16);
    po.pin(out, (size_t)n * 16);
    if (ierr) pe.pin(ierr, (size_t)n * 4);
    if (nz) pn.pin(nz, (size_t)n * 4);
    return slice_over_devices(devs, n, &g_last_launches, [&](int g, long long lo, long long hi, int *launches) {
        return airy_host_slice(get_ctx(devs[g]), which, kode, sem, z + lo, out + lo, ierr ? ierr + lo : nullptr,
                               nz ? nz + lo : nullptr, hi - lo, lo, launches);
    });
}

int spb_gamma(const double *x, double *out, int64_t n, const spb_exec *ex) {
    return run_map(ex, n, 8, 8, x, out, [](const void *i, void *o, long long m, cudaStream_t st) {
        launch_gamma(0, (const double *)i, (double *)o, m, st);
    });
}
int spb_lgamma(const double *x, double *out, int64_t n, const spb_exec *ex) {
    return run_map(ex, n, 8, 8, x, out, [](const void *i, void *o, long long m, cudaStream_t st) {
        launch_gamma(1, (const double *)i, (double *)o, m, st);
    });
}
int spb_cgamma(const spb_c128 *z, spb_c128 *out, int64_t n, const spb_exec *ex) {
    return run_map(ex, n, 16, 16, z, out, [](const void *i, void *o, long long m, cudaStream_t st) {
        launch_cgamma(0, (const double2 *)i, (double2 *)o, m, st);
    });
}
int spb_clgamma(const spb_c128 *z, spb_c128 *out, int64_t n, const spb_exec *ex) {
    return run_map(ex, n, 16, 16, z, out, [](const void *i, void *o, long long m, cudaStream_t st) {
        launch_cgamma(1, (const double2 *)i, (double2 *)o, m, st);
    });
}
int spb_sinc(const double *x, double *out, int64_t n, const spb_exec *ex) {
    return run_map(ex, n, 8, 8, x, out, [](const void *i, void *o, long long m, cudaStream_t st) {
        launch_sinc((const double *)i, (double *)o, m, st);
    });
}
int spb_devmath(int which, const double *x, double *out, int64_t n, const spb_exec *ex) {
    if (which < 0 || which > 4) return fail(SPB_E_ARG, "which must be 0 (exp), 1 (exp(-x)), 2 (sin), 3 (cos) or 4 (log)");
    return run_map(ex, n, 8, 8, x, out, [which](const void *i, void *o, long long m, cudaStream_t st) {
        launch_devmath(which, (const double *)i, (double *)o, m, st);
    });
}

// windows: nothing is read, m doubles are written (device pointer: on the caller's stream; host pointer: staged)
static int window_impl(int kind, int64_t m, int sym, const double *params, int n_params, double *out,
                       const spb_exec *ex) {
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    if (kind < SPB_WIN_BOXCAR || kind > SPB_WIN_KBD) return fail(SPB_E_ARG, "unknown window kind");
    if (m < 0) return fail(SPB_E_ARG, "negative window length");
    if (n_params < 0 || n_params > SPB_WIN_MAX_PARAMS || (n_params > 0 && !params))
        return fail(SPB_E_ARG, "bad window parameters");
    static const int need[] = {0, 0, 1, 0, 0, 0, 0, 0, 2, 1, 3, 1, 2, 1, 0, 1};
    if (n_params < need[kind]) return fail(SPB_E_ARG, "too few window parameters");
    WinParams p{};
    p.n = n_params;
    for (int k = 0; k < n_params; ++k) p.a[k] = params[k];
    if ((kind == SPB_WIN_KAISER || kind == SPB_WIN_KBD) && !(p.a[0] == p.a[0])) return fail(SPB_E_ARG, "beta is NaN");
    if (kind == SPB_WIN_TAYLOR && !(p.a[0] >= 2.0 && p.a[0] <= SPB_WIN_MAX_PARAMS + 1.0))
        return fail(SPB_E_ARG, "taylor: nbar must be 2..17");
    if (kind == SPB_WIN_KBD && m > 1 && (m % 2) != 0)
        return fail(SPB_E_ARG, "Kaiser-Bessel derived windows are only defined for an even number of points");
    if (m == 0) return SPB_OK;
    if (!out) return fail(SPB_E_ARG, "null buffer");
    // degenerate parameter values are other windows (windows.rs:423-427)
    if (kind == SPB_WIN_TUKEY && p.a[0] <= 0.0) kind = SPB_WIN_BOXCAR;
    if (kind == SPB_WIN_TUKEY && p.a[0] >= 1.0) {
        kind = SPB_WIN_GENERAL_COSINE;
        p.a[0] = 0.5;
        p.a[1] = 0.5;
        p.n = 2;
    }
    DeviceCtx *c = get_ctx(devs[0]);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    cudaStream_t s0 = dptr ? st : c->streams[0];
    // the Kaiser kinds share the per-device scalar I0(beta) and KBD the staging buffer: order against a device-pointer
    // call that may still be queued on another stream
    if (dptr)
        order_after_previous(c, s0);
    else if (c->busy_valid)
        cudaEventSynchronize(c->busy);
    double *dout = out;
    if (!dptr) {
        rc = ensure_stage(c->stage[0], std::max<long long>(1024, (m + 1) / 2 + 1), false, false, false, false);
        if (rc != SPB_OK) return rc;
        dout = (double *)c->stage[0].out;  // 16 bytes per staged element: room for m doubles
    }
    g_last_launches = 1;
    cudaError_t e = cudaSuccess;
    if (m <= 1) {
        // len_guards (windows.rs:624-626, 650-656): windows of length <= 1 are all ones
        static const double one = 1.0;
        e = cudaMemcpyAsync(dout, &one, 8, cudaMemcpyHostToDevice, s0);
    } else {
        // periodic windows (sym = false) are the first m points of the symmetric window of length m + 1
        // (extend / truncate, windows.rs:628-647)
        const long long m_ext = sym ? m : m + 1;
        if (kind == SPB_WIN_KAISER) {
            launch_kaiser(m_ext, m, p.a[0], c->d_scalar, dout, c->sms, s0);
            g_last_launches = 2;
        } else if (kind == SPB_WIN_LANCZOS) {
            launch_lanczos(m_ext, m, dout, s0);
        } else if (kind == SPB_WIN_KBD) {
            // Kaiser window of length m/2 + 1 into the z staging buffer, then the scan (sym is ignored: windows.rs:560)
            const long long h = m / 2;
            rc = ensure_stage(c->stage[1], std::max<long long>(1024, h + 1), false, false, false, false);
            if (rc != SPB_OK) return rc;
            double *kbuf = (double *)c->stage[1].z;
            launch_kaiser(h + 1, h + 1, p.a[0], c->d_scalar, kbuf, c->sms, s0);
            launch_kbd_finish(kbuf, h, dout, s0);
            g_last_launches = 3;
        } else {
            launch_window(kind, m_ext, m, p, dout, s0);
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && !dptr) {
        e = cudaMemcpyAsync(out, dout, m * 8, cudaMemcpyDeviceToHost, s0);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s0);
    }
    if (dptr) mark_busy(c, s0);
    if (e != cudaSuccess) return fail(SPB_E_CUDA, cudaGetErrorString(e));
    return SPB_OK;
}

int spb_kaiser(int64_t m, double beta, int sym, double *out, const spb_exec *ex) {
    return window_impl(SPB_WIN_KAISER, m, sym, &beta, 1, out, ex);
}
int spb_lanczos(int64_t m, int sym, double *out, const spb_exec *ex) {
    return window_impl(SPB_WIN_LANCZOS, m, sym, nullptr, 0, out, ex);
}
int spb_window(int kind, int64_t m, int sym, const double *params, int n_params, double *out, const spb_exec *ex) {
    return window_impl(kind, m, sym, params, n_params, out, ex);
}

int spb_first_error(const int32_t *ierr, int64_t n, int64_t *idx, int32_t *code, const spb_exec *ex) {
    if (!idx || !code) return fail(SPB_E_ARG, "null output");
    *idx = -1;
    *code = 0;
    if (n < 0) return fail(SPB_E_ARG, "negative n");
    if (n == 0) return SPB_OK;
    if (!ierr) return fail(SPB_E_ARG, "null buffer");
    if (!ex || !ex->device_ptrs) {
        // host array: plain scan (this is host-side bookkeeping, not part of the compute path)
        for (int64_t i = 0; i < n; ++i)
            if (ierr[i] != 0) {
                *idx = i;
                *code = ierr[i];
                break;
            }
        return SPB_OK;
    }
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    DeviceCtx *c = get_ctx(devs[0]);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    CK(cudaMemsetAsync(c->d_packed, 0xff, 8, st));
    launch_first_error(ierr, n, c->d_packed, st);
    unsigned long long packed = 0;
    CK(cudaMemcpyAsync(&packed, c->d_packed, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (packed != ~0ull) {
        *idx = (int64_t)(packed >> 8);
        *code = (int32_t)(packed & 0xff);
    }
    return SPB_OK;
}

int spb_synth(int cfg, uint64_t seed, int64_t offset, int64_t n, int64_t n_total, double *v, spb_c128 *z,
              const spb_exec *ex) {
    if (cfg < 1 || cfg > 5) return fail(SPB_E_ARG, "cfg must be 1..5");
    if (n < 0 || offset < 0) return fail(SPB_E_ARG, "negative n or offset");
    std::vector<int> devs;
    bool dptr;
    cudaStream_t st;
    int flags;
    int rc = check_exec(ex, devs, dptr, st, flags);
    if (rc != SPB_OK) return rc;
    if (n == 0) return SPB_OK;
    DeviceCtx *c = get_ctx(devs[0]);
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->dev));
    if (dptr) {
        launch_synth(cfg, seed, offset, n, n_total, v, (double2 *)z, st);
        CK(cudaGetLastError());
        return SPB_OK;
    }
    double *dv = nullptr;
    double2 *dz = nullptr;
    if (v && cudaMalloc(&dv, n * 8) != cudaSuccess) return fail(SPB_E_ALLOC, "cudaMalloc failed");
    if (z && cudaMalloc(&dz, n * 16) != cudaSuccess) {
        cudaFree(dv);
        return fail(SPB_E_ALLOC, "cudaMalloc failed");
    }
    launch_synth(cfg, seed, offset, n, n_total, dv, dz, c->streams[0]);
    cudaError_t e = cudaSuccess;
    if (v) e = cudaMemcpyAsync(v, dv, n * 8, cudaMemcpyDeviceToHost, c->streams[0]);
    if (z && e == cudaSuccess) e = cudaMemcpyAsync(z, dz, n * 16, cudaMemcpyDeviceToHost, c->streams[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->streams[0]);
    cudaFree(dv);
    cudaFree(dz);
    if (e != cudaSuccess) return fail(SPB_E_CUDA, cudaGetErrorString(e));
    return SPB_OK;
}

int spb_fp64_peak(int device, double seconds, double *tflops, double *sm_mhz) {
    int ndev = device_count();
    if (ndev <= 0) return fail(SPB_E_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev || !tflops) return fail(SPB_E_ARG, "bad device or null output");
    *tflops = run_fp64_peak(device, seconds, sm_mhz);
    return SPB_OK;
}

int spb_last_profile(int max_records, int *kernel_ids, double *ms, int64_t *points) {
    int n = (int)g_prof.size();
    if (n > max_records) n = max_records;
    for (int i = 0; i < n; ++i) {
        if (kernel_ids) kernel_ids[i] = g_prof[i].kernel;
        if (ms) ms[i] = g_prof[i].ms;
        if (points) points[i] = g_prof[i].points;
    }
    return n;
}

int spb_last_timing(double *kernel_ms, int *launches) {
    if (kernel_ms && g_pending_ev1) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_pending_ev1) == cudaSuccess &&
            cudaEventElapsedTime(&ms, g_pending_ev0, g_pending_ev1) == cudaSuccess)
            g_last_ms = ms;
        g_pending_ev0 = g_pending_ev1 = nullptr;
    }
    if (kernel_ms) *kernel_ms = g_last_ms;
    if (launches) *launches = g_last_launches;
    return SPB_OK;
}

int spb_last_first_error(int which, int64_t *idx, int32_t *code) {
    if (!idx || !code || which < 0 || which > 1) return fail(SPB_E_ARG, "bad argument");
    *idx = -1;
    *code = 0;
    unsigned long long best = ~0ull;
    for (int d : g_first_devs) {
        DeviceCtx *c = get_ctx(d);
        if (g_first_pending) {
            // device-pointer call: fetch the word behind the work queued on the caller's stream
            std::lock_guard<std::mutex> lk(c->mu);
            CK(cudaSetDevice(c->dev));
            CK(cudaMemcpyAsync(c->h_first, c->d_first, 16, cudaMemcpyDeviceToHost, g_first_stream));
            CK(cudaStreamSynchronize(g_first_stream));
        }
        best = std::min(best, c->h_first[which]);
    }
    g_first_pending = false;
    if (best != ~0ull) {
        *idx = (int64_t)(best >> 8);
        *code = (int32_t)(best & 0xff);
    }
    return SPB_OK;
}

int spb_device_topology(int device, char *buf, int buflen) {
    const int ndev = device_count();
    if (ndev <= 0) return fail(SPB_E_NO_DEVICE, "no CUDA device visible");
    if (device < 0 || device >= ndev || !buf || buflen <= 0) return fail(SPB_E_ARG, "bad device or buffer");
    const DevTopo t = device_topology(device);
    return snprintf(buf, (size_t)buflen, "bdf=%s numa=%d cpus=%s", t.bdf.c_str(), t.numa, t.cpulist.c_str());
}

int spb_device_count(void) { return device_count(); }
int spb_version(void) { return 100; }
const char *spb_last_error(void) { return g_err.c_str(); }

}  // extern "C"
