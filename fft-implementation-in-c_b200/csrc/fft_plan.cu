// fft_plan.cu - C-ABI of the engine (include/fftb200.h): device lifecycle, memory, plan construction
// (which tile-kernel variants run, in which order, through which buffers) and execution.
//
// Replaces gpu/fft_cuda.cu of the reference (cuFFT wrapper, never built): plan = cufftPlanMany
// (:138-163), exec = cufftExecZ2Z + device sync (:166-185), alloc/copies (:103-135).
// There is no CPU fallback anywhere in this file: without a CUDA device every call fails.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <cuda_runtime.h>
#include <ctype.h>
#include <sched.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/fftb200.h"
#include "fft_aux.cuh"
#include "fft_catalog.h"
#include "fft_fused.cuh"
#include "fft_pipe.cuh"
#include "fft_pipe13.cuh"
#include "fft_pipe13t.cuh"
#include "fft_lastpipe.cuh"

using namespace fftb200;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    if (getenv("FFTB200_VERBOSE")) fprintf(stderr, "fftb200: %s\n", g_err);
    return -1;
}
#define CU(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) return fail("%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" const char* fftb200_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// per-device state
// ---------------------------------------------------------------------------------------------
// A twiddle table in device memory, shared by the device's cache and by every plan that points at it: the memory goes
// away with the LAST reference, so fftb200_device_reset (fft_gpu_cleanup) or a larger table replacing it never pulls
// a table from under a live plan.
struct Table {
    cd* ptr = nullptr;
    int n = 0;
    int refs = 1;
};
struct DeviceState {
    bool init = false;
    int sms = 0;
    Table* tab = nullptr;   // largest reference-recurrence table uploaded so far
    Table* acc = nullptr;   // accurate (correctly rounded) stage tables, fixed size ACC_N
    char name[256] = "";
};
static DeviceState g_dev[64];
static std::mutex g_mu;
static int g_tables_alive = 0;   // diagnostics (fftb200_debug_tables_alive)

static void table_release_locked(Table* t) {
    if (t && --t->refs == 0) {
        cudaFree(t->ptr);
        delete t;
        g_tables_alive--;
    }
}
static void table_release(Table* t) {
    if (!t) return;
    std::lock_guard<std::mutex> lk(g_mu);
    table_release_locked(t);
}
extern "C" int fftb200_debug_tables_alive(void) { std::lock_guard<std::mutex> lk(g_mu); return g_tables_alive; }

// Plans remember their device: entry points that touch a plan switch to it for the duration of the call (a caller may
// have moved on with fft_gpu_set_device) and switch back.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev && dev >= 0) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

static int cur_device(DeviceState** ds) {
    int d = 0;
    CU(cudaGetDevice(&d));
    if (d < 0 || d >= 64) return fail("device index %d out of range", d);
    DeviceState& s = g_dev[d];
    if (!s.init) {
        cudaDeviceProp p;
        CU(cudaGetDeviceProperties(&p, d));
        s.sms = p.multiProcessorCount;
        snprintf(s.name, sizeof(s.name), "%s", p.name);
        s.init = true;
    }
    *ds = &s;
    return 0;
}

extern "C" int fftb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
extern "C" int fftb200_set_device(int device) { CU(cudaSetDevice(device)); return 0; }
extern "C" int fftb200_get_device(void) { int d = -1; if (cudaGetDevice(&d) != cudaSuccess) return -1; return d; }
extern "C" const char* fftb200_device_name(void) {
    DeviceState* s;
    std::lock_guard<std::mutex> lk(g_mu);
    if (cur_device(&s) != 0) return "";
    return s->name;
}
extern "C" int fftb200_sm_count(void) {
    DeviceState* s;
    std::lock_guard<std::mutex> lk(g_mu);
    if (cur_device(&s) != 0) return -1;
    return s->sms;
}
extern "C" int fftb200_mem_info(size_t* free_bytes, size_t* total_bytes) {
    size_t f = 0, t = 0;
    CU(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return 0;
}
extern "C" int fftb200_device_reset(void) {
    DeviceState* s;
    std::lock_guard<std::mutex> lk(g_mu);
    if (cur_device(&s) != 0) return -1;
    CU(cudaDeviceSynchronize());
    table_release_locked(s->tab);   // freed now unless a live plan still uses it
    table_release_locked(s->acc);
    s->tab = nullptr;
    s->acc = nullptr;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// memory
// ---------------------------------------------------------------------------------------------
extern "C" void* fftb200_malloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { fail("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void fftb200_free(void* p) { if (p) cudaFree(p); }
// CPUs of the NUMA node the current device hangs off (sysfs local_cpulist of its PCI function); false when unknown.
static bool device_local_cpus(cpu_set_t* set) {
    int dev = 0;
    char bus[32] = {0};
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), dev) != cudaSuccess) { cudaGetLastError(); return false; }
    for (char* c = bus; *c; c++) *c = (char)tolower((unsigned char)*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char line[1024] = {0};
    const bool ok = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!ok) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* p = line; *p && *p != '\n';) {       // "0-15,32-47"
        char* e;
        long a = strtol(p, &e, 10), b = a;
        if (e == p) break;
        if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET((int)c, set); count++; }
        p = (*e == ',') ? e + 1 : e;
    }
    return count > 0;
}

// Pinned host memory, placed on the NUMA node of the current device: the pages are allocated (first touched) inside
// cudaMallocHost by the calling thread, so the thread is moved onto the device's local CPUs for the duration of the call.
// Matters on multi-socket hosts, where a buffer on the far socket sends every PCIe transfer of the host pipeline
// (fftb200_plan_exec_host) across the inter-socket link; the boxes measured so far expose one NUMA node (no effect there:
// 2 GPUs, 121.6 ms per e2e step either way). FFTB200_NO_NUMA (set to anything) switches the placement off.
extern "C" void* fftb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    cpu_set_t old_set, dev_set;
    bool moved = false;
    if (!getenv("FFTB200_NO_NUMA") && sched_getaffinity(0, sizeof(old_set), &old_set) == 0 && device_local_cpus(&dev_set)) {
        cpu_set_t both;
        CPU_AND(&both, &old_set, &dev_set);             // stay inside what the process is allowed to use
        if (CPU_COUNT(&both) > 0 && !CPU_EQUAL(&both, &old_set)) moved = sched_setaffinity(0, sizeof(both), &both) == 0;
    }
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (moved) sched_setaffinity(0, sizeof(old_set), &old_set);
    if (e != cudaSuccess) { fail("cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void fftb200_host_free(void* p) { if (p) cudaFreeHost(p); }
// cudaMemcpy from pageable host memory may return while the DMA out of the driver's staging buffer is
// still in flight on the legacy stream; plan streams are non-blocking, so wait for it explicitly.
extern "C" int fftb200_memcpy_h2d(void* d, const void* s, size_t n) {
    CU(cudaMemcpy(d, s, n, cudaMemcpyHostToDevice));
    CU(cudaStreamSynchronize(cudaStreamLegacy));
    return 0;
}
extern "C" int fftb200_memcpy_d2h(void* d, const void* s, size_t n) { CU(cudaMemcpy(d, s, n, cudaMemcpyDeviceToHost)); return 0; }
extern "C" int fftb200_memcpy_d2d(void* d, const void* s, size_t n) { CU(cudaMemcpy(d, s, n, cudaMemcpyDeviceToDevice)); CU(cudaStreamSynchronize(cudaStreamLegacy)); return 0; }
extern "C" int fftb200_memset(void* d, int v, size_t n) { CU(cudaMemset(d, v, n)); CU(cudaStreamSynchronize(cudaStreamLegacy)); return 0; }

extern "C" int fftb200_fill_splitmix(void* dst, unsigned long long seed, unsigned long long first,
                                     unsigned long long count) {
    if (!dst) return fail("fill: null pointer");
    if (count == 0) return 0;
    fill_splitmix_kernel<<<(unsigned)((count + 255) / 256 > 1u << 20 ? 1u << 20 : (count + 255) / 256), 256>>>(
        (cd*)dst, seed, first, count);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int fftb200_pointwise_mul(void* y, const void* a, const void* b, size_t count) {
    if (!y || !a || !b) return fail("pointwise_mul: null pointer");
    if (count == 0) return 0;
    size_t blocks = (count + 255) / 256;
    if (blocks > (1u << 20)) blocks = 1u << 20;
    pointwise_mul_kernel<<<(unsigned)blocks, 256>>>((cd*)y, (const cd*)a, (const cd*)b, count, count);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
enum { BUF_IN = 0, BUF_OUT = 1, BUF_SCRATCH = 2 };

enum { ACC_N = 8192 };  // accurate tables cover stages m <= 8192 (hybrid twiddles)
// largest r2c size on the Hermitian schedule: mismatch vs the oracle 3.4e-13 there, 7.9e-13 at 2^17, 1.8e-12 at 2^18 (profiles/r02_real.md)
enum { R2C_HERM_MAX_LOG = 16 };

struct Pass {
    const KernelInfo* k;   // nullptr: persistent TMA kernel fft_pipe_kernel<log_p>, or the fused kernel when fused_lm > 0
    int fused_lm = 0, fused_lr = 0;   // fft_fused_kernel<fused_lm, fused_lr>: both passes in one launch
    int fused_cols = 0;               // column mode: stages 1 .. 16 of a larger transform, rows of 2^(log_n - 16) columns
    int tmem13 = 0;                   // N = 8192: the tensor-memory variant (fft_pipe13t.cuh) instead of fft_pipe13.cuh
    int lastpipe = 0;                 // LAST tile pass with a TMA-ring twin (fft_lastpipe.cuh), used when nothing rides on the pass
    int log_p;
    int log_m;
    int nt;        // whole-transform kernels: transforms per tile (tiles = ceil(nbatch / nt)); else 0
    int shift;     // strided / last kernels: log2 tiles per transform (tiles = nbatch << shift)
    int src, dst;
    int final_pass;
    int grid_max;  // persistent grid: SMs x resident CTAs
};

static long long pass_tiles(const Pass& ps, long long nbatch) {
    return ps.nt ? (nbatch + ps.nt - 1) / ps.nt : nbatch << ps.shift;
}

struct fftb200_plan {
    int device = 0;
    int n = 0, batch = 0, dir = -1, kind = FFTB200_C2C;
    int log_n = 0;             // of the power-of-two transform actually run (n, or Bluestein's m)
    int m = 0;                 // power-of-two length (== n for C2C / R2C)
    std::vector<Pass> passes;  // forward or inverse c2c of length m over `batch`
    const cd* tab = nullptr;
    const cd* acc = nullptr;   // accurate tables (nullptr: reference-recurrence tables everywhere)
    Table* tab_ref = nullptr;  // the shared tables `tab` / `acc` point into (one reference each, dropped by plan_destroy)
    Table* acc_ref = nullptr;
    cd* scratch = nullptr;     // ping-pong buffer for multi-pass plans, allocated on first use
    size_t scratch_elems = 0;
    cd* fscratch = nullptr;    // fused plans: L2-resident ring of `slots` groups of transforms
    size_t fscratch_elems = 0;
    int* fflags = nullptr;     // fused plans: per-group completion counters
    unsigned int* sched = nullptr;   // persistent kernels: tile hand-out counters (zero between launches, the kernels reset them)
    size_t fflags_count = 0;
    cd fdtw[3][16];            // fused plans: pass-B derived-twiddle constants (fft_fused.cuh: fused_twiddles)
    cd ldtw[3][16];            // the same for the last pass in the ring kernel (fft_lastpipe.cuh: derive)
    bool last_derive = false;
    cd* own_tab = nullptr;     // partial plans with a private (rank-specific) twiddle table
    cd** peers = nullptr;      // partial plans whose last pass stores into the peers' exchange buffers (device array of G pointers)
    int peer_lw = 0, peer_lrows = 0, peer_lg = 0, peer_me = 0;
    double scale = 0.0;        // 1/m, or the caller's value for partial plans of a distributed transform
    cd* work = nullptr;        // Bluestein / R2C: padded complex work array, m * batch
    fftb200_plan* child = nullptr;   // R2C of a non-power-of-two length: the Bluestein c2c plan that transforms the promoted input
    bool r2c_herm = false;     // R2C of 2^14 .. : the fused kernel transforms only the columns k <= M/2 in pass B (fft_fused.cuh, HERM)
    bool c2r_half = false;     // fused C2R that reads the half spectrum itself (no c2r_expand pass, no work array)
    bool fused_c2r = false;    // C2R of 2^14 .. 2^20 points: the fused kernel stores the real parts itself
    bool pipe_blue = false;    // Bluestein with m = 512 .. 4096: both transforms in the pipe kernel's Bluestein variants, no elementwise kernels
    bool fused_blue = false;   // Bluestein with padded length 2^14 .. 2^20: chirp / FB factors inside the two fused transforms (fft_fused.cuh BLUE)
    bool pipe_real = false;    // R2C / C2R of 512 .. 4096 points: the pipe kernel reads reals / half spectra itself (no work array)
    cd* chirp = nullptr;       // Bluestein: n entries
    cd* fb = nullptr;          // Bluestein: FFT_m of the wrapped chirp
    // host staging (exec_host)
    enum { NSTAGE = 3 };
    cd* d_ring[NSTAGE] = {nullptr, nullptr, nullptr};   // chunked host pipeline: in-place staging buffers
    int ring_batch = 0;                                  // transforms per staging buffer
    cudaEvent_t ev_up[NSTAGE] = {}, ev_run[NSTAGE] = {}, ev_down[NSTAGE] = {};
    char* h_zc = nullptr;      // small host transforms: page-locked staging the kernels read and write in place over PCIe (zero copy)
    char* d_zc = nullptr;      // its device address
    size_t zc_half = 0;        // bytes per direction
    cudaStream_t stream = nullptr, s_up = nullptr, s_down = nullptr;
    bool owns_stream = true;   // false after fftb200_plan_set_stream (plans chained on one stream, e.g. the two halves of a 2-D transform)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int launches = 0;
    std::string desc;
};

static const KernelInfo* find_kernel(int mode, int logp, int triv) {
    int cnt = 0;
    const KernelInfo* t = mode == MODE_CONTIG ? kernels_contig(&cnt) : mode == MODE_STRIDED ? kernels_strided(&cnt) : kernels_last(&cnt);
    for (int i = 0; i < cnt; i++)
        if (t[i].logp == logp && t[i].triv == triv) return &t[i];
    return nullptr;
}

static int ilog2(long long n) { int l = 0; while (n > 1) { n >>= 1; l++; } return l; }

// Split log2 N into per-pass sizes (each 6..9 bits, larger first); <= 13 bits is a single CONTIG pass.
static std::vector<int> split_passes(int log_n) {
    std::vector<int> v;
    if (log_n <= 13) { v.push_back(log_n); return v; }
    const char* force = getenv("FFTB200_SPLIT");  // e.g. "8,8" for tuning
    if (force) {
        int sum = 0;
        std::vector<int> f;
        for (const char* p = force; *p;) {
            int x = (int)strtol(p, (char**)&p, 10);
            if (x > 0) { f.push_back(x); sum += x; }
            if (*p == ',') p++;
        }
        if (sum == log_n) return f;
    }
    int np = (log_n + 8) / 9;
    if (np < 2) np = 2;
    int base = log_n / np, extra = log_n % np;
    for (int i = 0; i < np; i++) v.push_back(base + (i < extra ? 1 : 0));
    return v;
}

static int build_passes(fftb200_plan* p, DeviceState* ds) {
    const int L = p->log_n;
    if (L == 0) { p->desc += "identity (one point)"; return 0; }   // no passes: enqueue_c2c copies (radix2_dit.c:59-120 leaves n = 1 untouched)
    std::vector<int> sizes = split_passes(L);
    const int np = (int)sizes.size();
    int log_m = 0;
    const int persistent = getenv("FFTB200_PERSISTENT") ? atoi(getenv("FFTB200_PERSISTENT")) : 1;
    if (np == 1 && L >= 9 && L <= 12 && p->acc && !getenv("FFTB200_NO_PIPE")) {
        Pass ps;
        ps.k = nullptr; ps.log_p = L; ps.log_m = 0;
        ps.nt = PIPE_TILE >> L; ps.shift = 0;
        ps.src = BUF_IN; ps.dst = BUF_OUT; ps.final_pass = 1;
        for (int iv = 0; iv < 2; iv++)
            CU(cudaFuncSetAttribute(pipe_func(L, iv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
        ps.grid_max = ds->sms;
        p->passes.push_back(ps);
        char b[64];
        snprintf(b, sizeof(b), "P%d(tma ring %d x 64KB, 2x256 thr)", L, PIPE_STAGES);
        p->desc += b;
        return 0;
    }
    // (real transforms of 8192 points go to the fused kernel below: it reads reals / half spectra itself - one launch instead of
    // promote / extend + this kernel + extract)
    const bool real13 = (p->kind == FFTB200_R2C || p->kind == FFTB200_C2R) && !getenv("FFTB200_NO_FUSED") && !getenv("FFTB200_REAL13_PIPE");
    // (and so do the two transforms of a Bluestein plan with padded length 8192: the fused kernel carries the chirp / spectrum factors,
    // two launches instead of pre + this kernel + product + this kernel + post)
    const bool blue13 = p->kind == FFTB200_BLUESTEIN && !getenv("FFTB200_NO_FUSED") && !getenv("FFTB200_NO_FUSED_BLUE") && !getenv("FFTB200_BLUE13_PIPE");
    if (L == 13 && !real13 && !blue13 && p->acc && !getenv("FFTB200_NO_PIPE13") && !getenv("FFTB200_NO_PIPE")) {
        // one visit to shared memory: two 4096-point halves + stage 13 (fft_pipe13.cuh)
        Pass ps;
        ps.k = nullptr; ps.log_p = 13; ps.log_m = 0;
        ps.nt = 1; ps.shift = 0;
        ps.src = BUF_IN; ps.dst = BUF_OUT; ps.final_pass = 1;
        ps.tmem13 = getenv("FFTB200_PIPE13_TMEM") ? 1 : 0;   // measured 1.82 vs 1.77 ms per 2^28 points: opt-in only
        for (int iv = 0; iv < 2; iv++)
            CU(cudaFuncSetAttribute(ps.tmem13 ? pipe13t_func(iv) : pipe13_func(iv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM));
        ps.grid_max = ds->sms;
        p->passes.push_back(ps);
        p->desc += ps.tmem13 ? "P13t(tma ring 3 x 64KB halves, one 256-thread group per transform, radix-2 first, b and X[2k] parked in tensor memory)"
                             : "P13(tma ring 3 x 64KB halves, 2x256 thr, de-interleave on the first gather, stage 13 traded through shared memory)";
        return 0;
    }
    if (L >= 13 && L <= 20 && p->acc && !getenv("FFTB200_NO_FUSED")) {
        int lm = (L + 1) / 2, lr = L / 2;
        if (const char* e = getenv("FFTB200_FUSED_LM")) { lm = atoi(e); lr = L - lm; }
        if (fused_func(lm, lr, 0)) {
            Pass ps;
            ps.k = nullptr; ps.fused_lm = lm; ps.fused_lr = lr;
            ps.log_p = L; ps.log_m = 0; ps.nt = 0; ps.shift = 0;
            ps.src = BUF_IN; ps.dst = BUF_OUT; ps.final_pass = 1;
            for (int iv = 0; iv < 2; iv++)
                CU(cudaFuncSetAttribute(fused_func(lm, lr, iv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM));
            int occ = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_func(lm, lr, 0), FUSED_THREADS, FUSED_SMEM));
            if (occ < 1) return fail("fused kernel does not fit on an SM");
            ps.grid_max = ds->sms;   // every CTA must be resident: the passes synchronise through global counters
            p->passes.push_back(ps);
            char b[96];
            snprintf(b, sizeof(b), "Z%d+%d(fused two-pass, tma ring, L2-resident intermediate)", lm, lr);
            p->desc += b;
            return 0;
        }
    }
    // 2^21 (tail of 5 stages): c2c pays from 8 transforms per execution (same box, fused head + L5 / three tile passes: x4 0.154 / 0.155,
    // x8 0.280 / 0.290, x16 0.524 / 0.556, x64 2.00 / 2.11 ms); inside Bluestein the three-pass plan carries the chirp factor on its first
    // pass and stays ahead up to x16 (1.22 / 1.25 ms), level at x32
    const int l5_from = p->kind == FFTB200_C2C ? 8 : 32;
    // (a single 2^22-point transform is the one other case where three tile passes win: 0.090 vs 0.095 ms)
    const bool small22 = L == 22 && p->batch == 1 && !getenv("FFTB200_FORCE_L5");
    if (L >= ((getenv("FFTB200_NO_L5") || (p->batch < l5_from && !getenv("FFTB200_FORCE_L5"))) ? 22 : 21) && L <= 25 && !small22 && p->acc && !getenv("FFTB200_NO_FUSED") && !getenv("FFTB200_NO_FUSED_COLS") && !getenv("FFTB200_SPLIT")) {
        // stages 1 .. 16 in the fused kernel's column mode (intermediate in L2), then one LAST tile pass: two HBM round trips
        Pass fz;
        fz.k = nullptr; fz.fused_lm = 8; fz.fused_lr = 8; fz.fused_cols = 1;
        fz.log_p = 16; fz.log_m = 0; fz.nt = 0; fz.shift = 0;
        fz.src = BUF_IN; fz.dst = BUF_SCRATCH; fz.final_pass = 0;
        for (int iv = 0; iv < 2; iv++)
            CU(cudaFuncSetAttribute(fused_cols_func(iv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM));
        fz.grid_max = ds->sms;
        Pass lz;
        lz.log_p = L - 16; lz.log_m = 16;
        lz.k = find_kernel(MODE_LAST, lz.log_p, 0);
        if (!lz.k) return fail("no last-pass kernel for 2^%d points", lz.log_p);
        lz.nt = 0; lz.shift = 16 - lz.k->logc;
        lz.src = BUF_SCRATCH; lz.dst = BUF_OUT; lz.final_pass = 1;
        CU(cudaFuncSetAttribute(lz.k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz.k->smem));
        CU(cudaFuncSetAttribute(lz.k->func_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lz.k->smem));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lz.k->func, lz.k->threads, lz.k->smem));
        if (occ < 1) return fail("kernel variant does not fit on an SM");
        lz.grid_max = persistent ? ds->sms * occ : 0x7fffffff;
        p->passes.push_back(fz);
        p->passes.push_back(lz);
        char b[96];
        snprintf(b, sizeof(b), "Zc8+8(fused, column mode)+L%d(occ%d)", lz.log_p, occ);
        p->desc += b;
        return 0;
    }
    for (int i = 0; i < np; i++) {
        Pass ps;
        const int lp = sizes[i];
        ps.log_p = lp;
        int mode;
        if (np == 1) mode = MODE_CONTIG;
        else if (i == np - 1) mode = MODE_LAST;
        else mode = MODE_STRIDED;
        ps.k = find_kernel(mode, lp, mode == MODE_CONTIG ? 1 : (mode == MODE_STRIDED ? (i == 0) : 0));
        if (!ps.k) return fail("no kernel variant for mode %d, 2^%d points", mode, lp);
        ps.log_m = log_m;
        ps.nt = 0; ps.shift = 0;
        if (mode == MODE_CONTIG) {
            ps.nt = ps.k->nt;
        } else if (mode == MODE_STRIDED) {
            const int log_rest = L - log_m - lp;
            if (log_rest < ps.k->logc) return fail("pass split leaves too few columns");
            ps.shift = log_rest - ps.k->logc + log_m;
        } else {
            if (log_m < ps.k->logc) return fail("last pass too wide");
            ps.shift = log_m - ps.k->logc;
        }
        // ping-pong so that the final pass lands in OUT; the first pass always reads IN
        ps.src = (i == 0) ? BUF_IN : (((np - 1 - (i - 1)) % 2 == 0) ? BUF_OUT : BUF_SCRATCH);
        ps.dst = ((np - 1 - i) % 2 == 0) ? BUF_OUT : BUF_SCRATCH;
        ps.final_pass = (i == np - 1);
        CU(cudaFuncSetAttribute(ps.k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem));
        CU(cudaFuncSetAttribute(ps.k->func_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem));
        int occ = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.k->func, ps.k->threads, ps.k->smem));
        if (occ < 1) return fail("kernel variant does not fit on an SM");
        ps.grid_max = persistent ? ds->sms * occ : 0x7fffffff;
        p->passes.push_back(ps);
        char b[64];
        snprintf(b, sizeof(b), "%s%c%d(occ%d)", i ? "+" : "", mode == MODE_CONTIG ? 'C' : mode == MODE_STRIDED ? (i == 0 ? 'F' : 'M') : 'L', lp, occ);
        p->desc += b;
        log_m += lp;
    }
    return 0;
}

// The last pass of a multi-pass plan also exists as a persistent TMA-ring kernel (fft_lastpipe.cuh); the tile kernel stays for the
// executions where something rides on the pass (Bluestein factors, peer stores of the distributed transform).
static int mark_lastpipe(fftb200_plan* p, DeviceState* ds) {
    if (p->passes.size() < 2 || getenv("FFTB200_NO_LASTPIPE")) return 0;
    Pass& ps = p->passes.back();
    if (!ps.k || ps.k->mode != MODE_LAST || !ps.final_pass) return 0;
    // same-box A/B at 2^28 points (ms, ring / tile kernel): 2^21 4.03 / 4.04, 2^22 3.89 / 3.93, 2^23 3.96 / 4.30, 2^24 4.18 / 4.32, 2^25 4.82 / 4.61
    if (ps.log_p < 6 || ps.log_p > 8 || !lastpipe_func(ps.log_p, 0)) return 0;
    for (int iv = 0; iv < 2; iv++)
        for (int dv = 0; dv < 2; dv++)
            CU(cudaFuncSetAttribute(lastpipe_func(ps.log_p, iv, dv), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LASTPIPE_SMEM));
    ps.lastpipe = ds->sms;
    p->desc += "[tma ring]";
    return 0;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)f;
        else
            cudaGetLastError();
    });
    return fn;
}

// Tile hand-out counters of the persistent kernels (fft_pipe.cuh "Tile order"): allocated and zeroed once, the kernels leave them at zero.
// One set per plan: the kernels of a plan run one after the other on the plan's stream.
static unsigned int* sched_counters(fftb200_plan* p) {
    if (!p->sched) {
        p->sched = (unsigned int*)fftb200_malloc(sizeof(unsigned int) * 8);
        if (!p->sched) return nullptr;   // (fftb200_malloc has recorded the error)
        // (on the plan's stream - it is a non-blocking one, the legacy stream's memset would not be ordered with the kernels)
        if (cudaMemsetAsync(p->sched, 0, sizeof(unsigned int) * 8, p->stream) != cudaSuccess) {
            cudaFree(p->sched); p->sched = nullptr; cudaGetLastError();
            fail("cudaMemset failed for the tile counters");
            return nullptr;
        }
    }
    return p->sched;
}

// Enqueue the fused two-pass kernel (fft_fused.cuh) for `nbatch` transforms.
// blue = FUSED_BLUE_FWD: `in` is the caller's array (n elements per transform), `out` the work array; FUSED_BLUE_INV: `in` is the work
// array and `out` the caller's array.
static int enqueue_fused(fftb200_plan* p, const Pass& ps, const cd* in, cd* out, int inverse, long long nbatch_in, int r2c = 0, int blue = 0) {
    // column mode: a "transform" of the schedule is one 16-column block (2^20 points) of stages 1 .. 16
    const int cols = ps.fused_cols, log_rw = p->log_n - 16, log_cb = cols ? log_rw - 4 : 0;
    const int L = cols ? 20 : p->log_n, lm = ps.fused_lm, lr = ps.fused_lr;
    const long long nbatch = cols ? nbatch_in << log_cb : nbatch_in;
    const long long tpt = 1LL << (L - 12);                  // 64 KB tiles per transform and pass
    long long gt = tpt >= 32 ? 1 : 32 / tpt;                // transforms per group: >= 32 tiles (2 MB)
    if (const char* e = getenv("FFTB200_FUSED_GT")) gt = atol(e);
    if (gt > nbatch) gt = nbatch;
    if (gt < 1) gt = 1;
    const long long T = gt * tpt;
    long long lag = (3 * ps.grid_max + T - 1) / T;           // ~3 tiles per CTA between the end of A(g) and B(g)
    if (const char* e = getenv("FFTB200_FUSED_LAG")) lag = atol(e);
    if (lag < 1) lag = 1;
    // scratch ring: lag + 3 groups, but not more than 36 MB (about what stays resident in L2 next to the streaming traffic;
    // measured: 2^20 with 2 slots of 16 MB 2.90 ms, 3 slots 3.11 ms, 4 slots 3.27 ms)
    long long slots = lag + 3;
    while (slots > 2 && slots * T * PIPE_TILE * (long long)sizeof(cd) > (36LL << 20)) slots--;
    if (lag > slots - 1 && !getenv("FFTB200_FUSED_LAG")) lag = slots - 1;
    if (const char* e = getenv("FFTB200_FUSED_SLOTS")) slots = atol(e);
    if (slots < lag + 1) slots = lag + 1;
    const long long G = (nbatch + gt - 1) / gt;
    if (slots > G) slots = G > lag + 1 ? G : lag + 1;
    const size_t need = (size_t)(slots * gt) << L;
    if (p->fscratch_elems < need) {
        if (p->fscratch) { CU(cudaStreamSynchronize(p->stream)); cudaFree(p->fscratch); p->fscratch = nullptr; p->fscratch_elems = 0; }
        p->fscratch = (cd*)fftb200_malloc(sizeof(cd) * need);
        if (!p->fscratch) return -1;
        p->fscratch_elems = need;
        // r2c (Hermitian schedule): the last pass-B tile of a transform reads rows above k = M/2 that pass A never stores (their results are
        // dropped); c2r from the half spectrum: pass-B tiles load whole rows of which only the columns c <= R/2 (+ one tile) are ever written
        CU(cudaMemsetAsync(p->fscratch, 0, sizeof(cd) * need, p->stream));
    }
    const size_t nflags = (size_t)(2 * G + 1);
    if (p->fflags_count < nflags) {
        if (p->fflags) { CU(cudaStreamSynchronize(p->stream)); cudaFree(p->fflags); p->fflags = nullptr; p->fflags_count = 0; }
        p->fflags = (int*)fftb200_malloc(sizeof(int) * nflags);
        if (!p->fflags) return -1;
        p->fflags_count = nflags;
    }
    CU(cudaMemsetAsync(p->fflags, 0, sizeof(int) * nflags, p->stream));
    // tensor maps (complex double = 2 doubles): the input as a row-major [nbatch * M][R] array, a quarter of a pass-A tile is the
    // box C x M/4; the scratch ring the same way over slots * gt transforms; the output as [nbatch * R][M], box C2 x R/4
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
    CUtensorMap tm[4];
    memset(tm, 0, sizeof(tm));
    int promo = 12 - lm >= 4 ? 2 : 12 - lm >= 3 ? 1 : 0;      // rows of 256 / 128 / 64 bytes
    if (const char* e = getenv("FFTB200_FUSED_PROMO")) promo = atoi(e);
    const CUtensorMapL2promotion pr = promo >= 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                                                                   : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    if (r2c == 1) {
        // real input rows of R doubles; output bins 0 .. n/2 - 1 as [b][q < R/2][k] with n/2 + 1 elements per transform
        const cuuint64_t gdim[2] = {(cuuint64_t)1 << lr, (cuuint64_t)nbatch << lm};
        const cuuint64_t gstr[1] = {(cuuint64_t)sizeof(double) << lr};
        // a pass-A tile holds 2C real columns (= C complex columns of adjacent pairs, fft_fused.cuh PACK): a quarter is the box 2C x M/4
        const cuuint32_t box[2] = {(cuuint32_t)(FUSED_R2C_PACK ? 2 : 1) << (12 - lm), (cuuint32_t)1 << (lm - 2)};
        const cuuint32_t estr[3] = {1, 1, 1};
        CUresult r = enc(&tm[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, FUSED_R2C_PACK ? pr : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the real input", (int)r);
        const cuuint64_t odim[3] = {(cuuint64_t)2 << lm, (cuuint64_t)1 << (lr - 1), (cuuint64_t)nbatch};
        const cuuint64_t ostr[2] = {(cuuint64_t)sizeof(cd) << lm, (cuuint64_t)sizeof(cd) * (((cuuint64_t)1 << (L - 1)) + 1)};
        const cuuint32_t obox[3] = {(cuuint32_t)2 << (12 - lr), (cuuint32_t)1 << (lr - 2), 1};
        r = enc(&tm[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)out, odim, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the half spectrum", (int)r);
        // the same array cut off after column M/2: the last pass-B tile of a real transform holds one valid column (fft_fused.cuh, R2C)
        const cuuint64_t odim2[3] = {((cuuint64_t)1 << lm) + 2, (cuuint64_t)1 << (lr - 1), (cuuint64_t)nbatch};
        r = enc(&tm[3], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)out, odim2, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the cut-off half spectrum", (int)r);
    }
    const bool c2r_half = r2c == 2 && p->c2r_half;
    if (c2r_half) {
        // half spectra, N/2 + 1 bins per transform: rows t < M/2 of R bins each, a quarter of a pass-A tile is the box C x M/4 (fft_fused.cuh, C2R + HERM)
        const cuuint64_t hdim[3] = {(cuuint64_t)2 << lr, (cuuint64_t)1 << (lm - 1), (cuuint64_t)nbatch};
        const cuuint64_t hstr[2] = {(cuuint64_t)sizeof(cd) << lr, (cuuint64_t)sizeof(cd) * (((cuuint64_t)1 << (L - 1)) + 1)};
        const cuuint32_t hbox[3] = {(cuuint32_t)2 << (12 - lm), (cuuint32_t)1 << (lm - 2), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&tm[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)in, hdim, hstr, hbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the half-spectrum input", (int)r);
    }
    if (blue == FUSED_BLUE_FWD) {
        // the caller's rows of n elements as [transform][t < n / R][R]: a quarter of a pass-A tile is the box C x M/4; rows from n / R on are
        // out of range and arrive as zeros (the padding), the partial row n / R is read by the kernel itself
        const cuuint64_t hdim[3] = {(cuuint64_t)2 << lr, (cuuint64_t)(p->n >> lr), (cuuint64_t)nbatch};
        const cuuint64_t hstr[2] = {(cuuint64_t)sizeof(cd) << lr, (cuuint64_t)sizeof(cd) * (cuuint64_t)p->n};
        const cuuint32_t hbox[3] = {(cuuint32_t)2 << (12 - lm), (cuuint32_t)1 << (lm - 2), 1};
        const cuuint32_t estr[3] = {1, 1, 1};
        const CUresult r = enc(&tm[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)in, hdim, hstr, hbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, pr, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the Bluestein input", (int)r);
    }
    for (int i = 0; i < 3 && !cols; i++) {
        if (r2c == 1 && i != 1) continue;
        if (c2r_half && i == 0) continue;
        if (blue == FUSED_BLUE_FWD && i == 0) continue;
        const int lcols = i == 2 ? lm : lr, lrows = i == 2 ? lr : lm;          // row length / rows per transform (log2)
        const long long ntr = i == 1 ? slots * gt : nbatch;
        // (FUSED_BLUE_INV stores from registers: its output map is never used and points at the work array)
        void* base = i == 0 ? (void*)in : i == 1 ? (void*)p->fscratch : blue == FUSED_BLUE_INV ? (void*)p->work : (void*)out;
        const cuuint64_t gdim[2] = {(cuuint64_t)2 << lcols, (cuuint64_t)ntr << lrows};
        const cuuint64_t gstr[1] = {(cuuint64_t)sizeof(cd) << lcols};
        cuuint32_t box[2] = {(cuuint32_t)2 << (12 - lrows), (cuuint32_t)1 << (lrows - 2)};   // a quarter tile: 1024 elements
        if (r2c == 1 && i == 1 && FUSED_R2C_PACK) { box[0] *= 2; box[1] /= 2; }   // r2c: the unpacked rows k < M/2 of 2C columns, M/8 rows per quarter
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, i == 0 ? pr : CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for map %d", (int)r, i);
    }
    if (r2c == 2) {
        // c2r: the output as real rows [nbatch * R][M doubles], a quarter of a pass-B tile is the box C2 x R/4
        const cuuint64_t odim[2] = {(cuuint64_t)1 << lm, (cuuint64_t)nbatch << lr};
        const cuuint64_t ostr[1] = {(cuuint64_t)sizeof(double) << lm};
        const cuuint32_t obox[2] = {(cuuint32_t)1 << (12 - lr), (cuuint32_t)1 << (lr - 2)};
        const cuuint32_t estr[2] = {1, 1};
        const CUresult r = enc(&tm[2], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)out, odim, ostr, obox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for the real output", (int)r);
    }
    for (int i = 0; i < 3 && cols; i++) {
        // input / output: [b][t_hi or q][t_lo or k_hi][c] with rows of 2^log_rw columns; scratch ring: [slot][k_hi][t_lo][c16]
        const bool sc = i == 1;
        void* base = i == 0 ? (void*)in : sc ? (void*)p->fscratch : (void*)out;
        const cuuint64_t rowb = sc ? 16 * sizeof(cd) : (cuuint64_t)sizeof(cd) << log_rw;
        const cuuint64_t gdim[4] = {sc ? 32u : (cuuint64_t)2 << log_rw, 256, 256, (cuuint64_t)(sc ? slots * gt : nbatch_in)};
        const cuuint64_t gstr[3] = {rowb, 256 * rowb, 65536 * rowb};
        const cuuint32_t box[4] = {32, 1, 64, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        const CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 4, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (%d) for column-mode map %d", (int)r, i);
    }
    FusedArgs fa;
    fa.scratch = p->fscratch; fa.tab = p->tab; fa.acc = p->acc; fa.flags = p->fflags;
    fa.handout = nullptr;
#if FUSED_DYNAMIC
    {
        unsigned int* sc = sched_counters(p);
        if (!sc) return -1;
        fa.handout = reinterpret_cast<int*>(sc + 4);
        CU(cudaMemsetAsync(fa.handout, 0, sizeof(int), p->stream));
    }
#endif
    fa.nbatch = nbatch; fa.gt = (int)gt; fa.ngroups = (int)G; fa.lag = (int)lag; fa.slots = (int)slots;
    fa.inverse = inverse; fa.scale = cols ? 1.0 : p->scale; fa.log_cb = log_cb; fa.out = out; fa.half_in = in;
    fa.chirp = p->chirp; fa.fb = p->fb; fa.n_user = p->n; fa.y_scale = (blue && p->dir > 0) ? 1.0 / (double)p->n : 1.0;
    fa.user_in = in; fa.user_out = out;
    if (blue == FUSED_BLUE_INV) fa.out = p->work;
    fa.debug = getenv("FFTB200_FUSED_DEBUG") ? atoi(getenv("FFTB200_FUSED_DEBUG")) : 0;
    memcpy(fa.dtw, p->fdtw, sizeof(fa.dtw));
    fa.prof = nullptr;
#ifdef FUSED_PROF
    static long long* d_prof = nullptr;
    if (!d_prof) d_prof = (long long*)fftb200_malloc(sizeof(long long) * 16 * 1024);
    fa.prof = d_prof;
#endif
    if (r2c != 1) tm[3] = tm[2];
    const bool herm = r2c == 1 && p->r2c_herm;
    // packed real input: half the pass-A tiles; c2r from the half spectrum: the columns c <= R/2 only
    const long long tpa = (r2c == 1 && FUSED_R2C_PACK) ? tpt / 2 : (c2r_half && FUSED_C2R_HALFCOLS) ? tpt / 2 + 1 : tpt;
    const long long items = nbatch * (tpa + (herm ? tpt / 2 + 1 : tpt));   // Hermitian schedule: pass B on the columns k <= M/2 only
    const int grid = (int)(items < ps.grid_max ? items : ps.grid_max);
    const void* func = blue ? fused_blue_func(lm, lr, blue) : r2c == 1 ? fused_r2c_func(lm, lr, herm) : r2c == 2 ? fused_c2r_func(lm, lr, c2r_half)
                     : cols ? fused_cols_func(inverse) : fused_func(lm, lr, inverse);
    if (!func) return fail("no fused kernel for 2^%d x 2^%d", lm, lr);
    CU(launch_fused(func, fa, tm, grid, p->stream));
#ifdef FUSED_PROF
    if (getenv("FFTB200_FUSED_PROF_PRINT")) {
        CU(cudaStreamSynchronize(p->stream));
        std::vector<long long> h(8 * grid);
        CU(cudaMemcpy(h.data(), fa.prof, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost));
        double e = 0, f = 0, tot = 0, tiles = 0;
        for (int i = 0; i < 2 * grid; i++) { e += h[i * 4]; f += h[i * 4 + 1]; tot += h[i * 4 + 2]; tiles += h[i * 4 + 3]; }
        {
            std::vector<long long> hm(8 * 3 * grid);
            CU(cudaMemcpy(hm.data(), fa.prof + 8 * 1024, sizeof(long long) * 8 * 3 * grid, cudaMemcpyDeviceToHost));
            double v[5] = {0, 0, 0, 0, 0};
            for (int i = 0; i < 3 * grid; i++) for (int j = 0; j < 5; j++) v[j] += hm[i * 8 + j];
            fprintf(stderr, "fused prof managers: per tile cycles: load latency %.0f, loaded->staged %.0f, store read-out + next load issue %.0f, publish %.0f\n",
                    v[0] / v[4], v[1] / v[4], v[2] / v[4], v[3] / v[4]);
        }
        fprintf(stderr, "fused prof: per group-tile cycles: total %.0f, staged-wait %.0f, full-wait %.0f, busy %.0f (tiles/group %.1f)\n",
                tot / tiles, e / tiles, f / tiles, (tot - e - f) / tiles, tiles / (2 * grid));
        {
            std::vector<long long> hb(8 * grid);
            CU(cudaMemcpy(hb.data(), fa.prof + 12 * 1024, sizeof(long long) * 8 * grid, cudaMemcpyDeviceToHost));
            double ba = 0, ca = 0, bb = 0, cb = 0;
            for (int i = 0; i < 2 * grid; i++) { ba += hb[i * 4]; ca += hb[i * 4 + 1]; bb += hb[i * 4 + 2]; cb += hb[i * 4 + 3]; }
            fprintf(stderr, "fused prof: busy cycles per tile: pass A %.0f (%.0f tiles), pass B %.0f (%.0f tiles)\n", ca ? ba / ca : 0, ca, cb ? bb / cb : 0, cb);
        }
    }
#endif
    return 0;
}

static int ensure_scratch(fftb200_plan* p, long long nbatch) {
    if (p->passes.size() < 2) return 0;
    const size_t need = (size_t)nbatch << p->log_n;
    if (p->scratch_elems >= need) return 0;
    if (p->scratch) { CU(cudaStreamSynchronize(p->stream)); cudaFree(p->scratch); p->scratch = nullptr; p->scratch_elems = 0; }
    p->scratch = (cd*)fftb200_malloc(sizeof(cd) * need);
    if (!p->scratch) return -1;
    p->scratch_elems = need;
    return 0;
}

// Enqueue the power-of-two c2c passes: `inverse` selects conjugated twiddles and the 1/m scale.
// Elementwise factors fused into the first / last tile pass of a multi-pass plan (Bluestein, fft_tile.cuh MUL_*)
struct FuseSpec {
    int mode = MUL_NONE;       // MUL_PRE: `user` is the caller's input; MUL_FB; MUL_POST: `user` is the caller's output
    const cd* mul = nullptr;
    const cd* user = nullptr;
    int n = 0;
    double scale = 1.0;
};
static bool can_fuse_pre(const fftb200_plan* p) { return p->passes.size() >= 2 && p->passes.front().k && !getenv("FFTB200_NO_FUSED_CHIRP"); }
static bool can_fuse_post(const fftb200_plan* p) { return p->passes.size() >= 2 && p->passes.back().k && !p->peers && !getenv("FFTB200_NO_FUSED_CHIRP"); }

static int enqueue_c2c(fftb200_plan* p, const cd* in, cd* out, int inverse, long long nbatch, const FuseSpec* pre = nullptr,
                       const FuseSpec* post = nullptr) {
    if (nbatch <= 0) return 0;
    if (p->passes.empty()) {   // one-point transforms: X[0] = x[0] in both directions (1/n = 1)
        if (in != out) CU(cudaMemcpyAsync(out, in, sizeof(cd) * (size_t)nbatch, cudaMemcpyDeviceToDevice, p->stream));
        return 0;
    }
    if (ensure_scratch(p, nbatch) != 0) return -1;
    for (const Pass& ps : p->passes) {
        const long long ntiles = pass_tiles(ps, nbatch);
        const int grid = (int)(ntiles < ps.grid_max ? ntiles : ps.grid_max);
        if (ps.fused_lm) {
            const cd* fsrc = ps.src == BUF_IN ? in : ps.src == BUF_OUT ? out : p->scratch;
            cd* fdst = ps.dst == BUF_OUT ? out : p->scratch;
            if (enqueue_fused(p, ps, fsrc, fdst, inverse, nbatch) != 0) return -1;
            continue;
        }
        if (!ps.k) {
            PipeArgs pa;
            pa.in = in; pa.out = out; pa.tab = p->acc;
            pa.ntiles = ntiles; pa.batch = nbatch;
            pa.inverse = inverse; pa.scale = p->scale;
            pa.sched = sched_counters(p);
            if (!pa.sched) return -1;
            if (ps.log_p == 13) {
                CU(ps.tmem13 ? launch_pipe13t(pa, grid, p->stream) : launch_pipe13(pa, grid, p->stream));
                continue;
            }
            CU(launch_pipe(ps.log_p, pa, grid, p->stream));
            continue;
        }
        const cd* src = ps.src == BUF_IN ? in : ps.src == BUF_OUT ? out : p->scratch;
        cd* dst = ps.dst == BUF_OUT ? out : p->scratch;
        // With every twiddle from the table the ring kernel pays from 4 transforms per execution (below that the late-stage twiddles come from
        // HBM for most tiles and the tile kernel's two CTAs per SM hide that better: 2^24 x1 0.327 / 0.312 ms ring / tile). With derived twiddles
        // it pays from one transform for P = 128, 256 (2^24 x1 0.295 / 0.308, x2 0.538 / 0.566, x3 0.783 / 0.830; 2^23 x1 0.164 / 0.166,
        // x3 0.402 / 0.412) and from 8 for P = 64 (2^22 x2 0.161 / 0.156, x3 0.220 / 0.216).
        const long long ring_from = getenv("FFTB200_LASTPIPE_MIN") ? atoi(getenv("FFTB200_LASTPIPE_MIN")) : p->last_derive ? (ps.log_p <= 6 ? 8 : 1) : 4;
        if (ps.lastpipe && nbatch >= ring_from && !post && !p->peers) {
            LastPipeArgs la;
            la.in = src; la.out = dst; la.tab = p->tab;
            la.batch = nbatch; la.log_n = p->log_n; la.log_m = ps.log_m;
            la.ntiles = nbatch << (ps.log_m - (12 - ps.log_p));
            la.inverse = inverse; la.scale = p->scale;
            la.sched = sched_counters(p);
            if (!la.sched) return -1;
            la.derive = p->last_derive ? 1 : 0;
            memcpy(la.dtw, p->ldtw, sizeof(la.dtw));
            CU(launch_lastpipe(ps.log_p, la, (int)(la.ntiles < ps.lastpipe ? la.ntiles : ps.lastpipe), p->stream));
            continue;
        }
        TileArgs a;
        a.in = src; a.out = dst; a.tab = p->tab;
        a.ntiles = ntiles; a.batch = nbatch;
        a.log_n = p->log_n; a.log_m = ps.log_m;
        a.inverse = inverse; a.scale = p->scale; a.final_pass = ps.final_pass;
        a.peers = (ps.final_pass && p->peers) ? p->peers : nullptr;
        a.peer_lw = p->peer_lw; a.peer_lrows = p->peer_lrows; a.peer_lg = p->peer_lg; a.peer_me = p->peer_me;
        a.mul = nullptr; a.mul_mode = MUL_NONE; a.mul_n = 0; a.mul_scale = 1.0;
        if (pre && &ps == &p->passes.front()) {
            a.in = pre->user; a.mul = pre->mul; a.mul_mode = MUL_PRE; a.mul_n = pre->n;
        }
        if (post && &ps == &p->passes.back()) {
            a.mul = post->mul; a.mul_mode = post->mode; a.mul_n = post->n; a.mul_scale = post->scale;
            if (post->mode == MUL_POST) a.out = const_cast<cd*>(post->user);
        }
        ps.k->launch(a, grid, p->stream);
    }
    CU(cudaGetLastError());
    return 0;
}

static int upload_table(fftb200_plan* p, DeviceState* ds, const fftb200_plan_desc* d, int need_n) {
    if (need_n <= 1) { p->tab = nullptr; return 0; }
    std::lock_guard<std::mutex> lk(g_mu);
    if (ds->tab && ds->tab->n >= need_n) { p->tab_ref = ds->tab; ds->tab->refs++; p->tab = ds->tab->ptr; return 0; }
    if (!d->twiddles || d->table_n < need_n) return fail("plan needs a twiddle table for size %d", need_n);
    cd* t = nullptr;
    size_t bytes = sizeof(cd) * (size_t)(d->table_n - 1);
    CU(cudaMalloc(&t, bytes ? bytes : 16));
    if (cudaMemcpy(t, d->twiddles, bytes, cudaMemcpyHostToDevice) != cudaSuccess || cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) {
        cudaFree(t);
        return fail("twiddle table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    table_release_locked(ds->tab);   // plans created earlier keep their own reference to the smaller table
    ds->tab = new Table();
    ds->tab->ptr = t; ds->tab->n = d->table_n; ds->tab->refs = 2;   // the device cache + this plan
    g_tables_alive++;
    p->tab_ref = ds->tab;
    p->tab = t;
    return 0;
}

static int upload_accurate(fftb200_plan* p, DeviceState* ds, const fftb200_plan_desc* d) {
    p->acc = nullptr;
    if (!d->twiddles_accurate || d->accurate_n < 2) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!ds->acc) {
        if (d->accurate_n != ACC_N) return fail("accurate twiddle table must cover n = %d", (int)ACC_N);
        cd* t = nullptr;
        const size_t bytes = sizeof(cd) * (size_t)(ACC_N - 1);
        CU(cudaMalloc(&t, bytes));
        if (cudaMemcpy(t, d->twiddles_accurate, bytes, cudaMemcpyHostToDevice) != cudaSuccess || cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) {
            cudaFree(t);
            return fail("accurate table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
        }
        ds->acc = new Table();
        ds->acc->ptr = t; ds->acc->n = ACC_N;
        g_tables_alive++;
    }
    ds->acc->refs++;
    p->acc_ref = ds->acc;
    p->acc = ds->acc->ptr;
    return 0;
}

extern "C" int fftb200_plan_create(fftb200_plan** out, const fftb200_plan_desc* d) {
    if (!out || !d) return fail("plan_create: null argument");
    *out = nullptr;
    if (d->n <= 0 || d->batch <= 0) return fail("plan_create: n and batch must be positive");
    if (d->direction != -1 && d->direction != 1) return fail("plan_create: direction must be -1 or +1");
    DeviceState* ds;
    if (cur_device(&ds) != 0) return -1;
    fftb200_plan* p = new fftb200_plan();
    p->device = fftb200_get_device();
    p->n = d->n; p->batch = d->batch; p->dir = d->direction; p->kind = d->kind;
    const bool pow2 = (d->n & (d->n - 1)) == 0;
    int rc = 0;
    do {
        if (d->kind == FFTB200_C2C || d->kind == FFTB200_R2C || d->kind == FFTB200_C2R) {
            if (d->kind == FFTB200_C2R && d->direction != 1) { rc = fail("plan_create: a c2r plan is an inverse transform (direction +1)"); break; }
            if (!pow2 && d->kind == FFTB200_R2C && d->direction == -1) {
                // real input of any length (the reference plans it: fft_auto.c:391-403 promotes and routes non-powers of two to
                // Bluestein, :136-172): promote on the device, Bluestein c2c of length n, bins 0 .. n/2
                if (!d->chirp) { rc = fail("plan_create: r2c of a non-power-of-two length needs the host chirp table"); break; }
                fftb200_plan_desc cdsc = *d;
                cdsc.kind = FFTB200_BLUESTEIN;
                if (fftb200_plan_create(&p->child, &cdsc) != 0) { rc = -1; break; }
                p->work = (cd*)fftb200_malloc(sizeof(cd) * (size_t)d->n * (size_t)d->batch);
                if (!p->work) { rc = -1; break; }
                if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking) != cudaSuccess ||
                    cudaEventCreate(&p->ev0) != cudaSuccess || cudaEventCreate(&p->ev1) != cudaSuccess) { rc = fail("cudaStreamCreate failed"); break; }
                if (fftb200_plan_set_stream(p->child, p->stream) != 0) { rc = -1; break; }
                p->m = d->n; p->log_n = 0;
                p->launches = p->child->launches + 2;
                p->desc = "r2c n=" + std::to_string(d->n) + " b=" + std::to_string(d->batch) + ": promote + [" + p->child->desc + "] + bins 0 .. n/2";
                *out = p;
                return 0;
            }
            if (!pow2) { rc = fail("plan_create: kind %d needs a power-of-two n (got %d)", d->kind, d->n); break; }
            p->m = d->n;
        } else if (d->kind == FFTB200_BLUESTEIN) {
            long long m = 1;
            while (m < 2LL * d->n - 1) m <<= 1;
            if (m > (1LL << 30)) { rc = fail("plan_create: Bluestein length too large"); break; }
            p->m = (int)m;
            if (!d->chirp) { rc = fail("plan_create: Bluestein needs a host chirp table"); break; }
        } else { rc = fail("plan_create: unknown kind %d", d->kind); break; }
        p->log_n = ilog2(p->m);
        p->scale = 1.0 / (double)p->m;
        if ((rc = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking)) != 0) { rc = fail("cudaStreamCreate failed"); break; }
        if (cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking) != cudaSuccess) { rc = fail("cudaStreamCreate failed"); break; }
        if (cudaEventCreate(&p->ev0) != cudaSuccess || cudaEventCreate(&p->ev1) != cudaSuccess) { rc = fail("cudaEventCreate failed"); break; }
        if ((rc = upload_table(p, ds, d, p->m)) != 0) break;
        if ((rc = upload_accurate(p, ds, d)) != 0) break;
        char head[96];
        snprintf(head, sizeof(head), "%s n=%d b=%d dir=%d: ", d->kind == FFTB200_C2C ? "c2c" : d->kind == FFTB200_R2C ? "r2c" : d->kind == FFTB200_C2R ? "c2r" : "bluestein", d->n, d->batch, d->direction);
        p->desc = head;
        if ((rc = build_passes(p, ds)) != 0) break;
        if ((rc = mark_lastpipe(p, ds)) != 0) break;
        p->launches = (int)p->passes.size();
        if (!p->passes.empty() && p->passes[0].fused_lm) {
            // dtw[j][h] = table entry (h << a_tot_j) - 1: stage a_tot_j + s, index q << a_tot_j, h = 2^(s-1) + q
            if (!d->twiddles || d->table_n < p->m) { rc = fail("fused plan needs the host twiddle table"); break; }
            const int lm = p->passes[0].fused_lm, lr = p->passes[0].fused_lr;
            const int rb0 = lr >= 9 ? lr - 8 : lr - 4;
            const int atot[3] = {lm, lm + rb0, lm + lr - 4};   // first / middle (three sub-passes only) / last
            const int rad[3] = {rb0, 4, 4};
            const cd* ht = (const cd*)d->twiddles;
            for (int j = 0; j < 3; j++)
                for (int h = 0; h < 16; h++) {
                    p->fdtw[j][h] = make_double2(1.0, 0.0);
                    if (h >= 1 && h < (1 << rad[j]) && !(j == 1 && lr < 9)) p->fdtw[j][h] = ht[((size_t)h << atot[j]) - 1];
                }
        }
        // Derived twiddles in the last pass (fft_lastpipe.cuh `derive`): a quarter of the table traffic. The product T[kappa] * T[q << a] reproduces
        // the systematic drift of the reference's recurrence but not its rounding noise, which grows with the length of the recurrence: measured
        // against the all-table kernel 2.9e-14 at 2^22, 7.5e-14 at 2^23, 2.9e-13 at 2^24 - and past the 1e-12 bar at 2^26. Used up to 2^24.
        if (!p->passes.empty() && p->passes.back().lastpipe && p->log_n <= 24 && d->twiddles && d->table_n >= p->m && !getenv("FFTB200_LASTPIPE_TABLE")) {
            // ldtw[j][h] = table entry (h << a_j) - 1 (stage a_j + s, index q << a_j, h = 2^(s-1) + q)
            const Pass& lp = p->passes.back();
            const int lr = lp.log_p, rb0 = lr >= 9 ? lr - 8 : lr - 4;
            const int atot[3] = {lp.log_m, lp.log_m + rb0, lp.log_m + lr - 4};
            const int rad[3] = {rb0, 4, 4};
            const cd* ht = (const cd*)d->twiddles;
            for (int j = 0; j < 3; j++)
                for (int h = 0; h < 16; h++) {
                    p->ldtw[j][h] = make_double2(1.0, 0.0);
                    if (h >= 1 && h < (1 << rad[j]) && !(j == 1 && lr < 9)) p->ldtw[j][h] = ht[((size_t)h << atot[j]) - 1];
                }
            p->last_derive = true;
            p->desc += "[derived twiddles]";
        }
        if (d->kind == FFTB200_R2C) {
            const bool fused_r2c = p->passes.size() == 1 && p->passes[0].fused_lm && !getenv("FFTB200_NO_FUSED_R2C") &&
                                   fused_r2c_func(p->passes[0].fused_lm, p->passes[0].fused_lr, 0);
            if (fused_r2c) {
                // promotion and extraction happen inside the fused kernel (fft_fused.cuh, R2C). Up to R2C_HERM_MAX_LOG points pass B runs on
                // the columns k <= M/2 only and writes the other half of the bins as conjugates; above, the conjugate of the reference's
                // X[j] is no longer within 1e-12 of ITS X[N - j] (its twiddle recurrence is not conjugate-symmetric), so every column is transformed.
                p->r2c_herm = p->log_n <= R2C_HERM_MAX_LOG;
                if (const char* e = getenv("FFTB200_R2C_HERMITIAN")) p->r2c_herm = atoi(e) != 0;
                if (cudaFuncSetAttribute(fused_r2c_func(p->passes[0].fused_lm, p->passes[0].fused_lr, p->r2c_herm), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FUSED_SMEM) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
                p->desc += p->r2c_herm ? " [real in, half spectrum out, no promote / extract passes, pass B on the columns k <= M/2]"
                                       : " [real in, half spectrum out, no promote / extract passes]";
            } else if (p->passes.size() == 1 && !p->passes[0].k && p->passes[0].log_p <= 12 && !getenv("FFTB200_NO_PIPE_REAL")) {
                if (cudaFuncSetAttribute(pipe_real_func(p->passes[0].log_p, PIPE_R2C), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
                p->pipe_real = true;
                p->desc += " [real in, half spectrum out, no promote / extract passes]";
            } else {
                p->work = (cd*)fftb200_malloc(sizeof(cd) * (size_t)p->m * (size_t)p->batch);
                if (!p->work) { rc = -1; break; }
                p->launches += 2;
            }
        }
        if (d->kind == FFTB200_C2R && p->passes.size() == 1 && !p->passes[0].k && !p->passes[0].fused_lm && p->passes[0].log_p <= 12 && !getenv("FFTB200_NO_PIPE_REAL")) {
            if (cudaFuncSetAttribute(pipe_real_func(p->passes[0].log_p, PIPE_C2R), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
            p->pipe_real = true;
            p->desc += " [half spectrum in, real out, no extension / extraction passes]";
        } else
        if (d->kind == FFTB200_C2R) {
            // Hermitian extension -> inverse c2c of the full length with the reference's stage operators -> real parts
            p->fused_c2r = p->passes.size() == 1 && p->passes[0].fused_lm && !getenv("FFTB200_NO_FUSED_C2R") &&
                           fused_c2r_func(p->passes[0].fused_lm, p->passes[0].fused_lr, 0);
            // 2^14 .. 2^20: the fused kernel reads the half spectrum itself - the extension happens in its tile loads and first gather
            // (same values, same arithmetic: bit-identical to the separate c2r_expand pass, which FFTB200_C2R_HERMITIAN=0 keeps)
            p->c2r_half = p->fused_c2r && !(getenv("FFTB200_C2R_HERMITIAN") && atoi(getenv("FFTB200_C2R_HERMITIAN")) == 0);
            if (!p->c2r_half) {
                p->work = (cd*)fftb200_malloc(sizeof(cd) * (size_t)p->m * (size_t)p->batch);
                if (!p->work) { rc = -1; break; }
            }
            if (p->fused_c2r) {
                if (cudaFuncSetAttribute(fused_c2r_func(p->passes[0].fused_lm, p->passes[0].fused_lr, p->c2r_half), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)FUSED_SMEM) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
                p->launches += p->c2r_half ? 0 : 1;
                p->desc += p->c2r_half ? " [half spectrum in (extension in the tile loads), inverse c2c storing the real parts]"
                                       : " [hermitian extension + inverse c2c storing the real parts]";
            } else {
                p->launches += 2;
                p->desc += " [hermitian extension + inverse c2c + real parts]";
            }
        }
        if (d->kind == FFTB200_BLUESTEIN) {
            const size_t m = (size_t)p->m;
            p->work = (cd*)fftb200_malloc(sizeof(cd) * m * (size_t)p->batch);
            p->chirp = (cd*)fftb200_malloc(sizeof(cd) * (size_t)p->n);
            p->fb = (cd*)fftb200_malloc(sizeof(cd) * m);
            if (!p->work || !p->chirp || !p->fb) { rc = -1; break; }
            if (cudaMemcpy(p->chirp, d->chirp, sizeof(cd) * (size_t)p->n, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) { rc = fail("chirp upload failed"); break; }
            // b[k] = chirp[k], b[m-k] = chirp[k] (bluestein.c:116-121); FB = FFT_m(b), cached in the plan
            bluestein_wrap_kernel<<<(unsigned)((m + 255) / 256), 256, 0, p->stream>>>(p->fb, p->chirp, p->n, p->m);
            {   // (with every twiddle from the table: the spectrum of the chirp is computed once and multiplies every result)
                const bool keep = p->last_derive;
                p->last_derive = false;
                rc = enqueue_c2c(p, p->fb, p->fb, 0, 1);
                p->last_derive = keep;
            }
            if (rc == 0 && cudaStreamSynchronize(p->stream) != cudaSuccess) rc = fail("Bluestein kernel spectrum failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (rc != 0) break;
            p->launches = 2 * (int)p->passes.size() + (can_fuse_pre(p) ? 0 : 1) + (can_fuse_post(p) ? 0 : 2);
            if (p->passes.size() == 1 && !p->passes[0].k && !p->passes[0].fused_lm && p->passes[0].log_p <= 12 && !getenv("FFTB200_NO_PIPE_BLUE")) {
                bool ok = true;
                for (int kind = PIPE_BLUE_FWD; kind <= PIPE_BLUE_INV; kind++)
                    ok = ok && cudaFuncSetAttribute(pipe_real_func(p->passes[0].log_p, kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PIPE_SMEM) == cudaSuccess;
                if (!ok) { rc = fail("cudaFuncSetAttribute failed"); break; }
                p->pipe_blue = true;
                p->launches = 2;
                p->desc += " [chirp and spectrum factors inside the two transforms]";
            }
            // same box, ms per 2^27 padded points, five kernels / two: m = 2^13 3.56 / 2.27, 2^14 3.69 / 2.43, 2^15 3.59 / 2.45, 2^16 3.50 / 2.26,
            // 2^17 3.69 / 2.55, 2^18 3.99 / 3.11, 2^19 4.13 / 3.70, 2^20 4.38 / 3.84
            if (p->passes.size() == 1 && p->passes[0].fused_lm && !p->passes[0].fused_cols && !getenv("FFTB200_NO_FUSED_BLUE")) {
                const int lm = p->passes[0].fused_lm, lr = p->passes[0].fused_lr;
                bool ok = fused_blue_func(lm, lr, FUSED_BLUE_FWD) != nullptr;
                for (int kind = FUSED_BLUE_FWD; ok && kind <= FUSED_BLUE_INV; kind++)
                    ok = cudaFuncSetAttribute(fused_blue_func(lm, lr, kind), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM) == cudaSuccess;
                if (!ok) { rc = fail("cudaFuncSetAttribute failed for the fused Bluestein kernels"); break; }
                p->fused_blue = true;
                p->launches = 2;
                p->desc += " [chirp and spectrum factors inside the two fused transforms]";
            }
        }
    } while (0);
    if (rc != 0) { fftb200_plan_destroy(p); return -1; }
    *out = p;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// partial plans: stages [first_stage, first_stage + nstages) of the size-n Stockham transform
// ---------------------------------------------------------------------------------------------
// Building block of the distributed transform (N = R * M over G GPUs): between two all-to-all
// exchanges every GPU runs the first log2 M stages of one local array and, later, the last log2 R stages of
// another, the latter with a rank-specific table holding the reference's late-stage twiddles T[s][k + M q] for
// the k range the rank owns (host/ref_twiddle.c: fftb200_host_twiddles_dist). State before the first executed
// stage is the Stockham layout idx = c + (n / 2^first_stage) * k; the output is the layout after the last
// executed stage (natural order when the plan reaches stage log2 n).
static int split_partial(int cnt, std::vector<int>* v) {
    for (int np = 1; np <= 4; np++) {
        if (cnt < 6 * np || cnt > 9 * np) continue;
        const int base = cnt / np, extra = cnt % np;
        for (int i = 0; i < np; i++) v->push_back(base + (i < extra ? 1 : 0));
        return 0;
    }
    return -1;
}

extern "C" int fftb200_plan_create_partial(fftb200_plan** out, const fftb200_plan_desc* d, int first_stage, int nstages,
                                           int private_table, double inverse_scale) {
    if (!out || !d) return fail("plan_create_partial: null argument");
    *out = nullptr;
    if (d->n <= 0 || (d->n & (d->n - 1)) || d->batch <= 0) return fail("plan_create_partial: n must be a power of two, batch positive");
    if (d->direction != -1 && d->direction != 1) return fail("plan_create_partial: direction must be -1 or +1");
    const int L = ilog2(d->n);
    if (first_stage < 0 || nstages < 1 || first_stage + nstages > L) return fail("plan_create_partial: bad stage range");
    std::vector<int> sizes;
    if (split_partial(nstages, &sizes) != 0) return fail("plan_create_partial: %d stages cannot be split into passes of 6..9", nstages);
    DeviceState* ds;
    if (cur_device(&ds) != 0) return -1;
    fftb200_plan* p = new fftb200_plan();
    p->device = fftb200_get_device();
    p->n = d->n; p->batch = d->batch; p->dir = d->direction; p->kind = FFTB200_C2C;
    p->m = d->n; p->log_n = L; p->scale = inverse_scale;
    int rc = 0;
    do {
        if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->s_up, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&p->s_down, cudaStreamNonBlocking) != cudaSuccess) { rc = fail("cudaStreamCreate failed"); break; }
        if (cudaEventCreate(&p->ev0) != cudaSuccess || cudaEventCreate(&p->ev1) != cudaSuccess) { rc = fail("cudaEventCreate failed"); break; }
        const int need_n = 1 << (first_stage + nstages);
        if (private_table) {
            if (!d->twiddles || d->table_n < need_n) { rc = fail("plan_create_partial: private table too small"); break; }
            const size_t bytes = sizeof(cd) * (size_t)(need_n - 1);
            if (cudaMalloc(&p->own_tab, bytes) != cudaSuccess) { rc = fail("cudaMalloc(%zu) for the private table failed", bytes); cudaGetLastError(); break; }
            if (cudaMemcpy(p->own_tab, d->twiddles, bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
                cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) { rc = fail("private table upload failed"); break; }
            p->tab = p->own_tab;
        } else if ((rc = upload_table(p, ds, d, need_n)) != 0) break;
        char head[96];
        snprintf(head, sizeof(head), "partial n=%d b=%d dir=%d stages [%d,%d): ", d->n, d->batch, d->direction, first_stage, first_stage + nstages);
        p->desc = head;
        const int np = (int)sizes.size();
        int log_m = first_stage;
        for (int i = 0; i < np && rc == 0; i++) {
            Pass ps;
            const int lp = sizes[i];
            ps.log_p = lp;
            const bool last = (log_m + lp == L);
            const int mode = last ? MODE_LAST : MODE_STRIDED;
            ps.k = find_kernel(mode, lp, mode == MODE_STRIDED ? (log_m == 0) : 0);
            if (!ps.k) { rc = fail("no kernel variant for mode %d, 2^%d points", mode, lp); break; }
            ps.log_m = log_m;
            ps.nt = 0; ps.shift = 0;
            if (mode == MODE_STRIDED) {
                const int log_rest = L - log_m - lp;
                if (log_rest < ps.k->logc) { rc = fail("pass split leaves too few columns"); break; }
                ps.shift = log_rest - ps.k->logc + log_m;
            } else {
                if (log_m < ps.k->logc) { rc = fail("last pass too wide"); break; }
                ps.shift = log_m - ps.k->logc;
            }
            ps.src = (i == 0) ? BUF_IN : (((np - 1 - (i - 1)) % 2 == 0) ? BUF_OUT : BUF_SCRATCH);
            ps.dst = ((np - 1 - i) % 2 == 0) ? BUF_OUT : BUF_SCRATCH;
            ps.final_pass = (i == np - 1);
            if (cudaFuncSetAttribute(ps.k->func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem) != cudaSuccess ||
                cudaFuncSetAttribute(ps.k->func_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps.k->smem) != cudaSuccess) { rc = fail("cudaFuncSetAttribute failed"); break; }
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ps.k->func, ps.k->threads, ps.k->smem) != cudaSuccess || occ < 1) { rc = fail("kernel variant does not fit on an SM"); break; }
            ps.grid_max = ds->sms * occ;
            p->passes.push_back(ps);
            char b[64];
            snprintf(b, sizeof(b), "%s%c%d(occ%d)", i ? "+" : "", mode == MODE_LAST ? 'L' : (log_m == 0 ? 'F' : 'M'), lp, occ);
            p->desc += b;
            log_m += lp;
        }
        p->launches = (int)p->passes.size();
    } while (0);
    if (rc != 0) { fftb200_plan_destroy(p); return -1; }
    *out = p;
    return 0;
}

// dst[b][a][c] = src[a][b][c] over complex elements, c contiguous: the local half of the block transposes around
// the all-to-all exchanges of the distributed transform
__global__ void permute_bac_kernel(cd* __restrict__ dst, const cd* __restrict__ src, long long A, long long B, long long C) {
    const long long total = A * B * C, stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long c = i % C, ab = i / C, a = ab % A, b = ab / A;   // i indexes dst = [b][a][c]
        dst[i] = src[(a * B + b) * C + c];
    }
}
extern "C" int fftb200_permute_bac(void* dst, const void* src, long long A, long long B, long long C, void* stream) {
    if (!dst || !src || A < 1 || B < 1 || C < 1 || dst == src) return fail("permute_bac: bad argument");
    const long long total = A * B * C;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    permute_bac_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((cd*)dst, (const cd*)src, A, B, C);
    CU(cudaGetLastError());
    return 0;
}
extern "C" void* fftb200_plan_stream(fftb200_plan* p) { return p ? (void*)p->stream : nullptr; }

// Peer tables: device arrays of the 2^log_world base pointers of one exchange buffer on every rank (own buffer +
// IPC-opened peers), as seen from THIS process.
struct fftb200_peers {
    cd** d_bases = nullptr;
    int log_world = 0, rank = 0;
};
extern "C" int fftb200_peers_create(fftb200_peers** out, void* const* bases, int log_world, int rank) {
    if (!out || !bases || log_world < 0 || log_world > 6 || rank < 0 || rank >= (1 << log_world)) return fail("peers_create: bad argument");
    fftb200_peers* t = new fftb200_peers();
    t->log_world = log_world; t->rank = rank;
    if (cudaMalloc(&t->d_bases, sizeof(cd*) << log_world) != cudaSuccess ||
        cudaMemcpy(t->d_bases, bases, sizeof(cd*) << log_world, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); delete t; return fail("peers_create: device table failed");
    }
    *out = t;
    return 0;
}
extern "C" void fftb200_peers_destroy(fftb200_peers* t) {
    if (!t) return;
    if (t->d_bases) cudaFree(t->d_bases);
    delete t;
}
// The last pass of a partial plan writes into the exchange buffers of all ranks (fft_tile.cuh: peer_ptr).
extern "C" int fftb200_plan_set_peer_output(fftb200_plan* p, const fftb200_peers* t, int log_width, int log_rows_per_rank) {
    if (!p || !t) return fail("set_peer_output: null argument");
    p->peers = t->d_bases;   // borrowed: the table must outlive the plan's executions
    p->peer_lw = log_width; p->peer_lrows = log_rows_per_rank; p->peer_lg = t->log_world; p->peer_me = t->rank;
    return 0;
}

// dst_peer[g][(me * rows + t) * W + c] = src[t * (W << log_world) + g * W + c]: the first exchange (T0) as one kernel
// that pushes the rank's rows, split by destination column block, straight into the peers' buffers
__global__ void push_columns_kernel(cd* const* peers, const cd* __restrict__ src, long long rows, int lw, int lg, int me) {
    const long long total = rows << (lw + lg), stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const long long c = i & ((1LL << lw) - 1), g = (i >> lw) & ((1LL << lg) - 1), t = i >> (lw + lg);
        peers[g][(((long long)me * rows + t) << lw) + c] = src[i];
    }
}
extern "C" int fftb200_push_columns(const fftb200_peers* t, void* stream, const void* src, long long rows, int log_width) {
    if (!t || !src || rows < 1) return fail("push_columns: bad argument");
    const long long total = rows << (log_width + t->log_world);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    push_columns_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(t->d_bases, (const cd*)src, rows, log_width, t->log_world, t->rank);
    CU(cudaGetLastError());
    return 0;
}

// Stream-ordered barrier between the ranks of a distributed transform, through peer memory instead of a collective library:
// every rank owns a flag area (one 64-bit epoch per source rank, inside an IPC-exchanged buffer); the kernel stores the new
// epoch into its slot of every peer's area (system-scope release) and spins until all slots of its own area have reached
// it. Kernels enqueued before it on the stream have completed - their peer stores included - when it starts, and nothing
// enqueued after it starts before every rank has arrived. All ranks must enqueue the same sequence of barriers.
struct fftb200_barrier {
    unsigned long long** d_flags = nullptr;   // device array: flag area of every rank as seen from this process
    int world = 0, rank = 0;
    unsigned long long epoch = 0;
};
__global__ void peer_barrier_kernel(unsigned long long* const* flags, int world, int rank, unsigned long long epoch) {
    const int g = threadIdx.x;
    if (g >= world) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags[g] + rank), "l"(epoch) : "memory");
    const unsigned long long* mine = flags[rank] + g;
    for (long long it = 0;; it++) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
        if (v >= epoch) break;
        if (it > (1LL << 28)) __trap();   // ~1 min: a rank that never arrives fails the launch instead of hanging the GPU for good
        __nanosleep(200);
    }
}
extern "C" int fftb200_barrier_create(fftb200_barrier** out, void* const* flag_areas, int world, int rank) {
    if (!out || !flag_areas || world < 1 || world > 64 || rank < 0 || rank >= world) return fail("barrier_create: bad argument");
    fftb200_barrier* b = new fftb200_barrier();
    b->world = world; b->rank = rank;
    if (cudaMalloc(&b->d_flags, sizeof(void*) * world) != cudaSuccess ||
        cudaMemcpy(b->d_flags, flag_areas, sizeof(void*) * world, cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaGetLastError(); delete b; return fail("barrier_create: device table failed");
    }
    *out = b;
    return 0;
}
extern "C" int fftb200_barrier_enqueue(fftb200_barrier* b, void* stream) {
    if (!b) return fail("barrier_enqueue: null barrier");
    b->epoch++;
    peer_barrier_kernel<<<1, 64, 0, (cudaStream_t)stream>>>(b->d_flags, b->world, b->rank, b->epoch);
    CU(cudaGetLastError());
    return 0;
}
extern "C" void fftb200_barrier_destroy(fftb200_barrier* b) {
    if (!b) return;
    if (b->d_flags) cudaFree(b->d_flags);
    delete b;
}
extern "C" int fftb200_stream_sync(void* stream) { CU(cudaStreamSynchronize((cudaStream_t)stream)); return 0; }

// CUDA IPC: one process per GPU, so peers' exchange buffers are mapped through handles exchanged by the caller
extern "C" int fftb200_ipc_export(void* dptr, void* handle64) {
    if (!dptr || !handle64) return fail("ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, dptr));
    return 0;
}
extern "C" void* fftb200_ipc_open(const void* handle64) {
    void* p = nullptr;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { fail("cudaIpcOpenMemHandle: %s", cudaGetErrorString(e)); cudaGetLastError(); return nullptr; }
    return p;
}
// Ranks that are threads of one process reach each other's buffers directly: enable peer access from the current device.
extern "C" int fftb200_enable_peer_access(int peer_device) {
    int cur = -1;
    CU(cudaGetDevice(&cur));
    if (cur == peer_device) return 0;
    int can = 0;
    CU(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    if (!can) return fail("device %d cannot access device %d", cur, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail("cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
}
extern "C" int fftb200_ipc_close(void* p) {
    if (p) CU(cudaIpcCloseMemHandle(p));
    return 0;
}

static unsigned grid_for(size_t total) {
    const size_t b = (total + 255) / 256;
    return (unsigned)(b > (1u << 20) ? (1u << 20) : (b ? b : 1));
}

// Enqueue the plan's transform over the first `nbatch` transforms of d_in / d_out on the plan's stream.
static int exec_range(fftb200_plan* p, const void* d_in, void* d_out, long long nbatch) {
    const int inverse = p->dir > 0;
    if (p->kind == FFTB200_C2C) return enqueue_c2c(p, (const cd*)d_in, (cd*)d_out, inverse, nbatch);
    const size_t m = (size_t)p->m, n = (size_t)p->n, total = m * (size_t)nbatch;
    if (p->child) {   // r2c of a non-power-of-two length
        if (nbatch <= 0) return 0;
        const size_t nh = n / 2 + 1, tot_in = n * (size_t)nbatch, tot_out = nh * (size_t)nbatch;
        r2c_promote_kernel<<<grid_for(tot_in), 256, 0, p->stream>>>(p->work, (const double*)d_in, tot_in);
        if (exec_range(p->child, p->work, p->work, nbatch) != 0) return -1;
        r2c_extract_kernel<<<grid_for(tot_out), 256, 0, p->stream>>>((cd*)d_out, p->work, n, nh, tot_out);
        CU(cudaGetLastError());
        return 0;
    }
    if ((p->kind == FFTB200_R2C || p->kind == FFTB200_C2R) && (p->pipe_real || !p->work) && nbatch > 0) {
        // the single-kernel real transforms read packed rows that other CTAs' outputs would overwrite: out of place only
        const size_t half = sizeof(cd) * (n / 2 + 1), full = sizeof(double) * n;
        const char* i0 = (const char*)d_in; const char* o0 = (const char*)d_out;
        const size_t ib = (p->kind == FFTB200_R2C ? full : half) * (size_t)nbatch, ob = (p->kind == FFTB200_R2C ? half : full) * (size_t)nbatch;
        if (i0 < o0 + ob && o0 < i0 + ib) return fail("plan_exec: r2c / c2r plans of this size run out of place (input and output overlap)");
    }
    if (p->pipe_real) {
        if (nbatch <= 0) return 0;
        const Pass& ps = p->passes[0];
        const long long ntiles = (nbatch + 2 * ps.nt - 1) / (2 * ps.nt);   // two real transforms per complex transform: 2 nt rows per tile
        PipeArgs pa;
        pa.in = (const cd*)d_in; pa.out = (cd*)d_out; pa.tab = p->acc;
        pa.ntiles = ntiles; pa.batch = nbatch;
        pa.inverse = p->kind == FFTB200_C2R; pa.scale = p->scale;
        pa.sched = sched_counters(p);
        if (!pa.sched) return -1;
        CU(launch_pipe_real(ps.log_p, p->kind == FFTB200_R2C ? PIPE_R2C : PIPE_C2R, pa, (int)(ntiles < ps.grid_max ? ntiles : ps.grid_max), p->stream));
        return 0;
    }
    if (p->kind == FFTB200_C2R) {
        if (nbatch <= 0) return 0;
        const size_t nh = n / 2 + 1;
        if (p->c2r_half) {
            if (enqueue_fused(p, p->passes[0], (const cd*)d_in, (cd*)d_out, 1, nbatch, 2) != 0) return -1;
            CU(cudaGetLastError());
            return 0;
        }
        c2r_expand_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, (const cd*)d_in, n, nh, total);
        if (p->fused_c2r) {
            if (enqueue_fused(p, p->passes[0], p->work, (cd*)d_out, 1, nbatch, 2) != 0) return -1;
            CU(cudaGetLastError());
            return 0;
        }
        if (enqueue_c2c(p, p->work, p->work, 1, nbatch) != 0) return -1;
        c2r_real_kernel<<<grid_for(total), 256, 0, p->stream>>>((double*)d_out, p->work, total);
        CU(cudaGetLastError());
        return 0;
    }
    if (p->kind == FFTB200_R2C && !p->work) {
        if (nbatch <= 0) return 0;
        return enqueue_fused(p, p->passes[0], (const cd*)d_in, (cd*)d_out, 0, nbatch, 1);
    }
    if (p->kind == FFTB200_R2C) {
        r2c_promote_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, (const double*)d_in, total);
        if (enqueue_c2c(p, p->work, p->work, 0, nbatch) != 0) return -1;
        const size_t nh = n / 2 + 1, tot_out = nh * (size_t)nbatch;
        r2c_extract_kernel<<<grid_for(tot_out), 256, 0, p->stream>>>((cd*)d_out, p->work, n, nh, tot_out);
        CU(cudaGetLastError());
        return 0;
    }
    // Bluestein (bluestein.c:107-148): a = x * conj(chirp) zero-padded to m; A = FFT(a) * FB; inverse FFT;
    // y = a * conj(chirp) (and 1/n for the inverse direction).
    // Multi-pass plans (m >= 2^21) carry the three elementwise steps on their first / last tile pass: 6 launches and HBM
    // round trips per execution instead of 9 (n = 1000003, same box: x1 0.120 -> 0.109 ms, x16 1.68 -> 1.28 ms, x64 5.81 -> 4.83 ms).
    if (p->pipe_blue) {
        if (nbatch <= 0) return 0;
        const Pass& ps = p->passes[0];
        const long long ntiles = pass_tiles(ps, nbatch);
        const int grid = (int)(ntiles < ps.grid_max ? ntiles : ps.grid_max);
        PipeArgs pa;
        pa.tab = p->acc; pa.ntiles = ntiles; pa.batch = nbatch;
        pa.sched = sched_counters(p);
        if (!pa.sched) return -1;
        pa.chirp = p->chirp; pa.fb = p->fb; pa.n_user = p->n; pa.y_scale = inverse ? 1.0 / (double)p->n : 1.0;
        pa.in = (const cd*)d_in; pa.out = p->work; pa.inverse = 0; pa.scale = 1.0;
        CU(launch_pipe_real(ps.log_p, PIPE_BLUE_FWD, pa, grid, p->stream));
        pa.in = p->work; pa.out = (cd*)d_out; pa.inverse = 1; pa.scale = p->scale;
        CU(launch_pipe_real(ps.log_p, PIPE_BLUE_INV, pa, grid, p->stream));
        return 0;
    }
    if (p->fused_blue) {
        if (nbatch <= 0) return 0;
        const Pass& ps = p->passes[0];
        const int mode = getenv("FFTB200_FUSED_BLUE_MODE") ? atoi(getenv("FFTB200_FUSED_BLUE_MODE")) : 3;   // bit 0: forward fused, bit 1: inverse fused
        const double ys = inverse ? 1.0 / (double)p->n : 1.0;
        if (mode & 1) {
            if (enqueue_fused(p, ps, (const cd*)d_in, p->work, 0, nbatch, 0, FUSED_BLUE_FWD) != 0) return -1;
        } else {
            bluestein_pre_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, (const cd*)d_in, p->chirp, p->n, p->m, total);
            if (enqueue_c2c(p, p->work, p->work, 0, nbatch) != 0) return -1;
            pointwise_mul_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, p->work, p->fb, total, m);
        }
        if (mode & 2) {
            if (enqueue_fused(p, ps, p->work, (cd*)d_out, 1, nbatch, 0, FUSED_BLUE_INV) != 0) return -1;
        } else {
            if (enqueue_c2c(p, p->work, p->work, 1, nbatch) != 0) return -1;
            const size_t tot_out = n * (size_t)nbatch;
            bluestein_post_kernel<<<grid_for(tot_out), 256, 0, p->stream>>>((cd*)d_out, p->work, p->chirp, p->n, p->m, tot_out, ys);
        }
        CU(cudaGetLastError());
        return 0;
    }
    const bool fpre = can_fuse_pre(p), fpost = can_fuse_post(p);
    const double yscale = inverse ? 1.0 / (double)p->n : 1.0;
    FuseSpec pre, fb, post;
    pre.mode = MUL_PRE; pre.mul = p->chirp; pre.user = (const cd*)d_in; pre.n = p->n;
    fb.mode = MUL_FB; fb.mul = p->fb; fb.n = (int)m;
    post.mode = MUL_POST; post.mul = p->chirp; post.user = (const cd*)d_out; post.n = p->n; post.scale = yscale;
    if (!fpre) bluestein_pre_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, (const cd*)d_in, p->chirp, p->n, p->m, total);
    if (enqueue_c2c(p, p->work, p->work, 0, nbatch, fpre ? &pre : nullptr, fpost ? &fb : nullptr) != 0) return -1;
    if (!fpost) pointwise_mul_kernel<<<grid_for(total), 256, 0, p->stream>>>(p->work, p->work, p->fb, total, m);
    if (enqueue_c2c(p, p->work, p->work, 1, nbatch, nullptr, fpost ? &post : nullptr) != 0) return -1;
    if (!fpost) {
        const size_t tot_out = n * (size_t)nbatch;
        bluestein_post_kernel<<<grid_for(tot_out), 256, 0, p->stream>>>((cd*)d_out, p->work, p->chirp, p->n, p->m, tot_out, yscale);
    }
    CU(cudaGetLastError());
    return 0;
}

extern "C" int fftb200_plan_exec_async(fftb200_plan* p, const void* d_in, void* d_out) {
    if (!p || !d_in || !d_out) return fail("plan_exec: null argument");
    DeviceGuard dg(p->device);
    return exec_range(p, d_in, d_out, p->batch);
}

extern "C" int fftb200_plan_sync(fftb200_plan* p) {
    if (!p) return fail("plan_sync: null plan");
    DeviceGuard dg(p->device);
    CU(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int fftb200_plan_exec(fftb200_plan* p, const void* d_in, void* d_out) {
    if (fftb200_plan_exec_async(p, d_in, d_out) != 0) return -1;
    return fftb200_plan_sync(p);
}

// Host-pointer execution (the H2D -> execute -> D2H sequence of algorithms/auto/fft_auto.c:278-280 and
// gpu/fft_cuda.cu:214-252), pipelined: the batch is cut into chunks of whole transforms; chunk c is uploaded
// on one stream while chunk c-1 is transformed in place on the plan's stream and chunk c-2 is downloaded on
// a third stream, through a ring of three device staging buffers owned by the plan. PCIe is full duplex, so
// with pinned host memory (fft_alloc_complex / fftb200_host_alloc) the job takes max(upload, download)
// instead of their sum; pageable memory works too but serialises inside the driver.
extern "C" int fftb200_plan_exec_host(fftb200_plan* p, const void* h_in, void* h_out) {
    if (!p || !h_in || !h_out) return fail("plan_exec_host: null argument");
    DeviceGuard dg(p->device);
    const size_t half = sizeof(cd) * (size_t)(p->n / 2 + 1);
    const size_t in_per = p->kind == FFTB200_R2C ? sizeof(double) * (size_t)p->n : p->kind == FFTB200_C2R ? half : sizeof(cd) * (size_t)p->n;  // bytes / transform
    const size_t out_per = p->kind == FFTB200_R2C ? half : p->kind == FFTB200_C2R ? sizeof(double) * (size_t)p->n : sizeof(cd) * (size_t)p->n;
    // the fused r2c kernel cannot run in place: its staging buffers hold the real input followed by the half spectra
    const bool split = (p->kind == FFTB200_R2C || p->kind == FFTB200_C2R) && !p->work;   // (and neither can the real variants of the pipe kernel)
    const size_t per = split ? in_per + out_per + 16 : (in_per > out_per ? in_per : out_per);
    // Small single-kernel transforms (fft_auto on 1024 points: BASELINE config 1) are latency, not bandwidth: two cudaMemcpy
    // calls and three streams cost ~40 us around a 9 us kernel. Instead the caller's data goes through a page-locked, device-
    // mapped staging buffer of the plan and the kernel reads / writes it directly over PCIe: one launch, one synchronise.
    {
        const size_t in_bytes = in_per * (size_t)p->batch, out_bytes = out_per * (size_t)p->batch;
        // (kernels that address global memory through tensor maps - N >= 8192 - keep the staged path)
        const bool plain = p->passes.size() == 1 && !p->passes[0].fused_lm && p->passes[0].log_p <= 12;
        const size_t big = in_bytes > out_bytes ? in_bytes : out_bytes;
        if (plain && big <= (128u << 10) && !getenv("FFTB200_NO_ZEROCOPY")) {
            if (!p->h_zc) {
                const size_t half = (big + 255) & ~(size_t)255;
                void* h = nullptr; void* d = nullptr;
                if (cudaHostAlloc(&h, 2 * half, cudaHostAllocMapped) == cudaSuccess && cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
                    p->h_zc = (char*)h; p->d_zc = (char*)d; p->zc_half = half;
                } else {
                    if (h) cudaFreeHost(h);
                    cudaGetLastError();
                }
            }
            if (p->h_zc) {
                memcpy(p->h_zc, h_in, in_bytes);
                if (exec_range(p, p->d_zc, p->d_zc + p->zc_half, p->batch) != 0) return -1;
                CU(cudaStreamSynchronize(p->stream));
                memcpy(h_out, p->h_zc + p->zc_half, out_bytes);
                return 0;
            }
        }
    }
    if (!p->d_ring[0]) {
        size_t target = 32u << 20;                       // bytes per chunk
        if (const char* e = getenv("FFTB200_CHUNK_MB")) target = (size_t)atol(e) << 20;
        long long cb = (long long)(target / per);
        if (cb < 1) cb = 1;
        if (cb > p->batch) cb = p->batch;
        // all or nothing: a half-built ring must not survive into the next call (ring_batch == 0 would never advance the loop below)
        bool ok = true;
        for (int i = 0; i < fftb200_plan::NSTAGE && ok; i++) {
            p->d_ring[i] = (cd*)fftb200_malloc(per * (size_t)cb + 512);
            ok = p->d_ring[i] != nullptr &&
                 cudaEventCreateWithFlags(&p->ev_up[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->ev_run[i], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->ev_down[i], cudaEventDisableTiming) == cudaSuccess;
        }
        if (!ok) {
            const std::string why = g_err[0] ? g_err : "cudaEventCreate failed";
            for (int i = 0; i < fftb200_plan::NSTAGE; i++) {
                if (p->d_ring[i]) cudaFree(p->d_ring[i]);
                if (p->ev_up[i]) cudaEventDestroy(p->ev_up[i]);
                if (p->ev_run[i]) cudaEventDestroy(p->ev_run[i]);
                if (p->ev_down[i]) cudaEventDestroy(p->ev_down[i]);
                p->d_ring[i] = nullptr; p->ev_up[i] = nullptr; p->ev_run[i] = nullptr; p->ev_down[i] = nullptr;
            }
            p->ring_batch = 0;
            cudaGetLastError();
            return fail("plan_exec_host: staging ring setup failed (%s)", why.c_str());
        }
        p->ring_batch = (int)cb;
    }
    const long long cb = p->ring_batch;
    if (cb <= 0) return fail("plan_exec_host: staging ring is not set up");
    const char* src = (const char*)h_in;
    char* dst = (char*)h_out;
    int c = 0;
    for (long long b0 = 0; b0 < p->batch; b0 += cb, c++) {
        const long long nb = p->batch - b0 < cb ? p->batch - b0 : cb;
        const int r = c % fftb200_plan::NSTAGE;
        if (c >= fftb200_plan::NSTAGE) CU(cudaStreamWaitEvent(p->s_up, p->ev_down[r], 0));   // staging buffer drained
        CU(cudaMemcpyAsync(p->d_ring[r], src + (size_t)b0 * in_per, in_per * (size_t)nb, cudaMemcpyHostToDevice, p->s_up));
        CU(cudaEventRecord(p->ev_up[r], p->s_up));
        CU(cudaStreamWaitEvent(p->stream, p->ev_up[r], 0));
        char* const d_res = split ? (char*)p->d_ring[r] + ((in_per * (size_t)cb + 255) & ~(size_t)255) : (char*)p->d_ring[r];
        if (exec_range(p, p->d_ring[r], d_res, nb) != 0) return -1;
        CU(cudaEventRecord(p->ev_run[r], p->stream));
        CU(cudaStreamWaitEvent(p->s_down, p->ev_run[r], 0));
        CU(cudaMemcpyAsync(dst + (size_t)b0 * out_per, d_res, out_per * (size_t)nb, cudaMemcpyDeviceToHost, p->s_down));
        CU(cudaEventRecord(p->ev_down[r], p->s_down));
    }
    CU(cudaStreamSynchronize(p->s_down));
    CU(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" void fftb200_plan_destroy(fftb200_plan* p) {
    if (!p) return;
    DeviceGuard dg(p->device);
    if (p->stream) cudaStreamSynchronize(p->stream);
    if (p->s_down) cudaStreamSynchronize(p->s_down);
    if (p->child) fftb200_plan_destroy(p->child);   // borrows this plan's stream: goes first
    if (p->scratch) cudaFree(p->scratch);
    if (p->fscratch) cudaFree(p->fscratch);
    if (p->fflags) cudaFree(p->fflags);
    if (p->sched) cudaFree(p->sched);
    if (p->own_tab) cudaFree(p->own_tab);
    if (p->work) cudaFree(p->work);
    if (p->chirp) cudaFree(p->chirp);
    if (p->fb) cudaFree(p->fb);
    if (p->h_zc) cudaFreeHost(p->h_zc);
    for (int i = 0; i < fftb200_plan::NSTAGE; i++) {
        if (p->d_ring[i]) cudaFree(p->d_ring[i]);
        if (p->ev_up[i]) cudaEventDestroy(p->ev_up[i]);
        if (p->ev_run[i]) cudaEventDestroy(p->ev_run[i]);
        if (p->ev_down[i]) cudaEventDestroy(p->ev_down[i]);
    }
    if (p->ev0) cudaEventDestroy(p->ev0);
    if (p->ev1) cudaEventDestroy(p->ev1);
    if (p->stream && p->owns_stream) cudaStreamDestroy(p->stream);
    if (p->s_up) cudaStreamDestroy(p->s_up);
    if (p->s_down) cudaStreamDestroy(p->s_down);
    table_release(p->tab_ref);
    table_release(p->acc_ref);
    delete p;
}

extern "C" int fftb200_plan_launches(const fftb200_plan* p) { return p ? p->launches : 0; }
extern "C" const char* fftb200_plan_describe(const fftb200_plan* p) { return p ? p->desc.c_str() : ""; }

extern "C" int fftb200_plan_set_stream(fftb200_plan* p, void* stream) {
    if (!p || !stream) return fail("plan_set_stream: null argument");
    if (p->stream) CU(cudaStreamSynchronize(p->stream));
    if (p->stream && p->owns_stream) cudaStreamDestroy(p->stream);
    p->stream = (cudaStream_t)stream;
    p->owns_stream = false;
    return 0;
}

extern "C" int fftb200_transpose(void* dst, const void* src, long long rows, long long cols, long long batch, void* stream) {
    if (!dst || !src || rows < 1 || cols < 1 || batch < 1 || dst == src) return fail("transpose: bad argument");
    const long long gx = (cols + 31) / 32, gy = (rows + 31) / 32;
    if (gx > 0x7fffffffLL) return fail("transpose: too many columns");
    const dim3 grid((unsigned)gx, (unsigned)(gy > 65535 ? 65535 : gy), (unsigned)(batch > 65535 ? 65535 : batch));
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>((cd*)dst, (const cd*)src, rows, cols, batch);
    CU(cudaGetLastError());
    return 0;
}

extern "C" int fftb200_pointwise_mul_conj(void* y, const void* a, const void* b, size_t count) {
    if (!y || !a || !b) return fail("pointwise_mul_conj: null argument");
    if (count == 0) return 0;
    pointwise_mul_conj_kernel<<<grid_for(count), 256>>>((cd*)y, (const cd*)a, (const cd*)b, count);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return 0;
}

extern "C" int fftb200_timer_start(fftb200_plan* p) {
    if (!p) return fail("timer: null plan");
    DeviceGuard dg(p->device);
    CU(cudaEventRecord(p->ev0, p->stream));
    return 0;
}
extern "C" int fftb200_timer_stop(fftb200_plan* p, float* ms) {
    if (!p || !ms) return fail("timer: null argument");
    DeviceGuard dg(p->device);
    CU(cudaEventRecord(p->ev1, p->stream));
    CU(cudaEventSynchronize(p->ev1));
    CU(cudaEventElapsedTime(ms, p->ev0, p->ev1));
    return 0;
}
