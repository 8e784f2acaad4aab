/*
 * fft_auto.c - the planner API (include/fft_auto.h) on top of the B200 engine.
 *
 * Replaces the reference's algorithms/auto/fft_auto.c. The reference's select_algorithm (:136-172)
 * picks among CPU algorithms and only reaches the GPU under FFT_PREFER_GPU in a branch that gcc never
 * compiles (:220-229). Here the planner has one backend: every plan is an engine plan, chosen by n
 *   power of two  -> Stockham tile kernels (single pass up to 8192 points, otherwise 2-4 HBM passes)
 *   anything else -> Bluestein over the next power of two >= 2n-1 (the reference's choice for primes;
 *                    its "mixed radix" branch is an O(n^2) DFT, mixed_radix.c:107-124)
 * and there is NO CPU fallback: without a usable GPU the constructors return NULL, fft_auto returns -1.
 * Semantics kept: sign < 0 forward else inverse (:187), inverse scaled by 1/n, in/out borrowed and may
 * alias, fft_execute(NULL) is a no-op (:242), fft_execute_dft swaps arrays for one call (:287-302).
 */
#define _POSIX_C_SOURCE 200809L
#include "../../include/fft_auto.h"
#include "../../include/fft_gpu.h"
#include "../../include/fftb200.h"
#include "ref_twiddle.h"
#include <pthread.h>

fftb200_plan* fftb200_host_make_plan(int n, int batch, int direction, int kind); /* fft_gpu.c */

struct fft_plan {
    int n;
    complex_t* in;   /* borrowed */
    complex_t* out;  /* borrowed */
    double* real_in; /* borrowed, r2c only */
    double* real_out; /* borrowed, c2r only */
    fft_direction dir;
    unsigned flags;
    int kind;
    fftb200_plan* engine;
    /* 2-D plans (fft_plan_dft_2d): a device-resident image and the 2-D plan of the device API */
    int rows, cols;
    fft_gpu_plan_t plan2d;
    fft_gpu_memory_t image;
};

static int g_num_threads = 0;

static fft_plan_t make_plan(int n, complex_t* in, double* real_in, complex_t* out, int sign, unsigned flags, int kind) {
    fft_plan_t p = (fft_plan_t)calloc(1, sizeof(struct fft_plan));
    if (!p) return NULL;
    p->n = n; p->in = in; p->real_in = real_in; p->out = out;
    p->dir = sign < 0 ? FFT_FORWARD : FFT_INVERSE;
    p->flags = flags; p->kind = kind;
    p->engine = fftb200_host_make_plan(n, 1, (int)p->dir, kind);
    if (!p->engine) { free(p); return NULL; }
    return p;
}

fft_plan_t fft_plan_dft_1d(int n, complex_t* in, complex_t* out, int sign, unsigned flags) {
    if (n <= 0 || !in || !out) return NULL;
    return make_plan(n, in, NULL, out, sign, flags, is_power_of_two(n) ? FFTB200_C2C : FFTB200_BLUESTEIN);
}

fft_plan_t fft_plan_r2c_1d(int n, double* in, complex_t* out, unsigned flags) {
    if (n <= 0 || !in || !out) return NULL;
    /* any length, like the reference (fft_auto.c:391-403 promotes and plans a c2c, which routes other lengths to Bluestein) */
    return make_plan(n, NULL, in, out, -1, flags | FFT_REAL_INPUT, FFTB200_R2C);
}

void fft_execute(fft_plan_t plan) {
    if (!plan) return;
    if (plan->plan2d) { /* host -> device, rows + columns on the device, device -> host (fft_auto.c:278-280) */
        const size_t total = (size_t)plan->rows * (size_t)plan->cols;
        fft_gpu_copy_h2d(plan->image, plan->in, total);
        fft_gpu_execute(plan->plan2d, plan->image, plan->image);
        fft_gpu_copy_d2h(plan->out, plan->image, total);
        return;
    }
    if (plan->kind == FFTB200_C2R) {
        if (fftb200_plan_exec_host(plan->engine, plan->in, plan->real_out) != 0)
            fprintf(stderr, "fft_execute: %s\n", fftb200_last_error());
        return;
    }
    const void* src = plan->kind == FFTB200_R2C ? (const void*)plan->real_in : (const void*)plan->in;
    if (fftb200_plan_exec_host(plan->engine, src, plan->out) != 0)
        fprintf(stderr, "fft_execute: %s\n", fftb200_last_error());
}

void fft_execute_dft(fft_plan_t plan, complex_t* in, complex_t* out) {
    if (!plan || !in || !out) return;
    complex_t* keep_in = plan->in;
    complex_t* keep_out = plan->out;
    double* keep_real = plan->real_in;
    double* keep_real_out = plan->real_out;
    plan->in = in; plan->out = out;
    if (plan->kind == FFTB200_R2C) plan->real_in = (double*)in;
    if (plan->kind == FFTB200_C2R) plan->real_out = (double*)out;
    fft_execute(plan);
    plan->in = keep_in; plan->out = keep_out; plan->real_in = keep_real; plan->real_out = keep_real_out;
}

void fft_destroy_plan(fft_plan_t plan) {
    if (!plan) return;
    fftb200_plan_destroy(plan->engine);
    fft_gpu_destroy_plan(plan->plan2d);
    fft_gpu_free(plan->image);
    free(plan);
}

/* One-shot transform (reference fft_auto.c:325-333: plan, execute, destroy). The engine plan of the last
 * shape is kept by the host library, so a loop of fft_auto calls builds tables / streams / staging once. */
int fftb200_host_exec_cached(const void* in, void* out, int n, int batch, int direction, int kind); /* fft_gpu.c */

int fft_auto(complex_t* in, complex_t* out, int n, int sign) {
    if (n <= 0 || !in || !out) return -1;
    return fftb200_host_exec_cached(in, out, n, 1, sign < 0 ? -1 : 1, is_power_of_two(n) ? FFTB200_C2C : FFTB200_BLUESTEIN);
}

/* Stubs in the reference (fft_auto.c:405-415 return NULL); implemented here with the contracts of its header.
 * c2r (fft_auto.h:99-107): n/2 + 1 complex bins in, n reals out = the inverse of fft_plan_r2c_1d, scaled by 1/n like
 * every inverse of the library. `in` is read when the plan is executed. */
fft_plan_t fft_plan_c2r_1d(int n, complex_t* in, double* out, unsigned flags) {
    if (n <= 0 || !in || !out) return NULL;
    if (!is_power_of_two(n)) return NULL;
    fft_plan_t p = make_plan(n, in, NULL, NULL, 1, flags | FFT_REAL_OUTPUT, FFTB200_C2R);
    if (p) p->real_out = out;
    return p;
}
/* 2-D (fft_auto.h:109-121): row-major rows x cols, rows then columns (applications/image_fft.c:35-72); sign < 0
 * forward, otherwise inverse scaled by 1/(rows*cols). in/out are borrowed and may alias. */
fft_plan_t fft_plan_dft_2d(int rows, int cols, complex_t* in, complex_t* out, int sign, unsigned flags) {
    if (rows <= 0 || cols <= 0 || !in || !out) return NULL;
    fft_plan_t p = (fft_plan_t)calloc(1, sizeof(struct fft_plan));
    if (!p) return NULL;
    p->n = rows * cols; p->rows = rows; p->cols = cols; p->in = in; p->out = out;
    p->dir = sign < 0 ? FFT_FORWARD : FFT_INVERSE;
    p->flags = flags; p->kind = FFTB200_C2C;
    p->plan2d = fft_gpu_plan_2d(rows, cols, p->dir);
    p->image = p->plan2d ? fft_gpu_alloc((size_t)rows * (size_t)cols) : NULL;
    if (!p->plan2d || !p->image) { fft_destroy_plan(p); return NULL; }
    return p;
}

/* Wisdom: same header line as the reference's stub (fft_auto.c:417-421) followed by one line per planned shape; the
 * importer pre-builds the host tables for those shapes (host/fft_gpu.c). 1 on success, 0 on failure (fft_auto.h:137). */
char* fftb200_host_wisdom_export(void);            /* fft_gpu.c */
int fftb200_host_wisdom_import(const char* wisdom); /* fft_gpu.c */
char* fft_export_wisdom_to_string(void) { return fftb200_host_wisdom_export(); }
int fft_import_wisdom_from_string(const char* wisdom) { return fftb200_host_wisdom_import(wisdom); }

unsigned fft_get_hardware_capabilities(void) {
    unsigned caps = 0;
#if defined(__x86_64__) && defined(__GNUC__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("sse2")) caps |= FFT_HW_CPU_SSE;
    if (__builtin_cpu_supports("avx")) caps |= FFT_HW_CPU_AVX;
    if (__builtin_cpu_supports("avx2")) caps |= FFT_HW_CPU_AVX2;
    if (__builtin_cpu_supports("avx512f")) caps |= FFT_HW_CPU_AVX512;
#endif
    if (fft_gpu_available()) caps |= FFT_HW_GPU_CUDA;
    return caps;
}

void fft_plan_with_nthreads(int nthreads) { g_num_threads = nthreads; }

/* Allocators (reference fft_auto.c:352-384: 64-byte posix_memalign). With a GPU present the memory is
 * page-locked (cudaMallocHost, 256-byte aligned at least), so fft_execute / fft_gpu_dft_1d_batch on these
 * buffers upload and download by DMA at full PCIe rate and overlap the two directions; fft_free tells the
 * two kinds apart through a small registry. Without a GPU they are plain aligned allocations. */
struct pin_node { void* p; struct pin_node* next; };
static struct pin_node* g_pins = NULL;
static pthread_mutex_t g_pin_mu = PTHREAD_MUTEX_INITIALIZER;

static void* alloc_bytes(size_t bytes) {
    if (bytes == 0) bytes = 64;
    if (bytes >= (1u << 16) && fft_gpu_available() && !getenv("FFTB200_NO_PINNED")) {
        void* p = fftb200_host_alloc(bytes);
        if (p) {
            struct pin_node* nd = (struct pin_node*)malloc(sizeof(*nd));
            if (nd) {
                nd->p = p;
                pthread_mutex_lock(&g_pin_mu);
                nd->next = g_pins; g_pins = nd;
                pthread_mutex_unlock(&g_pin_mu);
                return p;
            }
            fftb200_host_free(p);
        }
    }
    void* p = NULL;
    if (posix_memalign(&p, 64, bytes) != 0) return NULL;
    return p;
}

complex_t* fft_alloc_complex(size_t n) { return (complex_t*)alloc_bytes(n * sizeof(complex_t)); }
double* fft_alloc_real(size_t n) { return (double*)alloc_bytes(n * sizeof(double)); }

void fft_free(void* p) {
    if (!p) return;
    pthread_mutex_lock(&g_pin_mu);
    for (struct pin_node** pp = &g_pins; *pp; pp = &(*pp)->next) {
        if ((*pp)->p == p) {
            struct pin_node* nd = *pp;
            *pp = nd->next;
            pthread_mutex_unlock(&g_pin_mu);
            free(nd);
            fftb200_host_free(p);
            return;
        }
    }
    pthread_mutex_unlock(&g_pin_mu);
    free(p);
}

const char* fft_version(void) { return "FFT Library v2.0.0 - Automatic Algorithm Selection (B200 sm_100a engine)"; }
