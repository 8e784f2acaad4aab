/*
 * benchmark_all_gpu.c - the reference's benchmark protocol (benchmarks/benchmark_all.c) run against the B200 library
 * through the PUBLIC API only. Same sizes and iteration counts (:274-279), same input (uniform rand()/RAND_MAX - 0.5
 * for both parts, default seed, :167-172), same table columns (:186-205) and the same verdict: PASS iff the
 * forward-then-inverse reconstruction is within 1e-10 max-abs of the input (:155). The reference benchmarks nine CPU
 * algorithms by calling them directly; this library has no CPU algorithm, so the rows are its three entry points:
 *   fft_auto       one-shot, host pointers (plan cache + H2D + kernel + D2H per call)
 *   plan+execute   fft_plan_dft_1d once, fft_execute per iteration (host pointers)
 *   gpu resident   fft_gpu_plan_1d / fft_gpu_execute on device buffers (what a batched caller pays per transform)
 * "Max/RMS Error" compare the forward output with an O(n^2) long-double DFT (the reference compares with its own
 * radix-2 code, :67-76), for n <= 4096 (the O(n^2) sum gets slow beyond). Larger sizes are added after the reference's list.
 * Exit code 0 iff every row passes.
 */
#define _POSIX_C_SOURCE 200809L
#include "fft_auto.h"
#include "fft_gpu.h"
#include <time.h>

static double now_ms(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static void truth_dft(const complex_t* in, complex_t* out, int n) {
    for (int k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (int j = 0; j < n; j++) {
            const long long kj = ((long long)k * j) % n;
            const long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)kj / n;
            const long double c = cosl(a), s = sinl(a);
            sr += creal(in[j]) * c - cimag(in[j]) * s;
            si += creal(in[j]) * s + cimag(in[j]) * c;
        }
        out[k] = (double)sr + I * (double)si;
    }
}

static void errors(const complex_t* a, const complex_t* b, int n, double* max_err, double* rms) {
    double m = 0, s = 0;
    for (int i = 0; i < n; i++) {
        const double e = cabs(a[i] - b[i]);
        if (e > m) m = e;
        s += e * e;
    }
    *max_err = m;
    *rms = sqrt(s / n);
}

typedef struct { const char* name; double fwd_ms, inv_ms, max_err, rms_err, recon; int ok; } row_t;

static void print_row(const row_t* r, int have_truth) {
    char me[32] = "N/A", re[32] = "N/A";
    if (have_truth) { snprintf(me, sizeof me, "%.2e", r->max_err); snprintf(re, sizeof re, "%.2e", r->rms_err); }
    printf("%-15s | %12.4f | %12.4f | %-12s | %-12s | %s (recon %.1e)\n", r->name, r->fwd_ms, r->inv_ms, me, re,
           r->ok ? "PASS" : "FAIL", r->recon);
}

int main(void) {
    printf("FFT Implementation Benchmark Suite - B200 library, public API\n");
    printf("==============================================================\n");
    printf("%s\n", fft_version());
    if (!fft_gpu_available() || fft_gpu_init(FFT_GPU_AUTO) != 0) {
        printf("no CUDA device: this library has no CPU fallback, nothing to benchmark\n");
        return 2;
    }
    printf("GPU Device: %s\n", fft_gpu_get_device_name());
    const int sizes[] = {16, 64, 256, 1024, 4096, 16384, 1 << 18, 1 << 20};
    const int iters[] = {10000, 5000, 1000, 100, 10, 1, 5, 5};
    int failures = 0;
    for (unsigned s = 0; s < sizeof sizes / sizeof sizes[0]; s++) {
        const int n = sizes[s], it = iters[s], have_truth = n <= 4096;
        complex_t* in = fft_alloc_complex(n); complex_t* work = fft_alloc_complex(n);
        complex_t* spec = fft_alloc_complex(n); complex_t* truth = fft_alloc_complex(n);
        if (!in || !work || !spec || !truth) return 3;
        for (int i = 0; i < n; i++) in[i] = ((double)rand() / RAND_MAX - 0.5) + I * ((double)rand() / RAND_MAX - 0.5);
        if (have_truth) truth_dft(in, truth, n);
        printf("\n=== Size: %d (%d iterations) ===\n", n, it);
        printf("%-15s | %-12s | %-12s | %-12s | %-12s | %s\n", "Implementation", "Forward (ms)", "Inverse (ms)", "Max Error", "RMS Error", "Status");
        printf("----------------|--------------|--------------|--------------|--------------|--------\n");
        row_t r;
        double t0;
        /* --- fft_auto --- */
        r.name = "fft_auto";
        fft_auto(in, spec, n, -1); /* warm-up, as the reference does (:119-121) */
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_auto(in, spec, n, -1);
        r.fwd_ms = (now_ms() - t0) / it;
        fft_auto(spec, work, n, 1);
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_auto(spec, work, n, 1);
        r.inv_ms = (now_ms() - t0) / it;
        if (have_truth) errors(spec, truth, n, &r.max_err, &r.rms_err);
        double rr; errors(work, in, n, &r.recon, &rr);
        r.ok = r.recon <= 1e-10; failures += !r.ok; print_row(&r, have_truth);
        /* --- plan + execute --- */
        r.name = "plan+execute";
        fft_plan_t pf = fft_plan_dft_1d(n, in, spec, -1, FFT_ESTIMATE), pi = fft_plan_dft_1d(n, spec, work, 1, FFT_ESTIMATE);
        if (!pf || !pi) return 4;
        fft_execute(pf);
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_execute(pf);
        r.fwd_ms = (now_ms() - t0) / it;
        fft_execute(pi);
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_execute(pi);
        r.inv_ms = (now_ms() - t0) / it;
        if (have_truth) errors(spec, truth, n, &r.max_err, &r.rms_err);
        errors(work, in, n, &r.recon, &rr);
        r.ok = r.recon <= 1e-10; failures += !r.ok; print_row(&r, have_truth);
        fft_destroy_plan(pf); fft_destroy_plan(pi);
        /* --- device resident --- */
        r.name = "gpu resident";
        fft_gpu_memory_t da = fft_gpu_alloc(n), db = fft_gpu_alloc(n);
        fft_gpu_plan_t gf = fft_gpu_plan_1d(n, 1, FFT_FORWARD), gi = fft_gpu_plan_1d(n, 1, FFT_INVERSE);
        if (!da || !db || !gf || !gi) return 5;
        fft_gpu_copy_h2d(da, in, n);
        fft_gpu_execute(gf, da, db);
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_gpu_execute(gf, da, db);
        r.fwd_ms = (now_ms() - t0) / it;
        fft_gpu_copy_d2h(spec, db, n);
        fft_gpu_execute(gi, db, da);
        t0 = now_ms();
        for (int k = 0; k < it; k++) fft_gpu_execute(gi, db, da);
        r.inv_ms = (now_ms() - t0) / it;
        fft_gpu_copy_d2h(work, da, n);
        if (have_truth) errors(spec, truth, n, &r.max_err, &r.rms_err);
        errors(work, in, n, &r.recon, &rr);
        r.ok = r.recon <= 1e-10; failures += !r.ok; print_row(&r, have_truth);
        fft_gpu_destroy_plan(gf); fft_gpu_destroy_plan(gi); fft_gpu_free(da); fft_gpu_free(db);
        fft_free(in); fft_free(work); fft_free(spec); fft_free(truth);
    }
    printf("\n%s\n", failures ? "SOME ROWS FAILED" : "ALL ROWS PASS");
    fft_gpu_cleanup();
    return failures ? 1 : 0;
}
