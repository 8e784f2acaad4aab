/*
 * demo_drop_in.c - every public entry point of the library in the order a user of the reference meets them
 * (README snippets, examples/demo_v2_features.c:65-181): hardware detection, fft_auto, plan / execute / destroy,
 * r2c and c2r, 2-D, the device handle API with a batch, the host-pointer batch call, and the convolution helper.
 * Each step checks itself (round trips, known peaks) and the program exits non-zero on the first failure, so it
 * doubles as a smoke test of the drop-in boundary (tests/test_gpu_apps.py runs it on the GPU box).
 */
#include "fft_auto.h"
#include "fft_gpu.h"
#include "fftb200_ext.h"

static int check(const char* what, double err, double tol) {
    printf("  %-58s %.2e  %s\n", what, err, err <= tol ? "ok" : "FAILED");
    return err <= tol ? 0 : 1;
}

static double max_abs_diff(const complex_t* a, const complex_t* b, size_t n) {
    double m = 0;
    for (size_t i = 0; i < n; i++) { const double e = cabs(a[i] - b[i]); if (e > m) m = e; }
    return m;
}

int main(void) {
    int bad = 0;
    printf("%s\n", fft_version());
    const unsigned caps = fft_get_hardware_capabilities();
    printf("hardware: %s%s%s%s%s\n", caps & FFT_HW_CPU_SSE ? "SSE " : "", caps & FFT_HW_CPU_AVX2 ? "AVX2 " : "",
           caps & FFT_HW_CPU_AVX512 ? "AVX-512 " : "", caps & FFT_HW_GPU_CUDA ? "CUDA " : "", caps & FFT_HW_GPU_CUDA ? "" : "(no GPU)");
    if (!fft_gpu_available()) { printf("no CUDA device: nothing runs on the CPU in this library\n"); return 2; }
    if (fft_gpu_init(FFT_GPU_AUTO) != 0) return 3;
    size_t tot = 0, avail = 0;
    fft_gpu_get_memory_info(&tot, &avail);
    printf("device: %s, %.1f GB total, %.1f GB free\n", fft_gpu_get_device_name(), tot / 1e9, avail / 1e9);

    /* 1. one-shot: a 50 Hz + 120 Hz signal sampled at 1 kHz (README quick start) */
    enum { N = 1024 };
    complex_t* sig = fft_alloc_complex(N); complex_t* spec = fft_alloc_complex(N); complex_t* back = fft_alloc_complex(N);
    for (int i = 0; i < N; i++) sig[i] = sin(TWO_PI * 50.0 * i / 1000.0) + 0.5 * sin(TWO_PI * 120.0 * i / 1000.0);
    printf("fft_auto\n");
    if (fft_auto(sig, spec, N, -1) != 0 || fft_auto(spec, back, N, 1) != 0) return 4;
    int peak = 1;
    for (int k = 1; k < N / 2; k++) if (cabs(spec[k]) > cabs(spec[peak])) peak = k;
    bad += check("spectral peak at 50 Hz (bin 51 of 1024 at 1 kHz)", fabs(peak * 1000.0 / N - 50.0), 1.0);
    bad += check("forward then inverse, max abs", max_abs_diff(back, sig, N), 1e-10);

    /* 2. plans, including a prime size (Bluestein) and execute_dft on other arrays */
    printf("fft_plan_dft_1d / fft_execute / fft_execute_dft\n");
    enum { P = 1009 };
    complex_t* a = fft_alloc_complex(P); complex_t* b = fft_alloc_complex(P); complex_t* c = fft_alloc_complex(P);
    for (int i = 0; i < P; i++) a[i] = cos(0.37 * i) + I * sin(0.11 * i * i);
    fft_plan_t pf = fft_plan_dft_1d(P, a, b, -1, FFT_MEASURE), pi = fft_plan_dft_1d(P, b, c, 1, FFT_ESTIMATE);
    if (!pf || !pi) return 5;
    fft_execute(pf); fft_execute(pi);
    bad += check("prime n = 1009 round trip, max abs", max_abs_diff(c, a, P), 1e-10);
    fft_execute_dft(pi, b, a);   /* same plan, other output array */
    bad += check("fft_execute_dft on other arrays agrees", max_abs_diff(a, c, P), 1e-13);
    fft_destroy_plan(pf); fft_destroy_plan(pi);

    /* 3. real transforms */
    printf("fft_plan_r2c_1d / fft_plan_c2r_1d\n");
    double* re = fft_alloc_real(N); double* re2 = fft_alloc_real(N);
    for (int i = 0; i < N; i++) re[i] = creal(sig[i]);
    fft_plan_t pr = fft_plan_r2c_1d(N, re, spec, 0), pc = fft_plan_c2r_1d(N, spec, re2, 0);
    if (!pr || !pc) return 6;
    fft_execute(pr); fft_execute(pc);
    double m = 0;
    for (int i = 0; i < N; i++) if (fabs(re2[i] - re[i]) > m) m = fabs(re2[i] - re[i]);
    bad += check("r2c then c2r, max abs", m, 1e-10);
    fft_destroy_plan(pr); fft_destroy_plan(pc);

    /* 4. 2-D */
    printf("fft_plan_dft_2d\n");
    enum { R = 64, Cc = 128 };
    complex_t* img = fft_alloc_complex(R * Cc); complex_t* img2 = fft_alloc_complex(R * Cc);
    for (int i = 0; i < R * Cc; i++) img[i] = ((i * 2654435761u) >> 8 & 0xffff) / 65536.0 - 0.5;
    fft_plan_t p2 = fft_plan_dft_2d(R, Cc, img, img2, -1, 0), p2i = fft_plan_dft_2d(R, Cc, img2, img2, 1, 0);
    if (!p2 || !p2i) return 7;
    fft_execute(p2);
    complex_t dc = 0;
    for (int i = 0; i < R * Cc; i++) dc += img[i];
    bad += check("2-D DC bin equals the pixel sum", cabs(img2[0] - dc), 1e-9);
    fft_execute(p2i);
    bad += check("2-D forward then inverse (in place), max abs", max_abs_diff(img2, img, R * Cc), 1e-10);
    fft_destroy_plan(p2); fft_destroy_plan(p2i);

    /* 5. device handle API with a batch, and the host-pointer batch convenience */
    printf("fft_gpu_plan_1d(n, batch) / fft_gpu_execute / fft_gpu_dft_1d_batch\n");
    enum { BN = 4096, BATCH = 64 };
    complex_t* hb = fft_alloc_complex((size_t)BN * BATCH); complex_t* hb2 = fft_alloc_complex((size_t)BN * BATCH);
    for (int i = 0; i < BN * BATCH; i++) hb[i] = sin(0.001 * i) + I * cos(0.003 * i);
    fft_gpu_memory_t dm = fft_gpu_alloc((size_t)BN * BATCH);
    fft_gpu_plan_t gf = fft_gpu_plan_1d(BN, BATCH, FFT_FORWARD), gi = fft_gpu_plan_1d(BN, BATCH, FFT_INVERSE);
    if (!dm || !gf || !gi) return 8;
    fft_gpu_copy_h2d(dm, hb, (size_t)BN * BATCH);
    fft_gpu_execute(gf, dm, dm); fft_gpu_execute(gi, dm, dm);
    fft_gpu_copy_d2h(hb2, dm, (size_t)BN * BATCH);
    bad += check("64 x 4096 forward + inverse on the device, max abs", max_abs_diff(hb2, hb, (size_t)BN * BATCH), 1e-10);
    if (fft_gpu_dft_1d_batch(hb, hb2, BN, BATCH, FFT_FORWARD) != 0 || fft_gpu_dft_1d_batch(hb2, hb2, BN, BATCH, FFT_INVERSE) != 0) return 9;
    bad += check("fft_gpu_dft_1d_batch forward + inverse, max abs", max_abs_diff(hb2, hb, (size_t)BN * BATCH), 1e-10);
    fft_gpu_destroy_plan(gf); fft_gpu_destroy_plan(gi); fft_gpu_free(dm);

    /* 6. convolution with a 5-tap moving average: interior samples are the local mean */
    printf("fft_gpu_convolution\n");
    complex_t taps[5], y[N + 4];
    for (int i = 0; i < 5; i++) taps[i] = 0.2;
    if (fft_gpu_convolution(sig, N, taps, 5, y) != 0) return 10;
    m = 0;
    for (int i = 4; i < N; i++) {
        const complex_t want = 0.2 * (sig[i] + sig[i - 1] + sig[i - 2] + sig[i - 3] + sig[i - 4]);
        if (cabs(y[i] - want) > m) m = cabs(y[i] - want);
    }
    bad += check("5-tap moving average through the FFT, max abs", m, 1e-12);

    char* w = fft_export_wisdom_to_string();
    printf("wisdom: %s", w ? w : "(none)\n");
    free(w);
    fft_free(sig); fft_free(spec); fft_free(back); fft_free(a); fft_free(b); fft_free(c); fft_free(re); fft_free(re2);
    fft_free(img); fft_free(img2); fft_free(hb); fft_free(hb2);
    fft_gpu_cleanup();
    printf("%s\n", bad ? "FAILED" : "all checks passed");
    return bad ? 1 : 0;
}
