/* TEST/BENCH INFRASTRUCTURE ONLY.
 * Wraps the reference's SIMD CPU path (optimizations/simd_fft.c: fft_radix2_sse2 :143-230, single precision, split
 * real / imaginary arrays) so bench.py can time it beside the GPU as BASELINE.json's north_star asks. SURVEY.md 6.2 / 8c-iii:
 * its output is numerically wrong (float, and the butterfly indexing is off), so it is TIMED ONLY, never compared.
 * The reference file carries an unguarded demo main() (:328); it is renamed away. Nothing is copied: the file is
 * #included from where it lies under /root/reference (path given by -DREF_SIMD_C). */
#define _GNU_SOURCE
#include <stdio.h>
#include <time.h>
#define main ref_simd_demo_main
#include REF_SIMD_C
#undef main

/* `reps` transforms of n points on one core; returns seconds per transform (CLOCK_MONOTONIC). */
double oracle_ref_sse2_time(int n, int reps) {
    complex_float_split_t* x = allocate_simd_complex(n);
    if (!x || !x->real || !x->imag) return -1.0;
    for (int i = 0; i < n; i++) { x->real[i] = (float)((i * 37 % 101) / 101.0 - 0.5); x->imag[i] = (float)((i * 53 % 103) / 103.0 - 0.5); }
    struct timespec t0, t1;
    fft_radix2_sse2(x, n, FFT_FORWARD);   /* warm */
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int r = 0; r < reps; r++) {
        fft_radix2_sse2(x, n, FFT_FORWARD);
        /* keep the values bounded: the unnormalised transform grows by sqrt(n) per call */
        const float s = 1.0f / (float)n;
        for (int i = 0; i < n; i += 997) { x->real[i] *= s; x->imag[i] *= s; }
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free_simd_complex(x);
    return ((double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec)) / (double)reps;
}
