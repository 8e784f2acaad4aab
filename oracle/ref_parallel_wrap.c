/* TEST/BENCH INFRASTRUCTURE ONLY.
 * Wraps the reference's multi-threaded CPU path (optimizations/parallel_fft.c: fft_radix2_parallel
 * :130-210, four_step_fft :213-272) so it can be timed as the CPU baseline. The reference file carries
 * an unguarded demo main() (:366); it is renamed away. Nothing is copied: the file is #included from
 * where it lies under /root/reference (path given by -DREF_PARALLEL_C). */
#define _GNU_SOURCE
#include <stdio.h>
#include <unistd.h>
#include <fcntl.h>
#define main ref_parallel_demo_main
#include REF_PARALLEL_C
#undef main

/* four_step_fft printf()s from inside the transform (parallel_fft.c:224): silence stdout around it. */
void oracle_ref_four_step(double* x, int n, int dir, int threads) {
    fflush(stdout);
    int saved = dup(1), nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    four_step_fft((complex_t*)x, n, (fft_direction)dir, threads);
    fflush(stdout);
    dup2(saved, 1); close(saved); close(nul);
}

void oracle_ref_radix2_parallel(double* x, int n, int dir, int threads) {
    fft_radix2_parallel((complex_t*)x, n, (fft_direction)dir, threads);
}
