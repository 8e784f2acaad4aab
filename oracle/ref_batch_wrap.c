/* TEST/BENCH INFRASTRUCTURE ONLY.
 * Times the UNMODIFIED reference over a batch of transforms through its own public API
 * (include/fft_auto.h of the reference: fft_plan_dft_1d :43, fft_execute_dft :60, fft_destroy_plan :67),
 * one plan per host thread, transforms dealt to threads by an OpenMP loop in THIS file. The reference has
 * no batched CPU entry point (its fft_gpu_dft_1d_batch is a loop of single transforms, gpu/fft_gpu.c:366-374);
 * this is that loop, spread over the host cores. Links against _ref/libfftref.so. */
#define _GNU_SOURCE
#include <complex.h>
#include <stdlib.h>
#include <omp.h>

typedef double complex complex_t;
typedef struct fft_plan* fft_plan_t;
fft_plan_t fft_plan_dft_1d(int n, complex_t* in, complex_t* out, int sign, unsigned flags);
void fft_execute_dft(fft_plan_t plan, complex_t* in, complex_t* out);
void fft_destroy_plan(fft_plan_t plan);

int oracle_ref_batch_execute(double* in, double* out, int n, long batch, int sign, int threads) {
    if (threads < 1) threads = 1;
    fft_plan_t* plans = (fft_plan_t*)calloc((size_t)threads, sizeof(fft_plan_t));
    if (!plans) return -1;
    complex_t* cin = (complex_t*)in;
    complex_t* cout = (complex_t*)out;
    int ok = 1;
    for (int t = 0; t < threads; t++) {   /* serial: the reference planner has unsynchronised globals */
        plans[t] = fft_plan_dft_1d(n, cin, cout, sign, 0);
        if (!plans[t]) ok = 0;
    }
    if (ok) {
        #pragma omp parallel for num_threads(threads) schedule(static)
        for (long b = 0; b < batch; b++)
            fft_execute_dft(plans[omp_get_thread_num()], cin + (size_t)b * (size_t)n, cout + (size_t)b * (size_t)n);
    }
    for (int t = 0; t < threads; t++) if (plans[t]) fft_destroy_plan(plans[t]);
    free(plans);
    return ok ? 0 : -1;
}
