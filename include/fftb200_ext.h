/*
 * fftb200_ext.h - additive helpers of the host library (not in the reference's headers; nothing in
 * fft_auto.h / fft_gpu.h changes because of them).
 *
 *  - access to the engine plan / device pointer behind the public opaque handles, so a caller can time
 *    with CUDA events on the plan's stream or enqueue asynchronously (include/fftb200.h);
 *  - the batch partition used when a batched job is spread over several GPUs, one process (or one
 *    fft_gpu_set_device + plan) per GPU: transforms are independent, so rank r of `world` simply owns a
 *    contiguous range of the batch and there is no exchange step (reference layout: transform b at
 *    [b*n, (b+1)*n), gpu/fft_cuda.cu:152-156);
 *  - the host-side table generators (the reference's twiddle recurrence, algorithms/core/radix2_dit.c:
 *    93,109, and Bluestein chirp, algorithms/core/bluestein.c:59-62) for inspection and tests.
 */
#ifndef FFTB200_EXT_H
#define FFTB200_EXT_H

#include "fft_gpu.h"
#include "fftb200.h"

fftb200_plan* fftb200_engine_of(fft_gpu_plan_t plan);
void* fftb200_devptr_of(fft_gpu_memory_t mem);

/* Contiguous partition of `batch` transforms over `world` ranks: the first (batch % world) ranks own one
 * transform more. Returns 0, or -1 on bad arguments. count may be 0 when world > batch. */
int fftb200_shard_range(long long batch, int world, int rank, long long* first, long long* count);

/* Multi-device fan-out of the batched host entry points (fft_gpu_dft_1d_batch and friends): with G > 1 the batch is cut
 * into G contiguous ranges (fftb200_shard_range), one host thread, plan and staging ring per device, starting at the current
 * device; no exchange step. Default 1, or the environment variable FFTB200_GPUS read at the first call. */
void fftb200_host_set_gpus(int gpus);
int fftb200_host_get_gpus(void);
/* Diagnostics: engine plans built / re-used by the cached host entry points (fft_auto, fft_gpu_dft_1d_batch, ...). */
void fftb200_host_cache_stats(long long* builds, long long* hits);

const double* fftb200_host_twiddles(int n);           /* n - 1 complex, stage s entry j at 2^(s-1) - 1 + j */
void fftb200_host_chirp(double* out, int n, int dir); /* n complex */
void fftb200_host_tables_release(void);
/* Late-stage reference twiddles owned by `rank` in a distributed transform (see host/ref_twiddle.c); out holds
 * 2^(log_total - log_world) - 1 complex numbers. */
int fftb200_host_twiddles_dist(double* out, int log_total, int log_world, int rank, int log_m);

/* FFT convolution / correlation with host pointers (host/fft_apps.c): GPU versions of the reference's CPU callers
 * applications/convolution.c:34-96 (fft_convolution: y has nx + nh - 1 samples; circular_convolution: n samples) and
 * applications/power_spectrum.c:133-192 (autocorrelation_fft / cross_correlation_fft: the first n lags of
 * IFFT(conj(X) * Y) over next_power_of_two(2n) points). 0 on success, -1 on bad arguments or without a GPU. */
int fft_gpu_convolution(const complex_t* x, int nx, const complex_t* h, int nh, complex_t* y);
int fft_gpu_circular_convolution(const complex_t* x, const complex_t* h, int n, complex_t* y);
int fft_gpu_cross_correlation(const complex_t* x, const complex_t* y, int n, complex_t* ccf);
int fft_gpu_autocorrelation(const complex_t* x, int n, complex_t* acf);

#endif /* FFTB200_EXT_H */
