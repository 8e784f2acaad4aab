/*
 * fft_gpu.h - device-resident handle API (init, device buffers, batched plans, execute).
 *
 * Drop-in for the reference's include/fft_gpu.h (reference lines in brackets): same names, signatures
 * and enum values. This is the only place the public API has a `batch` parameter, so it is the entry
 * point of the batched hot path: transform b of a plan occupies elements [b*n, (b+1)*n) of the buffer,
 * unit stride (the cufftPlanMany layout of the reference's gpu/fft_cuda.cu:152-156).
 *
 * Fixed relative to the reference while keeping signatures: fft_gpu_execute honours the plan's
 * direction (the reference hard-codes FFT_FORWARD, gpu/fft_gpu.c:252), and the inverse is scaled by
 * 1/n like every CPU algorithm of the reference (algorithms/core/radix2_dit.c:115-119).
 */
#ifndef FFT_GPU_H
#define FFT_GPU_H

#include "fft_common.h"

typedef enum { /* [14-20] */
    FFT_GPU_NONE = 0,
    FFT_GPU_CUDA = 1,
    FFT_GPU_METAL = 2,
    FFT_GPU_OPENCL = 3,
    FFT_GPU_AUTO = -1
} fft_gpu_backend_t;

typedef struct fft_gpu_memory* fft_gpu_memory_t; /* [23] opaque device buffer */
typedef struct fft_gpu_plan* fft_gpu_plan_t;     /* [26] opaque plan          */

int fft_gpu_init(fft_gpu_backend_t backend); /* [35] 0 on success; idempotent */
void fft_gpu_cleanup(void);                  /* [41] */
int fft_gpu_available(void);                 /* [47] 1 when a CUDA device is usable */
fft_gpu_backend_t fft_gpu_get_backend(void); /* [52] */

fft_gpu_memory_t fft_gpu_alloc(size_t size); /* [61] size in COMPLEX ELEMENTS */
void fft_gpu_free(fft_gpu_memory_t mem);     /* [67] */
void fft_gpu_copy_h2d(fft_gpu_memory_t dst, const complex_t* src, size_t size); /* [75] */
void fft_gpu_copy_d2h(complex_t* dst, fft_gpu_memory_t src, size_t size);       /* [83] */

/* [94] batched 1-D plan of any n >= 1 (power of two: Stockham kernels; otherwise Bluestein) */
fft_gpu_plan_t fft_gpu_plan_1d(int n, int batch, fft_direction direction);
/* [102] out may be the same buffer as in; returns when the result is visible */
void fft_gpu_execute(fft_gpu_plan_t plan, fft_gpu_memory_t in, fft_gpu_memory_t out);
void fft_gpu_destroy_plan(fft_gpu_plan_t plan); /* [108] */

/* [120, 131] host-pointer conveniences: one H2D, one batched execution, one D2H */
int fft_gpu_dft_1d(complex_t* in, complex_t* out, int n, fft_direction direction);
int fft_gpu_dft_1d_batch(complex_t* in, complex_t* out, int n, int batch, fft_direction direction);

/* [143, 154] 2-D plans over row-major rows x cols device buffers (execute with fft_gpu_execute, in place allowed) and
 * the host-pointer one-shot. Stubs in the reference (gpu/fft_gpu.c:377-394: NULL / -1); implemented here: one batched
 * row pass + strided column kernels (or two corner turns), forward unscaled, inverse scaled by 1/(rows*cols). */
fft_gpu_plan_t fft_gpu_plan_2d(int rows, int cols, fft_direction direction);
int fft_gpu_dft_2d(complex_t* in, complex_t* out, int rows, int cols, fft_direction direction);

const char* fft_gpu_get_device_name(void);                    /* [163] */
void fft_gpu_get_memory_info(size_t* total, size_t* available); /* [170] */
int fft_gpu_set_device(int device);                           /* [177] selects the device used by later calls */

#ifdef __CUDACC__
typedef struct { /* [191-197] declared by the reference, defined nowhere there */
    int block_size;
    int shared_memory_size;
    cudaStream_t stream;
} fft_cuda_options_t;
void fft_gpu_set_cuda_options(const fft_cuda_options_t* options);
#endif

#endif /* FFT_GPU_H */
