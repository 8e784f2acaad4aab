/*
 * fft_auto.h - planner API (plan / execute / destroy, one-shot fft_auto, r2c), v2.
 *
 * Drop-in for the reference's include/fft_auto.h: every prototype, enum value and typedef below keeps
 * the reference's name, signature and meaning (reference lines in brackets). Behind it the planner
 * (host/fft_auto.c) routes EVERY transform to the B200 engine - there is no CPU algorithm in this
 * library - choosing the kernel plan by n and batch.
 */
#ifndef FFT_AUTO_H
#define FFT_AUTO_H

#include "fft_common.h"

typedef struct fft_plan* fft_plan_t; /* [14] opaque */

typedef enum { /* [17-29] */
    FFT_ESTIMATE = 0,
    FFT_MEASURE = 1,
    FFT_PATIENT = 2,
    FFT_EXHAUSTIVE = 3,
    FFT_WISDOM_ONLY = 4,
    FFT_REAL_INPUT = 1 << 5,
    FFT_REAL_OUTPUT = 1 << 6,
    FFT_UNALIGNED = 1 << 7,
    FFT_CONSERVE_MEMORY = 1 << 8,
    FFT_PREFER_GPU = 1 << 9,
    FFT_THREADED = 1 << 10
} fft_flags_t;

/* [43] Plan an n-point complex transform; sign < 0 forward (unscaled), otherwise inverse (scaled 1/n).
 * in/out are borrowed host arrays of n elements and may alias. NULL on bad arguments or no GPU. */
fft_plan_t fft_plan_dft_1d(int n, complex_t* in, complex_t* out, int sign, unsigned flags);
/* [51] Run the plan on its own arrays: host -> device, kernels, device -> host. No-op on NULL. */
void fft_execute(fft_plan_t plan);
/* [60] Run the plan on other arrays of the same size. */
void fft_execute_dft(fft_plan_t plan, complex_t* in, complex_t* out);
/* [67] */
void fft_destroy_plan(fft_plan_t plan);
/* [85] plan + execute + destroy; 0 on success, -1 on error */
int fft_auto(complex_t* in, complex_t* out, int n, int sign);
/* [97] Real input of n doubles -> n/2 + 1 complex bins (power-of-two n). Unlike the reference
 * (which snapshots `in` at plan time and then frees the snapshot, fft_auto.c:391-403) the plan
 * reads `in` when it is executed. */
fft_plan_t fft_plan_r2c_1d(int n, double* in, complex_t* out, unsigned flags);
/* [107] n/2 + 1 complex bins -> n reals (power-of-two n): the inverse of fft_plan_r2c_1d, scaled by 1/n. A stub in the
 * reference (fft_auto.c:405-408 returns NULL); implemented here with the header's contract. `in` is read at execute time. */
fft_plan_t fft_plan_c2r_1d(int n, complex_t* in, double* out, unsigned flags);
/* [121] 2-D transform of a row-major rows x cols array, rows then columns (the decomposition of the reference's CPU
 * code, applications/image_fft.c:35-72); sign < 0 forward, otherwise inverse scaled by 1/(rows*cols). Any shape with
 * rows*cols <= 2^30. A stub in the reference (fft_auto.c:411-415 returns NULL); implemented here. */
fft_plan_t fft_plan_dft_2d(int rows, int cols, complex_t* in, complex_t* out, int sign, unsigned flags);

/* [130, 137] */
char* fft_export_wisdom_to_string(void);
int fft_import_wisdom_from_string(const char* wisdom);

typedef enum { /* [145-154] */
    FFT_HW_CPU_SSE = 1 << 0,
    FFT_HW_CPU_AVX = 1 << 1,
    FFT_HW_CPU_AVX2 = 1 << 2,
    FFT_HW_CPU_AVX512 = 1 << 3,
    FFT_HW_CPU_NEON = 1 << 4,
    FFT_HW_GPU_CUDA = 1 << 5,
    FFT_HW_GPU_MPS = 1 << 6,
    FFT_HW_GPU_OPENCL = 1 << 7
} fft_hardware_t;
unsigned fft_get_hardware_capabilities(void); /* [156] */

void fft_plan_with_nthreads(int nthreads); /* [164] kept; the GPU path has no host threads to size */

complex_t* fft_alloc_complex(size_t n); /* [173] 64-byte aligned */
double* fft_alloc_real(size_t n);       /* [180] */
void fft_free(void* p);                 /* [186] */

const char* fft_version(void); /* [194] */

#endif /* FFT_AUTO_H */
