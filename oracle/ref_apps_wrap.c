/* TEST INFRASTRUCTURE ONLY.
 * The reference's FFT callers - the "next" rows of SURVEY.md 8f - compiled UNMODIFIED from where they lie under
 * /root/reference (paths from -DREF_CONV_C / -DREF_IMAGE_C / -DREF_PSD_C) so that the oracle's restatements of
 * them (fft_oracle.c: oracle_convolution, oracle_fft2d, oracle_cross_correlation ...) are pinned to the real code:
 *   applications/convolution.c:34-66   fft_convolution          :71-96 circular_convolution
 *   applications/image_fft.c:35-72     fft_2d (array of row pointers, inverse scaled twice: once per 1-D
 *                                      transform inside radix2_dit_fft and once more at :63-71)
 *   applications/power_spectrum.c:133-159 autocorrelation_fft   :162-192 cross_correlation_fft
 * Each file carries its own demo main(); they are renamed away. Every file is one translation unit of its own
 * (this file is compiled three times with a different REF_APP_PART) because they share helper names. */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if REF_APP_PART == 1
#define main ref_convolution_demo_main
#include REF_CONV_C
#undef main
void oracle_ref_fft_convolution(const double* x, int nx, const double* h, int nh, double* y) {
    fft_convolution((complex_t*)x, nx, (complex_t*)h, nh, (complex_t*)y);
}
void oracle_ref_circular_convolution(const double* x, const double* h, int n, double* y) {
    circular_convolution((complex_t*)x, (complex_t*)h, n, (complex_t*)y);
}
#elif REF_APP_PART == 2
#define main ref_image_demo_main
#include REF_IMAGE_C
#undef main
/* row-major rows x cols in place, through the reference's array-of-row-pointers interface */
void oracle_ref_fft_2d(double* data, int rows, int cols, int dir) {
    complex_t** rp = (complex_t**)malloc(sizeof(complex_t*) * (size_t)rows);
    for (int i = 0; i < rows; i++) rp[i] = (complex_t*)data + (size_t)i * cols;
    fft_2d(rp, rows, cols, (fft_direction)dir);
    free(rp);
}
#else
#define main ref_psd_demo_main
#include REF_PSD_C
#undef main
void oracle_ref_autocorrelation(const double* x, int n, double* acf) {
    complex_t* r = autocorrelation_fft((complex_t*)x, n);
    memcpy(acf, r, sizeof(complex_t) * (size_t)n);
    free(r);
}
void oracle_ref_cross_correlation(const double* x, const double* y, int n, double* ccf) {
    complex_t* r = cross_correlation_fft((complex_t*)x, (complex_t*)y, n);
    memcpy(ccf, r, sizeof(complex_t) * (size_t)n);
    free(r);
}
#endif
