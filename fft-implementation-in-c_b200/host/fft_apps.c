/*
 * fft_apps.c - the FFT callers either side of the transform, on the GPU: convolution and correlation.
 *
 * The reference has these as CPU code inside its demo programs (each calls radix2_dit_fft directly):
 *   applications/convolution.c:34-66      fft_convolution       zero-pad to next_power_of_two(nx+nh-1), X*H, inverse
 *   applications/convolution.c:71-96      circular_convolution  length-n transforms, X*H, inverse
 *   applications/power_spectrum.c:162-192 cross_correlation_fft zero-pad to next_power_of_two(2n), conj(X)*Y, inverse
 *   applications/power_spectrum.c:133-159 autocorrelation_fft   the same with y = x
 * Here the two forward transforms are ONE batched engine plan (batch 2, both signals in one device buffer), the
 * spectral product is a device kernel (fftb200_pointwise_mul / _mul_conj - the kernel Bluestein uses), and the
 * inverse runs in place; only the inputs go up and the wanted samples come back. Same padding rules, same scaling
 * (inverse 1/n_fft), so results match the reference's functions to rounding (tests: <= 1e-12 relative L2).
 * Declared in include/fftb200_ext.h (additive: the reference's public headers have no such entry points).
 */
#include "../../include/fft_gpu.h"
#include "../../include/fftb200.h"
#include "../../include/fftb200_ext.h"

fftb200_plan* fftb200_host_make_plan(int n, int batch, int direction, int kind); /* fft_gpu.c */

/* out[0 .. n_out) = IFFT_nfft( op(FFT_nfft(a padded)) * FFT_nfft(b padded) ), op = conj when conj_a */
static int spectral_product(const complex_t* a, int na, const complex_t* b, int nb, int n_fft, int conj_a,
                            complex_t* out, int n_out) {
    if (!a || !b || !out || na <= 0 || nb <= 0 || na > n_fft || nb > n_fft || n_out > n_fft) return -1;
    const int kind = is_power_of_two(n_fft) ? FFTB200_C2C : FFTB200_BLUESTEIN;
    const int same = (a == b && na == nb);   /* autocorrelation: one transform is enough */
    int rc = -1;
    char* dev = NULL;
    fftb200_plan* fwd = fftb200_host_make_plan(n_fft, same ? 1 : 2, -1, kind);
    fftb200_plan* inv = fftb200_host_make_plan(n_fft, 1, 1, kind);
    const size_t bytes = sizeof(complex_t) * (size_t)n_fft;
    if (fwd && inv) dev = (char*)fftb200_malloc(2 * bytes);
    if (dev && fftb200_memset(dev, 0, 2 * bytes) == 0 &&
        fftb200_memcpy_h2d(dev, a, sizeof(complex_t) * (size_t)na) == 0 &&
        (same || fftb200_memcpy_h2d(dev + bytes, b, sizeof(complex_t) * (size_t)nb) == 0) &&
        fftb200_plan_exec(fwd, dev, dev) == 0) {
        const char* B = same ? dev : dev + bytes;
        rc = conj_a ? fftb200_pointwise_mul_conj(dev, dev, B, (size_t)n_fft) : fftb200_pointwise_mul(dev, dev, B, (size_t)n_fft);
        if (rc == 0) rc = fftb200_plan_exec(inv, dev, dev);
        if (rc == 0) rc = fftb200_memcpy_d2h(out, dev, sizeof(complex_t) * (size_t)n_out);
    }
    if (rc != 0) fprintf(stderr, "fft_gpu: spectral product: %s\n", fftb200_last_error());
    fftb200_free(dev);
    fftb200_plan_destroy(fwd);
    fftb200_plan_destroy(inv);
    return rc == 0 ? 0 : -1;
}

int fft_gpu_convolution(const complex_t* x, int nx, const complex_t* h, int nh, complex_t* y) {
    if (nx <= 0 || nh <= 0 || (long long)nx + nh - 1 > (1LL << 29)) return -1;
    return spectral_product(x, nx, h, nh, next_power_of_two(nx + nh - 1), 0, y, nx + nh - 1);
}

int fft_gpu_circular_convolution(const complex_t* x, const complex_t* h, int n, complex_t* y) {
    if (n <= 0) return -1;
    return spectral_product(x, n, h, n, n, 0, y, n);
}

int fft_gpu_cross_correlation(const complex_t* x, const complex_t* y, int n, complex_t* ccf) {
    if (n <= 0 || n > (1 << 28)) return -1;
    return spectral_product(x, n, y, n, next_power_of_two(2 * n), 1, ccf, n);
}

int fft_gpu_autocorrelation(const complex_t* x, int n, complex_t* acf) {
    return fft_gpu_cross_correlation(x, x, n, acf);
}
