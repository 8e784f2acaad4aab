/*
 * fft_common.h - shared types and small inline helpers of the public API.
 *
 * Drop-in for the reference's include/fft_common.h: same names, same types, same layout
 * (complex_t is C99 double _Complex = interleaved re, im doubles, reference :28; fft_direction
 * FFT_FORWARD = -1 / FFT_INVERSE = +1, reference :31-34). Programs written against the reference
 * header compile unchanged against this one. Not includable from C++/CUDA (C99 _Complex): the
 * device side sees the same bytes as double2 through include/fftb200.h.
 *
 * One deliberate difference: bit_reverse() here is correct for every width. The reference's
 * shortcut (:61-67) returns 0 for log2n <= 4, which breaks its own FFTs at N = 4, 8, 16.
 */
#ifndef FFT_COMMON_H
#define FFT_COMMON_H

#include <assert.h>
#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#ifdef __GNUC__
#define LIKELY(x) __builtin_expect(!!(x), 1)
#define UNLIKELY(x) __builtin_expect(!!(x), 0)
#define FORCE_INLINE __attribute__((always_inline)) inline
#else
#define LIKELY(x) (x)
#define UNLIKELY(x) (x)
#define FORCE_INLINE inline
#endif

#define PI 3.14159265358979323846
#define TWO_PI (2.0 * PI)

typedef double complex complex_t;

typedef enum { FFT_FORWARD = -1, FFT_INVERSE = 1 } fft_direction;

static FORCE_INLINE int is_power_of_two(int n) { return LIKELY(n > 0) && !(n & (n - 1)); }

static inline int next_power_of_two(int n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

static inline int log2_int(int n) {
    int l = 0;
    for (; n > 1; n >>= 1) l++;
    return l;
}

static inline unsigned int bit_reverse(unsigned int x, int log2n) {
    unsigned int r = 0;
    for (int b = 0; b < log2n; b++, x >>= 1) r = (r << 1) | (x & 1u);
    return r;
}

static inline complex_t* allocate_complex_array(int n) { return (complex_t*)calloc((size_t)n, sizeof(complex_t)); }
static inline void free_complex_array(complex_t* arr) { free(arr); }

/* exp(dir * 2*pi*i * k / n); the quarter-turn values are returned exactly, as in the reference (:89-98) */
static inline complex_t twiddle_factor(int k, int n, fft_direction dir) {
    if (k == 0) return 1.0;
    if (4 * k == n) return dir == FFT_FORWARD ? -I : I;
    if (2 * k == n) return -1.0;
    if (4 * k == 3 * n) return dir == FFT_FORWARD ? I : -I;
    double angle = dir * TWO_PI * k / n;
    return cexp(I * angle);
}

/* CPU-time stopwatch (clock()), kept for source compatibility; do not use it to time GPU work */
typedef struct { clock_t start; clock_t end; double elapsed_ms; } fft_timer_t;
static inline void timer_start(fft_timer_t* t) { t->start = clock(); }
static inline void timer_stop(fft_timer_t* t) {
    t->end = clock();
    t->elapsed_ms = 1000.0 * (double)(t->end - t->start) / CLOCKS_PER_SEC;
}

#define CHECK_NULL(ptr, msg) \
    if (!(ptr)) { fprintf(stderr, "Error: %s\n", msg); exit(EXIT_FAILURE); }
#define CHECK_POWER_OF_TWO(n) \
    if (!is_power_of_two(n)) { fprintf(stderr, "Error: Size %d is not a power of two\n", n); exit(EXIT_FAILURE); }

static inline void print_complex(complex_t c) {
    double re = fabs(creal(c)) < 1e-10 ? 0.0 : creal(c), im = fabs(cimag(c)) < 1e-10 ? 0.0 : cimag(c);
    printf("(%.3f, %.3fi)", re, im);
}
static inline void print_complex_array(const char* label, complex_t* arr, int n) {
    printf("%s: ", label);
    for (int i = 0; i < n; i++) { print_complex(arr[i]); printf(" "); }
    printf("\n");
}

static inline void generate_sine_wave(complex_t* signal, int n, double freq, double fs) {
    for (int i = 0; i < n; i++) signal[i] = sin(TWO_PI * freq * i / fs);
}
static inline void generate_square_wave(complex_t* signal, int n, double freq, double fs) {
    int period = (int)(fs / freq);
    for (int i = 0; i < n; i++) signal[i] = (i % period < period / 2) ? 1.0 : -1.0;
}
static inline void generate_impulse(complex_t* signal, int n) {
    memset(signal, 0, (size_t)n * sizeof(complex_t));
    signal[0] = 1.0;
}

static inline double* compute_magnitude(complex_t* x, int n) {
    double* m = (double*)malloc((size_t)n * sizeof(double));
    CHECK_NULL(m, "Failed to allocate magnitude array");
    for (int i = 0; i < n; i++) m[i] = cabs(x[i]);
    return m;
}
static inline double* compute_phase(complex_t* x, int n) {
    double* p = (double*)malloc((size_t)n * sizeof(double));
    CHECK_NULL(p, "Failed to allocate phase array");
    for (int i = 0; i < n; i++) p[i] = carg(x[i]);
    return p;
}
static inline double* compute_power_spectrum(complex_t* x, int n) {
    double* p = (double*)malloc((size_t)n * sizeof(double));
    CHECK_NULL(p, "Failed to allocate power spectrum array");
    for (int i = 0; i < n; i++) { double a = cabs(x[i]); p[i] = a * a / n; }
    return p;
}

#endif /* FFT_COMMON_H */
