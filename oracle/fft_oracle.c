/*
 * fft_oracle.c - CPU restatement of the reference FFT hot path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is linked, imported or executed by the product
 * (fft-implementation-in-c_b200/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the timed CPU baseline.
 *
 * Parity status: PINNED. tests/test_oracle.py checks every function here against the unmodified
 * reference compiled into oracle/_ref/libfftref.so (rel-L2 <= 1e-15, bit-exact in practice), against
 * committed fixtures in tests/golden/ generated from that library, and against the properties the
 * reference's own tests/test_all.c:64-351 assert (impulse, DC, linearity, Parseval, round trip, tone).
 *
 * Build with the reference's arithmetic flags (-O3 -ffast-math, FMA contraction; see oracle/Makefile):
 * the twiddle recurrence below is only bit-faithful when the complex multiply contracts the same way.
 *
 * All arrays are interleaved (re, im) doubles == the reference's complex_t (include/fft_common.h:28).
 * dir: -1 forward, +1 inverse (include/fft_common.h:31-34).
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double complex cplx;

#define ORACLE_PI 3.14159265358979323846 /* include/fft_common.h:24 */

static int ilog2(int n) { int l = 0; while (n >>= 1) l++; return l; }

/* Plain bit reversal: the general branch of include/fft_common.h:70-76. */
static unsigned bitrev_plain(unsigned x, int log2n) {
    unsigned r = 0;
    for (int i = 0; i < log2n; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* The reference's bit_reverse including its small-size shortcut (include/fft_common.h:59-77).
 * For log2n <= 4 the 16-bit byte swap is skipped but the shift is still 16-log2n, so the
 * function returns 0 for every index: N in {4, 8, 16} are NOT permuted by the reference. */
static unsigned bitrev_reference(unsigned x, int log2n) {
    if (log2n <= 8) {
        x = ((x & 0xAAAA) >> 1) | ((x & 0x5555) << 1);
        x = ((x & 0xCCCC) >> 2) | ((x & 0x3333) << 2);
        x = ((x & 0xF0F0) >> 4) | ((x & 0x0F0F) << 4);
        if (log2n > 4) x = ((x & 0xFF00) >> 8) | ((x & 0x00FF) << 8);
        return x >> (16 - log2n);
    }
    return bitrev_plain(x, log2n);
}

/* Stage root of unity: include/fft_common.h:89-98 with k = 1. Exact for m = 2 and m = 4. */
static cplx stage_root(int m, int dir) {
    if (4 == m) return (dir < 0) ? -I : I;
    if (2 == m) return -1.0;
    double angle = dir * (2.0 * ORACLE_PI) * 1 / m;
    return cexp(I * angle);
}

/*
 * Radix-2 DIT, in place: algorithms/core/radix2_dit.c:59-120. radix4.c:83-134 and
 * split_radix.c:23-70 run the identical loop, so this one function is the oracle for every
 * power-of-two size fft_auto can route (fft_auto.c:136-172, 250-262).
 * quirk != 0 reproduces the reference's missing permutation at N in {4, 8, 16}.
 * Returns -1 (and leaves x untouched) for non power-of-two n; the reference exit()s there
 * (include/fft_common.h:123-127).
 */
int oracle_fft_pow2(double* xd, int n, int dir, int quirk) {
    if (n <= 0 || (n & (n - 1))) return -1;
    cplx* x = (cplx*)xd;
    int log2n = ilog2(n);
    for (int i = 0; i < n; i++) {                                  /* radix2_dit.c:70-77 */
        int j = (int)(quirk ? bitrev_reference((unsigned)i, log2n) : bitrev_plain((unsigned)i, log2n));
        if (i < j) { cplx t = x[i]; x[i] = x[j]; x[j] = t; }
    }
    for (int stage = 1; stage <= log2n; stage++) {                 /* radix2_dit.c:84-112 */
        int m = 1 << stage, half = m >> 1;
        cplx w_m = stage_root(m, dir);
        for (int k = 0; k < n; k += m) {
            cplx w = 1.0;
            for (int j = 0; j < half; j++) {
                int t = k + j, u = t + half;
                cplx temp = x[u] * w;
                x[u] = x[t] - temp;
                x[t] = x[t] + temp;
                w *= w_m;                                          /* the serial twiddle recurrence */
            }
        }
    }
    if (dir > 0) for (int i = 0; i < n; i++) x[i] /= n;            /* radix2_dit.c:115-119 */
    return 0;
}

/* Batched form of the above: transform b occupies [b*n, (b+1)*n) (layout of gpu/fft_cuda.cu:152-156).
 * threads > 1 runs transforms in parallel (OpenMP; not in the reference - used for the CPU baseline). */
int oracle_fft_pow2_batch(double* x, int n, long batch, int dir, int threads) {
    if (n <= 0 || (n & (n - 1))) return -1;
    if (threads < 1) threads = 1;
    #pragma omp parallel for num_threads(threads) schedule(static)
    for (long b = 0; b < batch; b++) oracle_fft_pow2(x + 2 * (size_t)b * (size_t)n, n, dir, 0);
    return 0;
}

/*
 * Per-stage twiddle tables exactly as the recurrence at radix2_dit.c:89-109 produces them:
 * stage s (m = 2^s) contributes m/2 entries starting at flat offset (m/2 - 1); n - 1 entries in total.
 * Forward direction; the inverse tables are the bitwise conjugates.
 */
void oracle_twiddle_tables(double* td, int n) {
    cplx* T = (cplx*)td;
    int log2n = ilog2(n);
    for (int s = 1; s <= log2n; s++) {
        int half = 1 << (s - 1);
        cplx w_m = stage_root(1 << s, -1), w = 1.0;
        for (int j = 0; j < half; j++) { T[half - 1 + j] = w; w *= w_m; }
    }
}

/* Bluestein chirp: algorithms/core/bluestein.c:51-65 (phase evaluated left to right in double). */
void oracle_chirp(double* cd, int n, int dir) {
    cplx* c = (cplx*)cd;
    /* The source expression is -dir*PI*k*k/n; under the reference's -ffast-math gcc 13 evaluates it as
     * (k*k) * ((-dir*PI) * (1/n)) (disassembly of bluestein_fft in oracle/_ref/libfftref.so). At
     * n ~ 1e6 the phase reaches ~3e6 rad, so the association order decides bits at the 1e-10 level:
     * it is written out here, and `volatile` keeps this file's own -ffast-math from re-associating. */
    volatile double scale = (double)(-dir) * ORACLE_PI;
    volatile double rn = 1.0 / (double)n;
    volatile double c0 = scale * rn;
    for (int k = 0; k < n; k++) {
        volatile double k2 = (double)k * (double)k;
        double phase = k2 * c0;
        c[k] = cos(phase) + I * sin(phase);
    }
}

static int next_pow2(int n) {                                       /* include/fft_common.h:41-50 */
    n--; n |= n >> 1; n |= n >> 2; n |= n >> 4; n |= n >> 8; n |= n >> 16; n++;
    return n;
}

/* Bluestein chirp-z transform for arbitrary n, in place: algorithms/core/bluestein.c:79-155. */
int oracle_bluestein(double* xd, int n, int dir) {
    if (!xd || n <= 0) return -1;
    cplx* x = (cplx*)xd;
    int m = next_pow2(2 * n - 1);
    cplx* a = calloc((size_t)m, sizeof(cplx));
    cplx* b = calloc((size_t)m, sizeof(cplx));
    cplx* chirp = calloc((size_t)n, sizeof(cplx));
    if (!a || !b || !chirp) { free(a); free(b); free(chirp); return -1; }
    oracle_chirp((double*)chirp, n, dir);
    for (int k = 0; k < n; k++) a[k] = x[k] * conj(chirp[k]);      /* :107-109 */
    for (int k = 0; k < n; k++) {                                  /* :116-121 */
        b[k] = chirp[k];
        if (k > 0) b[m - k] = chirp[k];
    }
    /* the reference calls radix2_dit_fft here, i.e. with its small-size quirk (m <= 16 <=> n <= 8) */
    oracle_fft_pow2((double*)a, m, -1, 1);                         /* :124-125 */
    oracle_fft_pow2((double*)b, m, -1, 1);
    for (int k = 0; k < m; k++) a[k] *= b[k];                      /* :128-130 */
    oracle_fft_pow2((double*)a, m, +1, 1);                         /* :133 */
    for (int k = 0; k < n; k++) x[k] = a[k] * conj(chirp[k]);      /* :139-141 */
    if (dir > 0) for (int k = 0; k < n; k++) x[k] /= n;            /* :144-148 */
    free(a); free(b); free(chirp);
    return 0;
}

/* What fft_auto computes for size n (fft_auto.c:136-172): pow2 -> radix-2 DIT loop, else Bluestein
 * (mixed-radix sizes are out of scope: SURVEY.md section 2 row 6). quirk as in oracle_fft_pow2. */
int oracle_fft_auto(const double* in, double* out, int n, int sign, int quirk) {
    if (n <= 0 || !in || !out) return -1;
    if (in != out) memcpy(out, in, (size_t)n * 2 * sizeof(double));
    int dir = sign < 0 ? -1 : 1;
    if ((n & (n - 1)) == 0) return oracle_fft_pow2(out, n, dir, quirk);
    return oracle_bluestein(out, n, dir);
}

/* r2c as the reference intends it (fft_auto.c:391-399 + fft_auto.h:89-97): promote the real input to
 * complex, run the forward c2c, keep bins 0 .. n/2. (The reference itself frees the promoted buffer
 * before executing - use-after-free at fft_auto.c:400 - so this is the only meaningful reading.) */
int oracle_r2c(const double* in, double* out_half, int n) {
    cplx* t = malloc((size_t)n * sizeof(cplx));
    if (!t) return -1;
    for (int i = 0; i < n; i++) t[i] = in[i];
    int rc = oracle_fft_auto((double*)t, (double*)t, n, -1, 0);
    if (rc == 0) memcpy(out_half, t, (size_t)(n / 2 + 1) * sizeof(cplx));
    free(t);
    return rc;
}

/* ---- the callers either side of the transform (SURVEY.md 8f "next" rows) ------------------------------- */

/* c2r: the reference only declares it (fft_auto.h:99-107, stub at fft_auto.c:405-408). Defined here as the inverse
 * of oracle_r2c in the reference's own arithmetic: rebuild the Hermitian spectrum X[n-k] = conj(X[k]) from bins
 * 0 .. n/2, run the reference's inverse c2c (scaled 1/n, radix2_dit.c:115-119) and keep the real parts. */
int oracle_c2r(const double* in_half, double* out, int n) {
    if (n <= 0 || (n & (n - 1))) return -1;
    const cplx* h = (const cplx*)in_half;
    cplx* t = malloc((size_t)n * sizeof(cplx));
    if (!t) return -1;
    for (int k = 0; k < n; k++) t[k] = k <= n / 2 ? h[k] : conj(h[n - k]);
    int rc = oracle_fft_pow2((double*)t, n, 1, 0);
    if (rc == 0) for (int i = 0; i < n; i++) out[i] = creal(t[i]);
    free(t);
    return rc;
}

/* 2-D transform, row-column decomposition of applications/image_fft.c:35-72 on a row-major rows x cols array:
 * every row, then every column, with the 1-D oracle. double_scale != 0 reproduces the reference's inverse, which
 * divides by rows*cols a second time (:63-71) after each 1-D inverse already scaled; double_scale == 0 is the
 * contract of the public headers (inverse scaled by 1/(rows*cols) once), which is what the product implements.
 * quirk != 0 reproduces the reference's missing permutation when rows or cols is 4, 8 or 16 (see oracle_fft_pow2). */
int oracle_fft2d(double* data, int rows, int cols, int dir, int double_scale, int quirk) {
    if (rows <= 0 || cols <= 0 || (rows & (rows - 1)) || (cols & (cols - 1))) return -1;
    cplx* x = (cplx*)data;
    for (int i = 0; i < rows; i++) oracle_fft_pow2((double*)(x + (size_t)i * cols), cols, dir, quirk);   /* :41-43 */
    cplx* col = malloc((size_t)rows * sizeof(cplx));
    if (!col) return -1;
    for (int j = 0; j < cols; j++) {                                                                  /* :46-60 */
        for (int i = 0; i < rows; i++) col[i] = x[(size_t)i * cols + j];
        oracle_fft_pow2((double*)col, rows, dir, quirk);
        for (int i = 0; i < rows; i++) x[(size_t)i * cols + j] = col[i];
    }
    free(col);
    if (dir > 0 && double_scale) {
        double scale = 1.0 / (rows * cols);
        for (size_t i = 0; i < (size_t)rows * cols; i++) x[i] *= scale;
    }
    return 0;
}

/* Shared body of the spectral-product callers: zero-pad a and b to n_fft, forward transforms, pointwise product
 * (conj_a != 0: conj(A) * B), inverse transform, first n_out samples. */
static int spectral_product(const cplx* a, int na, const cplx* b, int nb, int n_fft, int conj_a, cplx* out, int n_out) {
    cplx* A = calloc((size_t)n_fft, sizeof(cplx));
    cplx* B = calloc((size_t)n_fft, sizeof(cplx));
    if (!A || !B) { free(A); free(B); return -1; }
    memcpy(A, a, (size_t)na * sizeof(cplx));
    memcpy(B, b, (size_t)nb * sizeof(cplx));
    int rc = oracle_fft_auto((double*)A, (double*)A, n_fft, -1, 0);
    if (rc == 0) rc = oracle_fft_auto((double*)B, (double*)B, n_fft, -1, 0);
    if (rc == 0) {
        for (int i = 0; i < n_fft; i++) A[i] = conj_a ? conj(A[i]) * B[i] : A[i] * B[i];
        rc = oracle_fft_auto((double*)A, (double*)A, n_fft, 1, 0);
    }
    if (rc == 0) memcpy(out, A, (size_t)n_out * sizeof(cplx));
    free(A); free(B);
    return rc;
}

/* applications/convolution.c:34-66: linear convolution, y has nx + nh - 1 samples */
int oracle_convolution(const double* x, int nx, const double* h, int nh, double* y) {
    if (nx <= 0 || nh <= 0) return -1;
    return spectral_product((const cplx*)x, nx, (const cplx*)h, nh, next_pow2(nx + nh - 1), 0, (cplx*)y, nx + nh - 1);
}
/* applications/convolution.c:71-96: circular convolution of two length-n signals (the reference needs a power of
 * two; any n here, through the Bluestein oracle) */
int oracle_circular_convolution(const double* x, const double* h, int n, double* y) {
    if (n <= 0) return -1;
    return spectral_product((const cplx*)x, n, (const cplx*)h, n, n, 0, (cplx*)y, n);
}
/* applications/power_spectrum.c:162-192: cross-correlation, conj(X) * Y, padded to next_pow2(2n), first n lags */
int oracle_cross_correlation(const double* x, const double* y, int n, double* ccf) {
    if (n <= 0) return -1;
    return spectral_product((const cplx*)x, n, (const cplx*)y, n, next_pow2(2 * n), 1, (cplx*)ccf, n);
}
/* applications/power_spectrum.c:133-159: autocorrelation, X * conj(X) */
int oracle_autocorrelation(const double* x, int n, double* acf) {
    return oracle_cross_correlation(x, x, n, acf);
}

/* Independent O(n^2) truth for small n (algorithms/dft/naive_dft.c:55-97): used where the reference's
 * own FFT is wrong (N in {4, 8, 16}). Angles reduced with k*j mod n so it stays accurate. */
void oracle_naive_dft(const double* ind, double* outd, int n, int dir) {
    const cplx* in = (const cplx*)ind; cplx* out = (cplx*)outd;
    for (int k = 0; k < n; k++) {
        long double sr = 0, si = 0;
        for (int j = 0; j < n; j++) {
            long long kj = ((long long)k * j) % n;
            long double ang = dir * 2.0L * 3.14159265358979323846264338327950288L * (long double)kj / n;
            long double c = cosl(ang), s = sinl(ang);
            sr += creal(in[j]) * c - cimag(in[j]) * s;
            si += creal(in[j]) * s + cimag(in[j]) * c;
        }
        if (dir > 0) { sr /= n; si /= n; }
        out[k] = (double)sr + I * (double)si;
    }
}

/* Counter-based synthetic input: element i of stream `seed` is
 * re = u(splitmix64(seed + 2i)), im = u(splitmix64(seed + 2i + 1)), u(z) = (z >> 11) * 2^-52 - 1. */
static uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
void oracle_fill(double* x, uint64_t seed, uint64_t first_elem, uint64_t count) {
    for (uint64_t i = 0; i < count; i++) {
        uint64_t e = first_elem + i;
        x[2 * i]     = (double)(splitmix64(seed + 2 * e) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
        x[2 * i + 1] = (double)(splitmix64(seed + 2 * e + 1) >> 11) * (1.0 / 4503599627370496.0) - 1.0;
    }
}
