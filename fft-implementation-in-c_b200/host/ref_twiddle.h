/* ref_twiddle.h - host-side generators for the tables the kernels consume (internal to the host library). */
#ifndef FFTB200_REF_TWIDDLE_H
#define FFTB200_REF_TWIDDLE_H
#include <stddef.h>

/* Forward per-stage twiddle tables for power-of-two n (n - 1 complex entries, interleaved doubles):
 * entry (stage s, j) at 2^(s-1) - 1 + j. The table of a larger n contains every smaller one as a
 * prefix, so one process-wide cache serves all plans. Returns NULL on allocation failure.
 * The returned pointer stays valid until fftb200_host_tables_release(). */
const double* fftb200_host_twiddles(int n);
/* Correctly rounded exp(-2*pi*i*j/2^s) in the same layout, for n = 8192 (stages m <= 8192): consumed by the
 * kernels for the early stages, where the reference's recurrence is still within 3e-14 of these values.
 * NULL when FFTB200_TWIDDLE=ref asks for the reference recurrence in every stage. */
const double* fftb200_host_twiddles_accurate(int* n_out);
/* Bluestein chirp c[k] = exp(i * (-dir * pi * k^2 / n)), k < n, for dir = -1 / +1. */
void fftb200_host_chirp(double* out, int n, int dir);
void fftb200_host_tables_release(void);
/* FFTB200_TWIDDLE=accurate switches the tables to correctly rounded exp(-2*pi*i*j/2^s) (ablation). */
int fftb200_host_twiddle_mode_accurate(void);

#endif
