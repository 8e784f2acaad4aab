/*
 * ref_twiddle.c - twiddle and chirp tables that make the GPU transform reproduce the reference's
 * numbers, not just the mathematically exact DFT.
 *
 * The reference never tabulates twiddles: inside every stage it advances w <- w * w_m serially
 * (algorithms/core/radix2_dit.c:93,109; same loop in radix4.c:110-125 and split_radix.c:39-54), with
 * w_m = twiddle_factor(1, m, dir) (include/fft_common.h:89-98). The rounded root w_m is raised to
 * powers up to m/2, so its rounding error grows linearly along the stage: the reference is off from
 * the exact DFT by 9e-12 (relative L2) at N = 2^20 and 1.6e-10 at 2^24 - far more than the 1e-12
 * parity bar. The kernels therefore read w from a table built here by the same recurrence.
 *
 * Bit-faithfulness: the reference is built with -O3 -ffast-math and FMA contraction (its Makefile:7).
 * gcc 13 compiles `w *= w_m` to
 *      re' = fma(re, m_re, -(im * m_im))        im' = fma(re, m_im, im * m_re)
 * (disassembly of radix2_dit_fft in oracle/_ref/libfftref.so), and cexp(I*angle) to sincos(angle).
 * That sequence is written out explicitly below and THIS FILE IS COMPILED WITHOUT -ffast-math and with
 * -ffp-contract=off, so the compiler cannot re-associate or re-contract it. tests/test_oracle.py checks
 * the tables bit-for-bit against the oracle's recurrence.
 *
 * Likewise the Bluestein chirp phase -dir*PI*k*k/n (algorithms/core/bluestein.c:59-62) is evaluated by
 * the reference's build as (k*k) * ((-dir*PI) * (1/n)); at n ~ 1e6 the phase is ~3e6 rad, so a
 * different association changes the result at the 1e-10 level. It is written out the same way here.
 */
#define _GNU_SOURCE
#include "ref_twiddle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#define REF_PI 3.14159265358979323846 /* include/fft_common.h:24 of the reference */

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static double* g_tab = NULL; /* interleaved, (g_n - 1) complex */
static int g_n = 0;
static int g_mode = -1;      /* 0 reference recurrence, 1 accurate */
static double* g_old[40];    /* superseded (smaller) tables: plans created earlier may still read them */
static int g_nold = 0;

int fftb200_host_twiddle_mode_accurate(void) {
    if (g_mode < 0) {
        const char* e = getenv("FFTB200_TWIDDLE");
        g_mode = (e && strcmp(e, "accurate") == 0) ? 1 : 0;
    }
    return g_mode;
}

/* stage root for the forward direction: twiddle_factor(1, m, FFT_FORWARD) */
static void stage_root(int m, double* re, double* im) {
    if (m == 2) { *re = -1.0; *im = 0.0; return; }
    if (m == 4) { *re = 0.0; *im = -1.0; return; }
    double angle = (-1.0 * (2.0 * REF_PI)) * 1.0 / (double)m;
    double s, c;
    sincos(angle, &s, &c);
    *re = c; *im = s;
}

static void fill_stage(double* t, int s) {
    const int half = 1 << (s - 1);
    if (fftb200_host_twiddle_mode_accurate()) {
        for (int j = 0; j < half; j++) {
            long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)(2 * half);
            t[2 * j] = (double)cosl(a);
            t[2 * j + 1] = (double)sinl(a);
        }
        return;
    }
    double mr, mi, wr = 1.0, wi = 0.0;
    stage_root(2 * half, &mr, &mi);
    for (int j = 0; j < half; j++) {
        t[2 * j] = wr;
        t[2 * j + 1] = wi;
        const double p_im_mi = wi * mi, p_im_mr = wi * mr;
        const double nr = fma(wr, mr, -p_im_mi);
        const double ni = fma(wr, mi, p_im_mr);
        wr = nr; wi = ni;
    }
}

#define ACC_N 8192
static double* g_acc = NULL;

const double* fftb200_host_twiddles_accurate(int* n_out) {
    const char* e = getenv("FFTB200_TWIDDLE");
    if (e && strcmp(e, "ref") == 0) return NULL;
    pthread_mutex_lock(&g_mu);
    if (!g_acc) {
        double* t = (double*)malloc(sizeof(double) * 2 * (ACC_N - 1));
        if (t) {
            for (int s = 1; (1 << s) <= ACC_N; s++) {
                const int half = 1 << (s - 1);
                double* ts = t + 2 * ((size_t)half - 1);
                for (int j = 0; j < half; j++) {
                    /* exact quarter-turn symmetry by construction: entry j + half/2 = -i * entry j */
                    if (s >= 2 && j >= half / 2) {
                        ts[2 * j] = ts[2 * (j - half / 2) + 1];
                        ts[2 * j + 1] = -ts[2 * (j - half / 2)];
                        continue;
                    }
                    long double a = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)(2 * half);
                    ts[2 * j] = (double)cosl(a);
                    ts[2 * j + 1] = (double)sinl(a);
                }
            }
            t[0] = 1.0; t[1] = 0.0;                 /* stage 1: w = 1 */
            if (ACC_N >= 4) { t[2] = 1.0; t[3] = 0.0; t[4] = 0.0; t[5] = -1.0; } /* stage 2: 1, -i exactly */
            g_acc = t;
        }
    }
    const double* r = g_acc;
    pthread_mutex_unlock(&g_mu);
    if (n_out) *n_out = ACC_N;
    return r;
}

const double* fftb200_host_twiddles(int n) {
    if (n < 1 || (n & (n - 1))) return NULL;
    pthread_mutex_lock(&g_mu);
    if (g_n < n) {
        double* t = (double*)malloc(sizeof(double) * 2 * (size_t)(n > 1 ? n - 1 : 1));
        if (!t) { pthread_mutex_unlock(&g_mu); return NULL; }
        int have = 0;
        if (g_tab && g_n > 1) { memcpy(t, g_tab, sizeof(double) * 2 * (size_t)(g_n - 1)); have = g_n; }
        for (int s = 1; (1 << s) <= n && s < 31; s++) {
            if ((1 << s) <= have) continue;
            fill_stage(t + 2 * ((size_t)(1 << (s - 1)) - 1), s);
        }
        /* Superseded tables stay allocated until release: a plan being created on another thread may
         * still be uploading from one. They are prefixes of the new table, at most as large in total. */
        if (g_tab && g_nold < 40) g_old[g_nold++] = g_tab;
        g_tab = t;
        g_n = n;
    }
    const double* r = g_tab;
    pthread_mutex_unlock(&g_mu);
    return r;
}

/* Rank-specific table for the tail plan of a distributed transform of 2^log_total points over 2^log_world ranks
 * with M = 2^log_m points in the first pass: laid out like the table of a 2^(log_total - log_world)-point
 * transform, but entry (stage s', j' = k_loc + M' q), M' = M / world, holds the reference's
 * T[s' + log_world][k_loc + rank * M' + M q] for every stage s' > log_m - log_world. Earlier stages are not read
 * by the tail plan and are left zero. The recurrence is serial per stage, so each rank walks all of it (about
 * 2^log_total steps in total) and keeps its 1/world share. Returns 0, -1 on bad arguments. */
int fftb200_host_twiddles_dist(double* out, int log_total, int log_world, int rank, int log_m) {
    if (!out || log_total < 2 || log_total > 30 || log_world < 0 || log_m < log_world || log_m >= log_total ||
        rank < 0 || rank >= (1 << log_world)) return -1;
    const int log_local = log_total - log_world;
    const long long mloc = 1LL << (log_m - log_world), m = 1LL << log_m, k0 = (long long)rank * mloc;
    memset(out, 0, sizeof(double) * 2 * (size_t)((1LL << log_local) - 1));
    for (int s = log_m + 1; s <= log_total; s++) {
        const long long half = 1LL << (s - 1);
        double* t = out + 2 * ((size_t)(1LL << (s - log_world - 1)) - 1);
        double mr, mi, wr = 1.0, wi = 0.0;
        stage_root((int)(2 * half), &mr, &mi);   /* 2 * half <= 2^30 fits an int */
        for (long long j = 0; j < half; j++) {
            const long long k = j & (m - 1);
            if (k >= k0 && k < k0 + mloc) {
                const long long jj = (k - k0) + mloc * (j >> log_m);
                t[2 * jj] = wr;
                t[2 * jj + 1] = wi;
            }
            const double p_im_mi = wi * mi, p_im_mr = wi * mr;
            const double nr = fma(wr, mr, -p_im_mi);
            const double ni = fma(wr, mi, p_im_mr);
            wr = nr; wi = ni;
        }
    }
    return 0;
}

void fftb200_host_tables_release(void) {
    pthread_mutex_lock(&g_mu);
    for (int i = 0; i < g_nold; i++) free(g_old[i]);
    g_nold = 0;
    free(g_tab);
    free(g_acc);
    g_acc = NULL;
    g_tab = NULL;
    g_n = 0;
    pthread_mutex_unlock(&g_mu);
}

void fftb200_host_chirp(double* out, int n, int dir) {
    const double scale = (double)(-dir) * REF_PI;
    const double rn = 1.0 / (double)n;
    const double c0 = scale * rn;
    for (int k = 0; k < n; k++) {
        const double k2 = (double)k * (double)k;
        const double phase = k2 * c0;
        double s, c;
        sincos(phase, &s, &c);
        out[2 * k] = c;
        out[2 * k + 1] = s;
    }
}
