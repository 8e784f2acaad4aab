/*
 * demo_dist.c - the distributed transform's C99 host API (include/fftb200_dist.h) called from plain C, one thread per GPU
 * inside one process: every thread owns a device and a block of the input, the all-gather callback the library asks for is a
 * pthread barrier over a shared array, and the result is checked against the single-GPU plan of the whole transform
 * (fft_gpu_plan_1d, itself checked against the reference oracle by the test suite).
 *
 * usage: demo_dist [log_n] [gpus]     (gpus: a power of two <= the devices present; default: 1)
 */
#define _POSIX_C_SOURCE 200809L
#include "fft_gpu.h"
#include "fftb200.h"
#include "fftb200_ext.h"
#include "fftb200_dist.h"
#include <pthread.h>

static int g_world = 1, g_log_n = 20, g_dir = -1;
static unsigned char g_board[64 * 256];
static pthread_barrier_t g_bar;
static void* g_full_in;    /* device 0: the whole input  */
static double g_err[64];

static int allgather(void* ctx, const void* send, void* recv, size_t bytes) {
    const int rank = *(int*)ctx;
    if (bytes > 256) return -1;   /* the library exchanges 160 bytes per rank */
    memcpy(g_board + 256 * rank, send, bytes);
    pthread_barrier_wait(&g_bar);
    for (int g = 0; g < g_world; g++) memcpy((char*)recv + bytes * g, g_board + 256 * g, bytes);
    pthread_barrier_wait(&g_bar);
    return 0;
}

static void* rank_main(void* arg) {
    int rank = *(int*)arg;
    g_err[rank] = -1.0;
    if (fft_gpu_set_device(rank) != 0) return NULL;
    const size_t n = (size_t)1 << g_log_n, nloc = n / (size_t)g_world;
    fftb200_dist* plan = NULL;
    if (fftb200_dist_create(&plan, g_log_n, g_world, rank, g_dir, 0, allgather, &rank) != 0) {
        fprintf(stderr, "rank %d: %s\n", rank, fftb200_last_error());
        return NULL;
    }
    if (rank == 0) printf("%s\n", fftb200_dist_describe(plan));
    void* d_in = fftb200_malloc(nloc * sizeof(complex_t));
    if (!d_in || fftb200_fill_splitmix(d_in, 45, (unsigned long long)rank * nloc, nloc) != 0) return NULL;
    void* d_out = NULL;
    for (int rep = 0; rep < 2; rep++) {   /* twice: the second run re-uses every buffer */
        if (fftb200_dist_exec_async(plan, d_in, &d_out) != 0 || fftb200_dist_sync(plan) != 0) {
            fprintf(stderr, "rank %d: %s\n", rank, fftb200_last_error());
            return NULL;
        }
    }
    complex_t* mine = (complex_t*)malloc(nloc * sizeof(complex_t));
    complex_t* want = (complex_t*)malloc(nloc * sizeof(complex_t));
    pthread_barrier_wait(&g_bar);   /* rank 0 has finished the single-GPU plan before anybody reads g_full_in */
    if (!mine || !want || fftb200_memcpy_d2h(mine, d_out, nloc * sizeof(complex_t)) != 0) return NULL;
    fft_gpu_set_device(0);
    if (fftb200_memcpy_d2h(want, (char*)g_full_in + (size_t)rank * nloc * sizeof(complex_t), nloc * sizeof(complex_t)) != 0) return NULL;
    fft_gpu_set_device(rank);
    double num = 0.0, den = 0.0;
    for (size_t i = 0; i < nloc; i++) {
        const double dr = creal(mine[i]) - creal(want[i]), di = cimag(mine[i]) - cimag(want[i]);
        num += dr * dr + di * di;
        den += creal(want[i]) * creal(want[i]) + cimag(want[i]) * cimag(want[i]);
    }
    g_err[rank] = sqrt(num / den);
    free(mine); free(want);
    fftb200_free(d_in);
    fftb200_dist_destroy(plan);
    return NULL;
}

int main(int argc, char** argv) {
    if (argc > 1) g_log_n = atoi(argv[1]);
    if (argc > 2) g_world = atoi(argv[2]);
    if (!fft_gpu_available()) { printf("no GPU: nothing to do (the library has no CPU fallback)\n"); return 2; }
    if (fft_gpu_init(FFT_GPU_AUTO) != 0) return 1;
    if (g_world < 1 || g_world > 64 || (g_world & (g_world - 1)) || g_world > fftb200_device_count()) { printf("gpus must be a power of two <= %d\n", fftb200_device_count()); return 1; }
    const size_t n = (size_t)1 << g_log_n;
    /* the yardstick: the single-GPU plan of the whole transform on device 0 */
    fft_gpu_set_device(0);
    fft_gpu_memory_t full = fft_gpu_alloc(n);
    fft_gpu_plan_t p1 = fft_gpu_plan_1d((int)n, 1, FFT_FORWARD);
    if (!full || !p1) { printf("setup failed: %s\n", fftb200_last_error()); return 1; }
    g_full_in = fftb200_devptr_of(full);
    fftb200_fill_splitmix(g_full_in, 45, 0, n);
    fft_gpu_execute(p1, full, full);
    pthread_barrier_init(&g_bar, NULL, (unsigned)g_world);
    pthread_t th[64];
    int ranks[64];
    for (int g = 0; g < g_world; g++) { ranks[g] = g; pthread_create(&th[g], NULL, rank_main, &ranks[g]); }
    double worst = 0.0;
    int ok = 1;
    for (int g = 0; g < g_world; g++) {
        pthread_join(th[g], NULL);
        printf("rank %d: rel L2 against the single-GPU plan %.3e\n", g, g_err[g]);
        if (g_err[g] < 0.0 || g_err[g] > 1e-12) ok = 0;
        if (g_err[g] > worst) worst = g_err[g];
    }
    fft_gpu_destroy_plan(p1);
    fft_gpu_free(full);
    printf("%s: one 2^%d-point transform over %d GPU(s) through fftb200_dist_*, worst block %.3e\n", ok ? "PASS" : "FAILED", g_log_n, g_world, worst);
    return ok ? 0 : 1;
}
