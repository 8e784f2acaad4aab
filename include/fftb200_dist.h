/*
 * fftb200_dist.h - one power-of-two c2c transform distributed over 2^g GPUs, one process (or thread) per GPU.
 *
 * Additive to the reference's API (BASELINE config 4: N = 2^30 over 8 B200). The reference has no
 * multi-device code; this is the four-step split N = R * M of ITS algorithm (radix-2 DIT stage product,
 * algorithms/core/radix2_dit.c:59-120, with the twiddle recurrence of :93,109 for the late stages), so the result is the
 * reference's result on the concatenated input. Rank s holds elements [s N/G, (s+1) N/G) of the input and receives the
 * same range of the natural-order output.
 *
 *   T0    push kernel: rank g gets the columns r in its range of the row-major [M][R] view (P2P stores)
 *   head  stages 1 .. log2 M on the local [M][R/G] array; the final scatter of its last pass stores into the peers (T1)
 *   tail  stages log2 M + 1 .. log2 N with the rank's share of the reference's late-stage twiddles; the final scatter
 *         of its last pass stores the natural-order blocks into the peers (T2)
 *
 * All data movement between GPUs is done by the kernels over NVLink peer memory (CUDA IPC mappings); the phases are
 * separated by a stream-ordered barrier through peer flags (fftb200_barrier_*). No collective library is linked: the only
 * thing the caller provides is an all-gather of a few hundred bytes at plan time (MPI_Allgather, torch.distributed,
 * a file, ... - see INTEGRATION.md). Ranks may also be threads of one process (one per GPU, programs/demo_dist.c): they are
 * recognised by their process id and use direct peer access instead of IPC mappings.
 */
#ifndef FFTB200_DIST_H
#define FFTB200_DIST_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftb200_dist fftb200_dist;

/* recv receives world * bytes: the `bytes` sent by rank 0, then rank 1, ... Returns 0 on success. Called at plan time and
 * once (as a barrier) by fftb200_dist_destroy - never on the data path. */
typedef int (*fftb200_allgather_fn)(void* ctx, const void* send, void* recv, size_t bytes);

/* log2 M of the head pass for a 2^log_n transform over 2^log_world ranks, -1 if no split exists (needs
 * log_n >= 2 * log_world + 12: both halves must split into passes of 6 .. 9 stages). */
int fftb200_dist_choose_split(int log_n, int log_world);

/* Collective over all ranks. direction -1 forward, +1 inverse (scaled by 1/N). log_m = 0 lets the library choose.
 * The calling thread's current device is the rank's GPU. 0 on success, -1 on failure (fftb200_last_error()). */
int fftb200_dist_create(fftb200_dist** out, int log_n, int world, int rank, int direction, int log_m,
                        fftb200_allgather_fn allgather, void* ctx);
/* Enqueue one transform on the plan's stream: d_in is this rank's block (N / world complex, device memory, not
 * modified; it must be ready when the call is made or be ordered before the plan's stream by the caller). *d_out is set
 * to the plan-owned buffer that holds this rank's block of the result once the stream has drained; it is overwritten
 * by the next execution. Every rank must call it the same number of times. */
int fftb200_dist_exec_async(fftb200_dist* plan, const void* d_in, void** d_out);
int fftb200_dist_sync(fftb200_dist* plan);
void* fftb200_dist_stream(fftb200_dist* plan);        /* cudaStream_t */
int fftb200_dist_log_m(const fftb200_dist* plan);
const char* fftb200_dist_describe(const fftb200_dist* plan);
/* Collective: synchronises the stream, meets the other ranks through the all-gather callback, then releases the mappings. */
void fftb200_dist_destroy(fftb200_dist* plan);

#ifdef __cplusplus
}
#endif
#endif /* FFTB200_DIST_H */
