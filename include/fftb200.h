/*
 * fftb200.h - C-ABI of the B200 (sm_100a) FFT engine: the boundary between the C99 host library
 * (fft_auto.c / fft_gpu.c, which keep the reference's public API) and the hand-written CUDA kernels.
 *
 * Plain pointers and sizes only: no C99 _Complex, no C++ or torch types. Complex data is interleaved
 * (re, im) doubles, byte-compatible with the reference's complex_t (include/fft_common.h:28 there) and
 * with CUDA's double2.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference tree):
 *   device lifecycle      gpu/fft_cuda.cu:53-101   fft_gpu_init_cuda / cleanup / available
 *   memory + copies       gpu/fft_cuda.cu:103-135  fft_gpu_alloc_cuda / free / copy_h2d / copy_d2h
 *   plan create / destroy gpu/fft_cuda.cu:138-163, 188-197  (cufftPlan1d / cufftPlanMany / cufftDestroy)
 *   plan exec             gpu/fft_cuda.cu:166-185  (cufftExecZ2Z + cudaDeviceSynchronize)
 *   bluestein / r2c kinds algorithms/core/bluestein.c:79-155, algorithms/auto/fft_auto.c:391-403
 *
 * All functions returning int give 0 on success and a negative value on failure; the message is
 * available from fftb200_last_error(). Nothing here falls back to the CPU.
 */
#ifndef FFTB200_H
#define FFTB200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fftb200_plan fftb200_plan;

enum fftb200_kind {
    FFTB200_C2C = 0,       /* power-of-two complex transform                                  */
    FFTB200_BLUESTEIN = 1, /* arbitrary n through a padded power-of-two circular convolution  */
    FFTB200_R2C = 2,       /* real input, n/2 + 1 complex outputs, power-of-two n             */
    FFTB200_C2R = 3        /* n/2 + 1 complex bins in, n real outputs scaled by 1/n (direction +1) */
};

typedef struct fftb200_plan_desc {
    int n;                  /* transform length                                                          */
    int batch;              /* transforms per execution; transform b occupies [b*n, (b+1)*n)             */
    int direction;          /* -1 forward (unscaled), +1 inverse (scaled by 1/n)                         */
    int kind;               /* enum fftb200_kind                                                         */
    const double* twiddles; /* HOST, forward stage tables for size `table_n` (table_n - 1 complex):      */
                            /* entry (stage s, j) at 2^(s-1) - 1 + j. table_n = n, or Bluestein's m      */
    int table_n;
    const double* chirp;    /* HOST, n complex, Bluestein only: chirp for `direction`                    */
    const double* twiddles_accurate; /* HOST, optional: correctly rounded stage tables, same layout, for  */
    int accurate_n;                  /* size accurate_n (must be 8192). When given, stages with m <= 4096 */
                                     /* of single-pass plans use them (3e-14 from the reference, see      */
                                     /* fft_pipe.cuh); NULL keeps the reference recurrence everywhere.    */
    unsigned flags;         /* reserved                                                                  */
} fftb200_plan_desc;

/* ---- device ---- */
int fftb200_device_count(void);              /* number of CUDA devices, 0 when none / no driver   */
int fftb200_set_device(int device);
int fftb200_get_device(void);
const char* fftb200_device_name(void);       /* name of the current device                        */
int fftb200_mem_info(size_t* free_bytes, size_t* total_bytes);
int fftb200_sm_count(void);
int fftb200_device_reset(void);              /* drops cached tables of the current device         */

/* ---- memory ---- */
void* fftb200_malloc(size_t bytes);          /* device memory, NULL on failure                    */
void fftb200_free(void* dptr);
void* fftb200_host_alloc(size_t bytes);      /* pinned host memory for staging                    */
void fftb200_host_free(void* hptr);
int fftb200_memcpy_h2d(void* dst, const void* src, size_t bytes);   /* synchronous               */
int fftb200_memcpy_d2h(void* dst, const void* src, size_t bytes);   /* synchronous               */
int fftb200_memcpy_d2d(void* dst, const void* src, size_t bytes);
int fftb200_memset(void* dst, int value, size_t bytes);
int fftb200_fill_splitmix(void* dst, unsigned long long seed, unsigned long long first_elem,
                          unsigned long long count);  /* synthetic input generated on the device */

/* ---- plans ---- */
int fftb200_plan_create(fftb200_plan** out, const fftb200_plan_desc* desc);
/* d_in / d_out are device pointers (complex; R2C: d_in is n*batch doubles, d_out (n/2+1)*batch complex; C2R the
 * other way round). In place (d_in == d_out) is allowed for C2C and BLUESTEIN; R2C / C2R plans run out of place (the
 * single-kernel real transforms read packed rows that other thread blocks' outputs would overwrite; overlapping
 * buffers are rejected where that applies). Returns after the result is visible. */
int fftb200_plan_exec(fftb200_plan* plan, const void* d_in, void* d_out);
/* Same, but only enqueues on the plan's stream. */
int fftb200_plan_exec_async(fftb200_plan* plan, const void* d_in, void* d_out);
int fftb200_plan_sync(fftb200_plan* plan);
/* Host-pointer execution: pinned staging, chunked H2D / kernels / D2H overlapped on two streams. */
int fftb200_plan_exec_host(fftb200_plan* plan, const void* h_in, void* h_out);
void fftb200_plan_destroy(fftb200_plan* plan);
int fftb200_plan_launches(const fftb200_plan* plan);     /* kernel launches per execution         */
const char* fftb200_plan_describe(const fftb200_plan* plan); /* e.g. "c2c n=4096 b=65536: C12"    */

/* ---- pieces of the distributed transform (one process per GPU, exchanges done by the caller) ----
 * A partial plan runs stages [first_stage, first_stage + nstages) of the size-n transform on the Stockham layout
 * idx = c + (n >> first_stage) * k; desc->twiddles must cover stages up to first_stage + nstages (table_n >= 2^that).
 * private_table != 0 uploads the table for this plan only (rank-specific late-stage tables, see
 * fftb200_host_twiddles_dist in fftb200_ext.h). inverse_scale is applied by the pass that ends the plan when
 * direction = +1 (pass 1.0 for the head plan and 1/N for the tail plan of an inverse transform). */
int fftb200_plan_create_partial(fftb200_plan** out, const fftb200_plan_desc* desc, int first_stage, int nstages,
                                int private_table, double inverse_scale);
/* dst[b][a][c] = src[a][b][c], complex elements, c contiguous; enqueued on `stream` (a cudaStream_t, may be NULL). */
int fftb200_permute_bac(void* d_dst, const void* d_src, long long A, long long B, long long C, void* stream);
void* fftb200_plan_stream(fftb200_plan* plan);   /* the plan's cudaStream_t */
/* Chain a plan behind another one: from now on it enqueues on `stream` (a cudaStream_t owned by someone else, e.g.
 * fftb200_plan_stream of the plan that runs before it), so no host synchronisation is needed between the two. */
int fftb200_plan_set_stream(fftb200_plan* plan, void* stream);
/* dst[b][c][r] = src[b][r][c], complex elements, enqueued on `stream`: the corner turn of the 2-D transform
 * (rows then columns, the decomposition of applications/image_fft.c:35-72) when the columns are too short or too
 * few for the strided column kernels. dst != src. */
int fftb200_transpose(void* d_dst, const void* d_src, long long rows, long long cols, long long batch, void* stream);
/* Fused exchange. A peer table holds the base pointers of one exchange buffer on all 2^log_world ranks as seen from
 * this process (own buffer and IPC-opened peers). With fftb200_plan_set_peer_output the last pass of a partial plan
 * stores over NVLink peer memory instead of into d_out: output index row * 2^log_width + col goes to rank
 * row >> log_rows_per_rank, element ((row mod 2^log_rows_per_rank) << (log_width + log_world)) + (rank << log_width) + col,
 * i.e. rows are split by destination and interleaved by source - the all-to-all and the block transpose in one.
 * The table is borrowed by the plan and must outlive its executions. */
typedef struct fftb200_peers fftb200_peers;
int fftb200_peers_create(fftb200_peers** out, void* const* bases, int log_world, int rank);
void fftb200_peers_destroy(fftb200_peers* peers);
int fftb200_plan_set_peer_output(fftb200_plan* plan, const fftb200_peers* peers, int log_width, int log_rows_per_rank);
/* First exchange: push `rows` rows of 2^(log_width + log_world) elements, split by destination column block, into the
 * peers: rank g gets [rank * rows + t][c]. Enqueued on `stream` (a cudaStream_t). */
int fftb200_push_columns(const fftb200_peers* peers, void* stream, const void* d_src, long long rows, int log_width);
/* Stream-ordered barrier between the ranks through peer memory (no collective library): flag_areas[g] is rank g's flag area
 * as seen from this process - at least 8 * world bytes inside an IPC-exchanged buffer, zeroed before the handles are
 * exchanged. fftb200_barrier_enqueue puts one arrive-and-wait kernel on `stream`; every rank enqueues the same sequence. */
typedef struct fftb200_barrier fftb200_barrier;
int fftb200_barrier_create(fftb200_barrier** out, void* const* flag_areas, int world, int rank);
int fftb200_barrier_enqueue(fftb200_barrier* barrier, void* stream);
void fftb200_barrier_destroy(fftb200_barrier* barrier);
int fftb200_stream_sync(void* stream);
/* CUDA IPC handles (64 bytes) of buffers from fftb200_malloc, exchanged by the caller between the per-GPU processes */
int fftb200_ipc_export(void* dptr, void* handle64);
void* fftb200_ipc_open(const void* handle64);
int fftb200_ipc_close(void* mapped);
/* ranks that are threads of ONE process use each other's pointers directly: enable peer access from the current device */
int fftb200_enable_peer_access(int peer_device);

/* ---- timing on the plan's stream (CUDA events) ---- */
int fftb200_timer_start(fftb200_plan* plan);
int fftb200_timer_stop(fftb200_plan* plan, float* elapsed_ms);   /* synchronises on the stop event */

/* ---- elementwise helpers used by the callers either side of the transform ---- */
/* y[i] = a[i] * b[i], count complex elements (FFT convolution, Bluestein's frequency-domain product) */
int fftb200_pointwise_mul(void* d_y, const void* d_a, const void* d_b, size_t count);
/* y[i] = conj(a[i]) * b[i] (cross-spectrum for FFT correlation, applications/power_spectrum.c:176-178) */
int fftb200_pointwise_mul_conj(void* d_y, const void* d_a, const void* d_b, size_t count);

const char* fftb200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FFTB200_H */
