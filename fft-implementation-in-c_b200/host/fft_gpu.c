/*
 * fft_gpu.c - the public device API (include/fft_gpu.h) as a thin C99 launcher over the engine's
 * C-ABI (include/fftb200.h).
 *
 * Replaces the reference's gpu/fft_gpu.c (a switch over backends whose CUDA branches are compiled out
 * under gcc, so fft_gpu_available() is constant 0 on Linux) and its callee gpu/fft_cuda.cu. Same
 * signatures and error conventions: constructors return NULL, int functions 0 / -1, void functions
 * silently ignore NULL handles (reference gpu/fft_gpu.c:247), errors print one line on stderr
 * (reference gpu/fft_cuda.cu:34-50). Nothing here computes on the CPU.
 */
#include "../../include/fft_gpu.h"
#include "../../include/fftb200.h"
#include "../../include/fftb200_ext.h"
#include "ref_twiddle.h"
#include <pthread.h>

struct fft_gpu_memory {
    void* dptr;
    size_t size; /* complex elements */
};

struct fft_gpu_plan {
    fftb200_plan* engine;  /* 1-D: the batched plan. 2-D: the row pass (n = cols, batch = rows)                      */
    fftb200_plan* engine2; /* 2-D: the column pass, chained on the row pass's stream (NULL for 1-D plans)          */
    void* scratch;         /* 2-D through corner turns: rows * cols complex                                         */
    int n, batch;          /* 1-D shape; for 2-D plans n * batch = rows * cols                                      */
    int rows, cols;        /* 2-D shape, 0 for 1-D plans                                                            */
    int turn;              /* 2-D: 1 = transpose, contiguous column transforms, transpose back; 0 = strided kernels */
    fft_direction direction;
};

static fft_gpu_backend_t g_backend = FFT_GPU_NONE;

static void report(const char* what) { fprintf(stderr, "fft_gpu: %s: %s\n", what, fftb200_last_error()); }

int fft_gpu_available(void) { return fftb200_device_count() > 0 ? 1 : 0; }

int fft_gpu_init(fft_gpu_backend_t backend) {
    if (backend != FFT_GPU_CUDA && backend != FFT_GPU_AUTO) return -1; /* Metal / OpenCL: not on this platform */
    if (g_backend == FFT_GPU_CUDA) return 0;                          /* idempotent, reference fft_cuda.cu:54 */
    if (!fft_gpu_available()) return -1;
    const char* dev = getenv("FFTB200_DEVICE");
    if (dev && fftb200_set_device(atoi(dev)) != 0) { report("set device"); return -1; }
    g_backend = FFT_GPU_CUDA;
    return 0;
}

static void cache_drop(void);

void fft_gpu_cleanup(void) {
    cache_drop();
    if (g_backend == FFT_GPU_CUDA) fftb200_device_reset();
    g_backend = FFT_GPU_NONE;
}

fft_gpu_backend_t fft_gpu_get_backend(void) { return g_backend; }

static int ensure_init(void) { return g_backend == FFT_GPU_CUDA ? 0 : fft_gpu_init(FFT_GPU_AUTO); }

fft_gpu_memory_t fft_gpu_alloc(size_t size) {
    if (ensure_init() != 0) return NULL;
    fft_gpu_memory_t m = (fft_gpu_memory_t)malloc(sizeof(*m));
    if (!m) return NULL;
    m->dptr = fftb200_malloc(size * sizeof(complex_t));
    if (!m->dptr) { report("alloc"); free(m); return NULL; }
    m->size = size;
    return m;
}

void fft_gpu_free(fft_gpu_memory_t mem) {
    if (!mem) return;
    fftb200_free(mem->dptr);
    free(mem);
}

void fft_gpu_copy_h2d(fft_gpu_memory_t dst, const complex_t* src, size_t size) {
    if (!dst || !src) return;
    if (size > dst->size) size = dst->size;
    if (fftb200_memcpy_h2d(dst->dptr, src, size * sizeof(complex_t)) != 0) report("copy_h2d");
}

void fft_gpu_copy_d2h(complex_t* dst, fft_gpu_memory_t src, size_t size) {
    if (!dst || !src) return;
    if (size > src->size) size = src->size;
    if (fftb200_memcpy_d2h(dst, src->dptr, size * sizeof(complex_t)) != 0) report("copy_d2h");
}

/* Wisdom (reference fft_auto.h:124-137; stubs in fft_auto.c:417-426): the planner here is a fixed function of
 * (n, batch), so there is nothing to tune - what is worth carrying between processes is WHICH shapes were planned, so
 * that an importer can build the host twiddle tables (the only expensive plan-time step: the reference's serial
 * recurrence, n - 1 complex multiplies) before the first plan is asked for. One line per distinct shape. */
#define WISDOM_MAX 128
static struct { int n, batch, dir, kind; char desc[160]; } g_wisdom[WISDOM_MAX];
static int g_wisdom_count = 0;
static pthread_mutex_t g_wisdom_mu = PTHREAD_MUTEX_INITIALIZER;

static void wisdom_note(int n, int batch, int dir, int kind, const char* desc) {
    pthread_mutex_lock(&g_wisdom_mu);
    int found = 0;
    for (int i = 0; i < g_wisdom_count; i++)
        if (g_wisdom[i].n == n && g_wisdom[i].batch == batch && g_wisdom[i].dir == dir && g_wisdom[i].kind == kind) found = 1;
    if (!found && g_wisdom_count < WISDOM_MAX) {
        g_wisdom[g_wisdom_count].n = n; g_wisdom[g_wisdom_count].batch = batch;
        g_wisdom[g_wisdom_count].dir = dir; g_wisdom[g_wisdom_count].kind = kind;
        snprintf(g_wisdom[g_wisdom_count].desc, sizeof(g_wisdom[0].desc), "%s", desc ? desc : "");
        g_wisdom_count++;
    }
    pthread_mutex_unlock(&g_wisdom_mu);
}

char* fftb200_host_wisdom_export(void) {
    pthread_mutex_lock(&g_wisdom_mu);
    const size_t cap = 64 + (size_t)g_wisdom_count * 224;
    char* out = (char*)malloc(cap);
    if (out) {
        size_t off = (size_t)snprintf(out, cap, "# FFT Wisdom v2.0.0\n");
        for (int i = 0; i < g_wisdom_count && off < cap; i++)
            off += (size_t)snprintf(out + off, cap - off, "plan %d %d %d %d # %s\n", g_wisdom[i].n, g_wisdom[i].batch, g_wisdom[i].dir,
                                    g_wisdom[i].kind, g_wisdom[i].desc);
    }
    pthread_mutex_unlock(&g_wisdom_mu);
    return out;
}

int fftb200_host_wisdom_import(const char* wisdom) {
    if (!wisdom || strncmp(wisdom, "# FFT Wisdom", 12) != 0) return 0;
    for (const char* line = wisdom; line && *line; line = strchr(line, '\n') ? strchr(line, '\n') + 1 : NULL) {
        int n, batch, dir, kind;
        if (sscanf(line, "plan %d %d %d %d", &n, &batch, &dir, &kind) != 4 || n <= 1) continue;
        long long m = n;
        if (kind == FFTB200_BLUESTEIN) { m = 1; while (m < 2LL * n - 1) m <<= 1; }
        if ((m & (m - 1)) == 0 && m <= (1LL << 30)) (void)fftb200_host_twiddles((int)m);   /* cached per process */
    }
    return 1;
}

/* shared with fft_auto.c: build an engine plan of the given kind for (n, batch, direction) */
fftb200_plan* fftb200_host_make_plan(int n, int batch, int direction, int kind) {
    if (n <= 0 || batch <= 0) return NULL;
    if (ensure_init() != 0) return NULL;
    fftb200_plan_desc d;
    memset(&d, 0, sizeof(d));
    d.n = n; d.batch = batch; d.direction = direction < 0 ? -1 : 1; d.kind = kind;
    double* chirp = NULL;
    if (kind == FFTB200_BLUESTEIN || (kind == FFTB200_R2C && !is_power_of_two(n))) {   /* real input of any length runs through Bluestein too */
        long long m = 1;
        while (m < 2LL * n - 1) m <<= 1;
        if (m > (1LL << 30)) return NULL;
        d.table_n = (int)m;
        chirp = (double*)malloc(sizeof(double) * 2 * (size_t)n);
        if (!chirp) return NULL;
        fftb200_host_chirp(chirp, n, d.direction);
        d.chirp = chirp;
    } else {
        d.table_n = n;
    }
    d.twiddles = fftb200_host_twiddles(d.table_n);
    d.twiddles_accurate = fftb200_host_twiddles_accurate(&d.accurate_n);
    fftb200_plan* p = NULL;
    if (!d.twiddles || fftb200_plan_create(&p, &d) != 0) { report("plan"); p = NULL; }
    if (p) wisdom_note(n, batch, d.direction, kind, fftb200_plan_describe(p));
    free(chirp);
    return p;
}

fft_gpu_plan_t fft_gpu_plan_1d(int n, int batch, fft_direction direction) {
    if (n <= 0 || batch <= 0) return NULL;
    fft_gpu_plan_t p = (fft_gpu_plan_t)malloc(sizeof(*p));
    if (!p) return NULL;
    memset(p, 0, sizeof(*p));
    p->engine = fftb200_host_make_plan(n, batch, (int)direction, is_power_of_two(n) ? FFTB200_C2C : FFTB200_BLUESTEIN);
    if (!p->engine) { free(p); return NULL; }
    p->n = n; p->batch = batch; p->direction = direction;
    return p;
}

/* 2-D execution on device pointers: rows first, then columns (the decomposition of the reference's CPU code,
 * applications/image_fft.c:35-72), everything enqueued on one stream and synchronised once. */
static int execute_2d(fft_gpu_plan_t plan, void* in, void* out) {
    void* st = fftb200_plan_stream(plan->engine);
    if (fftb200_plan_exec_async(plan->engine, in, out) != 0) return -1;
    if (!plan->turn) {
        if (fftb200_plan_exec_async(plan->engine2, out, out) != 0) return -1;
    } else {
        if (fftb200_transpose(plan->scratch, out, plan->rows, plan->cols, 1, st) != 0) return -1;
        if (fftb200_plan_exec_async(plan->engine2, plan->scratch, plan->scratch) != 0) return -1;
        if (fftb200_transpose(out, plan->scratch, plan->cols, plan->rows, 1, st) != 0) return -1;
    }
    return fftb200_plan_sync(plan->engine);
}

void fft_gpu_execute(fft_gpu_plan_t plan, fft_gpu_memory_t in, fft_gpu_memory_t out) {
    if (!plan || !in || !out) return;
    const size_t need = (size_t)plan->n * (size_t)plan->batch;
    if (in->size < need || out->size < need) { fprintf(stderr, "fft_gpu: execute: buffer smaller than n*batch\n"); return; }
    if (plan->engine2) {
        if (execute_2d(plan, in->dptr, out->dptr) != 0) report("execute 2d");
        return;
    }
    if (fftb200_plan_exec(plan->engine, in->dptr, out->dptr) != 0) report("execute");
}

void fft_gpu_destroy_plan(fft_gpu_plan_t plan) {
    if (!plan) return;
    if (plan->engine) fftb200_plan_sync(plan->engine);
    fftb200_plan_destroy(plan->engine2); /* borrows the first plan's stream: goes first */
    fftb200_plan_destroy(plan->engine);
    fftb200_free(plan->scratch);
    free(plan);
}

/* The host-pointer conveniences keep their engine plans (tables, streams, staging ring, Bluestein chirp and kernel
 * spectrum): a small LRU keyed on (device, n, batch, direction, kind). A caller that loops over fft_gpu_dft_1d_batch
 * with one shape - the reference's usage, gpu/fft_gpu.c:366-374 - or alternates forward and inverse transforms - the
 * convolution pattern, applications/convolution.c:34-96 - pays for plan construction once per shape. The mutex guards
 * the table only: an entry is marked busy while its plan runs and the lock is released, so threads working on
 * different shapes (or devices) run concurrently; a second thread asking for a shape that is busy builds a plan of its
 * own, which joins the cache afterwards if a slot is free. */
#define CACHE_SLOTS 16
static struct cache_entry { fftb200_plan* plan; int n, batch, dir, kind, device, busy; unsigned long long stamp; } g_cache[CACHE_SLOTS];
static unsigned long long g_cache_clock = 0;
static long long g_cache_builds = 0, g_cache_hits = 0;
static pthread_mutex_t g_cache_mu = PTHREAD_MUTEX_INITIALIZER;

static void cache_drop(void) {
    pthread_mutex_lock(&g_cache_mu);
    for (int i = 0; i < CACHE_SLOTS; i++) {
        if (g_cache[i].plan && !g_cache[i].busy) { fftb200_plan_destroy(g_cache[i].plan); g_cache[i].plan = NULL; }
    }
    pthread_mutex_unlock(&g_cache_mu);
}

/* diagnostics for tests: engine plans built / found by the cached host entry points since the process started */
void fftb200_host_cache_stats(long long* builds, long long* hits) {
    pthread_mutex_lock(&g_cache_mu);
    if (builds) *builds = g_cache_builds;
    if (hits) *hits = g_cache_hits;
    pthread_mutex_unlock(&g_cache_mu);
}

static int exec_cached_on_current_device(const void* in, void* out, int n, int batch, int direction, int kind) {
    const int dev = fftb200_get_device();
    struct cache_entry* e = NULL;
    pthread_mutex_lock(&g_cache_mu);
    for (int i = 0; i < CACHE_SLOTS && !e; i++) {
        struct cache_entry* c = &g_cache[i];
        if (c->plan && !c->busy && c->n == n && c->batch == batch && c->dir == direction && c->kind == kind && c->device == dev) e = c;
    }
    if (e) { e->busy = 1; e->stamp = ++g_cache_clock; g_cache_hits++; }
    pthread_mutex_unlock(&g_cache_mu);
    fftb200_plan* plan = e ? e->plan : fftb200_host_make_plan(n, batch, direction, kind);   /* built outside the lock */
    if (!plan) return -1;
    const int rc = fftb200_plan_exec_host(plan, in, out);
    if (rc != 0) report("host execute");
    pthread_mutex_lock(&g_cache_mu);
    if (e) {
        e->busy = 0;
    } else {
        g_cache_builds++;
        /* adopt the new plan: a free slot, else the least recently used idle one */
        struct cache_entry* victim = NULL;
        for (int i = 0; i < CACHE_SLOTS; i++) {
            struct cache_entry* c = &g_cache[i];
            if (c->busy) continue;
            if (!c->plan) { victim = c; break; }
            if (!victim || c->stamp < victim->stamp) victim = c;
        }
        fftb200_plan* old = NULL;
        if (victim) {
            old = victim->plan;
            victim->plan = plan; victim->n = n; victim->batch = batch; victim->dir = direction; victim->kind = kind;
            victim->device = dev; victim->busy = 0; victim->stamp = ++g_cache_clock;
            plan = NULL;
        }
        pthread_mutex_unlock(&g_cache_mu);
        if (old) fftb200_plan_destroy(old);
        if (plan) fftb200_plan_destroy(plan);   /* every slot was busy: a private plan */
        return rc == 0 ? 0 : -1;
    }
    pthread_mutex_unlock(&g_cache_mu);
    return rc == 0 ? 0 : -1;
}

/* Multi-device fan-out of a batched host job (behind the unchanged fft_gpu_dft_1d_batch signature):
 * FFTB200_GPUS=G (or fftb200_host_set_gpus) splits the batch into G contiguous ranges (fftb200_shard_range), one host
 * thread per device, each with its own cached plan, streams and staging ring on that device; transforms are independent,
 * so there is no exchange. The hook the reference offers for this is fft_gpu_set_device (include/fft_gpu.h:177). */
static int g_host_gpus = 0;   /* 0: read FFTB200_GPUS once */
void fftb200_host_set_gpus(int gpus) { g_host_gpus = gpus < 1 ? 1 : gpus; }
int fftb200_host_get_gpus(void) {
    if (g_host_gpus == 0) {
        const char* e = getenv("FFTB200_GPUS");
        int g = e ? atoi(e) : 1;
        const int have = fftb200_device_count();
        if (g > have) g = have;
        g_host_gpus = g < 1 ? 1 : g;
    }
    return g_host_gpus;
}

struct fan_job { const char* in; char* out; int n, direction, kind, device, rc; long long first, count; size_t in_per, out_per; };

static void* fan_worker(void* arg) {
    struct fan_job* j = (struct fan_job*)arg;
    j->rc = -1;
    if (fftb200_set_device(j->device) != 0) return NULL;
    j->rc = exec_cached_on_current_device(j->in + (size_t)j->first * j->in_per, j->out + (size_t)j->first * j->out_per, j->n, (int)j->count,
                                          j->direction, j->kind);
    return NULL;
}

int fftb200_host_exec_cached(const void* in, void* out, int n, int batch, int direction, int kind) {
    if (ensure_init() != 0) return -1;
    int gpus = fftb200_host_get_gpus();
    /* below ~4 MiB per device the second device's launch + sync costs more than it saves */
    const size_t half = sizeof(complex_t) * (size_t)(n / 2 + 1);
    const size_t in_per = kind == FFTB200_R2C ? sizeof(double) * (size_t)n : kind == FFTB200_C2R ? half : sizeof(complex_t) * (size_t)n;
    const size_t out_per = kind == FFTB200_R2C ? half : kind == FFTB200_C2R ? sizeof(double) * (size_t)n : sizeof(complex_t) * (size_t)n;
    while (gpus > 1 && (batch < gpus || in_per * (size_t)batch / (size_t)gpus < ((size_t)4 << 20))) gpus--;
    if (gpus <= 1) return exec_cached_on_current_device(in, out, n, batch, direction, kind);
    const int home = fftb200_get_device();
    struct fan_job jobs[64];
    pthread_t th[64];
    int spawned[64];
    if (gpus > 64) gpus = 64;
    int rc = 0;
    for (int g = 0; g < gpus; g++) {
        struct fan_job* j = &jobs[g];
        j->in = (const char*)in; j->out = (char*)out; j->n = n; j->direction = direction; j->kind = kind;
        j->device = (home + g) % fftb200_device_count();
        j->in_per = in_per; j->out_per = out_per; j->rc = -1;
        spawned[g] = 0;
        fftb200_shard_range(batch, gpus, g, &j->first, &j->count);
        if (j->count == 0) { j->rc = 0; continue; }
        if (g == gpus - 1) { fan_worker(j); continue; }   /* the calling thread takes the last range */
        if (pthread_create(&th[g], NULL, fan_worker, j) == 0) spawned[g] = 1;
        else fan_worker(j);
    }
    for (int g = 0; g < gpus; g++) {
        if (spawned[g]) pthread_join(th[g], NULL);
        if (jobs[g].rc != 0) rc = -1;
    }
    fftb200_set_device(home);
    return rc;
}

int fft_gpu_dft_1d_batch(complex_t* in, complex_t* out, int n, int batch, fft_direction direction) {
    if (!in || !out || n <= 0 || batch <= 0) return -1;
    return fftb200_host_exec_cached(in, out, n, batch, (int)direction < 0 ? -1 : 1,
                                    is_power_of_two(n) ? FFTB200_C2C : FFTB200_BLUESTEIN);
}

int fft_gpu_dft_1d(complex_t* in, complex_t* out, int n, fft_direction direction) {
    return fft_gpu_dft_1d_batch(in, out, n, 1, direction);
}

/* 2-D transforms of a row-major rows x cols array. Stubs in the reference (gpu/fft_gpu.c:377-394 return NULL / -1);
 * built here from the batched 1-D kernels the way the reference's CPU code does it (applications/image_fft.c:35-72:
 * every row, then every column), forward unscaled, inverse scaled by 1/(rows*cols) ONCE (the reference's fft_2d
 * scales the inverse twice, :63-71; the public headers promise 1/n). The row pass is one batched plan (n = cols,
 * batch = rows). The column pass is
 *   - for power-of-two shapes with >= 64 rows: stages 1 .. log2(rows) of a size rows*cols Stockham plan - exactly the
 *     column transforms, in place, through the strided tile kernels (no transposes, twiddles = the reference's
 *     size-`rows` tables); otherwise
 *   - transpose, one batched plan (n = rows, batch = cols, Bluestein when rows is not a power of two), transpose back.
 * A degenerate shape (one row or one column) is a 1-D plan. */
static fftb200_plan* make_column_plan(int rows, int cols, int direction) {
    if (!is_power_of_two(rows) || !is_power_of_two(cols) || rows < 64) return NULL;
    fftb200_plan_desc d;
    memset(&d, 0, sizeof(d));
    d.n = rows * cols; d.batch = 1; d.direction = direction < 0 ? -1 : 1; d.kind = FFTB200_C2C;
    d.table_n = rows;
    d.twiddles = fftb200_host_twiddles(rows);
    if (!d.twiddles) return NULL;
    fftb200_plan* p = NULL;
    if (fftb200_plan_create_partial(&p, &d, 0, log2_int(rows), 0, 1.0 / (double)rows) != 0) return NULL;
    return p;
}

fft_gpu_plan_t fft_gpu_plan_2d(int rows, int cols, fft_direction direction) {
    if (rows <= 0 || cols <= 0 || (long long)rows * cols > (1LL << 30)) return NULL;
    if (rows == 1) return fft_gpu_plan_1d(cols, 1, direction);
    if (cols == 1) return fft_gpu_plan_1d(rows, 1, direction);
    const int dir = (int)direction < 0 ? -1 : 1;
    fft_gpu_plan_t p = (fft_gpu_plan_t)malloc(sizeof(*p));
    if (!p) return NULL;
    memset(p, 0, sizeof(*p));
    p->n = cols; p->batch = rows; p->rows = rows; p->cols = cols; p->direction = direction;
    p->engine = fftb200_host_make_plan(cols, rows, dir, is_power_of_two(cols) ? FFTB200_C2C : FFTB200_BLUESTEIN);
    if (!p->engine) { free(p); return NULL; }
    if (!getenv("FFTB200_2D_TURN")) p->engine2 = make_column_plan(rows, cols, dir);
    if (!p->engine2) {
        p->turn = 1;
        p->engine2 = fftb200_host_make_plan(rows, cols, dir, is_power_of_two(rows) ? FFTB200_C2C : FFTB200_BLUESTEIN);
        p->scratch = fftb200_malloc(sizeof(complex_t) * (size_t)rows * (size_t)cols);
    }
    if (!p->engine2 || (p->turn && !p->scratch) ||
        fftb200_plan_set_stream(p->engine2, fftb200_plan_stream(p->engine)) != 0) {
        report("plan 2d");
        fft_gpu_destroy_plan(p);
        return NULL;
    }
    return p;
}

int fft_gpu_dft_2d(complex_t* in, complex_t* out, int rows, int cols, fft_direction direction) {
    if (!in || !out || rows <= 0 || cols <= 0) return -1;
    fft_gpu_plan_t p = fft_gpu_plan_2d(rows, cols, direction);
    if (!p) return -1;
    const size_t total = (size_t)rows * (size_t)cols;
    int rc = -1;
    fft_gpu_memory_t buf = fft_gpu_alloc(total);
    if (buf && fftb200_memcpy_h2d(buf->dptr, in, total * sizeof(complex_t)) == 0) {
        rc = p->engine2 ? execute_2d(p, buf->dptr, buf->dptr) : fftb200_plan_exec(p->engine, buf->dptr, buf->dptr);
        if (rc == 0) rc = fftb200_memcpy_d2h(out, buf->dptr, total * sizeof(complex_t));
    }
    if (rc != 0) report("dft_2d");
    fft_gpu_free(buf);
    fft_gpu_destroy_plan(p);
    return rc == 0 ? 0 : -1;
}

const char* fft_gpu_get_device_name(void) {
    if (ensure_init() != 0) return "No GPU";
    return fftb200_device_name();
}

void fft_gpu_get_memory_info(size_t* total, size_t* available) {
    size_t f = 0, t = 0;
    if (ensure_init() == 0) fftb200_mem_info(&f, &t);
    if (total) *total = t;
    if (available) *available = f;
}

int fft_gpu_set_device(int device) {
    if (!fft_gpu_available()) return -1;
    return fftb200_set_device(device) == 0 ? 0 : -1;
}

/* engine handle of a public plan, for the additive helpers in fftb200_ext.h (timing, async) */
fftb200_plan* fftb200_engine_of(fft_gpu_plan_t plan) { return plan ? plan->engine : NULL; }
void* fftb200_devptr_of(fft_gpu_memory_t mem) { return mem ? mem->dptr : NULL; }

int fftb200_shard_range(long long batch, int world, int rank, long long* first, long long* count) {
    if (batch < 0 || world <= 0 || rank < 0 || rank >= world || !first || !count) return -1;
    const long long base = batch / world, extra = batch % world;
    *first = base * rank + (rank < extra ? rank : extra);
    *count = base + (rank < extra ? 1 : 0);
    return 0;
}
