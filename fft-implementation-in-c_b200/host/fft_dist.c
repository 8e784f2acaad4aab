/*
 * fft_dist.c - C99 host side of the distributed transform (include/fftb200_dist.h): plan-time handle exchange and
 * the per-execution sequence push -> head -> tail over the engine's C-ABI (partial plans, peer tables, peer barrier).
 *
 * The reference has no counterpart (single-device cufftPlan1d, gpu/fft_cuda.cu:138-163); the arithmetic is its radix-2
 * DIT stage product (algorithms/core/radix2_dit.c:59-120) split as N = R * M, see the header.
 */
#include "../../include/fftb200_dist.h"
#include "../../include/fftb200.h"
#include "../../include/fftb200_ext.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define FLAG_BYTES 4096   /* barrier flag area appended to the first exchange buffer */

struct fftb200_dist {
    int log_n, log_world, world, rank, direction, log_m;
    long long nloc;
    fftb200_plan* head;
    fftb200_plan* tail;
    void* e0;   /* receives T0 (columns) and T2 (the result), followed by the barrier flags */
    void* e1;   /* scratch of the local passes */
    void* e2;   /* receives T1 */
    void* mapped[2 * 64];
    int nmapped;
    fftb200_peers* peers0;
    fftb200_peers* peers2;
    fftb200_barrier* barrier;
    fftb200_allgather_fn allgather;
    void* ctx;
    char desc[512];
};

static int feasible(int cnt) {
    for (int k = 1; k <= 4; k++)
        if (6 * k <= cnt && cnt <= 9 * k) return 1;
    return 0;
}

int fftb200_dist_choose_split(int log_n, int log_world) {
    int best = -1;
    for (int lm = log_world + 4; lm < log_n - log_world - 3; lm++) {
        const int lr = log_n - lm;
        if (!feasible(lm) || !feasible(lr)) continue;
        const int d = abs(lm - lr), db = best < 0 ? 1 << 30 : abs(best - (log_n - best));
        if (best < 0 || d < db || (d == db && lm > best)) best = lm;
    }
    return best;
}

static int ilog2i(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

int fftb200_dist_create(fftb200_dist** out, int log_n, int world, int rank, int direction, int log_m,
                        fftb200_allgather_fn allgather, void* ctx) {
    if (!out) return -1;
    *out = NULL;
    const int lw = ilog2i(world);
    if (world < 1 || world > 64 || (1 << lw) != world || rank < 0 || rank >= world || log_n < 12 || log_n > 30 ||
        (direction != -1 && direction != 1) || (world > 1 && !allgather)) return -1;
    if (log_m <= 0) log_m = fftb200_dist_choose_split(log_n, lw);
    if (log_m < lw + 4 || log_n - log_m < lw + 4 || !feasible(log_m) || !feasible(log_n - log_m)) return -1;
    fftb200_dist* d = (fftb200_dist*)calloc(1, sizeof(*d));
    if (!d) return -1;
    d->log_n = log_n; d->log_world = lw; d->world = world; d->rank = rank; d->direction = direction; d->log_m = log_m;
    d->nloc = 1LL << (log_n - lw);
    d->allgather = allgather; d->ctx = ctx;
    const size_t bytes = sizeof(double) * 2 * (size_t)d->nloc;
    double* ttab = NULL;
    int ok = 0;
    do {
        /* head: stages [0, log_m) of the local array with the standard reference table */
        fftb200_plan_desc pd;
        memset(&pd, 0, sizeof(pd));
        pd.n = (int)d->nloc; pd.batch = 1; pd.direction = direction; pd.kind = FFTB200_C2C;
        pd.twiddles = fftb200_host_twiddles(1 << log_m); pd.table_n = 1 << log_m;
        if (!pd.twiddles || fftb200_plan_create_partial(&d->head, &pd, 0, log_m, 0, 1.0) != 0) break;
        /* tail: stages [log_m - log_world, log_n - log_world) with the rank's share of the late-stage tables */
        ttab = (double*)malloc(bytes);
        if (!ttab || fftb200_host_twiddles_dist(ttab, log_n, lw, rank, log_m) != 0) break;
        pd.twiddles = ttab; pd.table_n = (int)d->nloc;
        if (fftb200_plan_create_partial(&d->tail, &pd, log_m - lw, log_n - log_m, 1, 1.0 / (double)(1LL << log_n)) != 0) break;
        free(ttab); ttab = NULL;
        /* one stream for the whole sequence */
        if (fftb200_plan_set_stream(d->tail, fftb200_plan_stream(d->head)) != 0) break;
        d->e0 = fftb200_malloc(bytes + FLAG_BYTES);
        d->e1 = fftb200_malloc(bytes);
        d->e2 = fftb200_malloc(bytes);
        if (!d->e0 || !d->e1 || !d->e2) break;
        if (fftb200_memset((char*)d->e0 + bytes, 0, FLAG_BYTES) != 0) break;
        /* exchange the addresses of e0 and e2: CUDA IPC handles for ranks in other processes; ranks that are threads of this
         * process (one thread per GPU) use the pointers themselves after enabling peer access */
        struct card { unsigned char h0[64], h2[64]; long long pid; void* p0; void* p2; int device; int pad; } mine, *all;
        memset(&mine, 0, sizeof(mine));
        all = (struct card*)malloc(sizeof(struct card) * (size_t)world);
        if (!all) break;
        mine.pid = (long long)getpid(); mine.p0 = d->e0; mine.p2 = d->e2; mine.device = fftb200_get_device();
        int xok = fftb200_ipc_export(d->e0, mine.h0) == 0 && fftb200_ipc_export(d->e2, mine.h2) == 0;
        if (xok && world > 1) xok = allgather(ctx, &mine, all, sizeof(mine)) == 0;
        else if (xok) memcpy(all, &mine, sizeof(mine));
        void* b0[64]; void* b2[64]; void* fl[64];
        for (int g = 0; g < world && xok; g++) {
            if (g == rank) { b0[g] = d->e0; b2[g] = d->e2; }
            else if (all[g].pid == mine.pid) {
                xok = fftb200_enable_peer_access(all[g].device) == 0;
                b0[g] = all[g].p0; b2[g] = all[g].p2;
            } else {
                b0[g] = fftb200_ipc_open(all[g].h0);
                if (b0[g]) d->mapped[d->nmapped++] = b0[g];
                b2[g] = fftb200_ipc_open(all[g].h2);
                if (b2[g]) d->mapped[d->nmapped++] = b2[g];
                xok = b0[g] && b2[g];
            }
            fl[g] = (char*)b0[g] + bytes;
        }
        free(all);
        if (!xok) break;
        if (fftb200_peers_create(&d->peers0, b0, lw, rank) != 0 || fftb200_peers_create(&d->peers2, b2, lw, rank) != 0) break;
        if (fftb200_barrier_create(&d->barrier, fl, world, rank) != 0) break;
        /* head output [k][r_loc] -> rank k / Ml, B[k_loc][rank * Rl + r_loc]; tail output [q][k_loc] -> rank q / Rl, X[q_loc][rank * Ml + k_loc] */
        const int lml = log_m - lw, lrl = log_n - log_m - lw;
        if (fftb200_plan_set_peer_output(d->head, d->peers2, lrl, lml) != 0 || fftb200_plan_set_peer_output(d->tail, d->peers0, lml, lrl) != 0) break;
        snprintf(d->desc, sizeof(d->desc), "dist 2^%d over %d ranks, M = 2^%d: push | %s | %s", log_n, world, log_m,
                 fftb200_plan_describe(d->head), fftb200_plan_describe(d->tail));
        ok = 1;
    } while (0);
    free(ttab);
    if (!ok) {
        char keep[512];
        snprintf(keep, sizeof(keep), "%s", fftb200_last_error());
        d->allgather = NULL;   /* a failed rank cannot meet the others */
        fftb200_dist_destroy(d);
        fprintf(stderr, "fftb200_dist_create: %s\n", keep[0] ? keep : "failed");
        return -1;
    }
    *out = d;
    return 0;
}

int fftb200_dist_exec_async(fftb200_dist* d, const void* d_in, void** d_out) {
    if (!d || !d_in || !d_out) return -1;
    void* st = fftb200_plan_stream(d->head);
    const long long ml = 1LL << (d->log_m - d->log_world);
    const int lrl = d->log_n - d->log_m - d->log_world;
    if (fftb200_barrier_enqueue(d->barrier, st) != 0) return -1;               /* the peers are done with the previous result */
    if (fftb200_push_columns(d->peers0, st, d_in, ml, lrl) != 0) return -1;    /* T0 */
    if (fftb200_barrier_enqueue(d->barrier, st) != 0) return -1;
    if (fftb200_plan_exec_async(d->head, d->e0, d->e1) != 0) return -1;        /* head; its last pass stores into the peers' e2 (T1) */
    if (fftb200_barrier_enqueue(d->barrier, st) != 0) return -1;
    if (fftb200_plan_exec_async(d->tail, d->e2, d->e1) != 0) return -1;        /* tail; its last pass stores into the peers' e0 (T2) */
    if (fftb200_barrier_enqueue(d->barrier, st) != 0) return -1;
    *d_out = d->e0;
    return 0;
}

int fftb200_dist_sync(fftb200_dist* d) { return d ? fftb200_plan_sync(d->head) : -1; }
void* fftb200_dist_stream(fftb200_dist* d) { return d ? fftb200_plan_stream(d->head) : NULL; }
int fftb200_dist_log_m(const fftb200_dist* d) { return d ? d->log_m : -1; }
const char* fftb200_dist_describe(const fftb200_dist* d) { return d ? d->desc : ""; }

void fftb200_dist_destroy(fftb200_dist* d) {
    if (!d) return;
    if (d->head) fftb200_plan_sync(d->head);
    if (d->allgather && d->world > 1) {   /* nobody unmaps while a peer may still store */
        unsigned char one = 1, *all = (unsigned char*)malloc((size_t)d->world);
        if (all) { d->allgather(d->ctx, &one, all, 1); free(all); }
    }
    fftb200_plan_destroy(d->tail);   /* borrows the head's stream: goes first */
    fftb200_plan_destroy(d->head);
    fftb200_barrier_destroy(d->barrier);
    fftb200_peers_destroy(d->peers0);
    fftb200_peers_destroy(d->peers2);
    for (int i = 0; i < d->nmapped; i++) fftb200_ipc_close(d->mapped[i]);
    fftb200_free(d->e0);
    fftb200_free(d->e1);
    fftb200_free(d->e2);
    free(d);
}
