// fft_pipe.cuh - the persistent, TMA-fed Stockham kernel for whole transforms of 512 .. 4096 points.
//
// This is the kernel behind BASELINE config 2 (N = 4096 x 65536 batch). It replaces the same reference
// code as fft_tile.cuh (the butterfly loop of algorithms/core/radix2_dit.c:70-119 and the never-built
// cufftExecZ2Z call of gpu/fft_cuda.cu:166-185) for the sizes where one transform fits a 64 KB tile.
//
// Structure (one CTA per SM, 512 threads = two groups of 256):
//   * a tile is 4096 consecutive complex doubles of the batch buffer = 4096/N whole transforms;
//   * three 64 KB shared-memory buffers form a ring. One thread issues a 1-D bulk async copy (TMA,
//     cp.async.bulk + mbarrier complete_tx) for the CTA's next tile as soon as a tile has left its buffer, so two
//     tiles (128 KB) are always in flight per SM while a third is being computed: HBM never waits for
//     the FP64 pipe and the other way round;
//   * tiles are handed out on demand from a global counter (round 2, "Tile order" in the kernel): the SMs of a B200 do not
//     all move data at the same rate, and a fixed tile-to-CTA assignment leaves 5-8 % of the HBM rate unused;
//   * the two thread groups work on alternate tiles and meet only through the ring, so one group's
//     shared-memory exchange overlaps the other group's butterflies;
//   * a thread holds 16 points in registers. Three sub-passes (radix N/256, 16, 16) regroup the
//     reference's log2 N radix-2 DIT stages; between sub-passes the tile is exchanged IN PLACE in the
//     buffer the TMA filled, with an XOR swizzle (element ^ ((element >> 4) & 7)) that makes every
//     128-bit shared-memory access pattern conflict-free without padding;
//   * results leave from registers with 128-bit stores, 512 contiguous bytes per warp instruction.
//
// Twiddles: all stages handled here have m <= 4096, where the reference's recurrence
// (radix2_dit.c:93,109) is still within 3e-14 (relative L2 of the whole transform) of correctly rounded
// twiddles - SURVEY.md 7.0 "hybrid". The kernel therefore reads the ACCURATE per-stage table (same layout
// as the reference-recurrence table), which has the symmetry w[q + m/4] = -i * w[q]; only 8 of the 15
// twiddles of a radix-16 butterfly are needed, the other 7 are free sign/swap variants. The middle
// sub-pass reads its 8 from a 2 KB shared table; the last sub-pass's twiddles depend only on the thread
// index: w^8, w^4, w^2, w are re-read per tile (L1 hits), the other four are w^2 * W8 and
// w * {W16, W8, W16^3}, rebuilt per tile (16 FP64 instructions).
//
// Measured alternatives (same box A/B, N = 4096 x 65536, ms): this form 1.31-1.32; all 8 last-pass twiddles
// in registers 1.51 (spills on the loop-carried path); the 4 powers in a per-thread shared table 1.43 (the
// LSU / MIO queue is the most loaded pipe: every extra LDS costs more than 16 extra DFMA). For N <= 2048 the four powers are
// re-read per tile from global memory (L1 hits) instead: these sizes are not HBM-bound and the freed registers are worth 8-9 %.
#pragma once
#include "fft_tile.cuh"

namespace fftb200 {

struct PipeArgs {
    const cd* in;
    cd* out;
    const cd* tab;     // accurate per-stage twiddle table (stage s, j) at 2^(s-1) - 1 + j
    long long ntiles;  // ceil(batch / (4096 / N))
    long long batch;
    int inverse;
    double scale;      // 1/N, applied when inverse
    // Bluestein variants (bluestein.c:107-148): chirp of the caller's length n_user, spectrum FB of the wrapped chirp (N entries)
    const cd* chirp;
    const cd* fb;
    int n_user;
    double y_scale;    // 1/n for the inverse direction of the caller's transform, else 1
    // dynamic tile hand-out (see "Tile order" in fft_pipe_kernel): sched[0] = tiles taken beyond the first three of every CTA,
    // sched[1] = CTAs that have finished; both are zero between launches.
    unsigned int* sched;
};

constexpr int PIPE_TILE = 4096;              // complex elements per tile
constexpr int PIPE_STAGES = 3;               // ring depth
constexpr int PIPE_GROUP = 256;              // threads per group
constexpr int PIPE_TW1 = 16 * 8;             // shared copy of the middle sub-pass twiddles: [kloc][8]
constexpr int PIPE_SLOT_C2R = PIPE_TILE + 16;   // c2r: a ring slot holds 2 NT half spectra of N/2 + 1 bins = 4096 + 2 NT elements
constexpr size_t PIPE_SMEM = (size_t)PIPE_STAGES * PIPE_SLOT_C2R * sizeof(cd) + PIPE_TW1 * sizeof(cd) + 128;   // + barriers, tile numbers

__device__ __forceinline__ int pipe_swz(int idx) { return idx ^ ((idx >> 4) & 7); }

// (lo, hi) <- (lo + (-i*w)*hi, lo - (-i*w)*hi): the q + m/4 twiddle of an accurate table
__device__ __forceinline__ void bfly_mi(cd& lo, cd& hi, const cd w) {
    double sx = fma(w.y, hi.x, lo.x);
    double sy = fma(w.y, hi.y, lo.y);
    sx = fma(w.x, hi.y, sx);
    sy = fma(-w.x, hi.x, sy);
    hi.x = fma(2.0, lo.x, -sx);
    hi.y = fma(2.0, lo.y, -sy);
    lo.x = sx;
    lo.y = sy;
}

// radix-2^LR DIT over w[] (bit-reversed placement) with the 2^(LR-1) stored twiddles tw[h],
// h = 2^(s-1) + q for q < max(1, 2^(s-2)); positions q >= 2^(s-2) use -i * tw[h - 2^(s-2)].
template <int LR, int S, int BASE, int Q>
struct SubStageSym {
    static __device__ __forceinline__ void run(cd* w, const cd* tw) {
        constexpr int HH = 1 << (S - 1), R = 1 << LR;
        if constexpr (S >= 2 && 2 * Q >= HH) bfly_mi(w[BASE + Q], w[BASE + Q + HH], tw[HH + Q - HH / 2]);
        else bfly(w[BASE + Q], w[BASE + Q + HH], tw[HH + Q]);
        if constexpr (Q + 1 < HH) SubStageSym<LR, S, BASE, Q + 1>::run(w, tw);
        else if constexpr (BASE + 2 * HH < R) SubStageSym<LR, S, BASE + 2 * HH, 0>::run(w, tw);
        else if constexpr (S < LR) SubStageSym<LR, S + 1, 0, 0>::run(w, tw);
    }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a lost bulk copy or a protocol error traps (the launch fails with an error the host reports) instead of
// hanging the GPU; try_wait suspends the thread in hardware for a time slice, so the loop costs nothing while data flows
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
    // (the counter lives in a PTX-scoped register: written as a C++ loop it cost the N = 4096 kernel 28 bytes of spills and 5 %)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .u32 c;\n"
        "mov.u32 c, 0;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "add.u32 c, c, 1;\n"
        "setp.lt.u32 p, c, 0x1000000;\n"
        "@p bra WAIT_%=;\n"
        "trap;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void group_sync(int g) {
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(PIPE_GROUP) : "memory");
}

__device__ __forceinline__ cd cmulc(const cd a, const double c, const double d) {  // a * (c + i d)
    return make_double2(fma(a.x, c, -(a.y * d)), fma(a.x, d, a.y * c));
}

// REAL = PIPE_R2C: the tile holds 8192 reals = 2 NT real transforms; rows 2j and 2j + 1 are transformed TOGETHER as the complex
// sequence z = x_a + i x_b (every stage here uses the accurate conjugate-symmetric tables, so this regrouping is exact to rounding)
// and separated at the end, X_a[k] = (Z[k] + conj Z[N - k]) / 2, X_b[k] = (Z[k] - conj Z[N - k]) / 2i - the partners Z[N - k] sit in
// another thread of the same transform and are traded through the tile. Only the bins 0 .. N/2 are stored, N/2 + 1 per transform
// (fft_auto.h:89-97): the reference's promote-then-c2c reading of fft_plan_r2c_1d (fft_auto.c:394-397) at half the arithmetic and
// without the promotion and extraction passes. REAL = PIPE_C2R (inverse only): the slot holds N/2 + 1 bins of 2 NT transforms; rows 2j
// and 2j + 1 enter ONE inverse transform as Z = X_a + i X_b (Hermitian halves rebuilt while gathering), whose real and imaginary
// parts are the two real results (fft_auto.h:99-107).
// REAL = PIPE_BLUE_FWD / PIPE_BLUE_INV: the two transforms of Bluestein's algorithm for padded lengths N = 512 .. 4096 with the
// elementwise steps riding on them: FWD reads the caller's n-point rows, multiplies by conj(chirp) and zero-pads while gathering,
// and multiplies the spectrum by FB before storing it; INV (inverse c2c, 1/N) multiplies by conj(chirp) * y_scale and stores the
// first n values of every row to the caller's array. Two launches and HBM round trips instead of five, the same arithmetic.
enum { PIPE_C2C = 0, PIPE_R2C = 1, PIPE_C2R = 2, PIPE_BLUE_FWD = 3, PIPE_BLUE_INV = 4 };

template <int LOGN, bool INV, int REAL = PIPE_C2C>
__global__ void __launch_bounds__(2 * PIPE_GROUP, 1) fft_pipe_kernel(const PipeArgs a) {
    static_assert(LOGN >= 9 && LOGN <= 12, "one transform must be 512 .. 4096 points");
    static_assert(REAL == PIPE_C2C || ((REAL == PIPE_R2C || REAL == PIPE_BLUE_FWD) && !INV) || ((REAL == PIPE_C2R || REAL == PIPE_BLUE_INV) && INV),
                  "r2c and Bluestein's first transform are forward, c2r and its second inverse");
    constexpr int NH = (1 << LOGN) / 2 + 1;   // bins per transform of a half spectrum
    constexpr int N = 1 << LOGN, NT = PIPE_TILE / N;  // (complex) transforms per tile
    constexpr bool PAIR = REAL == PIPE_R2C || REAL == PIPE_C2R;   // two real transforms per complex transform
    constexpr int NTR = PAIR ? 2 * NT : NT;                        // caller's transforms per tile
    constexpr int SLOT = REAL == PIPE_C2R ? PIPE_SLOT_C2R : PIPE_TILE;
    constexpr int LR0 = LOGN - 8, R0 = 1 << LR0, NB0 = 16 / R0;
    constexpr int LN16 = LOGN - 4;                    // log2 (N / 16)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    cd* const tw1s = bufs + (size_t)PIPE_STAGES * PIPE_SLOT_C2R;
    uint64_t* const full = reinterpret_cast<uint64_t*>(tw1s + PIPE_TW1);

    int* const tile_of = reinterpret_cast<int*>(full + 2 * PIPE_STAGES);   // [6]: the tile a barrier's current phase stands for, -1 = no more
    int* const next_of = tile_of + 2 * PIPE_STAGES;                        // [2]: the tile a group will load next
    int* const cur_of = next_of + 2;                                       // [2][3]: the tile a group works on (its own copy: tile_of[] is
                                                                           // rewritten as soon as the slot three ahead is loaded or closed)
    int* const told_of = cur_of + 2 * PIPE_STAGES;                         // [2]: a slot of the OTHER group has been closed (by this group, or at start-up)
    int* const leave_of = told_of + 2;                                     // [2]: thread 0's verdict for its group

    const int g = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const int first = blockIdx.x, stride = gridDim.x;

    // Tile order. The SMs of this chip do not move data at the same rate (a plain copy with a fixed tile-to-CTA assignment reaches
    // 6.0-6.4 TB/s, the same copy with tiles handed out on demand 6.9-7.0: tools/copybench.cu, profiles/r02_microbench.md), so a CTA
    // takes its first three tiles by position (CTA + k * grid: nothing to wait for at start-up) and every later one from a global
    // counter: ring slot s of the CTA (buffer s % 3, group s % 2) carries whatever tile its loader was handed and the number travels in
    // shared memory next to the barrier. The counter is read one tile ahead (after a tile's results have left, when no butterfly
    // registers are live) so that its latency stays off the critical path, and the last CTA to finish sets it back to zero for the
    // next launch.
    // End of work. A number past the end closes a slot (-1, the barrier's phase completed without data). Each group reads the counter
    // in the order of the slots it loads, and once the counter has run out it stays so; hence (a) after the first closed slot a group
    // finds, every later slot of its own is closed too: nothing left to compute; (b) but the number it holds for the other group's
    // next slot may have been read earlier and still be a tile. So a group keeps walking its slots - passing on what it holds, reading
    // the counter again - until it has both FOUND a closed slot and CLOSED one itself: then the other group is sure to find a closed
    // slot, and no tile that was handed out is left unloaded. (At most two empty turns per group.)
    // (tile numbers fit 32 bits: 2^31 tiles of 64 KB are more than any memory holds; clamped so that they never wrap)
    const bool counted = a.ntiles > 3LL * stride;   // small launches (the latency-bound one-shot calls) never touch the counter
    auto take = [&]() -> int {   // the tile of this group's next load
        if (!counted) return 0x7fffffff;
        const unsigned v = 3u * gridDim.x + atomicAdd(a.sched, 1u);
        return v < 0x7fffffffu ? (int)v : 0x7fffffff;
    };

    // Barriers: buffer b is filled for tiles b, b + 3, b + 6, ... which the two groups consume alternately, and the load
    // of tile k + 3 is issued by the group that consumed tile k - not by the group that will wait for it. With ONE barrier
    // per buffer a group that is early could test the parity of a phase whose predecessor has not completed yet (the
    // parity wait then passes on the phase before: stale data, a second expect_tx in the same phase and a launch failure;
    // seen on the B200 about once per 10^7 tiles). So each buffer has TWO barriers used in turn (bar = b + 3 * (round & 1),
    // parity (round >> 1) & 1): all phases of one barrier are waited for by the same group, in program order.
    auto issue = [&](int tile, int b, uint32_t rnd) -> bool {  // one thread: start the load of `tile` into buffer b; true = nothing left, slot closed
        uint64_t* const bar = &full[b + PIPE_STAGES * (rnd & 1)];
        if (tile >= a.ntiles) {   // nothing left: complete the phase without data
            tile_of[b + PIPE_STAGES * (rnd & 1)] = -1;
            mbar_arrive(bar);
            return true;
        }
        tile_of[b + PIPE_STAGES * (rnd & 1)] = tile;
        long long nvalid = a.batch - (long long)tile * NTR;
        if (nvalid > NTR) nvalid = NTR;
        const uint32_t per = REAL == PIPE_R2C ? N * (uint32_t)sizeof(double) : REAL == PIPE_C2R ? NH * (uint32_t)sizeof(cd)
                           : REAL == PIPE_BLUE_FWD ? (uint32_t)a.n_user * (uint32_t)sizeof(cd) : N * (uint32_t)sizeof(cd);
        const uint32_t bytes = (uint32_t)nvalid * per;
        mbar_expect_tx(bar, bytes);
        bulk_load(bufs + (size_t)b * SLOT, reinterpret_cast<const char*>(a.in) + (size_t)tile * NTR * per, bytes, bar);
        return false;
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2 * PIPE_STAGES; b++) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // middle sub-pass twiddles -> shared: stage (LR0 + s), position kappa = kloc: tab[(h << LR0) + kloc - 1]
    // for h in {1, 2, 4, 5, 8, 9, 10, 11} (the other 7 positions are -i times one of these)
    if (threadIdx.x < R0 * 8) {
        const int kl = threadIdx.x >> 3, e = threadIdx.x & 7;
        const int h = e == 0 ? 1 : e == 1 ? 2 : e < 4 ? 2 + e : 4 + e;
        tw1s[threadIdx.x] = __ldg(a.tab + ((h << LR0) + kl - 1));
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        told_of[0] = told_of[1] = 0;
        for (int k = 0; k < PIPE_STAGES; k++)   // (slot k belongs to group k % 2: closed here counts as closed by the other one)
            if (issue((long long)first + (long long)k * stride < a.ntiles ? first + k * stride : 0x7fffffff, k, 0)) told_of[(k + 1) & 1] = 1;
    }
    __syncthreads();
    int ahead = 0;   // thread 0 of a group: the tile its next load will fetch
    if (t == 0) ahead = take();

    // thread-constant butterfly coordinates of the two radix-16 sub-passes
    const int j = t >> LN16;              // transform within the tile
    const int v = t & ((1 << LN16) - 1);  // butterfly within the transform
    const int cp = v & 15, kloc1 = v >> 4;
    // last sub-pass: stage (LN16 + s), position kappa = v -> tab[(h << LN16) + v - 1]. The four powers
    // w^8, w^4, w^2, w (h = 1, 2, 4, 8) are re-read per tile (L1 hits); h = 5, 9, 10, 11 are w^2 * W8, w * W16, w * W8,
    // w * W16^3, rebuilt per tile (16 FP64 instructions). Keeping the four powers in registers for the whole kernel was worth
    // 4 % at N = 4096 while tiles were assigned statically (1.31 vs 1.36 ms); with the tile hand-out below it spills (24 bytes)
    // and loses: same box, N = 4096 x 65536, best of 60: registers 1.35-1.46 ms, re-read 1.27-1.29 ms (round-1 kernel 1.29-1.37).
    // The smaller sizes never kept them: N = 512 1.41 -> 1.35, N = 1024 1.46 -> 1.33, N = 2048 1.42 -> 1.30 ms.
    const int rd1 = j * N + cp + 256 * kloc1;   // sub-pass 1 gather base
    const int wr1 = j * N + v;                  // sub-pass 1 scatter base
    const int rd2 = j * N + 16 * v;             // sub-pass 2 gather base
    const cd* const tw1p = tw1s + kloc1 * 8;
    const double sc = a.scale;

    int b = g % PIPE_STAGES;   // buffer of ring slot k
    uint32_t round = 0;        // k / PIPE_STAGES
    for (;;) {
        cd* const sm = bufs + (size_t)b * SLOT;
        if (t == 0) next_of[g] = ahead;   // (parks the counter value read during the previous tile)
        mbar_wait_bounded(&full[b + PIPE_STAGES * (round & 1)], (round >> 1) & 1);
        {
            const int tile_k = tile_of[b + PIPE_STAGES * (round & 1)];
            if (tile_k < 0) {   // a closed slot: pass on what this group holds for the slot three ahead; leave once the other group is sure to find a closed slot
                if (t == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the buffer's last readers are ordered before this thread by the barrier it has just waited for)
                    if (issue(next_of[g], b, round + 1)) told_of[g] = 1;
                    leave_of[g] = told_of[g];
                }
                group_sync(g);
                const int leave = leave_of[g];
                group_sync(g);
                if (leave) break;
                if (t == 0) ahead = take();
                b += 2;
                if (b >= PIPE_STAGES) { b -= PIPE_STAGES; round++; }
                continue;
            }
            if (t == 0) cur_of[PIPE_STAGES * g + b] = tile_k;   // read by the whole group after its next barrier; rewritten three tiles later
        }
        cd x[16];
        int rows_left = NTR;
        if constexpr (REAL != PIPE_C2C) {
            const long long tile_k = tile_of[b + PIPE_STAGES * (round & 1)];   // (still intact: the slot three ahead is not loaded before this group's last gather)
            if (a.batch - tile_k * NTR < NTR) rows_left = (int)(a.batch - tile_k * NTR);
        }
        (void)rows_left;   // caller's transforms in this tile; a pair's
                                                                                                    // second row beyond them reads as zeros
        // ---- sub-pass 0: radix R0, exact constants, in place (each thread owns idx = t + 256 e) ----
#pragma unroll
        for (int e = 0; e < 16; e++) {
            cd y;
            if constexpr (REAL == PIPE_R2C) {
                // z = x_a + i x_b of the rows 2 jj, 2 jj + 1 (jj == j: t + 256 e stays inside the thread's transform)
                const int idx = t + 256 * e, jj = idx >> LOGN, i = idx & (N - 1);
                const double* sr = reinterpret_cast<const double*>(sm) + 2 * jj * N + i;
                y = make_double2(sr[0], 2 * jj + 1 < rows_left ? sr[N] : 0.0);
            } else if constexpr (REAL == PIPE_C2R) {
                // Z = X_a + i X_b, X[i] = bin i for i <= N/2 and conj(bin N - i) above; conjugated for the inverse
                const int idx = t + 256 * e, jj = idx >> LOGN, i = idx & (N - 1);
                const bool up = i > N / 2;
                const cd* sp = sm + 2 * jj * NH + (up ? N - i : i);
                const cd A = sp[0];
                cd B = sp[NH];
                if (2 * jj + 1 >= rows_left) B = make_double2(0.0, 0.0);
                y = up ? make_double2(A.x + B.y, A.y - B.x) : make_double2(A.x - B.y, -A.y - B.x);
            } else if constexpr (REAL == PIPE_BLUE_FWD) {
                // a = x * conj(chirp), zero-padded from n to N (bluestein.c:107-109)
                const int idx = t + 256 * e, jj = idx >> LOGN, i = idx & (N - 1);
                y = make_double2(0.0, 0.0);
                if (i < a.n_user) {
                    const cd xv = sm[jj * a.n_user + i], w = __ldg(a.chirp + i);
                    y = make_double2(fma(xv.x, w.x, xv.y * w.y), fma(xv.y, w.x, -(xv.x * w.y)));
                }
            } else {
                y = sm[t + 256 * e];
                if (INV) y.y = -y.y;
            }
            x[(e / R0) * R0 + bitrev_c<LR0>(e % R0)] = y;
        }
        if constexpr (REAL == PIPE_R2C || REAL == PIPE_C2R || REAL == PIPE_BLUE_FWD) group_sync(g);   // the complex tile overwrites other threads' packed inputs: gather everything first
#pragma unroll
        for (int bb = 0; bb < NB0; bb++) SubStageExact<LR0, 1, 0, 0>::run(&x[bb * R0]);
        __syncwarp();  // the swizzle moves a thread's slots within its warp's 32-element rows
#pragma unroll
        for (int e = 0; e < 16; e++) sm[pipe_swz(t + 256 * e)] = x[e];
        group_sync(g);
        // ---- sub-pass 1: radix 16, M = R0, S = 16 ----
        {
            cd y[16];
#pragma unroll
            for (int rho = 0; rho < 16; rho++) y[bitrev_c<4>(rho)] = sm[pipe_swz(rd1 + 16 * rho)];
            cd tw[16];
            tw[1] = tw1p[0]; tw[2] = tw1p[1]; tw[4] = tw1p[2]; tw[5] = tw1p[3];
            tw[8] = tw1p[4]; tw[9] = tw1p[5]; tw[10] = tw1p[6]; tw[11] = tw1p[7];
            SubStageSym<4, 1, 0, 0>::run(y, tw);
            group_sync(g);  // every gather of this sub-pass is done
#pragma unroll
            for (int q = 0; q < 16; q++) sm[pipe_swz(wr1 + (q << LN16))] = y[q];
        }
        group_sync(g);
        // ---- sub-pass 2: radix 16, M = N/16, S = 1 ----
#pragma unroll
        for (int rho = 0; rho < 16; rho++) x[bitrev_c<4>(rho)] = sm[pipe_swz(rd2 + rho)];
        const cd wa = __ldg(a.tab + (v - 1) + (1 << LN16)), wb = __ldg(a.tab + (v - 1) + (2 << LN16));
        const cd wc = __ldg(a.tab + (v - 1) + (4 << LN16)), wd = __ldg(a.tab + (v - 1) + (8 << LN16));
        group_sync(g);  // the buffer is free: refill it with this CTA's tile k + 3 (r2c: after the partner exchange below)
        if (REAL != PIPE_R2C && t == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (issue(next_of[g], b, round + 1)) told_of[g] = 1;
        }
        {
            constexpr double C8 = 0.70710678118654752440, C16 = 0.92387953251128675613, S16 = 0.38268343236508977173;
            cd tw[16];
            tw[1] = wa; tw[2] = wb; tw[4] = wc; tw[8] = wd;
            tw[5] = make_double2((wc.x + wc.y) * C8, (wc.y - wc.x) * C8);  // w^2 * (1 - i)/sqrt2
            tw[9] = cmulc(wd, C16, -S16);
            tw[10] = make_double2((wd.x + wd.y) * C8, (wd.y - wd.x) * C8);
            tw[11] = cmulc(wd, S16, -C16);
            SubStageSym<4, 1, 0, 0>::run(x, tw);
        }
        {
            const long long tile = cur_of[PIPE_STAGES * g + b];   // (re-read: one register less to carry through the butterflies)
            const bool valid = tile * NT + j < a.batch;
            if constexpr (REAL == PIPE_R2C) {
                // x[q] = Z[k], k = v + (q << LN16). The bins k < N/2 (q < 8) need the partner Z[N - k], held as q' = 15 - q by thread
                // N/16 - v of the same transform: park the upper half in natural order, fetch the partners into the registers it leaves
                const cd zmid = x[8];                      // Z[N/2] where v == 0
#pragma unroll
                for (int q = 8; q < 16; q++) sm[wr1 + (q << LN16)] = x[q];
                group_sync(g);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int kk = v + (q << LN16);
                    x[8 + q] = kk == 0 ? x[0] : sm[j * N + N - kk];
                }
                group_sync(g);  // now the buffer is free
                if (t == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    if (issue(next_of[g], b, round + 1)) told_of[g] = 1;
                }
                // X_a = (Z + conj W) / 2 -> row 2j, X_b = (Z - conj W) / 2i -> row 2j + 1; bins 0 .. N/2 - 1, and the Nyquist bins (v == 0)
                const long long row = tile * NTR + 2 * j;
                cd* pa = a.out + (size_t)row * NH + v;
                cd* pb = pa + NH;
                const bool va = row < a.batch, vb = row + 1 < a.batch;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const cd z = x[q], w = x[8 + q];
                    if (va) pa[q << LN16] = make_double2(0.5 * (z.x + w.x), 0.5 * (z.y - w.y));
                    if (vb) pb[q << LN16] = make_double2(0.5 * (z.y + w.y), 0.5 * (w.x - z.x));
                }
                if (v == 0) {
                    if (va) pa[8 << LN16] = make_double2(zmid.x, 0.0);
                    if (vb) pb[8 << LN16] = make_double2(zmid.y, 0.0);
                }
            } else if constexpr (REAL == PIPE_C2R) {
                // the inverse of Z = X_a + i X_b is x_a + i x_b (conjugated and scaled here): real part -> row 2j, imaginary part -> row 2j + 1
                const long long row = tile * NTR + 2 * j;
                double* pa = reinterpret_cast<double*>(a.out) + (size_t)row * N + v;
                double* pb = pa + N;
                const bool va = row < a.batch, vb = row + 1 < a.batch;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    if (va) pa[q << LN16] = x[q].x * sc;
                    if (vb) pb[q << LN16] = -(x[q].y * sc);
                }
            } else if constexpr (REAL == PIPE_BLUE_FWD) {
                // A * FB (bluestein.c:124-131); factors fetched four at a time ahead of the stores they feed
                cd* p = a.out + tile * PIPE_TILE + wr1;
                if (valid) {
#pragma unroll
                    for (int q0 = 0; q0 < 16; q0 += 4) {
                        cd w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) w[i] = __ldg(a.fb + v + ((q0 + i) << LN16));
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const cd r = x[q0 + i];
                            p[(q0 + i) << LN16] = make_double2(fma(r.x, w[i].x, -(r.y * w[i].y)), fma(r.x, w[i].y, r.y * w[i].x));
                        }
                    }
                }
            } else if constexpr (REAL == PIPE_BLUE_INV) {
                // y = a * conj(chirp) * y_scale for the first n values of the row (bluestein.c:139-148)
                cd* p = a.out + (size_t)(tile * NT + j) * a.n_user + v;
                const double s2 = a.y_scale;
                if (valid) {
#pragma unroll
                    for (int q0 = 0; q0 < 16; q0 += 4) {
                        cd w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int idx = v + ((q0 + i) << LN16);
                            w[i] = idx < a.n_user ? __ldg(a.chirp + idx) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            cd r = x[q0 + i];
                            r.x *= sc; r.y *= -sc;
                            if (v + ((q0 + i) << LN16) < a.n_user)
                                p[(q0 + i) << LN16] = make_double2(fma(r.x, w[i].x, r.y * w[i].y) * s2, fma(r.y, w[i].x, -(r.x * w[i].y)) * s2);
                        }
                    }
                }
            } else {
                cd* p = a.out + tile * PIPE_TILE + wr1;
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        cd r = x[q];
                        if (INV) { r.x *= sc; r.y *= -sc; }
                        p[q << LN16] = r;
                    }
                }
            }
        }
        if (t == 0) ahead = take();   // for the load this group issues while it works on its next tile
        // the other group consumed the next ring slot; this group's next tile is two slots ahead
        b += 2;
        if (b >= PIPE_STAGES) { b -= PIPE_STAGES; round++; }
    }
    // every counter read of this CTA has returned (its value was parked or used): the last CTA resets the counters
    if (counted) {
        __syncthreads();
        if (threadIdx.x == 0 && atomicInc(a.sched + 1, gridDim.x - 1) == gridDim.x - 1) a.sched[0] = 0;
    }
}

}  // namespace fftb200
