// fft_lastpipe.cuh - the last pass of multi-pass plans (N >= 2^21) as a persistent TMA-ring kernel: the pass-B dataflow of
// fft_fused.cuh fed like fft_pipe_kernel.
//
// Replaces fft_tile_kernel<LAST> where no elementwise factor or peer store rides on the pass (it keeps those). Same reference
// code as every other kernel here: the late stages of the butterfly loop of algorithms/core/radix2_dit.c:84-112 with the
// reference's own recurrence twiddles (host/ref_twiddle.c).
//
// A tile is C2 = 4096 / P adjacent k of one transform times all P = 2^LR rows: one contiguous 64 KB block of the [k][c]
// layout the head of the plan leaves behind. It is fetched by a 1-D bulk copy into a three-buffer ring two tiles ahead of
// the butterflies (the tile kernel waits for its gather in front of every tile: 28 % of its stall samples at 2^24), goes
// through the radix-2^(LR-4 or LR-8), [16,] 16 sub-passes of the fused kernel's pass B with in-place exchanges, and leaves
// from registers as rows of C2 contiguous elements at stride M, X[k + M q]. Tiles walk the batch first (tile = k block *
// batch + transform), so the late-stage twiddles of a k block - as many bytes as the tile - are read from HBM once per
// execution and from L2 by the other transforms. Every twiddle is loaded from the table (no derived products): results
// are the tile kernel's to the last bit for P = 128, 256 (same radix split) and within 1e-15 otherwise.
#pragma once
#include "fft_fused.cuh"

namespace fftb200 {

struct LastPipeArgs {
    const cd* in;
    cd* out;
    const cd* tab;      // reference-recurrence stage tables for size N
    long long ntiles;   // batch * M / C2
    long long batch;
    int log_n, log_m;   // N = 2^log_n, M = 2^log_m stages-worth of points done by earlier passes
    int inverse;
    double scale;       // applied with the final conjugation when inverse
    unsigned int* sched;   // tile hand-out counters (fft_pipe.cuh "Tile order"): [0] tiles taken beyond the first three per CTA, [1] CTAs finished
    // derive != 0: of the 2^R - 1 twiddles of a radix-2^R butterfly only the R entries T[a + s][kappa] are loaded, the others are the product
    // with dtw[j][h] = T[a + s][q << a], the table's own entry at kappa = 0 (fft_fused.cuh fused_twiddles: the reference's recurrence is
    // multiplicative up to rounding noise): a quarter of the table traffic, which for the last stages is as many bytes as the data
    int derive;
    cd dtw[3][16];         // sub-pass j (first / middle / last): a = log_m, log_m + RB0, log_m + LR - 4
};

template <int R>
__device__ __forceinline__ void table_twiddles(cd* tw, const cd* tp, const int a_tot) {
#pragma unroll
    for (int h = 1; h < (1 << R); h++) tw[h] = __ldg(tp + ((size_t)h << a_tot));
}
template <int R, bool DERIVE>
__device__ __forceinline__ void last_twiddles(cd* tw, const cd* tp, const int a_tot, const cd* d) {
    if constexpr (DERIVE) fused_twiddles<R>(tw, tp, a_tot, d);
    else table_twiddles<R>(tw, tp, a_tot);
}

// DERIVE is compiled in (as a run-time switch next to the table loads it cost both forms 72 - 172 bytes of spills)
template <int LR, bool INV, bool DERIVE>
__global__ void __launch_bounds__(2 * PIPE_GROUP, 1) fft_lastpipe_kernel(const LastPipeArgs a) {
    static_assert(LR >= 5 && LR <= 9, "last pass of 32 .. 512 points");
    constexpr int LC2 = 12 - LR;                 // log2 k's per tile
    constexpr int B3 = LR >= 9;                  // three sub-passes?
    constexpr int RB0 = B3 ? LR - 8 : LR - 4;
    constexpr int AL = LR - 4;                   // stages of this pass done before the last sub-pass

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    uint64_t* const full = reinterpret_cast<uint64_t*>(bufs + (size_t)PIPE_STAGES * PIPE_TILE);

    const int g2 = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const int first = blockIdx.x, stride = gridDim.x;
    const int lm = a.log_m;
    const double sc = a.scale;
    // tiles handed out on demand, exactly as in fft_pipe_kernel ("Tile order" / "End of work" there): the first three of a CTA by
    // position, the others from the global counter in the same batch-first order, read one tile ahead
    int* const tile_of = reinterpret_cast<int*>(full + 2 * PIPE_STAGES);   // [6]: the tile of a barrier's current phase, -1 = closed
    int* const next_of = tile_of + 2 * PIPE_STAGES;                        // [2]: the tile a group will load next
    int* const told_of = next_of + 2;                                      // [2]: a slot of the other group has been closed
    int* const leave_of = told_of + 2;                                     // [2]
    const bool counted = a.ntiles > 3LL * stride;
    auto take = [&]() -> int {
        if (!counted) return 0x7fffffff;
        const unsigned v = 3u * gridDim.x + atomicAdd(a.sched, 1u);
        return v < 0x7fffffffu ? (int)v : 0x7fffffff;
    };

    // same two-barriers-per-buffer scheme as fft_pipe_kernel
    auto issue = [&](int tile, int b, uint32_t rnd) -> bool {   // true: nothing left, the slot is closed
        uint64_t* const bar = &full[b + PIPE_STAGES * (rnd & 1)];
        if (tile >= a.ntiles) {
            tile_of[b + PIPE_STAGES * (rnd & 1)] = -1;
            mbar_arrive(bar);
            return true;
        }
        tile_of[b + PIPE_STAGES * (rnd & 1)] = tile;
        const long long tr = tile % a.batch, kb = tile / a.batch;
        mbar_expect_tx(bar, PIPE_TILE * (uint32_t)sizeof(cd));
        bulk_load(bufs + (size_t)b * PIPE_TILE, a.in + (tr << a.log_n) + (kb << 12), PIPE_TILE * (uint32_t)sizeof(cd), bar);
        return false;
    };
    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2 * PIPE_STAGES; b++) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        told_of[0] = told_of[1] = 0;
        for (int k = 0; k < PIPE_STAGES; k++)
            if (issue((long long)first + (long long)k * stride < a.ntiles ? first + k * stride : 0x7fffffff, k, 0)) told_of[(k + 1) & 1] = 1;
    }
    __syncthreads();
    int ahead = 0;
    if (t == 0) ahead = take();

    typedef typename SwzBlast<LR>::type SWL;
    typedef typename std::conditional<B3, SwzId, SWL>::type SW1;   // layout after sub-pass 0

    int b = g2 % PIPE_STAGES;
    uint32_t round = 0;
    for (;;) {
        cd* const sm = bufs + (size_t)b * PIPE_TILE;
        if (t == 0) next_of[g2] = ahead;
        mbar_wait_bounded(&full[b + PIPE_STAGES * (round & 1)], (round >> 1) & 1);
        const int tile = tile_of[b + PIPE_STAGES * (round & 1)];
        if (tile < 0) {   // a closed slot: pass on what this group holds; leave once it has closed a slot of the other group as well
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // (the buffer's last readers are ordered before this thread by the barrier it has just waited for)
                if (issue(next_of[g2], b, round + 1)) told_of[g2] = 1;
                leave_of[g2] = told_of[g2];
            }
            group_sync(g2);
            const int leave = leave_of[g2];
            group_sync(g2);
            if (leave) break;
            if (t == 0) ahead = take();
            b += 2;
            if (b >= PIPE_STAGES) { b -= PIPE_STAGES; round++; }
            continue;
        }
        const long long tr = tile % a.batch;
        const int kb = (int)(tile / a.batch);
        cd x[16];
        {
            typedef Geo<0, LR, LC2, 0, RB0, false> G0;
            constexpr int NB = 16 >> RB0, R0 = 1 << RB0;
#pragma unroll
            for (int bb = 0; bb < NB; bb++) {
                const G0 g(t + PIPE_GROUP * bb);
                fused_gather<G0, SwzId, RB0, INV>(&x[bb * R0], sm, g);
            }
            cd tw[R0];
#pragma unroll
            for (int bb = 0; bb < NB; bb++) {
                const G0 g(t + PIPE_GROUP * bb);
                last_twiddles<RB0, DERIVE>(tw, a.tab + ((kb << LC2) + g.hi - 1), lm, a.dtw[0]);
                SubStageGen<RB0, 1, 0, 0>::run(&x[bb * R0], tw);
            }
            group_sync(g2);   // every gather of sub-pass 0 is done (the layout changes)
#pragma unroll
            for (int bb = 0; bb < NB; bb++) {
                const G0 g(t + PIPE_GROUP * bb);
                fused_scatter<G0, SW1, RB0>(&x[bb * R0], sm, g);
            }
        }
        group_sync(g2);
        if constexpr (B3) {
            typedef Geo<0, LR, LC2, RB0, 4, false> G1;
            const G1 g1(t);
            fused_gather<G1, SW1, 4, false>(x, sm, g1);
            {
                cd tw[16];
                last_twiddles<4, DERIVE>(tw, a.tab + ((kb << LC2) + g1.hi + ((size_t)g1.kloc << lm) - 1), lm + RB0, a.dtw[1]);
                SubStageGen<4, 1, 0, 0>::run(x, tw);
            }
            group_sync(g2);
            fused_scatter<G1, SWL, 4>(x, sm, g1);
            group_sync(g2);
        }
        typedef Geo<0, LR, LC2, AL, 4, true> GL;
        const GL gl(t);
        fused_gather<GL, SWL, 4, false>(x, sm, gl);
        group_sync(g2);   // the buffer is free: refill it with this CTA's tile k + 3
        if (t == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (issue(next_of[g2], b, round + 1)) told_of[g2] = 1;
        }
        {
            // (asking for these 15 entries before the exchange settles - they only depend on the tile and the thread - spills
            // 190 bytes and is slower: 4.49 vs 4.21 ms at 2^24 x 16)
            cd tw[16];
            last_twiddles<4, DERIVE>(tw, a.tab + ((kb << LC2) + gl.hi + ((size_t)gl.kloc << lm) - 1), lm + AL, a.dtw[2]);
            SubStageGen<4, 1, 0, 0>::run(x, tw);
        }
        {
            // X[k + M q], k = (kb << LC2) + hi, q = kloc + (q' << AL): lanes run over hi first, so a warp instruction writes
            // 32 / C2 rows of C2 contiguous elements
            cd* p = a.out + (tr << a.log_n) + ((size_t)kb << LC2) + gl.hi + ((size_t)gl.kloc << lm);
#pragma unroll
            for (int q = 0; q < 16; q++) {
                cd r = x[q];
                if (INV) { r.x *= sc; r.y *= -sc; }
                p[(size_t)q << (AL + lm)] = r;
            }
        }
        if (t == 0) ahead = take();
        b += 2;
        if (b >= PIPE_STAGES) { b -= PIPE_STAGES; round++; }
    }
    if (counted) {   // every counter read of this CTA has returned: the last CTA to finish resets the counters
        __syncthreads();
        if (threadIdx.x == 0 && atomicInc(a.sched + 1, gridDim.x - 1) == gridDim.x - 1) a.sched[0] = 0;
    }
}

constexpr size_t LASTPIPE_SMEM = (size_t)PIPE_STAGES * PIPE_TILE * sizeof(cd) + 128;   // + barriers, tile numbers
const void* lastpipe_func(int lr, int inverse, int derive = 0);   // fft_kernels_lastpipe.cu
cudaError_t launch_lastpipe(int lr, const LastPipeArgs& a, int grid, cudaStream_t s);

}  // namespace fftb200
