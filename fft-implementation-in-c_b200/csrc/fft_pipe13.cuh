// fft_pipe13.cuh - whole transforms of 8192 points in ONE visit to shared memory (HBM sees the algorithmic 32 bytes
// per point and nothing passes through L2 twice): the 2^13 member of the fft_pipe.cuh family.
//
// Replaces the same reference code as fft_pipe.cuh (the butterfly loop of algorithms/core/radix2_dit.c:70-119 and the
// never-built cufftExecZ2Z call of gpu/fft_cuda.cu:166-185) for N = 8192, where one transform is two 64 KB tiles.
//
// Decomposition N = 4096 * 2 in the reference's DIT order: stages 1 .. 12 are two independent 4096-point transforms
// of the even and the odd samples x[c + 2 t] (c = 0, 1), stage 13 combines X[k], X[k + 4096] = Y0[k] +- w^k Y1[k].
//   * the two halves of a transform arrive as they lie in memory (x[0 .. 4095], x[4096 .. 8191]: full-rate bulk copies into
//     two ring buffers) and group c takes its samples x[c + 2 t] out of BOTH while gathering its first sub-pass: 16-byte
//     reads at a 32-byte stride (two shared-memory wavefronts per quarter warp instead of one). Letting the TMA do the
//     de-interleave (a [transform][t][c] tensor, 16 boxes of 256 rows of 16 bytes per half) was measured first: it keeps the
//     LSU out of it but writes 4096 16-byte rows per half into shared memory and doubles the L2 -> SM traffic: 1.92 vs
//     1.87 ms, and far less stable;
//   * after that first gather (and one CTA-wide barrier: both groups have read both buffers) group c runs the unchanged
//     4096-point dataflow of fft_pipe_kernel<12> in its own buffer: the halves h = 2 k + c of this CTA's transforms k go
//     through the same three-buffer ring with the same two-barriers-per-buffer scheme (tile index = h; both groups wait
//     for every fill);
//   * stage 13: each thread holds Y_c[t + 256 q], q < 16. Group 0 does the butterflies q < 8, group 1 q >= 8: the
//     halves trade 8 values per thread through group 1's buffer (32 KB each way), and since
//     w13[k + 2048] = -i w13[k] both groups use w13[t + 256 e] = w13[t] * W32^e (one table entry per thread, re-read from
//     L1 per transform to keep the register file free of spills). Group 0's buffer is not involved, so it is refilled
//     as early as in fft_pipe_kernel (after the last gather): that load - the second half of the NEXT transform - is the
//     one on the critical path, the first half of the transform after it goes into group 1's buffer after the trade;
//   * results leave from registers, X[k] and X[k + 4096], 512 contiguous bytes per warp instruction.
// Twiddles are the accurate tables for all 13 stages (SURVEY.md 7.0 hybrid rule; 8192-point mismatch vs the
// reference recurrence 2.6e-14, checked by the GPU parity tests at 1e-12).
#pragma once
#include "fft_fused.cuh"

namespace fftb200 {

__device__ __forceinline__ void bar_sync_n(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive_n(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// exp(-2 pi i e / 32), e < 8
template <int E> struct W32 {
    static constexpr double c = E == 0 ? 1.0 : E == 1 ? 0.98078528040323044913 : E == 2 ? 0.92387953251128675613 : E == 3 ? 0.83146961230254523708
                              : E == 4 ? 0.70710678118654752440 : E == 5 ? 0.55557023301960222474 : E == 6 ? 0.38268343236508977173 : 0.19509032201612826785;
    static constexpr double s = W32<8 - E>::c;
};
template <> struct W32<8> { static constexpr double c = 0.0, s = 1.0; };

// stage 13 for the 8 positions k = k0 + 256 e of one thread: w13[t + 256 e] = w13[t] * W32^e; the upper half of the
// positions (group 1, MI) uses -i times that. own[e] is this group's Y, z[e] the other group's.
template <bool INV, bool MI, int E>
struct Stage13 {
    static __device__ __forceinline__ void run(const cd* own, const cd* z, const cd w13, cd* p, const double sc) {
        const cd w = E == 0 ? w13 : cmulc(w13, W32<E>::c, -W32<E>::s);
        cd lo, hi;
        if constexpr (!MI) { lo = own[E]; hi = z[E]; bfly(lo, hi, w); }
        else { lo = z[E]; hi = own[E]; bfly_mi(lo, hi, w); }
        if (INV) { lo.x *= sc; lo.y *= -sc; hi.x *= sc; hi.y *= -sc; }
        p[256 * E] = lo;
        p[256 * E + 4096] = hi;
        if constexpr (E < 7) Stage13<INV, MI, E + 1>::run(own, z, w13, p, sc);
    }
};

template <bool INV>
__global__ void __launch_bounds__(2 * PIPE_GROUP, 1) fft_pipe13_kernel(const PipeArgs a) {
    constexpr int N = 8192, H = 4096;
    constexpr int LN16 = 8;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    cd* const tw1s = bufs + (size_t)PIPE_STAGES * PIPE_TILE;
    uint64_t* const full = reinterpret_cast<uint64_t*>(tw1s + PIPE_TW1);

    int* const tr_of = reinterpret_cast<int*>(full + 2 * PIPE_STAGES);   // [4]: the transform behind this CTA's k-th turn, k & 3

    const int g = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const int first = blockIdx.x, stride = gridDim.x;

    // Transforms are handed out on demand (fft_pipe.cuh "Tile order": the SMs do not all move data at the same rate). The two groups
    // of this kernel work on the same transform in lock step, so one thread keeps the list: the first two transforms of a CTA by
    // position, every later one from the global counter, read at the end of turn k - 1 for turn k + 2, parked in shared memory at the
    // top of turn k (in front of that turn's CTA-wide barriers) and used from turn k's refills on. A number past the end (the counter
    // stays exhausted once it is) ends the CTA at the turn that would have worked on it; halves of such a transform are never loaded.
    const bool counted = a.ntiles > 2LL * stride;
    auto take = [&]() -> int {
        if (!counted) return 0x7fffffff;
        const unsigned v = 2u * gridDim.x + atomicAdd(a.sched, 1u);
        return v < 0x7fffffffu ? (int)v : 0x7fffffff;
    };

    // half h = 2 k + (h & 1) of this CTA's k-th transform = elements [4096 (h & 1), + 4096) -> buffer b (= h % 3)
    auto issue = [&](int h, int b, uint32_t rnd) {
        const long long tr = tr_of[(h >> 1) & 3];
        if (tr >= a.ntiles) return;
        uint64_t* const bar = &full[b + PIPE_STAGES * (rnd & 1)];
        mbar_expect_tx(bar, H * (uint32_t)sizeof(cd));
        bulk_load(bufs + (size_t)b * PIPE_TILE, a.in + tr * N + (h & 1) * H, H * (uint32_t)sizeof(cd), bar);
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < 2 * PIPE_STAGES; b++) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // middle sub-pass twiddles -> shared: stage (4 + s), position kloc (fft_pipe.cuh)
    if (threadIdx.x < 16 * 8) {
        const int kl = threadIdx.x >> 3, e = threadIdx.x & 7;
        tw1s[threadIdx.x] = __ldg(a.tab + ((sym_h(e) << 4) + kl - 1));
    }
    int ahead = 0;   // thread 0: the transform of turn k + 2, until it is parked
    if (threadIdx.x == 0) {
        tr_of[0] = first < a.ntiles ? first : 0x7fffffff;
        tr_of[1] = (long long)first + stride < a.ntiles ? first + stride : 0x7fffffff;
        ahead = take();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int h = 0; h < PIPE_STAGES; h++) issue(h, h, 0);
    }

    const int cp = t & 15, kloc1 = t >> 4;
    const int rd1 = cp + 256 * kloc1;
    const cd* const tw1p = tw1s + kloc1 * 8;
    const double sc = a.scale;

    int b = g;                 // ring slot of half h (h % 3)
    uint32_t round = 0;        // h / 3
    for (int h = g;; h += 2) {
        cd* const sm = bufs + (size_t)b * PIPE_TILE;
        if (threadIdx.x == 0) tr_of[((h >> 1) + 2) & 3] = ahead;
        if (tr_of[(h >> 1) & 3] >= a.ntiles) break;   // (written two turns ago, or before the kernel's first barrier)
        mbar_wait_bounded(&full[b + PIPE_STAGES * (round & 1)], (round >> 1) & 1);   // a lost load traps instead of hanging the GPU
        cd x[16];
        // ---- sub-pass 0: radix 16, exact constants, in place ----
        {
            // the halves arrive as they lie in memory (x[0 .. 4095] and x[4096 .. 8191], full-rate bulk copies); group g takes
            // the samples x[g + 2 t] out of BOTH: 16-byte reads at a 32-byte stride (two wavefronts per quarter warp instead of
            // one) - cheaper than letting the TMA write 4096 16-byte rows per half into shared memory
            const int bp = g == 0 ? (b + 1 == PIPE_STAGES ? 0 : b + 1) : (b == 0 ? PIPE_STAGES - 1 : b - 1);
            const uint32_t rp = g == 0 ? round + (b + 1 == PIPE_STAGES ? 1 : 0) : round - (b == 0 ? 1 : 0);
            mbar_wait_bounded(&full[bp + PIPE_STAGES * (rp & 1)], (rp >> 1) & 1);
            const cd* const other = bufs + (size_t)bp * PIPE_TILE;
            const cd* const lo = g == 0 ? sm : other;
            const cd* const hi = g == 0 ? other : sm;
#pragma unroll
            for (int i = 0; i < 16; i++) {
                const int e = bitrev_c<4>(i);
                cd y = (e < 8 ? lo : hi)[g + 2 * (t + 256 * (e & 7))];
                if (INV) y.y = -y.y;
                x[i] = y;
            }
        }
        SubStageExact<4, 1, 0, 0>::run(x);
        bar_sync_n(6, 2 * PIPE_GROUP);   // both groups have read both halves: each scatters into its own buffer
#pragma unroll
        for (int e = 0; e < 16; e++) sm[pipe_swz(t + 256 * e)] = x[e];
        group_sync(g);
        // ---- sub-pass 1: radix 16 after 4 stages ----
        {
            cd y[16];
#pragma unroll
            for (int i = 0; i < 16; i++) y[i] = sm[pipe_swz(rd1 + 16 * bitrev_c<4>(i))];
            cd tw[16];
            load_sym(tw, tw1p);
            SubStageSym<4, 1, 0, 0>::run(y, tw);
            group_sync(g);
#pragma unroll
            for (int q = 0; q < 16; q++) sm[pipe_swz(t + (q << LN16))] = y[q];
        }
        group_sync(g);
        // ---- sub-pass 2: radix 16 after 8 stages ----
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = sm[pipe_swz(16 * t + bitrev_c<4>(i))];
        // the four last-sub-pass twiddle powers are re-read per half (L1 hits) instead of living in registers for the whole
        // kernel: no spills on the loop-carried path (same box: 1.874 -> 1.772 ms)
        const cd wa = __ldg(a.tab + (t - 1) + (1 << LN16)), wb = __ldg(a.tab + (t - 1) + (2 << LN16));
        const cd wc = __ldg(a.tab + (t - 1) + (4 << LN16)), wd = __ldg(a.tab + (t - 1) + (8 << LN16));
        group_sync(g);
        // group 0's buffer is free now: refill it with half h + 3 (the odd half of the next transform). Group 1's buffer
        // hosts the stage-13 trade first: tell group 0 that every gather from it is done.
        if (g == 0) {
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(h + PIPE_STAGES, b, round + 1);
            }
        } else {
            bar_arrive_n(3, 2 * PIPE_GROUP);
        }
        {
            constexpr double C8 = 0.70710678118654752440, C16 = 0.92387953251128675613, S16 = 0.38268343236508977173;
            cd tw[16];
            tw[1] = wa; tw[2] = wb; tw[4] = wc; tw[8] = wd;
            tw[5] = make_double2((wc.x + wc.y) * C8, (wc.y - wc.x) * C8);
            tw[9] = cmulc(wd, C16, -S16);
            tw[10] = make_double2((wd.x + wd.y) * C8, (wd.y - wd.x) * C8);
            tw[11] = cmulc(wd, S16, -C16);
            SubStageSym<4, 1, 0, 0>::run(x, tw);
        }
        // ---- stage 13: x[q] = Y_g[t + 256 q]. Trade through group 1's buffer: group 0 leaves its q >= 8 in the lower
        //      32 KB, group 1 its q < 8 in the upper 32 KB. Measured alternatives (same box, ms at 2^28 points): this form
        //      2.05; group 0's half in a separate 32 KB area and only the refilling warp waiting for the reads (one
        //      blocking CTA-wide barrier instead of three) 2.13; barrier 5 waited for by the refilling warp alone 2.06;
        //      group 1 held back by 200 / 400 / 800 ns per transform to take the groups out of phase 2.17 / 2.17 / 2.19.
        //      (Gathers issued in the order the butterflies consume them: 1.95 -> 1.90; the four last-sub-pass twiddles
        //      re-read per transform instead of living in registers - no spills - 1.95.) ----
        cd z[8];
        {
            cd* const xb = bufs + (size_t)(g == 0 ? (b + 1 == PIPE_STAGES ? 0 : b + 1) : b) * PIPE_TILE;   // buffer of half 2k + 1
            if (g == 0) {
                bar_sync_n(3, 2 * PIPE_GROUP);   // group 1 has gathered its last sub-pass
#pragma unroll
                for (int e = 0; e < 8; e++) xb[256 * e + t] = x[8 + e];
            } else {
#pragma unroll
                for (int e = 0; e < 8; e++) xb[2048 + 256 * e + t] = x[e];
            }
            bar_sync_n(4, 2 * PIPE_GROUP);
#pragma unroll
            for (int e = 0; e < 8; e++) z[e] = xb[(g == 0 ? 2048 : 0) + 256 * e + t];
            bar_sync_n(5, 2 * PIPE_GROUP);       // the trade has been read: group 1's buffer is free
            if (g == 1 && t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue(h + PIPE_STAGES, b, round + 1);
            }
        }
        {
            const long long tr = tr_of[(h >> 1) & 3];
            cd* const p = a.out + tr * N + t + (g ? 2048 : 0);
            const cd w13 = __ldg(a.tab + (H - 1) + t);   // stage 13, position t (L1-resident: the same entry for every transform)
            if (g == 0) Stage13<INV, false, 0>::run(x, z, w13, p, sc);
            else Stage13<INV, true, 0>::run(x + 8, z, w13, p, sc);
        }
        if (threadIdx.x == 0) ahead = take();
        b += 2;
        if (b >= PIPE_STAGES) { b -= PIPE_STAGES; round++; }
    }
    if (counted) {   // every counter read of this CTA has returned: the last CTA to finish resets the counters
        __syncthreads();
        if (threadIdx.x == 0 && atomicInc(a.sched + 1, gridDim.x - 1) == gridDim.x - 1) a.sched[0] = 0;
    }
}

const void* pipe13_func(int inverse);
cudaError_t launch_pipe13(const PipeArgs& a, int grid, cudaStream_t s);

}  // namespace fftb200
