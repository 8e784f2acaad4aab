// fft_fused.cuh - one persistent kernel for whole transforms of 2^13 .. 2^20 points: both passes of the
// two-pass (four-step) decomposition N = M * R run in the SAME launch and the intermediate array never
// leaves the L2 cache, so HBM sees the algorithmic 32 bytes per point (16 read + 16 written) once.
//
// Replaces, for these sizes, the same reference code as fft_tile.cuh / fft_pipe.cuh: the butterfly loop of
// algorithms/core/radix2_dit.c:70-119 (bit-reversal + log2 N radix-2 DIT stages) and the never-built
// cufftExecZ2Z call of gpu/fft_cuda.cu:166-185.
//
// Dataflow (Stockham autosort regrouping of the reference's radix-2 DIT stages, see fft_tile.cuh):
//   pass A  stages 1 .. log2 M: for every residue c in [0, R) the M-point transform of x[c + R t]; a tile is
//           C = 4096 / M adjacent residues x all M rows, fetched by one TMA tensor copy (box C x M out of the
//           row-major [batch * M][R] view of the input); result Y[c][k] goes to scratch[c + R k].
//   pass B  stages log2 M + 1 .. log2 N: for every k the R-point combination over c with the twiddles
//           T[stage][k + M q]; a tile is C2 = 4096 / R adjacent k, i.e. ONE contiguous 64 KB block of the
//           scratch array (1-D bulk copy); outputs X[k + M q] leave as rows of C2 contiguous elements.
// Both kinds of tile flow through the same 3 x 64 KB shared-memory ring. Three producer warps (one per ring
// buffer) issue the TMA copies; pass-B tiles of a group of transforms are issued only after every pass-A tile of that group
// has signalled completion through a global counter (release / acquire), and the work list interleaves
// A(g + lag) with B(g) so that the producer practically never waits. Scratch is a ring of `slots` groups
// (a few tens of MB): its lines are written and re-read while still in L2 (stores / loads carry evict_last /
// evict_first policies) - measured: no DRAM traffic beyond the algorithmic bytes up to ~24 MB of scratch.
//
// Twiddles: pass A covers stages m <= 1024 and uses the accurate tables with the w[q + m/4] = -i w[q]
// symmetry exactly like fft_pipe.cuh (hybrid rule of SURVEY.md 7.0); pass B reads the reference-recurrence
// table (host/ref_twiddle.c), which is what keeps large N within 1e-12 of the reference: 4 entries per thread
// and radix-16 butterfly, the other 11 derived with per-stage constants (fused_twiddles below).
#pragma once
#include <cuda.h>

#include <type_traits>

#include "fft_pipe.cuh"

namespace fftb200 {

struct FusedArgs {
    const cd* in;
    cd* out;
    cd* scratch;       // slots * gt * N elements
    const cd* tab;     // reference-recurrence stage tables for size N
    const cd* acc;     // accurate stage tables (stages m <= 8192)
    int* flags;        // [0, G): finished pass-A tiles per group; [G, 2G): finished pass-B tiles; [2G]: error
    long long nbatch;
    int gt;            // transforms per group
    int ngroups;       // G = ceil(nbatch / gt)
    int lag;           // pass B of group g is scheduled after pass A of group g + lag
    int slots;         // scratch ring depth in groups (>= lag + 1)
    int inverse;       // selects the INV instantiation (conjugate in, conjugate + scale out)
    long long* prof;   // development only (FUSED_PROF builds): per CTA and group {empty wait, full wait, total, tiles} in cycles
    int debug;         // development only: 1 = pass A alone, 2 = pass B alone, 4 = ignore the dependency counters
    double scale;      // 1/N for the inverse
    cd dtw[3][16];     // pass B, sub-pass j: dtw[j][h] = T[stage][q << a_tot] (table entry at kappa = 0), h = 2^(s-1) + q
};

constexpr int FUSED_THREADS = 2 * PIPE_GROUP + 32 * PIPE_STAGES + 32;   // two compute groups, one producer warp per ring buffer, one signal warp
constexpr int FUSED_TW1 = 16 * 8, FUSED_TW2 = 64 * 8;
constexpr int FUSED_NSB = 4;   // depth of the "tile stored" barrier ring between a compute group and the signal warp
// what a producer tells the compute group about the tile it put into a ring buffer
struct __align__(16) FusedDesc {
    int is_b;          // 0: pass-A tile, 1: pass-B tile
    int g;             // group (index of the completion counters)
    int kb;            // pass B: block of k values, k0 = kb << (12 - LR)
    int war_need;      // pass A: pass-B tiles of group g - slots that must have completed before the scratch stores (0: none)
    long long goff;    // element offset of the tile's output: A -> scratch, B -> out
    long long pad;
};
constexpr size_t FUSED_SMEM = (size_t)PIPE_STAGES * PIPE_TILE * sizeof(cd) + (FUSED_TW1 + FUSED_TW2) * sizeof(cd) +
                              PIPE_STAGES * sizeof(FusedDesc) + 256;

// ---- tile geometry: logical index I = lo + 2^LB * f + 2^(LB+LP) * hi, f = LP-bit transform field -------------
template <int LB, int LP, int LH, int A, int R, bool HIGH>
struct Geo {
    static constexpr int NCPP = LP - A - R;
    static constexpr int GSTRIDE = 1 << (NCPP + LB);      // gather: + rho * GSTRIDE
    static constexpr int SSTRIDE = 1 << (LP - R + LB);    // scatter: + q * SSTRIDE
    int lo, cpp, kloc, hi;
    __device__ __forceinline__ explicit Geo(int u) {
        if constexpr (HIGH) {
            hi = u & ((1 << LH) - 1); u >>= LH;
            lo = u & ((1 << LB) - 1); u >>= LB;
            cpp = u & ((1 << NCPP) - 1); u >>= NCPP;
            kloc = u;
        } else {
            lo = u & ((1 << LB) - 1); u >>= LB;
            cpp = u & ((1 << NCPP) - 1); u >>= NCPP;
            kloc = u & ((1 << A) - 1); u >>= A;
            hi = u;
        }
    }
    __device__ __forceinline__ int gbase() const { return lo + ((cpp + (kloc << (LP - A))) << LB) + (hi << (LB + LP)); }
    __device__ __forceinline__ int sbase() const { return lo + ((cpp + (kloc << NCPP)) << LB) + (hi << (LB + LP)); }
};

// physical position = I ^ (bit(I,B0) | bit(I,B1) << 1 | bit(I,B2) << 2); B0 < 0: identity (tools/swizzle_search.py)
template <int B0, int B1, int B2>
struct Swz {
    static __device__ __forceinline__ int f(int i) {
        if constexpr (B0 < 0) return i;
        else if constexpr (B1 == B0 + 1 && B2 == B0 + 2) return i ^ ((i >> B0) & 7);
        else return i ^ (((i >> B0) & 1) | (((i >> B1) & 1) << 1) | (((i >> B2) & 1) << 2));
    }
};
typedef Swz<-1, -1, -1> SwzId;

// exchange layouts found by tools/swizzle_search.py (conflict-free for every 128-bit access pattern)
template <int LM> struct SwzA2 { typedef SwzId type; };                 // pass A, before its third sub-pass
template <> struct SwzA2<10> { typedef Swz<4, 5, 6> type; };
template <int LR> struct SwzBlast { typedef Swz<LR, LR + 1, LR + 2> type; };   // pass B, before its last sub-pass
template <> struct SwzBlast<10> { typedef Swz<4, 10, 11> type; };

// ---- small PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol error traps (kernel fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity) {
#pragma unroll 1
    for (int it = 0; it < (1 << 24); it++) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
    }
    __trap();
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_count(const int* p, long long need) {
#pragma unroll 1
    for (int it = 0; it < (1 << 24); it++) {
        if (ld_acquire_gpu(p) >= need) return;
        __nanosleep(100);
    }
    __trap();
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint(cd* p, const cd v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_load_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// ---- work list ------------------------------------------------------------------------------------------------
// round rho = [pass-A tiles of group rho][pass-B tiles of group rho - lag]; every CTA walks the same list and
// takes items first, first + grid, ...; the list is a topological order of the A(g) -> B(g) dependencies.
struct FusedSched {
    long long nbatch;
    int gt, G, L, log_tpt, a_on, b_on;
    __device__ __forceinline__ long long tiles_of(int g) const {
        long long nb = nbatch - (long long)g * gt;
        if (nb > gt) nb = gt;
        return nb << log_tpt;
    }
    __device__ __forceinline__ long long round_len(int rho) const {
        long long n = 0;
        if (a_on && rho < G) n += tiles_of(rho);
        if (b_on && rho >= L && rho - L < G) n += tiles_of(rho - L);
        return n;
    }
};
struct FusedItem {
    int is_b, g;
    long long tau;   // tile within the group: (transform within group) * tiles_per_transform + block
};
struct FusedCursor {
    long long base = 0;
    int rho = 0;
    __device__ __forceinline__ FusedItem locate(const FusedSched& s, long long i) {
        long long len = s.round_len(rho);
        while (i >= base + len) { base += len; rho++; len = s.round_len(rho); }
        const long long off = i - base;
        const long long ta = (s.a_on && rho < s.G) ? s.tiles_of(rho) : 0;
        FusedItem it;
        if (off < ta) { it.is_b = 0; it.g = rho; it.tau = off; }
        else { it.is_b = 1; it.g = rho - s.L; it.tau = off - ta; }
        return it;
    }
};

// ---- sub-pass building blocks ---------------------------------------------------------------------------------
template <class G, class SW, int R, bool CONJ>
__device__ __forceinline__ void fused_gather(cd* x, const cd* sm, const G& g) {
    const int base = g.gbase();
#pragma unroll
    for (int rho = 0; rho < (1 << R); rho++) {
        cd y = sm[SW::f(base + rho * G::GSTRIDE)];
        if (CONJ) y.y = -y.y;
        x[bitrev_c<R>(rho)] = y;
    }
}
template <class G, class SW, int R>
__device__ __forceinline__ void fused_scatter(const cd* x, cd* sm, const G& g) {
    const int base = g.sbase();
#pragma unroll
    for (int q = 0; q < (1 << R); q++) sm[SW::f(base + q * G::SSTRIDE)] = x[q];
}
// the 8 stored twiddles of a symmetric radix-16 butterfly: h = 1, 2, 4, 5, 8, 9, 10, 11
__device__ __forceinline__ int sym_h(int e) { return e == 0 ? 1 : e == 1 ? 2 : e < 4 ? 2 + e : 4 + e; }
__device__ __forceinline__ void load_sym(cd* tw, const cd* p) {
    tw[1] = p[0]; tw[2] = p[1]; tw[4] = p[2]; tw[5] = p[3];
    tw[8] = p[4]; tw[9] = p[5]; tw[10] = p[6]; tw[11] = p[7];
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, int count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// End of a tile. The group leader first makes sure the signal warp has consumed the tile that used this barrier
// slot FUSED_NSB tiles ago (a parity wait is only valid one phase ahead) and posts the counter to increment
// (idx < 0: nothing to publish). PUBLISH: every warp arrives (release) after its last global store.
template <bool PUBLISH>
__device__ __forceinline__ void fused_signal(uint64_t* stored, volatile int* mail, volatile int* ack, int g2, int t, int m, int idx) {
    const int slot = g2 * FUSED_NSB + (m & (FUSED_NSB - 1));
    if (t == 0) {
#pragma unroll 1
        for (int it = 0; ack[g2] + FUSED_NSB <= m; it++)
            if (it > (1 << 26)) __trap();
        mail[slot] = idx;
    }
    if (PUBLISH) {
        __syncwarp();
        if ((t & 31) == 0) mbar_arrive_cnt(&stored[slot], 32);
    } else {
        if (t == 0) mbar_arrive_cnt(&stored[slot], PIPE_GROUP);
    }
}

// (a + ib)(c + id)
__device__ __forceinline__ cd cmul2(const cd a, const cd b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// Twiddles of a radix-2^R DIT butterfly after `a_tot` stages at position kappa (tp = table + kappa - 1). The R
// entries T[a_tot + s][kappa] (h = 2^(s-1)) are loaded; the others, T[a_tot + s][kappa + (q << a_tot)], are the
// product with d[h] = T[a_tot + s][q << a_tot], the table's own entry at kappa = 0. The reference recurrence
// w_j = fl(w_(j-1) w_m) makes its table multiplicative up to rounding noise, so this reproduces the reference's
// accumulated twiddle drift: whole-transform mismatch 1.3e-14 at N = 2^20 (tools: derived-twiddle check in DESIGN.md).
template <int R>
__device__ __forceinline__ void fused_twiddles(cd* tw, const cd* tp, const int a_tot, const cd* d) {
#pragma unroll
    for (int s = 1; s <= R; s++) {
        const int h0 = 1 << (s - 1);
        tw[h0] = __ldg(tp + ((size_t)h0 << a_tot));
    }
#pragma unroll
    for (int s = 2; s <= R; s++) {
        const int h0 = 1 << (s - 1);
#pragma unroll
        for (int q = 1; q < h0; q++) tw[h0 + q] = cmul2(tw[h0], d[h0 + q]);
    }
}

template <int LM, int LR, bool INV>
__global__ void __launch_bounds__(FUSED_THREADS, 1) fft_fused_kernel(const FusedArgs a, const __grid_constant__ CUtensorMap tmap) {
    static_assert(LM >= 6 && LM <= 10 && LR >= 6 && LR <= 10, "pass sizes 64 .. 1024");
    constexpr int LOGN = LM + LR, LOG_TPT = LOGN - 12;
    constexpr int LC = 12 - LM, LC2 = 12 - LR;          // log2 columns per A tile / k's per B tile
    constexpr int A3 = LM >= 9, B3 = LR >= 9;            // three sub-passes?
    constexpr int RA0 = A3 ? LM - 8 : LM - 4, RB0 = B3 ? LR - 8 : LR - 4;
    constexpr int NBOX = LM > 8 ? 1 << (LM - 8) : 1, BOXROWS = LM > 8 ? 256 : 1 << LM;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    cd* const tw1s = bufs + (size_t)PIPE_STAGES * PIPE_TILE;
    cd* const tw2s = tw1s + FUSED_TW1;
    FusedDesc* const descs = reinterpret_cast<FusedDesc*>(tw2s + FUSED_TW2);
    uint64_t* const full = reinterpret_cast<uint64_t*>(descs + PIPE_STAGES);
    uint64_t* const empty = full + PIPE_STAGES;
    uint64_t* const stored = empty + PIPE_STAGES;                          // [2][NSB] per compute group: all threads stored their tile
    volatile int* const mail = reinterpret_cast<volatile int*>(stored + 2 * FUSED_NSB);   // [2][NSB] counter to increment (-1: none)
    volatile int* const ack = mail + 2 * FUSED_NSB;                        // [2] tiles of the group the signal warp has consumed

    FusedSched sch;
    sch.nbatch = a.nbatch; sch.gt = a.gt; sch.G = a.ngroups; sch.L = a.lag; sch.log_tpt = LOG_TPT;
    sch.a_on = !(a.debug & 2); sch.b_on = !(a.debug & 1);
    const bool nowait = (a.debug & 7) != 0;
    const long long total = (a.nbatch << LOG_TPT) * (sch.a_on + sch.b_on);
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tiles = first < total ? (int)((total - first + stride - 1) / stride) : 0;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < PIPE_STAGES; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], 1); }
        for (int i = 0; i < 2 * FUSED_NSB; i++) mbar_init(&stored[i], PIPE_GROUP);
        ack[0] = 0; ack[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // pass-A twiddle tables (accurate, symmetric): sub-pass 1 after RA0 stages, sub-pass 2 after RA0 + 4
    for (int i = threadIdx.x; i < (8 << RA0); i += FUSED_THREADS) tw1s[i] = __ldg(a.acc + ((sym_h(i & 7) << RA0) + (i >> 3) - 1));
    if constexpr (A3)
        for (int i = threadIdx.x; i < (8 << (RA0 + 4)); i += FUSED_THREADS)
            tw2s[i] = __ldg(a.acc + ((sym_h(i & 7) << (RA0 + 4)) + (i >> 3) - 1));
    __syncthreads();

    // =========================== signal warp ===========================
    // Publishes pass-A tile completions (gpu-scope fence + counter increment) so that no compute thread ever waits
    // for a fence: compute warps arrive on `stored[group][tile & 3]` right after their last store and move on. One
    // fence covers every tile found complete, so the warp keeps up however long a fence takes under store traffic.
    if (threadIdx.x >= 2 * PIPE_GROUP + 32 * PIPE_STAGES) {
        if ((threadIdx.x & 31) == 0) {
            int done[2] = {0, 0};
            const int cnt[2] = {(my_tiles + 1) >> 1, my_tiles >> 1};
            int idle = 0;
            while (done[0] < cnt[0] || done[1] < cnt[1]) {
                int c[2] = {0, 0};
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int i = 0; i < FUSED_NSB; i++) {
                        const int m = done[g] + i;
                        if (c[g] == i && m < cnt[g] && mbar_test(&stored[g * FUSED_NSB + (m & (FUSED_NSB - 1))], (m / FUSED_NSB) & 1)) c[g]++;
                    }
                if (c[0] + c[1] == 0) {
                    if (++idle > (1 << 26)) __trap();
                    continue;
                }
                idle = 0;
                int idx[2][FUSED_NSB];
                bool any = false;
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int i = 0; i < FUSED_NSB; i++) {
                        idx[g][i] = i < c[g] ? mail[g * FUSED_NSB + ((done[g] + i) & (FUSED_NSB - 1))] : -1;
                        any |= idx[g][i] >= 0;
                    }
                if (any) __threadfence();
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int i = 0; i < FUSED_NSB; i++)
                        if (idx[g][i] >= 0) atomicAdd(a.flags + idx[g][i], 1);
                done[0] += c[0]; done[1] += c[1];
                ack[0] = done[0]; ack[1] = done[1];
            }
        }
        return;
    }

    // =========================== producer warps ===========================
    // One warp (one lane) per ring buffer: items k = w, w + 3, ... The dependency counter of the NEXT item is
    // polled while the consumers still work on the buffer, so a refill is issued the moment the buffer is free.
    if (threadIdx.x >= 2 * PIPE_GROUP) {
        const int w = (threadIdx.x - 2 * PIPE_GROUP) >> 5;   // 0 .. 2
        if ((threadIdx.x & 31) == 0) {
            const uint64_t pol_in = policy_evict_first(), pol_sc = policy_evict_last();
            FusedCursor cur;
            cd* const dst = bufs + (size_t)w * PIPE_TILE;
            for (int k = w, n = 0; k < my_tiles; k += PIPE_STAGES, n++) {
                const FusedItem it = cur.locate(sch, first + (long long)k * stride);
                FusedDesc d;
                d.is_b = it.is_b; d.g = it.g; d.pad = 0;
                const int blk = (int)(it.tau & ((1 << LOG_TPT) - 1));
                if (!it.is_b) {
                    const long long tr = (long long)it.g * a.gt + (it.tau >> LOG_TPT);
                    const long long trl = (long long)(it.g % a.slots) * a.gt + (it.tau >> LOG_TPT);
                    d.kb = 0;
                    d.war_need = (it.g >= a.slots && !nowait) ? (int)sch.tiles_of(it.g - a.slots) : 0;
                    d.goff = (trl << LOGN) + (blk << LC);
                    if (n >= 1) mbar_wait_bounded(&empty[w], (n - 1) & 1);
                    descs[w] = d;
                    mbar_expect_tx(&full[w], PIPE_TILE * (uint32_t)sizeof(cd));
#pragma unroll
                    for (int bx = 0; bx < NBOX; bx++)
                        tma_load_2d(dst + (size_t)bx * (BOXROWS << LC), &tmap, 2 * (blk << LC), (int)((tr << LM) + bx * BOXROWS),
                                    &full[w], pol_in);
                } else {
                    if (!nowait) wait_count(a.flags + it.g, sch.tiles_of(it.g));
                    const cd* src = a.scratch + (((size_t)(it.g % a.slots) * a.gt) << LOGN) + (size_t)it.tau * PIPE_TILE;
                    const long long tr = (long long)it.g * a.gt + (it.tau >> LOG_TPT);
                    d.kb = blk; d.war_need = 0;
                    d.goff = (tr << LOGN) + (blk << LC2);
                    if (n >= 1) mbar_wait_bounded(&empty[w], (n - 1) & 1);
                    descs[w] = d;
                    asm volatile("fence.proxy.async;" ::: "memory");
                    mbar_expect_tx(&full[w], PIPE_TILE * (uint32_t)sizeof(cd));
                    bulk_load_hint(dst, src, PIPE_TILE * (uint32_t)sizeof(cd), &full[w], pol_sc);
                }
            }
        }
        return;
    }

    // =========================== compute groups ===========================
    const int g2 = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
    const double sc = a.scale;
    int b = g2, n = 0;   // ring buffer and use count of tile k = g2, g2 + 2, ... (b = k % 3, n = k / 3)
    int m = 0;           // tiles this group has completed
#ifdef FUSED_PROF
    long long pr_e = 0, pr_f = 0, pr_s = 0;
    long long pr_ph[2][6] = {{0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0}};
    long long pr_last = 0;
    const long long pr_t0 = clock64();
#define PROF_MARK(kind, i) { const long long c_ = clock64(); pr_ph[kind][i] += c_ - pr_last; pr_last = c_; }
#else
#define PROF_MARK(kind, i)
#endif
    for (int k = g2; k < my_tiles; k += 2, m++) {
        cd* const sm = bufs + (size_t)b * PIPE_TILE;
#ifdef FUSED_PROF
        const long long c0 = clock64();
#endif
        // A parity wait is only valid one phase ahead: first make sure the previous tile of this buffer (handled by
        // the other group) has been consumed - its load may complete later than the load of this group's last tile.
        if (n >= 1) mbar_wait_bounded(&empty[b], (n - 1) & 1);
#ifdef FUSED_PROF
        const long long c1 = clock64();
#endif
        mbar_wait_bounded(&full[b], n & 1);
#ifdef FUSED_PROF
        const long long c2 = clock64();
        pr_e += c1 - c0; pr_f += c2 - c1; pr_last = c2;
#endif
        const FusedDesc it = descs[b];
        cd x[16];
        if (!it.is_b) {
            // write-after-read on the scratch ring: the slot of this group was last read by pass B of group g - slots.
            // Poll its completion counter now, look at the answer just before the stores.
            const int* war_p = nullptr;
            int war_seen = 0;
            if (t == 0 && it.war_need) {
                war_p = a.flags + a.ngroups + (it.g - a.slots);
                war_seen = ld_acquire_gpu(war_p);
            }
            // ------------------------------ pass A: stages 1 .. LM over C = 2^LC columns ------------------------------
            // sub-pass 0: radix 2^RA0, exact constants, in place per thread
            {
                typedef Geo<LC, LM, 0, 0, RA0, false> G0;
                constexpr int NB = 16 >> RA0, R0 = 1 << RA0;
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_gather<G0, SwzId, RA0, INV>(&x[bb * R0], sm, g);
                    SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                    fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, g);
                }
            }
            PROF_MARK(0, 0)
            group_sync(g2);
            PROF_MARK(0, 1)
            // sub-pass 1: radix 16 after RA0 stages
            typedef Geo<LC, LM, 0, RA0, 4, false> G1;
            const G1 g1(t);
            fused_gather<G1, SwzId, 4, false>(x, sm, g1);
            {
                cd tw[16];
                load_sym(tw, tw1s + g1.kloc * 8);
                SubStageSym<4, 1, 0, 0>::run(x, tw);
            }
            int kq = g1.kloc, lo = g1.lo;   // output k = kq + (q << stages so far)
            if constexpr (A3) {
                typedef typename SwzA2<LM>::type SW2;
                group_sync(g2);   // every gather of sub-pass 1 is done
                fused_scatter<G1, SW2, 4>(x, sm, g1);
                group_sync(g2);
                typedef Geo<LC, LM, 0, RA0 + 4, 4, false> G2;
                const G2 gg(t);
                fused_gather<G2, SW2, 4, false>(x, sm, gg);
                cd tw[16];
                load_sym(tw, tw2s + gg.kloc * 8);
                SubStageSym<4, 1, 0, 0>::run(x, tw);
                kq = gg.kloc; lo = gg.lo;
            }
            PROF_MARK(0, 2)
            if (war_p && war_seen < it.war_need) wait_count(war_p, it.war_need);
            group_sync(g2);   // the ring buffer is free, the scratch slot is drained
            PROF_MARK(0, 3)
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&empty[b]);
            }
            {
                cd* p = a.scratch + it.goff + lo + ((size_t)kq << LR);
#pragma unroll
                for (int q = 0; q < 16; q++) st_hint(p + ((size_t)q << (LM - 4 + LR)), x[q], pol_keep);
            }
            PROF_MARK(0, 4)
            fused_signal<true>(stored, mail, ack, g2, t, m, it.g);
            PROF_MARK(0, 5)
        } else {
            // ------------------------------ pass B: stages LM + 1 .. LM + LR for C2 = 2^LC2 values of k ------------------------------
            const int kb = it.kb;                                       // k0 = kb << LC2
            // the scratch block is in shared memory now: its ring slot may be overwritten (no fence needed for a read)
            if (t == 0) atomicAdd(a.flags + a.ngroups + it.g, 1);
            typedef typename SwzBlast<LR>::type SWL;
            typedef typename std::conditional<B3, SwzId, SWL>::type SW1;   // layout after sub-pass 0
            {
                typedef Geo<0, LR, LC2, 0, RB0, false> G0;
                constexpr int NB = 16 >> RB0, R0 = 1 << RB0;
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_gather<G0, SwzId, RB0, false>(&x[bb * R0], sm, g);
                }
                cd tw[R0];
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_twiddles<RB0>(tw, a.tab + ((kb << LC2) + g.hi - 1), LM, a.dtw[0]);
                    SubStageGen<RB0, 1, 0, 0>::run(&x[bb * R0], tw);
                }
                PROF_MARK(1, 0)
                group_sync(g2);   // every gather of sub-pass 0 is done (the layout changes)
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_scatter<G0, SW1, RB0>(&x[bb * R0], sm, g);
                }
            }
            group_sync(g2);
            PROF_MARK(1, 1)
            if constexpr (B3) {
                typedef Geo<0, LR, LC2, RB0, 4, false> G1;
                const G1 g1(t);
                fused_gather<G1, SW1, 4, false>(x, sm, g1);
                {
                    cd tw[16];
                    fused_twiddles<4>(tw, a.tab + ((kb << LC2) + g1.hi + (g1.kloc << LM) - 1), LM + RB0, a.dtw[1]);
                    SubStageGen<4, 1, 0, 0>::run(x, tw);
                }
                group_sync(g2);
                fused_scatter<G1, SWL, 4>(x, sm, g1);
                group_sync(g2);
            }
            constexpr int AL = LR - 4;   // stages of this pass done before the last sub-pass
            typedef Geo<0, LR, LC2, AL, 4, true> GL;
            const GL gl(t);
            fused_gather<GL, SWL, 4, false>(x, sm, gl);
            PROF_MARK(1, 2)
            group_sync(g2);   // the ring buffer is free
            PROF_MARK(1, 3)
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(&empty[b]);
            }
            {
                cd tw[16];
                fused_twiddles<4>(tw, a.tab + ((kb << LC2) + gl.hi + (gl.kloc << LM) - 1), LM + AL, a.dtw[2]);
                SubStageGen<4, 1, 0, 0>::run(x, tw);
            }
            {
                cd* p = a.out + it.goff + gl.hi + ((size_t)gl.kloc << LM);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    cd r = x[q];
                    if (INV) { r.x *= sc; r.y *= -sc; }
                    st_hint(p + ((size_t)q << (LM + AL)), r, pol_stream);
                }
            }
            PROF_MARK(1, 4)
            fused_signal<false>(stored, mail, ack, g2, t, m, -1);
            PROF_MARK(1, 5)
        }
        b += 2;
        if (b >= PIPE_STAGES) { b -= PIPE_STAGES; n++; }
    }
#ifdef FUSED_PROF
    if (t == 0 && a.prof) {
        long long* q = a.prof + (blockIdx.x * 2 + g2) * 16;
        q[0] = pr_e; q[1] = pr_f; q[2] = clock64() - pr_t0; q[3] = m;
        for (int i = 0; i < 6; i++) { q[4 + i] = pr_ph[0][i]; q[10 + i] = pr_ph[1][i]; }
    }
#endif
}

// defined in fft_kernels_fused{0..3}.cu: nullptr / false when the (lm, lr) pair is not compiled in that unit
#define FUSED_UNIT_DECL(I)                                                    \
    const void* fused_func_##I(int lm, int lr, int inverse);                  \
    bool launch_fused_##I(int lm, int lr, const FusedArgs& a, const CUtensorMap& tmap, int grid, cudaStream_t s);
FUSED_UNIT_DECL(0) FUSED_UNIT_DECL(1) FUSED_UNIT_DECL(2) FUSED_UNIT_DECL(3)
#undef FUSED_UNIT_DECL
inline const void* fused_func(int lm, int lr, int inverse) {
    const void* f = fused_func_0(lm, lr, inverse);
    if (!f) f = fused_func_1(lm, lr, inverse);
    if (!f) f = fused_func_2(lm, lr, inverse);
    if (!f) f = fused_func_3(lm, lr, inverse);
    return f;
}
inline bool launch_fused(int lm, int lr, const FusedArgs& a, const CUtensorMap& tmap, int grid, cudaStream_t s) {
    return launch_fused_0(lm, lr, a, tmap, grid, s) || launch_fused_1(lm, lr, a, tmap, grid, s) ||
           launch_fused_2(lm, lr, a, tmap, grid, s) || launch_fused_3(lm, lr, a, tmap, grid, s);
}

}  // namespace fftb200
