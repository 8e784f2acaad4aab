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
//           row-major [batch * M][R] view of the input); the result Y[c][k] goes to scratch[c + R k] through
//           the mirror-image TMA tensor store.
//   pass B  stages log2 M + 1 .. log2 N: for every k the R-point combination over c with the twiddles
//           T[stage][k + M q]; a tile is C2 = 4096 / R adjacent k, i.e. ONE contiguous 64 KB block of the
//           scratch array (1-D bulk copy); outputs X[k + M q] leave as a TMA tensor store of R rows of C2
//           contiguous elements into the [batch * R][M] view of the output.
// All global traffic is asynchronous (TMA both ways): the compute threads only touch shared memory. A tile
// lives in one of three 64 KB ring buffers from its load to the end of its store; each buffer has a manager
// warp (one lane) that issues the load, waits until a compute group has staged the result in place, issues
// the store, and - once the store has read the buffer - the next load. Two groups of 256 compute threads
// take the loaded tiles alternately (16 points per thread, radix-16 sub-passes, in-place exchanges with the
// conflict-free layouts found by tools/swizzle_search.py).
//
// Dependencies: pass-B tiles of a group of transforms may be loaded only after every pass-A tile of that
// group has been stored: the manager that completes an A store publishes it (gpu-scope fence + counter
// increment), the manager of a B tile acquires the counter before issuing the load. The work list
// interleaves A(g + lag) with B(g), so counters are normally long satisfied when they are looked at. Scratch
// is a ring of `slots` groups (a few tens of MB, stores / loads carry evict_last / evict_first policies):
// its lines are rewritten while still in L2 - measured DRAM traffic = the algorithmic bytes. A slot is
// reused for group g + slots once pass B of group g has been LOADED (second counter, no fence needed).
//
// Twiddles: pass A covers stages m <= 1024 and uses the accurate tables with the w[q + m/4] = -i w[q]
// symmetry exactly like fft_pipe.cuh (hybrid rule of SURVEY.md 7.0); pass B reads the reference-recurrence
// table (host/ref_twiddle.c), which is what keeps large N within 1e-12 of the reference: 4 entries per thread
// and radix-16 butterfly, the other 11 derived with per-stage constants (fused_twiddles below).
//
// Measured (B200, N = 2^16 x 4096, ms per execution; HBM floor 1.32): this form 2.28; results stored from registers
// with st.global + a signal warp doing the fences 2.49 (the compute warps stall on the store queue, a gpu-scope
// fence takes as long as the outstanding stores need to drain); all 15 twiddles of a pass-B butterfly loaded
// instead of 4 + 11 derived 2.68; whole-buffer instead of quarter-wise store -> load hand-over 2.28; L2 prefetch
// (cp.async.bulk.prefetch.tensor) of the next pass-A tile 2.37; 104 instead of 96 registers for the compute
// warps (setmaxnreg) 2.32 at 2^16 in round 1 - the rebalancing is in the product since the end of round 2 (FUSED_REGS below: 2^16 is the
// one size it does not move; 2^14, 2^15, 2^17, 2^19 gain 3-8 %). What bounds it: a 64 KB tile occupies shared memory for its load latency + ~4500-7000
// cycles of compute + its store read-out, and only three tiles fit, so the SM <-> L2 interface (measured with
// tools/l2bench.cu: TMA stores 26 B/clk/SM, loads 69 B/clk/SM) idles about half of the time.
#pragma once
#include <cuda.h>
#include <string.h>

#include <type_traits>

#include "fft_pipe.cuh"

#ifndef FUSED_BDIRECT
#define FUSED_BDIRECT 1
#endif
#ifndef FUSED_R2C_MIRROR
#define FUSED_R2C_MIRROR 1
#endif
#ifndef FUSED_R2C_PACK
#define FUSED_R2C_PACK 1
#endif
#ifndef FUSED_C2R_HALFCOLS
#define FUSED_C2R_HALFCOLS 1
#endif
#ifndef FUSED_DYNAMIC
#define FUSED_DYNAMIC 0
#endif
// FUSED_REGS > 0: the manager warps give registers to the compute warps (setmaxnreg). The CTA gets a fourth (idle) manager warp so that the
// managers are a whole warpgroup and starts at 96 registers per thread (640 threads); the managers drop to FUSED_MGR_REGS and the compute
// warps rise to FUSED_REGS. The registers come out of the CTA's own pool (640 x 96 = 61440: an increase beyond what the managers released
// blocks for ever - 112 / 40 did), so 128 x FUSED_MGR_REGS + 512 x FUSED_REGS <= 61440. At 104 the compute code of every variant but a few
// real / Bluestein ones is free of spills (24-88 bytes of stack at 96). Same box, ms per 2^28 points, 96 -> 104 registers: 2^14 2.34 -> 2.16,
// 2^15 2.37 -> 2.17, 2^16 2.25 -> 2.26, 2^17 2.41 -> 2.27, 2^18 2.66 -> 2.60, 2^19 2.78 -> 2.66, 2^20 2.90 -> 2.85, 2^24 x 16 4.07 -> 3.98.
// 112 with managers at 32 (they spill): the same within noise. Round 2's first half tried 104 by other means and saw nothing.
#ifndef FUSED_REGS
#define FUSED_REGS 104
#endif
#ifndef FUSED_MGR_REGS
#define FUSED_MGR_REGS 56
#endif
static_assert(FUSED_REGS == 0 || 128 * FUSED_MGR_REGS + 512 * FUSED_REGS <= 640 * 96, "setmaxnreg: the compute warps can only take what the managers release");
// FUSED_TW_AHEAD: the four table twiddles of a pass-B sub-pass are requested right after the butterflies of the sub-pass before (one
// exchange ahead of their use); lost at 96 registers per thread (round 2, first half) and again at 104 / 112 (2^15 2.17 -> 2.23, 2^17 2.27 -> 2.30,
// 2^18 2.63 -> 2.56, 2^19 2.66 -> 2.70, the rest equal): off
#ifndef FUSED_TW_AHEAD
#define FUSED_TW_AHEAD 0
#endif
#define FUSED_STR2(x) #x
#define FUSED_STR(x) FUSED_STR2(x)

namespace fftb200 {

struct FusedArgs {
    const cd* scratch;  // slots * gt * N elements (pass-B loads; stores go through the tensor map)
    const cd* tab;      // reference-recurrence stage tables for size N
    const cd* acc;      // accurate stage tables (stages m <= 8192)
    int* flags;         // [0, G): stored pass-A tiles per group; [G, 2G): loaded pass-B tiles per group
    int* handout;       // FUSED_DYNAMIC builds: work items handed out beyond the first three of every CTA (its own cache line; zero at launch)
    long long* prof;    // development only (FUSED_PROF builds)
    long long nbatch;
    int gt;             // transforms per group
    int ngroups;        // G = ceil(nbatch / gt)
    int lag;            // pass B of group g is scheduled after pass A of group g + lag
    int slots;          // scratch ring depth in groups (>= lag + 1)
    int inverse;        // selects the INV instantiation (conjugate in, conjugate + scale out)
    cd* out;            // r2c mode: output base (the Nyquist bin of every transform is stored by a plain bulk copy)
    const cd* half_in;  // c2r from the half spectrum: input base (the Nyquist bin of every transform is read by one thread)
    int log_cb;         // column mode: log2 of the 16-column blocks per transform (row length / 16)
    int debug;          // development only: 1 = pass A alone, 2 = pass B alone, 4 = ignore the dependency counters
    double scale;       // 1/N for the inverse
    // Bluestein variants (BLUE): chirp of the caller's length n_user, spectrum FB of the wrapped chirp (N entries), the caller's arrays
    const cd* chirp;
    const cd* fb;
    const cd* user_in;  // BLUE_FWD: n_user elements per transform (the partial last row of a transform is read from here)
    cd* user_out;       // BLUE_INV: n_user elements per transform
    int n_user;
    double y_scale;     // 1/n_user for the inverse direction of the caller's transform, else 1
    cd dtw[3][16];      // pass B, sub-pass j: dtw[j][h] = T[stage][q << a_tot] (table entry at kappa = 0), h = 2^(s-1) + q
};

constexpr int FUSED_THREADS = 2 * PIPE_GROUP + 32 * (FUSED_REGS ? 4 : PIPE_STAGES);   // two compute groups + one manager warp per ring buffer
constexpr int FUSED_TW1 = 16 * 8, FUSED_TW2 = 64 * 8;
constexpr int FUSED_ROWH = 64;   // r2c: the pass-A row k = M/2 of a tile (2C <= 64 complex), staged beside each ring buffer
constexpr size_t FUSED_SMEM = (size_t)PIPE_STAGES * PIPE_TILE * sizeof(cd) + (FUSED_TW1 + FUSED_TW2 + PIPE_STAGES * FUSED_ROWH) * sizeof(cd) + 256;

// ---- tile geometry: logical index I = lo + 2^LB * f + 2^(LB+LP) * hi, f = LP-bit transform field -------------
// Sub-pass (A stages done, radix 2^R): butterfly (lo, c'', kloc, hi) gathers f = c'' + 2^(LP-A-R) rho + 2^(LP-A) kloc
// and scatters f = c'' + 2^(LP-A-R) (kloc + 2^A q). HIGH selects which field runs fastest across the lanes.
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
__device__ __forceinline__ void mbar_arrive_cnt(uint64_t* bar, int count) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_count(const int* p, int need) {
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
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int x, int y, const void* src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                 ::"l"(tm), "r"(x), "r"(y), "r"(smem_u32(src)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5}], [%6], %7;"
        ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tm, int c0, int c1, int c2, int c3, const void* src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3, %4}], [%5], %6;"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_hint(void* dst, const void* src, uint32_t bytes, uint64_t pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes), "l"(pol) : "memory");
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
    int tpa;       // pass-A tiles per transform: 2^log_tpt; half of it when a tile holds twice the columns (packed real input);
                   // half of it + 1 when only the columns c <= R/2 are transformed (c2r from the half spectrum)
    int tpb;   // pass-B tiles per transform: 2^log_tpt, or 2^(log_tpt - 1) + 1 when only the columns k <= M/2 are transformed (real input)
    __device__ __forceinline__ long long transforms_of(int g) const {
        long long nb = nbatch - (long long)g * gt;
        return nb > gt ? gt : nb;
    }
    __device__ __forceinline__ long long tiles_of(int g) const { return transforms_of(g) * tpa; }   // pass A
    __device__ __forceinline__ long long tiles_b(int g) const { return transforms_of(g) * tpb; }        // pass B
    __device__ __forceinline__ long long round_len(int rho) const {
        long long n = 0;
        if (a_on && rho < G) n += tiles_of(rho);
        if (b_on && rho >= L && rho - L < G) n += tiles_b(rho - L);
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
// the tile holds 4096 REAL values (first half of the buffer): promote to complex while gathering (fft_auto.c:394-397)
template <class G, int R>
__device__ __forceinline__ void fused_gather_real(cd* x, const cd* sm, const G& g) {
    const double* sr = reinterpret_cast<const double*>(sm);
    const int base = g.gbase();
#pragma unroll
    for (int rho = 0; rho < (1 << R); rho++) x[bitrev_c<R>(rho)] = make_double2(sr[base + rho * G::GSTRIDE], 0.0);
}
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
// r2c, first gather of a pass-B tile that may hold the rows M - k of its columns (flip = the row field complemented): Y[M - k] = conj Y[k]
template <class G, int R>
__device__ __forceinline__ void fused_gather_mirror(cd* x, const cd* sm, const G& g, const int flip, const bool conj) {
    const int base = g.gbase() ^ flip;
#pragma unroll
    for (int rho = 0; rho < (1 << R); rho++) {
        cd y = sm[base + rho * G::GSTRIDE];
        if (conj) y.y = -y.y;
        x[bitrev_c<R>(rho)] = y;
    }
}
// c2r, first gather of a pass-B tile whose rows hold the columns c <= R/2 only: Y[c][k] for c > R/2 is conj(w_M^k) conj(Y[R - c][k])
template <class G, int R, int LR_>
__device__ __forceinline__ void fused_gather_halfcols(cd* x, const cd* sm, const G& g, const cd wk) {
    const int base = g.gbase();
    const int row = base & ~((1 << LR_) - 1), c0 = base & ((1 << LR_) - 1);
#pragma unroll
    for (int rho = 0; rho < (1 << R); rho++) {
        const int c = c0 + rho * G::GSTRIDE;
        cd y;
        if (rho < (1 << (R - 1)) || (rho == (1 << (R - 1)) && c0 == 0)) {
            y = sm[row + c];
        } else {
            const cd v = sm[row + (1 << LR_) - c];
            // conj(wk) * conj(v) = conj(wk * v)
            y = make_double2(fma(wk.x, v.x, -(wk.y * v.y)), -fma(wk.x, v.y, wk.y * v.x));
        }
        x[bitrev_c<R>(rho)] = y;
    }
}
// c2r, first gather of pass A from a tile that holds the half spectrum (see the manager's load): conj(X[c + R t]) for the inverse
// transform = conj(H) of the direct half, H itself read backwards out of the mirrored half. Column 0 of a transform mirrors onto
// itself one row down, X[R t] = conj(H[R (M - t)]), and its row M/2 is the Nyquist bin, the one element the tile does not hold.
template <class G, int R>
__device__ __forceinline__ void fused_gather_half(cd* x, const cd* sm, const G& g, const bool col0, const cd* nyquist) {
    const int base = g.gbase();
#pragma unroll
    for (int rho = 0; rho < (1 << R); rho++) {
        const int i = base + rho * G::GSTRIDE;   // logical [t][c] index of the tile; rho >= 2^(R-1) <=> t >= M/2 <=> i >= PIPE_TILE / 2
        cd y;
        if (rho < (1 << (R - 1))) {
            y = sm[i];
            y.y = -y.y;
        } else if (!col0) {
            y = sm[PIPE_TILE + PIPE_TILE / 2 - 1 - i];
        } else {
            const int tm = PIPE_TILE - i;          // (M - t) * C for column 0
            y = tm == PIPE_TILE / 2 ? __ldg(nyquist) : sm[tm];
        }
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
// (a + ib)(c + id)
__device__ __forceinline__ cd cmul2(const cd a, const cd b) {
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}
// Twiddles of a radix-2^R DIT butterfly after `a_tot` stages at position kappa (tp = table + kappa - 1). The R
// entries T[a_tot + s][kappa] (h = 2^(s-1)) are loaded; the others, T[a_tot + s][kappa + (q << a_tot)], are the
// product with d[h] = T[a_tot + s][q << a_tot], the table's own entry at kappa = 0. The reference recurrence
// w_j = fl(w_(j-1) w_m) (radix2_dit.c:93,109) makes its table multiplicative up to rounding noise, so the product
// reproduces the reference's accumulated twiddle drift: whole-transform mismatch 1.3e-14 at N = 2^20.
// Measured and not kept (r02, same-box A/B at 2^28 points): issuing the loads one exchange ahead of the butterflies they feed
// (right after the previous sub-pass's butterflies, 16 registers in flight across two group barriers, a scatter and a gather).
// Pass-B tiles are 2 200 - 3 300 cycles slower than pass-A tiles (FUSED_PROF), but the early loads cost more in register
// pressure than they hide: 2^14 2.34 -> 2.29 ms, 2^16 2.24 -> 2.30, 2^17 2.40 -> 2.57, 2^18 2.60 -> 2.87, 2^20 2.92 -> 3.29.
template <int R>
__device__ __forceinline__ void fused_tw_load(cd* raw, const cd* tp, const int a_tot) {
#pragma unroll
    for (int s = 1; s <= R; s++) raw[s - 1] = __ldg(tp + ((size_t)1 << (s - 1 + a_tot)));
}
template <int R>
__device__ __forceinline__ void fused_tw_expand(cd* tw, const cd* raw, const cd* d) {
#pragma unroll
    for (int s = 1; s <= R; s++) tw[1 << (s - 1)] = raw[s - 1];
#pragma unroll
    for (int s = 2; s <= R; s++) {
        const int h0 = 1 << (s - 1);
#pragma unroll
        for (int q = 1; q < h0; q++) tw[h0 + q] = cmul2(tw[h0], d[h0 + q]);
    }
}
template <int R>
__device__ __forceinline__ void fused_twiddles(cd* tw, const cd* tp, const int a_tot, const cd* d) {
    cd raw[R];
    fused_tw_load<R>(raw, tp, a_tot);
    fused_tw_expand<R>(tw, raw, d);
}
// a compute warp has written its part of the result tile into the ring buffer: make the writes visible to the
// async proxy (the TMA store reads them) and arrive once per warp on the buffer's `staged` barrier
__device__ __forceinline__ void fused_stage_done(uint64_t* bar, int t) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if ((t & 31) == 0) mbar_arrive_cnt(bar, 32);
}
// a pass-A store has been issued earlier by this thread: wait for its completion and publish it
__device__ __forceinline__ void fused_publish(int* counter) {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence();
    atomicAdd(counter, 1);
}

// COLS (column mode, LM = LR = 8): the 2^16-point transforms run along t of a row-major [t][c] array with rows of
// 16 * 2^log_cb contiguous columns c - stages 1 .. 16 of a larger transform N = 2^16 * row length (head of a plan
// that ends with one LAST tile pass). Tiles are 16 columns x 256 points in both passes, a "virtual transform" is one
// 16-column block (2^20 points, 256 tiles per pass); the pass-B twiddles T[8 + s][k_hi + 256 q] are uniform per tile.
// R2C (forward only): the input is real (n doubles per transform, promoted in the first gather) and only the bins
// 0 .. n/2 are stored (n/2 + 1 per transform, fft_auto.h:89-97) - the reference's promote-then-c2c reading of
// fft_plan_r2c_1d without the separate promotion and extraction passes.
// C2R + HERM: the input is the half spectrum itself (n/2 + 1 bins per transform); the Hermitian extension happens in the tile load (two
// boxes: the tile's columns and the mirrored columns) and the first gather of pass A - no c2r_expand pass, no full-length work array.
// C2R (inverse only): the input is the Hermitian-extended spectrum (c2r_expand_kernel), only the real parts of the result are
// staged and stored (n doubles per transform, fft_auto.h:99-107): the separate real-part pass and its 24 bytes per point go away.
// BLUE = FUSED_BLUE_FWD / FUSED_BLUE_INV: the two transforms of Bluestein's algorithm (bluestein.c:107-148) for padded lengths
// N = 2^14 .. 2^20 with the elementwise steps riding on them, as in fft_pipe_kernel for N <= 4096. FWD: pass A reads the caller's rows
// of n_user elements through a 3-D tensor map whose rows end at the last full row of R elements (the rows behind it arrive as zeros:
// the padding; the partial row is read from the caller's array in the first gather) and multiplies by conj(chirp) in the first gather.
// INV (inverse transform, 1/N): pass A multiplies the spectrum by FB in its first gather, pass B multiplies by conj(chirp) * y_scale and
// stores the first n_user values of every row from registers to the caller's array. Two launches and HBM round trips instead of five,
// the same arithmetic as bluestein_pre / pointwise_mul / bluestein_post (fft_aux.cuh).
enum { FUSED_BLUE_NONE = 0, FUSED_BLUE_FWD = 1, FUSED_BLUE_INV = 2 };
template <int LM, int LR, bool INV, bool COLS = false, bool R2C = false, bool C2R = false, bool HERM = false, int BLUE = FUSED_BLUE_NONE>
__global__ void __launch_bounds__(FUSED_THREADS, 1)
fft_fused_kernel(const FusedArgs a, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_sc,
                 const __grid_constant__ CUtensorMap tm_out, const __grid_constant__ CUtensorMap tm_out2) {
    static_assert(LM >= 6 && LM <= 10 && LR >= 6 && LR <= 10, "pass sizes 64 .. 1024");
    static_assert(!COLS || (LM == 8 && LR == 8), "column mode is built for 256 x 256");
    static_assert(!R2C || (!INV && !COLS), "r2c is a forward transform of whole arrays");
    static_assert(!C2R || (INV && !COLS && !R2C), "c2r is an inverse transform of whole arrays");
    static_assert(!HERM || R2C || C2R, "the Hermitian variants belong to the real transforms");
    static_assert(BLUE == FUSED_BLUE_NONE || (!COLS && !R2C && !C2R && INV == (BLUE == FUSED_BLUE_INV)), "Bluestein: forward transform first, inverse second");
    constexpr int LOGN = COLS ? 20 : LM + LR;             // points per (virtual) transform
    constexpr int LOG_TPT = LOGN - 12;
    // HERM: r2c with a Hermitian-aware schedule: the pass-A outputs of a real column satisfy Y_c[M - k] = conj(Y_c[k]), so only the
    // rows k <= M/2 go to scratch and pass B transforms only those columns: TPB = tiles/2 + 1 tiles per transform instead of
    // 2^LOG_TPT (the last one holds the single column k = M/2). A column k yields the bins k + M q directly for q < R/2 and, as
    // conjugates, the bins (M - k) + M (R - 1 - q) of the column M - k for q >= R/2 (X[N - j] = conj(X[j])): every bin 0 .. N/2 is
    // produced with half of the pass-B work and scratch traffic. The mirrored bins are NOT the reference's own X[N - j]: its twiddle
    // recurrence (radix2_dit.c:93,109) is not conjugate-symmetric, so its output for real input is Hermitian only to the accuracy of
    // its late-stage twiddles. Measured mismatch against the oracle (profiles/r02_real.md) decides which sizes run this schedule;
    // the others keep the full pass B (R2C without HERM: every column transformed, rows q < R/2 stored).
    constexpr bool RH = R2C && HERM;   // r2c on the Hermitian schedule
    constexpr bool CH = C2R && HERM;   // c2r reading the half spectrum
    // RM: r2c with the full pass B (every size above R2C_HERM_MAX_LOG). Pass A still stores only the rows k <= M/2: its outputs ARE
    // Hermitian in k to rounding (stages m <= 1024, conjugate-symmetric tables), so a pass-B tile of columns k > M/2 loads the block of
    // rows M - k instead and reads it backwards and conjugated in its first gather. Pass B itself - the late stages with the reference's
    // twiddles, where the symmetry does not hold to 1e-12 - runs on every column as before: half of the pass-A scratch stores for free.
    constexpr bool RM = R2C && !HERM && FUSED_R2C_MIRROR;
    // PACK: a pass-A tile of a real transform holds 2C real columns = C complex columns z = x[2c'] + i x[2c' + 1] as they lie in memory
    // (a row of 2C doubles IS a row of C complex numbers): one complex M-point transform per pair, then the unpack step
    // Y[2c'][k] = (Z[k] + conj Z[M - k]) / 2, Y[2c' + 1][k] = (Z[k] - conj Z[M - k]) / 2i for k <= M/2 - half of the pass-A tiles and
    // arithmetic. Pass A uses the accurate tables, so this regrouping is exact to rounding like every other use of them (hybrid rule).
    constexpr bool PACK = R2C && FUSED_R2C_PACK && (RH || RM);
    // HC: c2r from the half spectrum transforms only the columns c <= R/2 in pass A. The input is Hermitian, so for the conjugated input u
    // the column R - c is u[(R - c) + R (M - 1 - t)] = conj(u[c + R t]) and its transform is Y[R - c][k] = w_M^-k conj(Y[c][k]): pass B
    // rebuilds the columns above R/2 of its rows in the first gather (one complex multiply with an accurate-table entry per element, exact
    // to rounding like everything else in pass A's domain). Half + 1 of the pass-A tiles; their scratch stores go with them.
    constexpr bool HC = CH && FUSED_C2R_HALFCOLS;
    constexpr int TPA = PACK ? (1 << (LOG_TPT - 1)) : HC ? (1 << (LOG_TPT - 1)) + 1 : (1 << LOG_TPT);
    constexpr int TPB = RH ? (1 << (LOG_TPT - 1)) + 1 : (1 << LOG_TPT);
    constexpr int LC = 12 - LM, LC2 = 12 - LR;          // log2 columns per A tile / k's per B tile
    constexpr int A3 = LM >= 9, B3 = LR >= 9;            // three sub-passes?
    constexpr int RA0 = A3 ? LM - 8 : LM - 4, RB0 = B3 ? LR - 8 : LR - 4;
    // BDIRECT: pass-B results leave from registers (st.global, rows of C2 contiguous elements) instead of being staged in
    // the ring buffer for a TMA store - nobody waits for them (no counter to publish) - and the buffer goes back to its
    // manager as soon as the last gather has read it. That shortens a pass-B tile's buffer residency by a third but moves
    // the store drain into the compute group's time. Same-box A/B at 2^28 points (ms, direct vs staged): 2^13 2.61 / 2.24,
    // 2^14 2.34 / 2.62, 2^15 2.34 / 2.59, 2^16 2.35 / 2.26, 2^17 2.55 / 2.50, 2^18 3.04 / 2.77, 2^19 3.00 / 2.79,
    // 2^20 3.39 / 2.95: it pays only for LR = 7 (rows of 32 elements, two sub-passes), which is where it is used.
#ifndef FUSED_BDIRECT_MAXLR
#define FUSED_BDIRECT_MAXLR 7
#endif
    constexpr bool BDIRECT = (FUSED_BDIRECT && !COLS && !R2C && !C2R && LR >= 7 && LR <= FUSED_BDIRECT_MAXLR) || BLUE == FUSED_BLUE_INV;   // (the caller's rows of n_user elements cannot be a tensor box)

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    cd* const tw1s = bufs + (size_t)PIPE_STAGES * PIPE_TILE;
    cd* const tw2s = tw1s + FUSED_TW1;
    uint64_t* const full = reinterpret_cast<uint64_t*>(smem_raw + FUSED_SMEM - 256);   // [3] tile loaded (the last 256 bytes: barriers + kinds)
    uint64_t* const staged = full + PIPE_STAGES;                            // [3] result staged in place by a compute group
    volatile int* const kinds = reinterpret_cast<volatile int*>(staged + PIPE_STAGES);   // [3][4]: is_b, kb, g of the loaded tile
    cd* const rowh = reinterpret_cast<cd*>(smem_raw + FUSED_SMEM - 256) - PIPE_STAGES * FUSED_ROWH;   // [3][64] (r2c, PACK)

    FusedSched sch;
    sch.nbatch = a.nbatch; sch.gt = a.gt; sch.G = a.ngroups; sch.L = a.lag; sch.log_tpt = LOG_TPT; sch.tpa = TPA; sch.tpb = TPB;
    sch.a_on = !(a.debug & 2); sch.b_on = !(a.debug & 1);
    const bool nowait = (a.debug & 7) != 0;
    const long long total = a.nbatch * ((long long)sch.a_on * TPA + (long long)sch.b_on * TPB);
    const int first = blockIdx.x, stride = gridDim.x;
#if FUSED_DYNAMIC
    // Experimental (not the product build): work items handed out on demand. A manager takes the first item of its buffer by position
    // and every later one from a global counter at the moment the buffer's current tile has been staged; a manager that draws a number
    // past the end closes its buffer twice (once for each compute group, the second time after the first has been acknowledged); a
    // group skips a buffer it has seen closed and leaves when it has seen all three closed.
    int* const handout = a.handout;
#else
    const int my_tiles = first < total ? (int)((total - first + stride - 1) / stride) : 0;
#endif

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < PIPE_STAGES; b++) { mbar_init(&full[b], 1); mbar_init(&staged[b], PIPE_GROUP); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // pass-A twiddle tables (accurate, symmetric): sub-pass 1 after RA0 stages, sub-pass 2 after RA0 + 4
    for (int i = threadIdx.x; i < (8 << RA0); i += FUSED_THREADS) tw1s[i] = __ldg(a.acc + ((sym_h(i & 7) << RA0) + (i >> 3) - 1));
    if constexpr (A3)
        for (int i = threadIdx.x; i < (8 << (RA0 + 4)); i += FUSED_THREADS)
            tw2s[i] = __ldg(a.acc + ((sym_h(i & 7) << (RA0 + 4)) + (i >> 3) - 1));
    __syncthreads();

    // =========================== buffer managers ===========================
    // Warp w (one lane) owns ring buffer w and the tiles k = w, w + 3, ...: load -> (compute group stages the
    // result in place) -> store -> next load. Tiles move in four 16 KB quarters, so the load of the next tile
    // follows the store of the previous one quarter by quarter instead of waiting for the whole buffer.
    // Dependency counters are polled early (while the buffer is busy) and only waited for at points where
    // nothing this CTA owes to others is pending in the buffer.
    if (threadIdx.x >= 2 * PIPE_GROUP) {
#if FUSED_REGS
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " FUSED_STR(FUSED_MGR_REGS) ";");
#endif
        const int w = (threadIdx.x - 2 * PIPE_GROUP) >> 5;
        if ((threadIdx.x & 31) != 0 || w >= PIPE_STAGES) return;
        const uint64_t pol_first = policy_evict_first(), pol_last = policy_evict_last();
        cd* const buf = bufs + (size_t)w * PIPE_TILE;
        constexpr int QT = PIPE_TILE / 4;                          // elements per quarter
        constexpr uint32_t QBYTES = QT * (uint32_t)sizeof(cd);
        FusedCursor cur;
#if FUSED_DYNAMIC
        auto close_twice = [&](int n) {   // uses n and n + 1 of this buffer: one for each group
            kinds[4 * w] = -1;
            mbar_arrive(&full[w]);
            mbar_wait_bounded(&staged[w], n & 1);   // (the group's acknowledgement)
            mbar_arrive(&full[w]);
        };
        if (first + (long long)w * stride >= total) { close_twice(0); return; }
#else
        if (w >= my_tiles) return;
#endif
        FusedItem it = cur.locate(sch, first + (long long)w * stride);
        // issue the four quarter loads of tile `x` into the buffer; WAITQ: quarter q only after the store of quarter q has been read
        auto load = [&](const FusedItem& x, auto waitq) {
            const bool half_b = RH && x.is_b;   // pass-B tiles of a real transform are numbered 0 .. TPB - 1 per transform
            const int per = x.is_b ? TPB : TPA;     // tiles per transform of this pass (compile-time constants: the divisions are cheap)
            const int blk = (int)(x.tau % per);
            const long long trg = x.tau / per;
            (void)half_b;
            kinds[4 * w] = x.is_b; kinds[4 * w + 1] = blk; kinds[4 * w + 2] = x.g; kinds[4 * w + 3] = (int)trg;
            mbar_expect_tx(&full[w], (R2C && !PACK && !x.is_b) ? PIPE_TILE * (uint32_t)sizeof(double) : PIPE_TILE * (uint32_t)sizeof(cd));
            if (x.is_b) {
                asm volatile("fence.proxy.async;" ::: "memory");
                // (RM: a tile of columns k0 .. k0 + C2 - 1 above M/2 loads the rows M - k0 - C2 + 1 .. M - k0 instead: the same bytes count, read backwards)
                const size_t row0 = (RM && blk >= (1 << (LOG_TPT - 1))) ? ((size_t)1 << LM) - ((size_t)blk << LC2) - ((size_t)1 << LC2) + 1 : (size_t)blk << LC2;
                const cd* src = a.scratch + ((((size_t)(x.g % a.slots) * a.gt) + (size_t)trg) << LOGN) + (row0 << LR);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    waitq(q);
                    bulk_load_hint(buf + q * QT, src + q * QT, QBYTES, &full[w], pol_last);
                }
            } else {
                const long long tr = (long long)x.g * a.gt + trg;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    waitq(q);
                    if constexpr (BLUE == FUSED_BLUE_FWD) {
                        // the caller's rows: [transform][t < n_user / R][R]; rows from n_user / R on are out of range and arrive as zeros
                        tma_load_3d(buf + q * QT, &tm_in, 2 * (blk << LC), q * (QT >> LC), (int)tr, &full[w], pol_first);
                    } else if constexpr (CH) {
                        // the input is the HALF spectrum (N/2 + 1 bins per transform, rows of R): quarters 0, 1 = rows t < M/2 of the tile's columns;
                        // quarters 2, 3 = the same rows of the mirrored columns R - c0 - C + 1 .. R - c0 - the rows t >= M/2 of the Hermitian
                        // extension read backwards, X[c + R t] = conj(H[(R - c) + R (M - 1 - t)]) (column R of the first tile is out of range: zeros)
                        if (q < 2) tma_load_3d(buf + q * QT, &tm_in, 2 * (blk << LC), q * (QT >> LC), (int)tr, &full[w], pol_first);
                        else tma_load_3d(buf + q * QT, &tm_in, 2 * ((1 << LR) - (blk << LC) - (1 << LC) + 1), (q - 2) * (QT >> LC), (int)tr, &full[w], pol_first);
                    } else if constexpr (R2C && !PACK)    // real rows: a quarter is 1024 doubles
                        tma_load_2d(reinterpret_cast<double*>(buf) + q * QT, &tm_in, blk << LC, (int)((tr << LM) + q * (QT >> LC)), &full[w], pol_first);
                    else if constexpr (COLS)   // [b][t_hi][t_lo][c]: 16 columns of block cb, t_lo = blk, a quarter of the t_hi range
                        tma_load_4d(buf + q * QT, &tm_in, 32 * (int)(tr & ((1 << a.log_cb) - 1)), blk, q * 64, (int)(tr >> a.log_cb), &full[w], pol_first);
                    else
                        tma_load_2d(buf + q * QT, &tm_in, 2 * (blk << LC), (int)((tr << LM) + q * (QT >> LC)), &full[w], pol_first);
                }
            }
        };
        auto nowaitq = [](int) {};
        auto storeq = [](int q) {   // all but the 3 - q most recent store groups have been read
            if (q == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
            else if (q == 1) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else if (q == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        };
        if (it.is_b && !nowait) wait_count(a.flags + it.g, (int)sch.tiles_of(it.g));   // every pass-A tile of the group has been stored
        load(it, nowaitq);
#ifdef FUSED_PROF
        long long mp[6] = {0, 0, 0, 0, 0, 0};   // load latency, load->staged, store read-out (incl. chase), publish, tiles
#endif
        for (int k = w, n = 0;; k += PIPE_STAGES, n++) {
            const FusedItem cur_it = it;
#ifdef FUSED_PROF
            const long long m0 = clock64();
            mbar_wait_bounded(&full[w], n & 1);
            const long long m1 = clock64();
            mp[0] += m1 - m0;
#endif
            // ---- look ahead: next tile of this buffer, early poll of its dependency; and of this tile's WAR counter ----
#if !FUSED_DYNAMIC
            const bool have = k + PIPE_STAGES < my_tiles;
            int seen = 0, need = 0;
            if (have) {
                it = cur.locate(sch, first + (long long)(k + PIPE_STAGES) * stride);
                if (it.is_b && !nowait) { need = (int)sch.tiles_of(it.g); seen = ld_acquire_gpu(a.flags + it.g); }
            }
#endif
            int war_need = 0, war_seen = 0;
            const int* war_p = nullptr;
            if (!cur_it.is_b && cur_it.g >= a.slots && !nowait) {
                // the scratch slot of this group was last read by pass B of group g - slots
                war_p = a.flags + a.ngroups + (cur_it.g - a.slots);
                war_need = (int)sch.tiles_b(cur_it.g - a.slots);
                war_seen = ld_acquire_gpu(war_p);
            }
            // ---- store, one bulk group per quarter ----
            mbar_wait_bounded(&staged[w], n & 1);
#if FUSED_DYNAMIC
            const long long next_item = 3LL * stride + (long long)atomicAdd(handout, 1);   // (first needed after the stores have been issued)
#endif
#ifdef FUSED_PROF
            const long long m2 = clock64();
            mp[1] += m2 - m1;
#endif
            {
                const bool half_b = RH && cur_it.is_b;
                const int per = cur_it.is_b ? TPB : TPA;
                const int blk = (int)(cur_it.tau % per);
                const long long trg = cur_it.tau / per;
                (void)half_b;
                if (!cur_it.is_b) {
                    if (war_seen < war_need) wait_count(war_p, war_need);
                    const long long trl = (long long)(cur_it.g % a.slots) * a.gt + trg;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if constexpr (COLS) tma_store_4d(&tm_sc, 0, blk, q * 64, (int)trl, buf + q * QT, pol_last);   // [slot][k_hi][t_lo][c16]
                        else if constexpr (PACK) {
                            // the unpacked rows k < M/2 of the tile's 2C columns ([k][2C], four quarters of M/8 rows) and, with the last
                            // quarter, the row k = M/2 from its side buffer (2C contiguous elements of scratch[c + R k])
                            tma_store_2d(&tm_sc, 2 * (blk << (LC + 1)), (int)((trl << LM) + q * (QT >> (LC + 1))), buf + q * QT, pol_last);
                            if (q == 3)
                                bulk_store_hint(const_cast<cd*>(a.scratch) + ((size_t)trl << LOGN) + ((size_t)1 << (LOGN - 1)) + ((size_t)blk << (LC + 1)),
                                                rowh + w * FUSED_ROWH, (uint32_t)sizeof(cd) << (LC + 1), pol_last);
                        } else if constexpr (RH || RM) {
                            // rows k < M/2 (two quarters) and the row k = M/2 (C contiguous elements of scratch[c + R k]); the rest is never read
                            if (q < 2) tma_store_2d(&tm_sc, 2 * (blk << LC), (int)((trl << LM) + q * (QT >> LC)), buf + q * QT, pol_last);
                            else if (q == 2)
                                bulk_store_hint(const_cast<cd*>(a.scratch) + ((size_t)trl << LOGN) + ((size_t)1 << (LOGN - 1)) + ((size_t)blk << LC), buf + 2 * QT,
                                                (uint32_t)sizeof(cd) << LC, pol_last);
                        } else tma_store_2d(&tm_sc, 2 * (blk << LC), (int)((trl << LM) + q * (QT >> LC)), buf + q * QT, pol_last);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else if constexpr (!BDIRECT) {
                    const long long tr = (long long)cur_it.g * a.gt + trg;
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        if constexpr (R2C && !HERM) {   // rows q < R/2 (bins below n/2) are the first two quarters; then the Nyquist bin X[M R/2]
                            if (q < 2) tma_store_3d(&tm_out, 2 * (blk << LC2), q * (QT >> LC2), (int)tr, buf + q * QT, pol_first);
                            else if (q == 2 && blk == 0) bulk_store(a.out + (size_t)tr * ((1 << (LOGN - 1)) + 1) + (1 << (LOGN - 1)), buf + 2 * QT, sizeof(cd));
                        } else if constexpr (RH) {
                            // quarters 0, 1: the bins k + M q, q < R/2, of the tile's columns; quarters 2, 3: the mirrored rows, conj(X[k + M q]) for
                            // q >= R/2 staged as [R - 1 - q][C2 - 1 - (k - k0)] = the bins of the columns M - k0 - C2 + 1 .. M - k0 (column M of the
                            // first tile - the mirror of k = 0, which that tile produces directly - falls off the tensor and is dropped).
                            // The last tile (k0 = M/2) holds one valid column: its direct half goes through the map that ends at column M/2.
                            if (blk == TPB - 1) {
                                if (q < 2) tma_store_3d(&tm_out2, 2 * (blk << LC2), q * (QT >> LC2), (int)tr, buf + q * QT, pol_first);
                            } else {
                                if (q < 2) tma_store_3d(&tm_out, 2 * (blk << LC2), q * (QT >> LC2), (int)tr, buf + q * QT, pol_first);
                                else tma_store_3d(&tm_out, 2 * ((1 << LM) - (blk << LC2) - (1 << LC2) + 1), (q - 2) * (QT >> LC2), (int)tr, buf + q * QT, pol_first);
                                // the Nyquist bin X[M R/2] (k = 0, q = R/2) sits at the end of the mirrored block of the first tile
                                if (q == 3 && blk == 0) bulk_store(a.out + (size_t)tr * ((1 << (LOGN - 1)) + 1) + (1 << (LOGN - 1)), buf + PIPE_TILE - 1, sizeof(cd));
                            }
                        } else if constexpr (C2R) {   // real rows: a quarter is 1024 doubles
                            tma_store_2d(&tm_out, blk << LC2, (int)((tr << LR) + q * (QT >> LC2)), reinterpret_cast<double*>(buf) + q * QT, pol_first);
                        } else if constexpr (COLS)   // [b][q][k_hi][c]
                            tma_store_4d(&tm_out, 32 * (int)(tr & ((1 << a.log_cb) - 1)), blk, q * 64, (int)(tr >> a.log_cb), buf + q * QT, pol_first);
                        else
                            tma_store_2d(&tm_out, 2 * (blk << LC2), (int)((tr << LR) + q * (QT >> LC2)), buf + q * QT, pol_first);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            }
            // ---- next load, chasing the store quarter by quarter; then publish the store ----
#if FUSED_DYNAMIC
            const bool have = next_item < total;
            int seen = 0, need = 0;
            if (have) {
                it = cur.locate(sch, next_item);
                if (it.is_b && !nowait) { need = (int)sch.tiles_of(it.g); seen = ld_acquire_gpu(a.flags + it.g); }
            }
#endif
            if (have) {
                if (seen >= need) {
                    if (BDIRECT && cur_it.is_b) load(it, nowaitq);   // nothing was stored from this buffer: it is free now
                    else load(it, storeq);
#ifdef FUSED_PROF
                    const long long m3 = clock64();
                    mp[2] += m3 - m2;
#endif
                    if (!cur_it.is_b) fused_publish(a.flags + cur_it.g);
#ifdef FUSED_PROF
                    mp[3] += clock64() - m3; mp[4]++;
#endif
                } else {
                    // the counter may depend on the tile just stored: publish it before waiting
                    if (!cur_it.is_b) fused_publish(a.flags + cur_it.g);
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    wait_count(a.flags + it.g, need);
                    load(it, nowaitq);
                }
            } else {
                if (!cur_it.is_b) fused_publish(a.flags + cur_it.g);
#if FUSED_DYNAMIC
                close_twice(n + 1);
#endif
                break;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef FUSED_PROF
        if (a.prof) { long long* q = a.prof + 8 * 1024 + (blockIdx.x * 3 + w) * 8; for (int i = 0; i < 5; i++) q[i] = mp[i]; }
#endif
        return;
    }

    // =========================== compute groups ===========================
#if FUSED_REGS
    asm volatile("setmaxnreg.inc.sync.aligned.u32 " FUSED_STR(FUSED_REGS) ";");
#endif
    const int g2 = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const double sc = a.scale;
    int b = g2, n = 0;   // ring buffer and use count of tile k = g2, g2 + 2, ... (b = k % 3, n = k / 3)
#ifdef FUSED_PROF
    long long pr_e = 0, pr_f = 0, pr_busy[2] = {0, 0}, pr_cnt[2] = {0, 0};
    const long long pr_t0 = clock64();
#endif
#if FUSED_DYNAMIC
    int closed = 0;   // buffers seen closed (bit b)
    for (;;) {
        if ((closed >> b) & 1) {
            b += 2;
            if (b >= PIPE_STAGES) { b -= PIPE_STAGES; n++; }
            continue;
        }
#else
    for (int k = g2; k < my_tiles; k += 2) {
#endif
        cd* const sm = bufs + (size_t)b * PIPE_TILE;
#ifdef FUSED_PROF
        const long long c0 = clock64();
#endif
        // A parity wait is only valid one phase ahead: first make sure the previous tile of this buffer (handled by
        // the other group) has been staged - its load may complete later than the load of this group's last tile.
        if (n >= 1) mbar_wait_bounded(&staged[b], (n - 1) & 1);
#ifdef FUSED_PROF
        const long long c1 = clock64();
#endif
        mbar_wait_bounded(&full[b], n & 1);
#ifdef FUSED_PROF
        const long long c2 = clock64();
        pr_e += c1 - c0; pr_f += c2 - c1;
#endif
        const int is_b = kinds[4 * b], kb = kinds[4 * b + 1];
#if FUSED_DYNAMIC
        if (is_b < 0) {   // closed: acknowledge (the manager closes the buffer once more for the other group), leave after the third
            closed |= 1 << b;
            __syncwarp();
            if ((t & 31) == 0) mbar_arrive_cnt(&staged[b], 32);
            if (closed == (1 << PIPE_STAGES) - 1) break;
            b += 2;
            if (b >= PIPE_STAGES) { b -= PIPE_STAGES; n++; }
            continue;
        }
#endif
        const int kg = kinds[4 * b + 2], ktr = kinds[4 * b + 3];   // group and transform within it (read before the buffer is handed back)
        cd x[16];
        if (!is_b) {
            // ------------------------------ pass A: stages 1 .. LM over C = 2^LC columns ------------------------------
            // sub-pass 0: radix 2^RA0, exact constants, in place per thread
            {
                typedef Geo<LC, LM, 0, 0, RA0, false> G0;
                constexpr int NB = 16 >> RA0, R0 = 1 << RA0;
                if constexpr (R2C && !PACK) {
                    // the complex results overwrite other threads' real inputs: gather everything first
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) fused_gather_real<G0, RA0>(&x[bb * R0], sm, G0(t + PIPE_GROUP * bb));
                    group_sync(g2);
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                        fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, G0(t + PIPE_GROUP * bb));
                    }
                } else if constexpr (CH) {
                    // the complex results overwrite other threads' half-spectrum inputs (the mirrored half is read backwards): gather everything first
                    const cd* nyq = a.half_in + ((size_t)kg * a.gt + ktr) * (((size_t)1 << (LOGN - 1)) + 1) + ((size_t)1 << (LOGN - 1));
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        fused_gather_half<G0, RA0>(&x[bb * R0], sm, g, kb == 0 && g.lo == 0, nyq);
                    }
                    group_sync(g2);
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                        fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, G0(t + PIPE_GROUP * bb));
                    }
                } else if constexpr (BLUE == FUSED_BLUE_FWD) {
                    // a = x * conj(chirp), zero-padded from n_user to N (bluestein.c:107-109): element (t, c) of the tile is x[c + R t]
                    const cd* const urow = a.user_in + ((size_t)kg * a.gt + ktr) * (size_t)a.n_user;
                    const int t_full = a.n_user >> LR;   // rows the tensor map holds; row t_full is the partial one
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        const int base = g.gbase();
#pragma unroll
                        for (int rho = 0; rho < R0; rho++) {
                            const int I = base + rho * G0::GSTRIDE, tt = I >> LC;
                            const int i = (kb << LC) + (I & ((1 << LC) - 1)) + (tt << LR);
                            cd y = make_double2(0.0, 0.0);
                            if (i < a.n_user) {
                                const cd xv = tt == t_full ? __ldg(urow + i) : sm[I], w = __ldg(a.chirp + i);
                                y = make_double2(fma(xv.x, w.x, xv.y * w.y), fma(xv.y, w.x, -(xv.x * w.y)));
                            }
                            x[bb * R0 + bitrev_c<RA0>(rho)] = y;
                        }
                        SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                        fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, g);
                    }
                } else if constexpr (BLUE == FUSED_BLUE_INV) {
                    // A * FB (bluestein.c:124-131) on the way in, then the conjugate of the inverse transform: element (t, c) of the tile is A[c + R t].
                    // (The product rides here and not on the forward transform's staging: there it costs 270 bytes of spills at 2^10 x 2^9.)
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        const int base = g.gbase();
#pragma unroll
                        for (int rho = 0; rho < R0; rho++) {
                            const int I = base + rho * G0::GSTRIDE;
                            const int i = (kb << LC) + (I & ((1 << LC) - 1)) + ((I >> LC) << LR);
                            const cd y = cmul2(sm[I], __ldg(a.fb + i));
                            x[bb * R0 + bitrev_c<RA0>(rho)] = make_double2(y.x, -y.y);
                        }
                        SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                        fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, g);
                    }
                } else {
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        fused_gather<G0, SwzId, RA0, INV>(&x[bb * R0], sm, g);
                        SubStageExact<RA0, 1, 0, 0>::run(&x[bb * R0]);
                        fused_scatter<G0, SwzId, RA0>(&x[bb * R0], sm, g);
                    }
                }
            }
            group_sync(g2);
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
            group_sync(g2);   // every gather is done: stage Y[c][k] in place as [k][c], the box the tensor store expects
            if constexpr (PACK) {
                // x[q] = Z[k][lo], k = kq + (q << S): the transform of the complex column lo = the real columns 2 lo, 2 lo + 1. Unpack for k < M/2
                // (q < 8) with the partner Z[M - k] (q >= 8 of another thread): park the upper half as [k - M/2][C], fetch the partners into
                // the registers it leaves, then write [k][2C] - rows twice as wide, hence the second barrier before anything is overwritten.
                constexpr int S = LM - 4;
                const cd zmid = x[8];                      // Z[M/2] where kq == 0
#pragma unroll
                for (int q = 8; q < 16; q++) sm[lo + ((kq + ((q - 8) << S)) << LC)] = x[q];
                group_sync(g2);
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const int k = kq + (q << S);
                    x[8 + q] = k == 0 ? x[0] : sm[lo + (((1 << (LM - 1)) - k) << LC)];   // Z[M - k] sits in row M/2 - k of the parked half
                }
                group_sync(g2);
                cd* p = sm + 2 * lo + (kq << (LC + 1));
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    const cd z = x[q], w = x[8 + q];
                    p[q << (S + LC + 1)] = make_double2(0.5 * (z.x + w.x), 0.5 * (z.y - w.y));        // (Z + conj W) / 2
                    p[(q << (S + LC + 1)) + 1] = make_double2(0.5 * (z.y + w.y), 0.5 * (w.x - z.x));  // (Z - conj W) / 2i
                }
                if (kq == 0) {   // row k = M/2: Z[M/2] is its own partner
                    cd* r = rowh + b * FUSED_ROWH + 2 * lo;
                    r[0] = make_double2(zmid.x, 0.0);
                    r[1] = make_double2(zmid.y, 0.0);
                }
            } else {
                cd* p = sm + lo + (kq << LC);
#pragma unroll
                for (int q = 0; q < 16; q++) p[q << (LM - 4 + LC)] = x[q];
            }
            fused_stage_done(&staged[b], t);
        } else {
            // ------------------------------ pass B: stages LM + 1 .. LM + LR for C2 = 2^LC2 values of k ------------------------------
            // the scratch block is in shared memory now: its ring slot may be overwritten (a completed read needs no fence)
            if (t == 0 && !nowait) atomicAdd(a.flags + a.ngroups + kg, 1);
            if constexpr (COLS) {
                // the tile is [t_lo][c16], the geometry of a pass-A tile; stages 9 .. 16 with T[8 + s][kb + 256 q]
                typedef Geo<4, 8, 0, 0, 4, false> G0;
                typedef Geo<4, 8, 0, 4, 4, false> G1;
                const G0 g0(t);
                const G1 g1(t);
                fused_gather<G0, SwzId, 4, false>(x, sm, g0);
                {
                    cd tw[16];
                    fused_twiddles<4>(tw, a.tab + (kb - 1), 8, a.dtw[0]);
                    SubStageGen<4, 1, 0, 0>::run(x, tw);
                }
                fused_scatter<G0, SwzId, 4>(x, sm, g0);   // in place per thread
                group_sync(g2);
                fused_gather<G1, SwzId, 4, false>(x, sm, g1);
                {
                    cd tw[16];
                    fused_twiddles<4>(tw, a.tab + (kb + (g1.kloc << 8) - 1), 12, a.dtw[2]);
                    SubStageGen<4, 1, 0, 0>::run(x, tw);
                }
                group_sync(g2);   // every gather is done: stage [q][c16]
                cd* p = sm + g1.lo + (g1.kloc << 4);
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    cd r = x[q];
                    if (INV) r.y = -r.y;   // conjugate out; the scale belongs to the pass that ends the plan
                    p[q << 8] = r;
                }
                fused_stage_done(&staged[b], t);
                b += 2;
                if (b >= PIPE_STAGES) { b -= PIPE_STAGES; n++; }
                continue;
            }
            typedef typename SwzBlast<LR>::type SWL;
            typedef typename std::conditional<B3, SwzId, SWL>::type SW1;   // layout after sub-pass 0
            constexpr int AL = LR - 4;   // stages of this pass done before the last sub-pass
            typedef Geo<0, LR, LC2, RB0, 4, false> G1;    // middle sub-pass (three sub-passes only)
            typedef Geo<0, LR, LC2, AL, 4, true> GL;      // last sub-pass
            const G1 g1(t);
            const GL gl(t);
#if FUSED_TW_AHEAD
            cd raw_next[4];   // table twiddles of the next sub-pass, in flight across the exchange
#endif
            {
                typedef Geo<0, LR, LC2, 0, RB0, false> G0;
                constexpr int NB = 16 >> RB0, R0 = 1 << RB0;
                if constexpr (HC) {
                    // columns c <= R/2 as stored; c > R/2 rebuilt as w_M^-k conj(Y[R - c][k]) from the same row (k = the tile's row)
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        const int k = (kb << LC2) + g.hi;
                        cd wk = __ldg(a.acc + ((1 << (LM - 1)) - 1 + (k & ((1 << (LM - 1)) - 1))));   // w_M^(k mod M/2); w_M^(k + M/2) = -w_M^k
                        if (k >> (LM - 1)) { wk.x = -wk.x; wk.y = -wk.y; }
                        fused_gather_halfcols<G0, RB0, LR>(&x[bb * R0], sm, g, wk);
                    }
                } else if constexpr (RM) {
                    // 0: the tile's own rows; 1: the rows M - k, backwards and conjugated; 2: the tile that starts at k = M/2 (its first row is its own)
                    const int mode = kb < (1 << (LOG_TPT - 1)) ? 0 : kb == (1 << (LOG_TPT - 1)) ? 2 : 1;
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        fused_gather_mirror<G0, RB0>(&x[bb * R0], sm, g, mode ? ((1 << LC2) - 1) << LR : 0, mode == 1 || (mode == 2 && g.hi != 0));
                    }
                } else {
#pragma unroll
                    for (int bb = 0; bb < NB; bb++) {
                        const G0 g(t + PIPE_GROUP * bb);
                        fused_gather<G0, SwzId, RB0, false>(&x[bb * R0], sm, g);
                    }
                }
                cd tw[R0];
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_twiddles<RB0>(tw, a.tab + ((kb << LC2) + g.hi - 1), LM, a.dtw[0]);
                    SubStageGen<RB0, 1, 0, 0>::run(&x[bb * R0], tw);
                }
#if FUSED_TW_AHEAD
                if constexpr (B3) fused_tw_load<4>(raw_next, a.tab + ((kb << LC2) + g1.hi + (g1.kloc << LM) - 1), LM + RB0);
                else fused_tw_load<4>(raw_next, a.tab + ((kb << LC2) + gl.hi + (gl.kloc << LM) - 1), LM + AL);
#endif
                group_sync(g2);   // every gather of sub-pass 0 is done (the layout changes)
#pragma unroll
                for (int bb = 0; bb < NB; bb++) {
                    const G0 g(t + PIPE_GROUP * bb);
                    fused_scatter<G0, SW1, RB0>(&x[bb * R0], sm, g);
                }
            }
            group_sync(g2);
            if constexpr (B3) {
                fused_gather<G1, SW1, 4, false>(x, sm, g1);
                {
                    cd tw[16];
#if FUSED_TW_AHEAD
                    fused_tw_expand<4>(tw, raw_next, a.dtw[1]);
#else
                    fused_twiddles<4>(tw, a.tab + ((kb << LC2) + g1.hi + (g1.kloc << LM) - 1), LM + RB0, a.dtw[1]);
#endif
                    SubStageGen<4, 1, 0, 0>::run(x, tw);
                }
#if FUSED_TW_AHEAD
                fused_tw_load<4>(raw_next, a.tab + ((kb << LC2) + gl.hi + (gl.kloc << LM) - 1), LM + AL);
#endif
                group_sync(g2);
                fused_scatter<G1, SWL, 4>(x, sm, g1);
                group_sync(g2);
            }
            fused_gather<GL, SWL, 4, false>(x, sm, gl);
            if constexpr (BDIRECT) fused_stage_done(&staged[b], t);   // this warp has read its part: the buffer goes back when all have
            {
                cd tw[16];
#if FUSED_TW_AHEAD
                fused_tw_expand<4>(tw, raw_next, a.dtw[2]);
#else
                fused_twiddles<4>(tw, a.tab + ((kb << LC2) + gl.hi + (gl.kloc << LM) - 1), LM + AL, a.dtw[2]);
#endif
                SubStageGen<4, 1, 0, 0>::run(x, tw);
            }
            if constexpr (BDIRECT) {
                // X[k + M q], k = (kb << LC2) + hi, q = kloc + (q' << AL): lanes run over hi first, so a warp instruction writes
                // 32 / C2 rows of C2 contiguous elements (>= 64 bytes each)
                const long long tr = (long long)kg * a.gt + ktr;
                if constexpr (BLUE == FUSED_BLUE_INV) {
                    // y = a * conj(chirp) * y_scale for the first n_user values of the row (bluestein.c:139-148), factors fetched four at a time
                    const int j0 = (kb << LC2) + gl.hi + (gl.kloc << LM);
                    cd* const p = a.user_out + (size_t)tr * (size_t)a.n_user + j0;
                    const double s2 = a.y_scale;
#pragma unroll
                    for (int q0 = 0; q0 < 16; q0 += 4) {
                        cd w[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const int j = j0 + ((q0 + i) << (AL + LM));
                            w[i] = j < a.n_user ? __ldg(a.chirp + j) : make_double2(0.0, 0.0);
                        }
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            cd r = x[q0 + i];
                            r.x *= sc; r.y *= -sc;
                            if (j0 + ((q0 + i) << (AL + LM)) < a.n_user)
                                p[(size_t)(q0 + i) << (AL + LM)] = make_double2(fma(r.x, w[i].x, r.y * w[i].y) * s2, fma(r.y, w[i].x, -(r.x * w[i].y)) * s2);
                        }
                    }
                } else {
                    cd* p = a.out + ((size_t)tr << LOGN) + ((size_t)kb << LC2) + gl.hi + ((size_t)gl.kloc << LM);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        cd r = x[q];
                        if (INV) { r.x *= sc; r.y *= -sc; }
                        p[(size_t)q << (AL + LM)] = r;
                    }
                }
            } else {
                group_sync(g2);   // every gather is done: stage X[k + M q] in place as [q][k], the box the tensor store expects
                if constexpr (C2R) {
                    double* p = reinterpret_cast<double*>(sm) + gl.hi + (gl.kloc << LC2);
#pragma unroll
                    for (int q = 0; q < 16; q++) p[q << (AL + LC2)] = x[q].x * sc;
                } else if constexpr (RH) {
                    // q' < 8 (q < R/2): [q][k] as ever; q' >= 8: the conjugate at [R - 1 - q][C2 - 1 - k] of the second half of the buffer
                    cd* p = sm + gl.hi + (gl.kloc << LC2);
                    cd* pm = sm + PIPE_TILE / 2 + (((1 << LR) - 1 - gl.kloc) << LC2) + ((1 << LC2) - 1 - gl.hi);
#pragma unroll
                    for (int q = 0; q < 8; q++) p[q << (AL + LC2)] = x[q];
#pragma unroll
                    for (int q = 8; q < 16; q++) pm[-(q << (AL + LC2))] = make_double2(x[q].x, -x[q].y);
                } else {
                    cd* p = sm + gl.hi + (gl.kloc << LC2);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        cd r = x[q];
                        if (INV) { r.x *= sc; r.y *= -sc; }
                        p[q << (AL + LC2)] = r;
                    }
                }
                fused_stage_done(&staged[b], t);
            }
        }
#ifdef FUSED_PROF
        pr_busy[is_b] += clock64() - c2; pr_cnt[is_b]++;
#endif
        b += 2;
        if (b >= PIPE_STAGES) { b -= PIPE_STAGES; n++; }
    }
#ifdef FUSED_PROF
    if (t == 0 && a.prof) {
        long long* q = a.prof + (blockIdx.x * 2 + g2) * 4;
        q[0] = pr_e; q[1] = pr_f; q[2] = clock64() - pr_t0; q[3] = pr_cnt[0] + pr_cnt[1];
        long long* q2 = a.prof + 12 * 1024 + (blockIdx.x * 2 + g2) * 4;
        q2[0] = pr_busy[0]; q2[1] = pr_cnt[0]; q2[2] = pr_busy[1]; q2[3] = pr_cnt[1];
    }
#endif
}

// defined in fft_kernels_fused{0..3}.cu: nullptr when the (lm, lr) pair is not compiled in that unit
const void* fused_func_0(int lm, int lr, int inverse);
const void* fused_func_1(int lm, int lr, int inverse);
const void* fused_func_2(int lm, int lr, int inverse);
const void* fused_func_3(int lm, int lr, int inverse);
inline const void* fused_func(int lm, int lr, int inverse) {
    const void* f = fused_func_0(lm, lr, inverse);
    if (!f) f = fused_func_1(lm, lr, inverse);
    if (!f) f = fused_func_2(lm, lr, inverse);
    if (!f) f = fused_func_3(lm, lr, inverse);
    return f;
}
const void* fused_cols_func(int inverse);   // column mode (fft_kernels_fused1.cu)
const void* fused_blue_func_0(int lm, int lr, int kind);
const void* fused_blue_func_1(int lm, int lr, int kind);
const void* fused_blue_func_2(int lm, int lr, int kind);
const void* fused_blue_func_3(int lm, int lr, int kind);
inline const void* fused_blue_func(int lm, int lr, int kind) {   // kind = FUSED_BLUE_FWD / FUSED_BLUE_INV
    const void* f = fused_blue_func_0(lm, lr, kind);
    if (!f) f = fused_blue_func_1(lm, lr, kind);
    if (!f) f = fused_blue_func_2(lm, lr, kind);
    if (!f) f = fused_blue_func_3(lm, lr, kind);
    return f;
}
const void* fused_r2c_func_0(int lm, int lr, int herm);
const void* fused_r2c_func_1(int lm, int lr, int herm);
const void* fused_r2c_func_2(int lm, int lr, int herm);
const void* fused_r2c_func_3(int lm, int lr, int herm);
const void* fused_c2r_func_0(int lm, int lr, int herm);
const void* fused_c2r_func_1(int lm, int lr, int herm);
const void* fused_c2r_func_2(int lm, int lr, int herm);
const void* fused_c2r_func_3(int lm, int lr, int herm);
inline const void* fused_c2r_func(int lm, int lr, int herm) {
    const void* f = fused_c2r_func_0(lm, lr, herm);
    if (!f) f = fused_c2r_func_1(lm, lr, herm);
    if (!f) f = fused_c2r_func_2(lm, lr, herm);
    if (!f) f = fused_c2r_func_3(lm, lr, herm);
    return f;
}
inline const void* fused_r2c_func(int lm, int lr, int herm) {
    const void* f = fused_r2c_func_0(lm, lr, herm);
    if (!f) f = fused_r2c_func_1(lm, lr, herm);
    if (!f) f = fused_r2c_func_2(lm, lr, herm);
    if (!f) f = fused_r2c_func_3(lm, lr, herm);
    return f;
}

// tm[0..3]: tensor maps of the input (pass-A loads), the scratch ring (pass-A stores), the output (pass-B stores) and - r2c
// only, a copy of tm[2] elsewhere - the output cut off after column M/2 (the single valid column of the last pass-B tile).
// The CTAs synchronise through global counters, so all of them must be resident: a cooperative launch makes the
// driver guarantee that (or fail) even when other kernels compete for the SMs.
inline cudaError_t launch_fused(const void* func, const FusedArgs& a, const CUtensorMap* tm, int grid, cudaStream_t s) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(FUSED_THREADS); cfg.dynamicSmemBytes = FUSED_SMEM; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    void* args[5] = {(void*)&a, (void*)&tm[0], (void*)&tm[1], (void*)&tm[2], (void*)&tm[3]};
    return cudaLaunchKernelExC(&cfg, func, args);
}

}  // namespace fftb200
