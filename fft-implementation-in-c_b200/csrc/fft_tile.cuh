// fft_tile.cuh - the sm_100a Stockham tile kernel behind every power-of-two transform.
//
// Replaces the reference's butterfly hot loop (algorithms/core/radix2_dit.c:70-119, run identically by
// radix4.c:108-125 and split_radix.c:39-54) and its never-built cuFFT call (gpu/fft_cuda.cu:166-185).
//
// One CTA owns a "tile": a flat array of NP = P*C complex doubles (P points of C adjacent columns, or
// one whole transform when C = 1) that lives in registers, 16 (or fewer) points per thread, and is
// exchanged through shared memory between sub-passes. The dataflow is the reference's radix-2
// decimation-in-time stage sequence regrouped into radix-2^r register butterflies and kept in natural
// order (Stockham autosort, so no bit-reversal pass): state after `a` stages is A[c][k] stored at
// idx = c + (NP / 2^a) * k. Because the regrouping is algebraically the same product of stage
// matrices, feeding it the reference's own per-stage twiddle values (host-built table, flat offset
// Mt*h + kappa - 1, see host/ref_twiddle.c) reproduces the reference's output - including its
// recurrence error - to ~3e-16. tools/emulate_plan.py is the numpy statement of the same algebra.
//
// Global addressing (elements are double2 = complex_t, include/fft_common.h:28):
//   CONTIG   one or more whole transforms per CTA, unit stride (N <= 8192)
//   STRIDED  first / middle pass of a multi-pass plan: rows of C contiguous elements, strided rows
//   LAST     last pass: reads a contiguous block, writes rows of C contiguous elements at stride M
// Every thread loads and stores 16-byte elements; a warp touches >= 128 contiguous bytes per request.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fftb200 {

typedef double2 cd;

enum { MODE_CONTIG = 0, MODE_STRIDED = 1, MODE_LAST = 2 };

struct TileArgs {
    const cd* in;
    cd* out;
    const cd* tab;        // per-stage twiddle table, forward direction (entry (stage s, j) at 2^(s-1)-1+j)
    long long ntiles;     // tiles in this launch
    long long batch;      // transforms (CONTIG: validity bound)
    int log_n;            // log2 of the full transform length N
    int log_m;            // log2 M: stages completed by earlier passes
    int inverse;          // 0 forward, 1 inverse (conjugate in, conjugate out, scale)
    double scale;         // applied on the final store of the last pass when inverse
    int final_pass;       // this launch writes the user-visible result
    // Distributed transform: the last pass of a partial plan stores straight into the peers' exchange buffers
    // (P2P over NVLink) in the layout the next local pass wants, so the all-to-all is fused into the kernel.
    // Output index idx = row * W + col (W = 2^peer_lw) goes to rank row >> peer_lrows at
    // ((row & (2^peer_lrows - 1)) << (peer_lw + peer_lg)) + (peer_me << peer_lw) + col. peers == nullptr: local.
    cd* const* peers;
    int peer_lw, peer_lrows, peer_lg, peer_me;
    // Bluestein (bluestein.c:107-148): the chirp multiplications and the spectral product ride on the first / last pass
    // instead of running as elementwise kernels of their own (three HBM round trips less per transform).
    //   MUL_PRE   first pass: reads the caller's x (n per transform), element idx < n is x[idx] * conj(mul[idx]), the rest 0
    //   MUL_FB    last pass of the forward transform: X[idx] * mul[idx] (mul = FFT of the wrapped chirp, m entries)
    //   MUL_POST  last pass of the inverse transform: y[idx] = a[idx] * conj(mul[idx]) * mul_scale for idx < n, stored to
    //             the caller's y (n per transform)
    const cd* mul;
    int mul_mode, mul_n;
    double mul_scale;
};
enum { MUL_NONE = 0, MUL_PRE = 1, MUL_FB = 2, MUL_POST = 3 };

__device__ __forceinline__ cd* peer_ptr(const TileArgs& a, long long idx) {
    const long long row = idx >> a.peer_lw, col = idx & ((1LL << a.peer_lw) - 1);
    const int dest = (int)(row >> a.peer_lrows);
    const long long rloc = row & ((1LL << a.peer_lrows) - 1);
    return a.peers[dest] + ((rloc << (a.peer_lw + a.peer_lg)) + ((long long)a.peer_me << a.peer_lw) + col);
}

// ---------------------------------------------------------------------------------------------
// arithmetic
// ---------------------------------------------------------------------------------------------

// Radix-2 DIT butterfly, general twiddle: (lo, hi) <- (lo + w*hi, lo - w*hi) in 6 FP64 instructions.
__device__ __forceinline__ void bfly(cd& lo, cd& hi, const cd w) {
    double sx = fma(w.x, hi.x, lo.x);
    double sy = fma(w.x, hi.y, lo.y);
    sx = fma(-w.y, hi.y, sx);
    sy = fma(w.y, hi.x, sy);
    hi.x = fma(2.0, lo.x, -sx);
    hi.y = fma(2.0, lo.y, -sy);
    lo.x = sx;
    lo.y = sy;
}

// Butterfly with the exact twiddle exp(-2*pi*i*Q/(2*HH)), used for the first stages of a transform
// where the reference's table holds (within 1 ulp) these constants: w = 1 and w = -i cost 4 adds.
template <int HH, int Q>
__device__ __forceinline__ void bfly_exact(cd& lo, cd& hi) {
    if constexpr (Q == 0) {
        cd s = make_double2(lo.x + hi.x, lo.y + hi.y);
        hi = make_double2(lo.x - hi.x, lo.y - hi.y);
        lo = s;
    } else if constexpr (2 * Q == HH) {  // -i
        cd s = make_double2(lo.x + hi.y, lo.y - hi.x);
        hi = make_double2(lo.x - hi.y, lo.y + hi.x);
        lo = s;
    } else if constexpr (4 * Q == HH) {  // (1 - i)/sqrt2
        constexpr double c = 0.70710678118654752440;
        const double a = hi.x + hi.y, b = hi.y - hi.x;
        cd s = make_double2(fma(c, a, lo.x), fma(c, b, lo.y));
        hi = make_double2(fma(-c, a, lo.x), fma(-c, b, lo.y));
        lo = s;
    } else if constexpr (4 * Q == 3 * HH) {  // (-1 - i)/sqrt2
        constexpr double c = 0.70710678118654752440;
        const double a = hi.y - hi.x, b = -(hi.x + hi.y);
        cd s = make_double2(fma(c, a, lo.x), fma(c, b, lo.y));
        hi = make_double2(fma(-c, a, lo.x), fma(-c, b, lo.y));
        lo = s;
    } else {
        static_assert(HH == 8, "exact twiddles are tabulated up to radix 16");
        constexpr double c1 = 0.92387953251128675613, s1 = 0.38268343236508977173;
        constexpr double wr = (Q == 1 || Q == 7) ? (Q == 1 ? c1 : -c1) : (Q == 3 ? s1 : -s1);
        constexpr double wi = (Q == 1 || Q == 7) ? -s1 : -c1;
        bfly(lo, hi, make_double2(wr, wi));
    }
}

template <int LR> __host__ __device__ constexpr int bitrev_c(int x) {
    int r = 0;
    for (int i = 0; i < LR; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

// In-register radix-2^LR DIT over w[0..R) (inputs already placed in bit-reversed order) with one
// general twiddle per (sub-stage, position): tw[h], h = 2^(s-1) + q.
template <int LR, int S, int BASE, int Q>
struct SubStageGen {
    static __device__ __forceinline__ void run(cd* w, const cd* tw) {
        constexpr int HH = 1 << (S - 1), R = 1 << LR;
        bfly(w[BASE + Q], w[BASE + Q + HH], tw[HH + Q]);
        if constexpr (Q + 1 < HH) SubStageGen<LR, S, BASE, Q + 1>::run(w, tw);
        else if constexpr (BASE + 2 * HH < R) SubStageGen<LR, S, BASE + 2 * HH, 0>::run(w, tw);
        else if constexpr (S < LR) SubStageGen<LR, S + 1, 0, 0>::run(w, tw);
    }
};
template <int LR, int S, int BASE, int Q>
struct SubStageExact {
    static __device__ __forceinline__ void run(cd* w) {
        constexpr int HH = 1 << (S - 1), R = 1 << LR;
        bfly_exact<HH, Q>(w[BASE + Q], w[BASE + Q + HH]);
        if constexpr (Q + 1 < HH) SubStageExact<LR, S, BASE, Q + 1>::run(w);
        else if constexpr (BASE + 2 * HH < R) SubStageExact<LR, S, BASE + 2 * HH, 0>::run(w);
        else if constexpr (S < LR) SubStageExact<LR, S + 1, 0, 0>::run(w);
    }
};

// ---------------------------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------------------------
template <int LOGP_, int LOGC_, int LOGE_, int NT_, int MODE_, bool TRIV_, int LR0_, int LR1_, int LR2_,
          int LR3_, int MINB_, int PSH_>
struct TileCfg {
    static constexpr int LOGP = LOGP_, LOGC = LOGC_, LOGE = LOGE_, NT = NT_, MODE = MODE_;
    static constexpr bool TRIV = TRIV_;           // first sub-pass uses exact constants (stages 1..LR0 of N)
    static constexpr int MINB = MINB_, PSH = PSH_;  // PSH: one pad element every 2^PSH (bank-conflict padding)
    static constexpr int LOGNP = LOGP + LOGC;     // flat tile size
    static constexpr int NP = 1 << LOGNP, E = 1 << LOGE, T = NP / E, THREADS = T * NT;
    static constexpr int NSUB = (LR0_ > 0) + (LR1_ > 0) + (LR2_ > 0) + (LR3_ > 0);
    static constexpr int SM_STRIDE = NP + (NP >> PSH) + 2;  // elements per sub-tile in shared memory
    static constexpr size_t SMEM_BYTES = NSUB > 1 ? (size_t)SM_STRIDE * NT * sizeof(cd) : 0;
    static __host__ __device__ constexpr int lr(int i) { return i == 0 ? LR0_ : i == 1 ? LR1_ : i == 2 ? LR2_ : LR3_; }
    static __host__ __device__ constexpr int lm(int i) {  // log2 of points already combined before sub-pass i
        int s = 0;
        for (int j = 0; j < i; j++) s += lr(j);
        return s;
    }
    static_assert(LR0_ + LR1_ + LR2_ + LR3_ == LOGP_, "radices must multiply to P");
    static_assert(LR0_ <= LOGE_ && LR1_ <= LOGE_ && LR2_ <= LOGE_ && LR3_ <= LOGE_, "radix larger than E");
    static_assert(MODE_ == MODE_CONTIG ? LOGC_ == 0 : NT_ == 1, "CONTIG has no columns; others one tile per CTA");
};

template <class C> __device__ __forceinline__ int smpos(int idx) { return idx + (idx >> C::PSH); }

struct TileCtx {
    long long in_base, out_base, in_rs, out_rs;  // element offsets / row strides
    long long tb;                                // transform index (fused Bluestein factors address per transform)
    int kap_base, kap_col;
    bool valid;
};

// ---------------------------------------------------------------------------------------------
// one sub-pass: gather -> radix-2^LR butterflies -> scatter
// ---------------------------------------------------------------------------------------------
template <class C, int I, bool INV>
__device__ __forceinline__ void subpass(cd (&v)[C::E], cd* __restrict__ sm, const int t, const TileArgs& a,
                                        const TileCtx& cx) {
    constexpr int LR = C::lr(I), R = 1 << LR, NB = C::E / R;
    constexpr int LM = C::lm(I);                    // log2 M_loc
    constexpr int LS = C::LOGNP - LM - LR;          // log2 S, S = NP / (M_loc * R)
    constexpr int S = 1 << LS;
    constexpr bool FIRST = (I == 0), FINAL = (I == C::NSUB - 1);
    constexpr bool GATHER_LAST = FIRST && C::MODE == MODE_LAST;
    constexpr int CM = (1 << C::LOGC) - 1;

    int ub[NB];  // butterfly ids
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int tau = t + C::T * b;
        if constexpr (GATHER_LAST) {
            // coalesced order: consecutive threads walk down one column of the contiguous input block
            const int rowlow = tau & ((1 << (C::LOGP - LR)) - 1), col = tau >> (C::LOGP - LR);
            ub[b] = col + (rowlow << C::LOGC);
        } else {
            ub[b] = tau;
        }
    }

    // ---- gather ----
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int u = ub[b];
        if constexpr (FIRST) {
            if constexpr (GATHER_LAST) {
                const int col = u & CM, rowlow = u >> C::LOGC;
                const cd* p = a.in + cx.in_base + ((long long)col << C::LOGP) + rowlow;
#pragma unroll
                for (int rho = 0; rho < R; rho++) {
                    cd x = p[rho << (C::LOGP - LR)];
                    if (INV) x.y = -x.y;
                    v[b * R + bitrev_c<LR>(rho)] = x;
                }
            } else {
                // f = u + rho * NP/R ; (row, col) = (f >> LOGC, f & CM) ; col is fixed per thread
                const cd* p = a.in + cx.in_base + (u & CM) + (long long)(u >> C::LOGC) * cx.in_rs;
                const long long step = (long long)(C::NP >> (LR + C::LOGC)) * cx.in_rs;
                if (C::MODE == MODE_STRIDED && a.mul_mode == MUL_PRE) {
                    // a = x * conj(chirp), zero-padded from n to m (bluestein.c:107-109)
                    const long long i0 = cx.in_base - (cx.tb << a.log_n) + (u & CM) + (long long)(u >> C::LOGC) * cx.in_rs;
                    const cd* xin = a.in + cx.tb * a.mul_n;
#pragma unroll
                    for (int rho = 0; rho < R; rho++) {
                        const long long idx = i0 + rho * step;
                        cd x = make_double2(0.0, 0.0);
                        if (idx < a.mul_n) {
                            const cd xv = xin[idx], w = __ldg(a.mul + idx);
                            x = make_double2(fma(xv.x, w.x, xv.y * w.y), fma(xv.y, w.x, -(xv.x * w.y)));
                        }
                        v[b * R + bitrev_c<LR>(rho)] = x;
                    }
                } else {
#pragma unroll
                    for (int rho = 0; rho < R; rho++) {
                        cd x = make_double2(0.0, 0.0);
                        if (cx.valid) x = p[rho * step];
                        if (INV) x.y = -x.y;
                        v[b * R + bitrev_c<LR>(rho)] = x;
                    }
                }
            }
        } else {
            const int cp = u & (S - 1), kloc = u >> LS;
            const int base = cp + (kloc << (C::LOGNP - LM));
#pragma unroll
            for (int rho = 0; rho < R; rho++) v[b * R + bitrev_c<LR>(rho)] = sm[smpos<C>(base + rho * S)];
        }
    }

    // ---- butterflies ----
#pragma unroll
    for (int b = 0; b < NB; b++) {
        if constexpr (FIRST && C::TRIV) {
            SubStageExact<LR, 1, 0, 0>::run(&v[b * R]);
        } else {
            const int u = ub[b];
            const int col = u & CM, kloc = FIRST ? 0 : (u >> LS);
            const int kappa = cx.kap_base + cx.kap_col * col + (kloc << a.log_m);
            const cd* tp = a.tab + (kappa - 1);
            const int lmt = a.log_m + LM;  // log2 Mt
            cd tw[R];
#pragma unroll
            for (int h = 1; h < R; h++) tw[h] = __ldg(tp + ((long long)h << lmt));
            SubStageGen<LR, 1, 0, 0>::run(&v[b * R], tw);
        }
    }

    // ---- scatter ----
    if constexpr (FINAL) {
        constexpr bool conj_out = INV;
        const double sc = (INV && a.final_pass) ? a.scale : 1.0;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const int u = ub[b];
            const long long o = cx.out_base + (u & CM) + (long long)(u >> C::LOGC) * cx.out_rs;
            cd* p = a.out + o;
            const long long step = (long long)(C::NP >> (LR + C::LOGC)) * cx.out_rs;
            if (C::MODE == MODE_LAST && a.mul_mode != MUL_NONE) {
                // factors are fetched four at a time ahead of the stores they feed (a load cannot be hoisted over a store
                // by the compiler: the pointers may alias)
                const long long i0 = o - (cx.tb << a.log_n);   // index within the transform
                const bool fb = a.mul_mode == MUL_FB;
                const long long lim = fb ? (1LL << a.log_n) : a.mul_n;
                const double s2 = a.mul_scale;
                cd* const yo = a.out + cx.tb * a.mul_n;
#pragma unroll
                for (int q0 = 0; q0 < R; q0 += 4) {
                    cd w[4];
#pragma unroll
                    for (int j = 0; j < 4 && q0 + j < R; j++) {
                        const long long idx = i0 + (q0 + j) * step;
                        w[j] = idx < lim ? __ldg(a.mul + idx) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int j = 0; j < 4 && q0 + j < R; j++) {
                        cd x = v[b * R + q0 + j];
                        if (conj_out) { x.x *= sc; x.y *= -sc; }
                        const long long idx = i0 + (q0 + j) * step;
                        if (fb) p[(q0 + j) * step] = make_double2(fma(x.x, w[j].x, -(x.y * w[j].y)), fma(x.x, w[j].y, x.y * w[j].x));
                        else if (idx < lim) yo[idx] = make_double2(fma(x.x, w[j].x, x.y * w[j].y) * s2, fma(x.y, w[j].x, -(x.x * w[j].y)) * s2);
                    }
                }
                continue;
            }
#pragma unroll
            for (int q = 0; q < R; q++) {
                cd x = v[b * R + q];
                if (conj_out) { x.x *= sc; x.y *= -sc; }
                if (a.peers) *peer_ptr(a, o + q * step) = x;
                else if (cx.valid) p[q * step] = x;
            }
        }
    } else {
        __syncthreads();  // every reader of the previous contents is done
#pragma unroll
        for (int b = 0; b < NB; b++) {
#pragma unroll
            for (int q = 0; q < R; q++) sm[smpos<C>(ub[b] + q * (C::NP >> LR))] = v[b * R + q];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// the kernel: persistent loop over tiles
// ---------------------------------------------------------------------------------------------
// INV: conjugate in, conjugate (and scale, on the pass that ends the plan) out - compiled in, not selected per element
template <class C, bool INV>
__global__ void __launch_bounds__(C::THREADS, C::MINB) fft_tile_kernel(const TileArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd* sm_all = reinterpret_cast<cd*>(smem_raw);
    const int st = threadIdx.x / C::T, t = threadIdx.x % C::T;
    cd* sm = sm_all + (size_t)st * C::SM_STRIDE;

    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        TileCtx cx;
        const int log_n = a.log_n, log_m = a.log_m;
        if constexpr (C::MODE == MODE_CONTIG) {
            const long long b = tile * C::NT + st;
            cx.valid = b < a.batch;
            cx.tb = b;
            cx.in_base = cx.out_base = b << log_n;
            cx.in_rs = cx.out_rs = 1;
            cx.kap_base = 0; cx.kap_col = 0;
        } else if constexpr (C::MODE == MODE_STRIDED) {
            // rest = N / (M * P) columns c' per (transform, k); tiles cover C of them
            const int log_rest = log_n - log_m - C::LOGP;
            const int log_tpk = log_rest - C::LOGC;                 // tiles per k
            const int log_tpx = log_tpk + log_m;                    // tiles per transform
            const long long b = tile >> log_tpx;
            const long long r = tile & ((1LL << log_tpx) - 1);
            const long long k = r >> log_tpk;
            const long long c0 = (r & ((1LL << log_tpk) - 1)) << C::LOGC;
            cx.valid = true;
            cx.tb = b;
            cx.in_base = (b << log_n) + c0 + (k << (log_n - log_m));
            cx.in_rs = 1LL << log_rest;
            cx.out_base = (b << log_n) + c0 + (k << log_rest);
            cx.out_rs = 1LL << (log_rest + log_m);
            cx.kap_base = (int)k; cx.kap_col = 0;
        } else {
            // transform index fastest: the batch walks through the same k block back to back, so the block's late-stage
            // twiddles (as many bytes as the tile itself, every table entry used once per transform) come from L2 for all
            // but the first transform. 2^24 x 16: DRAM reads 8.63 -> 4.6 GB per execution.
            const long long b = tile % a.batch;
            const long long k0 = (tile / a.batch) << C::LOGC;
            cx.valid = true;
            cx.tb = b;
            cx.in_base = (b << log_n) + (k0 << C::LOGP);
            cx.in_rs = 0;
            cx.out_base = (b << log_n) + k0;
            cx.out_rs = 1LL << log_m;
            cx.kap_base = (int)k0; cx.kap_col = 1;
        }
        cd v[C::E];
        subpass<C, 0, INV>(v, sm, t, a, cx);
        if constexpr (C::NSUB > 1) subpass<C, 1, INV>(v, sm, t, a, cx);
        if constexpr (C::NSUB > 2) subpass<C, 2, INV>(v, sm, t, a, cx);
        if constexpr (C::NSUB > 3) subpass<C, 3, INV>(v, sm, t, a, cx);
    }
}

}  // namespace fftb200
