// fft_pipe13t.cuh - whole transforms of 8192 points, one visit to shared memory, with TENSOR MEMORY as the parking space that
// shared memory and the register file do not have (round 2; replaces fft_pipe13.cuh as the planner's choice for N = 8192).
//
// Replaces the same reference code as fft_pipe.cuh (the butterfly loop of algorithms/core/radix2_dit.c:70-119 and the never-built
// cufftExecZ2Z call of gpu/fft_cuda.cu:166-185) for N = 8192, where one transform is two 64 KB tiles.
//
// All 13 stages of this size may use the accurate twiddle tables (SURVEY.md 7.0 hybrid rule: 2.6e-14 from the reference's
// recurrence, far inside the 1e-12 bar), so the regrouping is free. This kernel takes the top radix-2 FIRST (decimation in
// frequency), because then the two halves of a transform are used exactly as they lie in memory:
//     a[t] = x[t] + x[t + 4096],   b[t] = (x[t] - x[t + 4096]) w^t,   X[2k] = FFT_4096(a)[k],   X[2k + 1] = FFT_4096(b)[k]
//   * ONE group of 256 threads owns a whole transform (the two groups of a CTA work on different transforms and never meet: no
//     CTA-wide rendezvous, no trade between groups, no de-interleaving gather - fft_pipe13.cuh has all three);
//   * a thread reads its 16 points of both halves (conflict-free 128-bit reads), keeps a in registers and PARKS b in tensor
//     memory (tcgen05.st, 64 columns of its warp's lane quadrant): 64 registers and no shared memory for 4096 waiting points;
//   * the first half's ring buffer is released at once; a and then b go through the unchanged 4096-point dataflow of
//     fft_pipe_kernel<12> in the second half's buffer; X[2k] waits in tensor memory for X[2k + 1], and each thread stores 32
//     contiguous bytes per bin pair (a warp: 1 KB contiguous).
// Per transform and thread: 32 + 2 x 64 = 160 shared-memory accesses against 2 x 112 in fft_pipe13.cuh, 9 group barriers and no
// CTA-wide one against 8 + 8. Tensor memory: 512 columns (the whole of it: one CTA per SM), 2 x 64 per warp.
// Ring: halves h = 2k, 2k + 1 of the CTA's k-th transform go to buffer h % 3; a buffer's fills are counted on FOUR barriers used
// in turn (round & 3), so that all phases of one barrier are waited for by the same group in program order (the two-barrier
// scheme of fft_pipe.cuh, generalised: the groups take turns every two fills here).
#pragma once
#ifdef PIPE13T_PROF
#include <cstdio>
#endif
#include "fft_pipe13.cuh"

namespace fftb200 {

// one complex double <-> 4 columns of the warp's lane quadrant (.x4: four consecutive registers - wider shapes make the register
// allocator shuffle 16-register tuples around and spill: 288 bytes of spills with .x16)
__device__ __forceinline__ void tmem_st1(uint32_t taddr, const cd v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(__double2loint(v.x)), "r"(__double2hiint(v.x)), "r"(__double2loint(v.y)), "r"(__double2hiint(v.y)) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const cd* v) {
    tmem_st1(taddr, v[0]); tmem_st1(taddr + 4, v[1]); tmem_st1(taddr + 8, v[2]); tmem_st1(taddr + 12, v[3]);
}
__device__ __forceinline__ void tmem_ld1(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
    tmem_ld1(taddr, r); tmem_ld1(taddr + 4, r + 4); tmem_ld1(taddr + 8, r + 8); tmem_ld1(taddr + 12, r + 12);
}
__device__ __forceinline__ cd tmem_cd(const uint32_t* r) {
    return make_double2(__hiloint2double((int)r[1], (int)r[0]), __hiloint2double((int)r[3], (int)r[2]));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// exp(-2 pi i e / 32) for e < 16
template <int E> struct W32f {
    static constexpr double c = E <= 8 ? W32<E>::c : -W32<16 - E>::c;
    static constexpr double s = E <= 8 ? W32<E>::s : W32<16 - E>::s;
};

// the 4096-point dataflow of fft_pipe_kernel<12> from registers to registers: x = the 16 points idx = t + 256 e (bit-reversed
// placement already done), through the buffer sm; on return x[q] = FFT[t + 256 q]. `release` runs once every gather is done.
template <class Release>
__device__ __forceinline__ void fft4096_in_buffer(cd* x, cd* sm, const int g, const int t, const cd* tw1p, const cd* tab, Release release) {
    constexpr int LN16 = 8;
    const int cp = t & 15, kloc1 = t >> 4;
    const int rd1 = cp + 256 * kloc1;
    SubStageExact<4, 1, 0, 0>::run(x);
#pragma unroll
    for (int e = 0; e < 16; e++) sm[pipe_swz(t + 256 * e)] = x[e];
    group_sync(g);
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
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = sm[pipe_swz(16 * t + bitrev_c<4>(i))];
    const cd wa = __ldg(tab + (t - 1) + (1 << LN16)), wb = __ldg(tab + (t - 1) + (2 << LN16));
    const cd wc = __ldg(tab + (t - 1) + (4 << LN16)), wd = __ldg(tab + (t - 1) + (8 << LN16));
    group_sync(g);   // every gather is done: the buffer may be written again
    release();
    constexpr double C8 = 0.70710678118654752440, C16 = 0.92387953251128675613, S16 = 0.38268343236508977173;
    cd tw[16];
    tw[1] = wa; tw[2] = wb; tw[4] = wc; tw[8] = wd;
    tw[5] = make_double2((wc.x + wc.y) * C8, (wc.y - wc.x) * C8);
    tw[9] = cmulc(wd, C16, -S16);
    tw[10] = make_double2((wd.x + wd.y) * C8, (wd.y - wd.x) * C8);
    tw[11] = cmulc(wd, S16, -C16);
    SubStageSym<4, 1, 0, 0>::run(x, tw);
}

template <bool INV, int E0>
struct DifChunk {   // points e = E0 .. E0 + 3 of the thread: a into x (bit-reversed slot), b = (x0 - x1) w^(t + 256 e) into v
    static __device__ __forceinline__ void run(cd* x, cd* v, const cd* s0, const cd* s1, const int t, const cd w13) {
        one<0>(x, v, s0, s1, t, w13); one<1>(x, v, s0, s1, t, w13); one<2>(x, v, s0, s1, t, w13); one<3>(x, v, s0, s1, t, w13);
    }
    template <int I>
    static __device__ __forceinline__ void one(cd* x, cd* v, const cd* s0, const cd* s1, const int t, const cd w13) {
        constexpr int E = E0 + I;
        cd p = s0[t + 256 * E], q = s1[t + 256 * E];
        if (INV) { p.y = -p.y; q.y = -q.y; }
        x[bitrev_c<4>(E)] = make_double2(p.x + q.x, p.y + q.y);
        const cd d = make_double2(p.x - q.x, p.y - q.y);
        const cd w = E == 0 ? w13 : cmulc(w13, W32f<E>::c, -W32f<E>::s);   // w^(t + 256 e) = w^t W32^e
        v[I] = make_double2(fma(d.x, w.x, -(d.y * w.y)), fma(d.x, w.y, d.y * w.x));
    }
};

template <bool INV>
__global__ void __launch_bounds__(2 * PIPE_GROUP, 1) fft_pipe13t_kernel(const PipeArgs a) {
    constexpr int N = 8192, H = 4096, NBAR = 4 * PIPE_STAGES;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    cd* const bufs = reinterpret_cast<cd*>(smem_raw);
    cd* const tw1s = bufs + (size_t)PIPE_STAGES * PIPE_TILE;
    uint64_t* const full = reinterpret_cast<uint64_t*>(tw1s + PIPE_TW1);      // [4][3]: barrier (round & 3) of buffer b
    uint32_t* const tmem_base_p = reinterpret_cast<uint32_t*>(full + NBAR);

    const int g = threadIdx.x / PIPE_GROUP, t = threadIdx.x % PIPE_GROUP;
    const int first = blockIdx.x, stride = gridDim.x;
    const int my_tr = first < a.ntiles ? (int)((a.ntiles - first + stride - 1) / stride) : 0;   // transforms of this CTA
    const int my_halves = 2 * my_tr;

    auto issue = [&](int h, int b, uint32_t rnd) {   // half h of this CTA (transform h / 2) -> buffer b
        const long long tr = first + (long long)(h >> 1) * stride;
        uint64_t* const bar = &full[b + PIPE_STAGES * (rnd & 3)];
        mbar_expect_tx(bar, H * (uint32_t)sizeof(cd));
        bulk_load(bufs + (size_t)b * PIPE_TILE, a.in + tr * N + (h & 1) * H, H * (uint32_t)sizeof(cd), bar);
    };

    if (threadIdx.x == 0) {
#pragma unroll
        for (int b = 0; b < NBAR; b++) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (threadIdx.x < 32) {   // warp 0 allocates all 512 columns of tensor memory (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_base_p)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // middle sub-pass twiddles -> shared: stage (4 + s), position kloc (fft_pipe.cuh)
    if (threadIdx.x >= 32 && threadIdx.x < 32 + 16 * 8) {
        const int i = threadIdx.x - 32, kl = i >> 3, e = i & 7;
        tw1s[i] = __ldg(a.tab + ((sym_h(e) << 4) + kl - 1));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (threadIdx.x == 0) {
        for (int h = 0; h < PIPE_STAGES && h < my_halves; h++) issue(h, h, 0);
    }
    // this warp's tensor-memory window: lanes of quadrant (warp % 4), columns [128 (warp / 4), + 128): b in the first 64, X[2k] in the rest
    const int warp = threadIdx.x >> 5;
    const uint32_t tm_b = *tmem_base_p + ((uint32_t)(32 * (warp & 3)) << 16) + 128u * (uint32_t)(warp >> 2);
    const uint32_t tm_e = tm_b + 64;

    const cd* const tw1p = tw1s + (t >> 4) * 8;
    const double sc = a.scale;

#ifdef PIPE13T_PROF
    long long pf[6] = {0, 0, 0, 0, 0, 0};
#define PF(i) { const long long now_ = clock64(); pf[i] += now_ - pf_t; pf_t = now_; }
#else
#define PF(i)
#endif
    for (int k = g; k < my_tr; k += 2) {
#ifdef PIPE13T_PROF
        long long pf_t = clock64();
#endif
        const int h0 = 2 * k, h1 = h0 + 1;
        const int b0 = h0 % PIPE_STAGES, b1 = h1 % PIPE_STAGES;
        const uint32_t r0 = h0 / PIPE_STAGES, r1 = h1 / PIPE_STAGES;
        cd* const s0 = bufs + (size_t)b0 * PIPE_TILE;
        cd* const s1 = bufs + (size_t)b1 * PIPE_TILE;
        mbar_wait_bounded(&full[b0 + PIPE_STAGES * (r0 & 3)], (r0 >> 2) & 1);
        mbar_wait_bounded(&full[b1 + PIPE_STAGES * (r1 & 3)], (r1 >> 2) & 1);
        PF(0)
        cd x[16];
        const cd w13 = __ldg(a.tab + (H - 1) + t);   // stage 13, position t: w^t (re-read per transform: an L1 hit, four registers less to carry)
        __syncwarp();   // the waits above leave their loops lane by lane; tcgen05 instructions are warp-wide
        // ---- top radix-2 (decimation in frequency): a stays, b is parked in tensor memory ----
        {
            cd v[4];
            DifChunk<INV, 0>::run(x, v, s0, s1, t, w13);  tmem_st4(tm_b, v);
            DifChunk<INV, 4>::run(x, v, s0, s1, t, w13);  tmem_st4(tm_b + 16, v);
            DifChunk<INV, 8>::run(x, v, s0, s1, t, w13);  tmem_st4(tm_b + 32, v);
            DifChunk<INV, 12>::run(x, v, s0, s1, t, w13); tmem_st4(tm_b + 48, v);
        }
        PF(1)
        group_sync(g);   // both halves have been read: the first one's buffer goes back to the ring, the second hosts the exchanges
        if (t == 0 && h0 + PIPE_STAGES < my_halves) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(h0 + PIPE_STAGES, b0, r0 + 1);
        }
        // ---- X[2k] = FFT_4096(a), then X[2k + 1] = FFT_4096(b): one body, two trips (two inlined copies spill twice as much) ----
        {
            // (the refill of the exchange buffer is prepared here so that only three values live through the transforms)
            const bool refill = t == 0 && h1 + PIPE_STAGES < my_halves;
            const long long trn = first + (long long)((h1 + PIPE_STAGES) >> 1) * stride;
            const cd* const src = a.in + trn * N + ((h1 + PIPE_STAGES) & 1) * H;
            const uint32_t bar = smem_u32(&full[b1 + PIPE_STAGES * ((r1 + 1) & 3)]), dst = smem_u32(s1);
#pragma unroll 1
            for (int pass = 0; pass < 2; pass++) {
                PF(2 + pass)
                if (pass) {   // park X[2k], fetch b (natural order e -> bit-reversed slot)
                    tmem_st4(tm_e, x); tmem_st4(tm_e + 16, x + 4); tmem_st4(tm_e + 32, x + 8); tmem_st4(tm_e + 48, x + 12);
                    tmem_wait_st();   // (covers the b stores as well)
#pragma unroll
                    for (int c = 0; c < 4; c++) {   // four values at a time: X[2k] has just left these registers
                        uint32_t r[16];
                        tmem_ld4(tm_b + 16 * c, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 4; i++) x[bitrev_c<4>(4 * c + i)] = tmem_cd(&r[4 * i]);
                    }
                }
                fft4096_in_buffer(x, s1, g, t, tw1p, a.tab, [=] {
                    if (pass && refill) {   // the buffer is free after the last gather of the second transform
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(H * (uint32_t)sizeof(cd)) : "memory");
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                                     "r"(H * (uint32_t)sizeof(cd)), "r"(bar) : "memory");
                    }
                });
            }
        }
        PF(4)
        // ---- bins 2k, 2k + 1 leave together: 32 contiguous bytes per thread and q ----
        {
            tmem_wait_st();
            const long long tr = first + (long long)k * stride;
            cd* const p = a.out + tr * N + 2 * t;
#pragma unroll
            for (int c = 0; c < 4; c++) {
                uint32_t r[16];
                tmem_ld4(tm_e + 16 * c, r);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int q = 4 * c + i;
                    cd ev = tmem_cd(&r[4 * i]), od = x[q];
                    if (INV) { ev.x *= sc; ev.y *= -sc; od.x *= sc; od.y *= -sc; }
                    p[512 * q] = ev;
                    p[512 * q + 1] = od;
                }
            }
        }
    }
#ifdef PIPE13T_PROF
    if (blockIdx.x == 3 && t == 0 && my_tr)
        printf("pipe13t prof group %d, %d transforms: per transform cycles: wait loads %lld, dif+park %lld, [sync+refill .. before pass1] %lld, fftA->park/unpark.. %lld, fftB %lld\n", g,
               (my_tr - g + 1) / 2, pf[0] / ((my_tr - g + 1) / 2), pf[1] / ((my_tr - g + 1) / 2), pf[2] / ((my_tr - g + 1) / 2), pf[3] / ((my_tr - g + 1) / 2), pf[4] / ((my_tr - g + 1) / 2));
#endif
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_base_p) : "memory");
    }
}

const void* pipe13t_func(int inverse);
cudaError_t launch_pipe13t(const PipeArgs& a, int grid, cudaStream_t s);

}  // namespace fftb200
