// l2bench.cu - measures what a two-pass-through-L2 dataflow can reach on this GPU (development tool).
//   mode 0: big -> big copy (HBM copy rate, the roofline denominator)
//   mode 1: small -> small copy, L2 resident (LTS throughput)
//   mode 2: big_in -> ring (L2 resident, reused) -> big_out, each CTA re-reading what it wrote `lag` tiles ago:
//           the upper bound for the fused two-pass FFT (32 B/pt of HBM traffic + 32 B/pt of L2-only traffic)
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2bench tools/l2bench.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int TILE = 4096;  // double2 per tile (64 KB)
constexpr int THREADS = 512;

__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ double2 ldh(const double2* p, uint64_t pol) {
    double2 v; asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol)); return v;
}
__device__ __forceinline__ void sth(double2* p, double2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1,%2}, %3;" :: "l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
}

template <bool HINT>
__device__ __forceinline__ void copy_tile(double2* dst, const double2* src, uint64_t pl, uint64_t ps) {
    double2 v[TILE / THREADS];
#pragma unroll
    for (int e = 0; e < TILE / THREADS; e++) v[e] = HINT ? ldh(src + threadIdx.x + e * THREADS, pl) : src[threadIdx.x + e * THREADS];
#pragma unroll
    for (int e = 0; e < TILE / THREADS; e++) { v[e].x += 1.0; if (HINT) sth(dst + threadIdx.x + e * THREADS, v[e], ps); else dst[threadIdx.x + e * THREADS] = v[e]; }
}

// mode 0/1: tiles [0, ntiles) of src -> dst, `iters` sweeps
template <bool HINT>
__global__ void __launch_bounds__(THREADS) copy_kernel(double2* dst, const double2* src, long long ntiles, int iters) {
    const uint64_t pf = pol_first();
    for (int it = 0; it < iters; it++)
        for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) copy_tile<HINT>(dst + t * TILE, src + t * TILE, pf, pf);
}

// mode 2
template <bool HINT>
__global__ void __launch_bounds__(THREADS) twopass_kernel(double2* out, const double2* in, double2* ring, long long ntiles,
                                                          long long ring_tiles, int lag) {
    const uint64_t pf = pol_first(), pl = pol_last();
    const long long mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (long long k = 0; k < mine + lag; k++) {
        if (k < mine) {
            const long long t = blockIdx.x + k * gridDim.x;
            copy_tile<HINT>(ring + (t % ring_tiles) * TILE, in + t * TILE, pf, pl);
        }
        if (k >= lag) {
            const long long t = blockIdx.x + (k - lag) * gridDim.x;
            __syncthreads();
            copy_tile<HINT>(out + t * TILE, ring + (t % ring_tiles) * TILE, pl, pf);
        }
    }
}

// mode 3: per-SM TMA (bulk async copy) throughput: NB ring buffers of 64 KB per CTA, one thread drives them.
// what: 1 = loads only, 2 = stores only, 3 = load then store of every tile (copy through shared memory)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int NB>
__global__ void __launch_bounds__(128) tma_kernel(double2* out, const double2* in, long long ntiles, long long wrap, int what, long long wrap_out = 0, int spin = 0) {
    extern __shared__ __align__(128) unsigned char smem[];
    double2* bufs = (double2*)smem;
    uint64_t* full = (uint64_t*)(smem + (size_t)NB * TILE * 16);
    if (threadIdx.x == 0) {
        for (int b = 0; b < NB; b++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= NB) return;
    // thread b drives buffer b (as the manager warps of the fused kernel do, here lanes of one warp)
    const int b = threadIdx.x;
    const long long mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    int n = 0;
    for (long long k = b; k < mine; k += NB, n++) {
        const long long t = (blockIdx.x + k * gridDim.x) % wrap;
        if (what & 1) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[b])), "r"(TILE * 16) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(bufs + (size_t)b * TILE)),
                         "l"(in + t * TILE), "r"(TILE * 16), "r"(s32(&full[b])) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(&full[b])), "r"(n & 1) : "memory");
        }
        if (spin) { const long long t0 = clock64(); while (clock64() - t0 < spin) {} }   // stands for the compute time of a tile
        if (what & 2) {
            const long long to = wrap_out ? (blockIdx.x + k * gridDim.x) % wrap_out : t;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + to * TILE), "r"(s32(bufs + (size_t)b * TILE)), "r"(TILE * 16) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const long long big_mb = 4096;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int l2 = 0, pers = 0; CK(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, 0)); CK(cudaDeviceGetAttribute(&pers, cudaDevAttrMaxPersistingL2CacheSize, 0));
    printf("sms %d l2 %d MB maxPersisting %d MB\n", sms, l2 >> 20, pers >> 20);
    double2 *a, *b, *r;
    const long long big_tiles = (big_mb << 20) / (TILE * 16);
    CK(cudaMalloc(&a, big_mb << 20)); CK(cudaMalloc(&b, big_mb << 20)); CK(cudaMalloc(&r, 256ll << 20));
    CK(cudaMemset(a, 0, big_mb << 20)); CK(cudaMemset(b, 0, big_mb << 20)); CK(cudaMemset(r, 0, 256ll << 20));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto time = [&](auto f) { float best = 1e9; for (int i = 0; i < 4; i++) { CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i && ms < best) best = ms; } CK(cudaGetLastError()); return best; };
    const int grid = sms * 2;
    if (mode == 0 || mode == 9) {
        for (int h = 0; h < 2; h++) {
            float ms = time([&] { if (h) copy_kernel<true><<<grid, THREADS>>>(b, a, big_tiles, 1); else copy_kernel<false><<<grid, THREADS>>>(b, a, big_tiles, 1); });
            printf("mode0 big->big hint=%d: %.3f ms, %.0f GB/s (r+w)\n", h, ms, 2.0 * (big_mb << 20) / ms * 1e-6);
        }
    }
    if (mode == 1 || mode == 9) {
        for (int mb : {4, 8, 16, 32, 48, 64}) {
            const long long tiles = ((long long)mb << 20) / (TILE * 16); const int iters = 4096 / mb;
            float ms = time([&] { copy_kernel<false><<<grid, THREADS>>>(r + tiles * TILE, r, tiles, iters); });
            printf("mode1 L2 copy %d MB -> %d MB x%d: %.3f ms, %.0f GB/s (r+w)\n", mb, mb, iters, ms, 2.0 * ((double)mb * 1048576) * iters / ms * 1e-6);
        }
    }
    if (mode == 2 || mode == 9) {
        for (int h = 0; h < 2; h++)
            for (int mb : {8, 16, 32, 48, 64, 96}) {
                const long long rt = ((long long)mb << 20) / (TILE * 16);
                const int lag = (int)(rt / grid / 2) > 0 ? (int)(rt / grid / 2) : 1;
                float ms = time([&] { if (h) twopass_kernel<true><<<grid, THREADS>>>(b, a, r, big_tiles, rt, lag); else twopass_kernel<false><<<grid, THREADS>>>(b, a, r, big_tiles, rt, lag); });
                printf("mode2 twopass ring %d MB lag %d hint=%d: %.3f ms, strict %.0f GB/s (32 B/pt)\n", mb, lag, h, ms, 2.0 * (big_mb << 20) / ms * 1e-6);
            }
    }
    if (mode == 3 || mode == 9) {
        CK(cudaFuncSetAttribute(tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE * 16 + 64));
        CK(cudaFuncSetAttribute(tma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TILE * 16 + 64));
        for (int l2 = 0; l2 < 2; l2++)
            for (int what = 1; what <= 3; what++)
                for (int nb = 2; nb <= 3; nb++) {
                    const long long wrap = l2 ? (24ll << 20) / (TILE * 16) : big_tiles;
                    float ms = time([&] { if (nb == 3) tma_kernel<3><<<sms, 128, 3 * TILE * 16 + 64>>>(b, a, big_tiles, wrap, what);
                                          else tma_kernel<2><<<sms, 128, 2 * TILE * 16 + 64>>>(b, a, big_tiles, wrap, what); });
                    const double bytes = (double)(big_mb << 20) * ((what & 1) + ((what >> 1) & 1));
                    printf("mode3 tma %s %s bufs=%d: %.3f ms, %.0f GB/s total, %.1f B/clk/SM @1.9GHz\n", l2 ? "L2(24MB)" : "HBM(4GiB)",
                           what == 1 ? "load " : what == 2 ? "store" : "copy ", nb, ms, bytes / ms * 1e-6, bytes / ms * 1e-6 / sms / 1.9);
                }
    }
    if (mode == 4 || mode == 9) {
        CK(cudaFuncSetAttribute(tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE * 16 + 64));
        const long long rw = (24ll << 20) / (TILE * 16);
        for (int spin : {0, 2000, 4000, 6000}) {
            float ms = time([&] { tma_kernel<3><<<sms, 128, 3 * TILE * 16 + 64>>>(r, a, big_tiles, big_tiles, 3, rw, spin); });
            printf("mode4 tma copy HBM -> L2 ring (pass A shape), 3 bufs, compute %d cyc: %.3f ms\n", spin, ms);
            ms = time([&] { tma_kernel<3><<<sms, 128, 3 * TILE * 16 + 64>>>(b, r, big_tiles, rw, 3, 0, spin); });
            printf("mode4 tma copy L2 ring -> HBM (pass B shape), 3 bufs, compute %d cyc: %.3f ms\n", spin, ms);
        }
    }
    return 0;
}
