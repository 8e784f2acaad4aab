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
    return 0;
}
