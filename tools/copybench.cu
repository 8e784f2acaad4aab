// copybench.cu - which plain copy shape reaches the highest HBM rate on this chip? (development tool: cuFFT's 1024-point kernel moves
// 6.99 TB/s where torch's copy - the roofline denominator - and this library's TMA ring reach 6.5.)
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/copybench tools/copybench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef double2 cd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// a CTA of T threads moves K * T consecutive elements per tile: all loads first, then all stores
template <int K, bool PERSIST, bool CS>
__global__ void copy_k(const cd* __restrict__ in, cd* __restrict__ out, long long ntiles) {
    const int T = blockDim.x;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const cd* p = in + tile * K * T + threadIdx.x;
        cd* q = out + tile * K * T + threadIdx.x;
        cd v[K];
#pragma unroll
        for (int i = 0; i < K; i++) v[i] = CS ? __ldcs(p + i * T) : p[i * T];
#pragma unroll
        for (int i = 0; i < K; i++) { if (CS) __stcs(q + i * T, v[i]); else q[i * T] = v[i]; }
        if (!PERSIST) break;
    }
}

// persistent, blocked: CTA b owns the tiles [b * per, (b + 1) * per)
template <int K>
__global__ void copy_blocked(const cd* __restrict__ in, cd* __restrict__ out, long long ntiles) {
    const int T = blockDim.x;
    const long long per = (ntiles + gridDim.x - 1) / gridDim.x;
    long long t0 = blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    for (long long tile = t0; tile < t1; tile++) {
        const cd* p = in + tile * K * T + threadIdx.x;
        cd* q = out + tile * K * T + threadIdx.x;
        cd v[K];
#pragma unroll
        for (int i = 0; i < K; i++) v[i] = p[i * T];
#pragma unroll
        for (int i = 0; i < K; i++) q[i * T] = v[i];
    }
}
// persistent, dynamic: tiles handed out by an atomic counter (the next index is fetched before the current tile's loads are waited for)
template <int K>
__global__ void copy_dynamic(const cd* __restrict__ in, cd* __restrict__ out, long long ntiles, unsigned long long* counter) {
    const int T = blockDim.x;
    __shared__ long long next[2];
    if (threadIdx.x == 0) next[0] = atomicAdd(counter, 1ULL);
    __syncthreads();
    int ph = 0;
    for (;;) {
        const long long tile = next[ph];
        if (tile >= ntiles) break;
        if (threadIdx.x == 0) next[ph ^ 1] = atomicAdd(counter, 1ULL);
        const cd* p = in + tile * K * T + threadIdx.x;
        cd* q = out + tile * K * T + threadIdx.x;
        cd v[K];
#pragma unroll
        for (int i = 0; i < K; i++) v[i] = p[i * T];
#pragma unroll
        for (int i = 0; i < K; i++) q[i * T] = v[i];
        __syncthreads();
        ph ^= 1;
    }
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// one-shot TMA: a CTA (one thread) moves one tile of BYTES through shared memory with bulk copies
__global__ void __launch_bounds__(32) copy_tma_oneshot(const char* __restrict__ in, char* __restrict__ out, unsigned bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    if (threadIdx.x != 0) return;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(in + (size_t)blockIdx.x * bytes), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (size_t)blockIdx.x * bytes), "r"(smem_u32(smem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// one-shot, TMA load into shared, results leave through the threads (the pipe kernel's shape): T threads, tile of BYTES
__global__ void copy_tma_ld_thread_st(const char* __restrict__ in, cd* __restrict__ out, unsigned bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem)), "l"(in + (size_t)blockIdx.x * bytes), "r"(bytes), "r"(smem_u32(&bar)) : "memory");
    }
    __syncthreads();
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
    const cd* s = reinterpret_cast<const cd*>(smem);
    cd* q = out + (size_t)blockIdx.x * (bytes / 16);
    for (unsigned i = threadIdx.x; i < bytes / 16; i += blockDim.x) q[i] = s[i];
}

template <int K, bool CS>
static void run(const cd* in, cd* out, long long total, int T, int persist_ctas, cudaEvent_t e0, cudaEvent_t e1) {
    const long long ntiles = total / ((long long)K * T);
    float best = 1e9, sum = 0;
    const int reps = 7;
    for (int rep = 0; rep < reps; rep++) {
        CK(cudaEventRecord(e0));
        if (persist_ctas) copy_k<K, true, CS><<<persist_ctas, T>>>(in, out, ntiles);
        else copy_k<K, false, CS><<<(unsigned)ntiles, T>>>(in, out, ntiles);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep) { sum += ms; if (ms < best) best = ms; }
    }
    CK(cudaGetLastError());
    printf("K=%2d T=%4d %s %s tile=%6d B: best %.3f ms %.0f GB/s, mean %.3f ms\n", K, T, persist_ctas ? "persistent" : "one-shot  ", CS ? "cs" : "  ",
           K * T * 16, best, 32.0 * total / best * 1e-6, sum / (reps - 1));
}

int main() {
    const long long total = 1LL << 28;
    cd *in, *out;
    CK(cudaMalloc(&in, sizeof(cd) * total)); CK(cudaMalloc(&out, sizeof(cd) * total));
    CK(cudaMemset(in, 1, sizeof(cd) * total)); CK(cudaMemset(out, 0, sizeof(cd) * total));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    { float best = 1e9; for (int r = 0; r < 6; r++) { CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(out, in, sizeof(cd) * total, cudaMemcpyDeviceToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
      printf("cudaMemcpy D2D: %.3f ms %.0f GB/s\n", best, 32.0 * total / best * 1e-6); }
    for (int T = 64; T <= 1024; T *= 2) {
        run<4, false>(in, out, total, T, 0, e0, e1);
        run<8, false>(in, out, total, T, 0, e0, e1);
        run<16, false>(in, out, total, T, 0, e0, e1);
    }
    run<8, true>(in, out, total, 128, 0, e0, e1);
    run<16, true>(in, out, total, 64, 0, e0, e1);
    run<16, true>(in, out, total, 128, 0, e0, e1);
    for (int c = 1; c <= 8; c *= 2) {
        run<8, false>(in, out, total, 256, 148 * c, e0, e1);
        run<16, false>(in, out, total, 256, 148 * c, e0, e1);
        run<8, false>(in, out, total, 512, 148 * c, e0, e1);
    }
    for (int c = 1; c <= 4; c *= 2) {
        for (int rep = 0; rep < 2; rep++) {
            float best = 1e9;
            for (int r = 0; r < 6; r++) {
                CK(cudaEventRecord(e0));
                if (rep == 0) copy_blocked<8><<<148 * c, 512>>>(in, out, total / (8 * 512));
                else copy_blocked<4><<<148 * c, 512>>>(in, out, total / (4 * 512));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            printf("persistent blocked K=%d T=512 ctas/sm=%d: %.3f ms %.0f GB/s\n", rep ? 4 : 8, c, best, 32.0 * total / best * 1e-6);
        }
    }
    unsigned long long* counter; CK(cudaMalloc(&counter, 8));
    for (int c = 1; c <= 4; c *= 2) {
        for (int rep = 0; rep < 2; rep++) {
            float best = 1e9;
            for (int r = 0; r < 6; r++) {
                CK(cudaMemsetAsync(counter, 0, 8));
                CK(cudaEventRecord(e0));
                if (rep == 0) copy_dynamic<8><<<148 * c, 512>>>(in, out, total / (8 * 512), counter);
                else copy_dynamic<4><<<148 * c, 512>>>(in, out, total / (4 * 512), counter);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            printf("persistent dynamic K=%d T=512 ctas/sm=%d: %.3f ms %.0f GB/s\n", rep ? 4 : 8, c, best, 32.0 * total / best * 1e-6);
        }
    }
    for (unsigned bytes = 16384; bytes <= 65536; bytes *= 2) {
        CK(cudaFuncSetAttribute(copy_tma_oneshot, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(copy_tma_ld_thread_st, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        for (int v = 0; v < 3; v++) {
            float best = 1e9;
            for (int r = 0; r < 6; r++) {
                CK(cudaEventRecord(e0));
                if (v == 0) copy_tma_oneshot<<<(unsigned)(total * 16 / bytes), 32, bytes>>>((const char*)in, (char*)out, bytes);
                else copy_tma_ld_thread_st<<<(unsigned)(total * 16 / bytes), v == 1 ? 256 : 512, bytes>>>((const char*)in, out, bytes);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("one-shot TMA %s tile=%u B: %.3f ms %.0f GB/s\n", v == 0 ? "load + TMA store      " : v == 1 ? "load + 256-thread store" : "load + 512-thread store", bytes, best, 32.0 * total / best * 1e-6);
        }
    }
    return 0;
}
