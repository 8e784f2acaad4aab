// dsbench.cu - what a bulk-copy (TMA) push through distributed shared memory costs, alone and next to global TMA traffic.
// Development tool behind the cluster kernel decision (profiles/r02_microbench.md), not product code.
//   what & 1: every CTA of a cluster of C pushes the C - 1 remote 64/C KB chunks of its 64 KB tile into its peers' receive
//             buffers (cp.async.bulk.shared::cluster.shared::cta + complete_tx on the peer's mbarrier), double-buffered
//   what & 2: 64 KB bulk stores shared -> global (2 in flight)
//   what & 4: 64 KB... (16 KB x 2 ring) bulk loads global -> shared
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/dsbench tools/dsbench.cu
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int TILE_B = 65536, LD_B = 16384;
constexpr size_t SMEM = 3 * TILE_B + 2 * LD_B + 256;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, int peer) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(peer)); return r; }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* b, uint32_t ph) {   // acquire at cluster scope: remote arrivals / remote bulk writes
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void remote_arrive(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ void push(uint32_t dst_cl, uint32_t src_cta, uint32_t bytes, uint32_t bar_cl) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cl), "r"(src_cta), "r"(bytes), "r"(bar_cl) : "memory");
}

__global__ void __launch_bounds__(128, 1) ds_kernel(char* gout, const char* gin, long long gtiles, long long* cycles, int C, int iters, int what, int chunk_split) {
    extern __shared__ __align__(128) unsigned char raw[];
    unsigned char* src = raw;                    // 64 KB tile this CTA produced
    unsigned char* rcv = raw + TILE_B;           // 2 x 64 KB receive buffers
    unsigned char* ldb = raw + 3 * TILE_B;       // 2 x 16 KB global load ring
    uint64_t* full = reinterpret_cast<uint64_t*>(raw + 3 * TILE_B + 2 * LD_B);   // [2] chunks arrived
    uint64_t* cons = full + 2;                                                    // [2] peers consumed what I sent
    uint64_t* ldf = cons + 2;                                                     // [2] global loads
    cg::cluster_group cl = cg::this_cluster();
    const int r = (int)cl.block_rank();
    if (threadIdx.x == 0) {
        for (int b = 0; b < 2; b++) { mbar_init(&full[b], 1); mbar_init(&cons[b], C > 1 ? C - 1 : 1); mbar_init(&ldf[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < TILE_B / 8; i += blockDim.x) reinterpret_cast<double*>(src)[i] = i + r;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    cl.sync();
    const long long t0 = clock64();
    const uint32_t chunk = TILE_B / C;
    if (threadIdx.x == 0 && (what & 1) && C > 1) {
        for (int it = 0; it < iters; it++) {
            const int b = it & 1, n = it >> 1;
            if (n >= 1) mbar_wait_cl(&cons[b], (n - 1) & 1);      // the peers have consumed buffer b of iteration it - 2
            mbar_expect(&full[b], chunk * (C - 1));               // my own receive buffer: C - 1 chunks will arrive
            const uint32_t sub = chunk / chunk_split;
            for (int d = 1; d < C; d++) {
                const int peer = (r + d) % C;
                for (int s = 0; s < chunk_split; s++)
                    push(mapa(s32(rcv + b * TILE_B + r * chunk + s * sub), peer), s32(src + peer * chunk + s * sub), sub, mapa(s32(&full[b]), peer));
            }
            mbar_wait_cl(&full[b], n & 1);                        // everything for me has arrived
            for (int d = 1; d < C; d++) remote_arrive(mapa(s32(&cons[b]), (r + d) % C));   // tell the senders
        }
    }
    if (threadIdx.x == 32 && (what & 2)) {
        for (int it = 0; it < iters; it++) {
            const long long t = ((long long)blockIdx.x + (long long)it * gridDim.x) % gtiles;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gout + t * TILE_B), "r"(s32(src)), "r"(TILE_B) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    if (threadIdx.x == 64 && (what & 4)) {
        // 64 KB per iteration as four 16 KB loads through a 2-deep ring
        const int nl = iters * 4;
        for (int k = 0; k < 2 && k < nl; k++) {
            const long long t = ((long long)blockIdx.x * 4 + k + 0) % (gtiles * 4);
            mbar_expect(&ldf[k], LD_B);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ldb + k * LD_B)), "l"(gin + t * LD_B), "r"(LD_B), "r"(s32(&ldf[k])) : "memory");
        }
        for (int k = 0; k < nl; k++) {
            const int b = k & 1;
            mbar_wait(&ldf[b], (k >> 1) & 1);
            if (k + 2 < nl) {
                const long long t = (((long long)blockIdx.x + (long long)((k + 2) >> 2) * gridDim.x) * 4 + ((k + 2) & 3)) % (gtiles * 4);
                mbar_expect(&ldf[b], LD_B);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(ldb + b * LD_B)), "l"(gin + t * LD_B), "r"(LD_B), "r"(s32(&ldf[b])) : "memory");
            }
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main(int argc, char** argv) {
    const long long gbytes = 2LL << 30, gtiles = gbytes / TILE_B;
    char *gin, *gout; long long* cyc;
    CK(cudaMalloc(&gin, gbytes)); CK(cudaMalloc(&gout, gbytes)); CK(cudaMalloc(&cyc, 4096 * 8));
    CK(cudaMemset(gin, 0, gbytes)); CK(cudaMemset(gout, 0, gbytes));
    CK(cudaFuncSetAttribute(ds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    CK(cudaFuncSetAttribute(ds_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    int clk = 0; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    const int iters = 400;
    for (int C : {1, 2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(C * 148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = SMEM;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
        int maxcl = 0;
        CK(cudaOccupancyMaxActiveClusters(&maxcl, ds_kernel, &cfg));
        const int grid = maxcl * C;
        cfg.gridDim = dim3(grid);
        for (int what : {1, 2, 4, 6, 3, 5, 7}) {
            if (C == 1 && (what & 1)) continue;
            for (int split : {1, 4}) {
                if (split > 1 && !(what & 1)) continue;
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    CK(cudaLaunchKernelEx(&cfg, ds_kernel, gout, (const char*)gin, gtiles, cyc, C, iters, what, split));
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                long long h[4096]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
                double mean = 0, mx = 0; for (int i = 0; i < grid; i++) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
                mean /= grid;
                const double per = mean / iters;
                printf("C=%2d SMs=%3d what=%d%s%s%s split=%d: %.0f cycles per 64 KB tile (mean; max %.0f), %.3f ms for %d tiles/SM", C, grid, what,
                       (what & 1) ? " push" : "", (what & 2) ? " gstore" : "", (what & 4) ? " gload" : "", split, per, mx / iters, best, iters);
                if (what & 1) printf(", push %.1f B/clk/SM", (double)TILE_B * (C - 1) / C / per);
                if (what & 2) printf(", gstore %.1f B/clk/SM = %.0f GB/s", TILE_B / per, (double)TILE_B * iters * grid / best * 1e-6);
                if (what & 4) printf(", gload %.1f B/clk/SM = %.0f GB/s", TILE_B / per, (double)TILE_B * iters * grid / best * 1e-6);
                printf("\n");
            }
        }
    }
    return 0;
}
