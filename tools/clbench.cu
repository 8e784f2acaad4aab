// clbench.cu - microbenchmarks behind the cluster kernel (fft_cluster.cuh): what a thread-block cluster of C CTAs can
// pull out of HBM when CTA r gathers column r of a row-major [T][C] array of complex doubles (16-byte elements at a stride
// of 16 C bytes), and what the distributed-shared-memory exchange costs. Development tool, not product code.
//   mode 0: gather through TMA tensor loads, box = 1 element x 256 rows (16 loads per 64 KB tile), 3-stage ring
//   mode 1: gather through 16-byte cp.async (LDGSTS), 3-stage ring
//   mode 2: gather through plain 16-byte loads into registers
//   mode 3: DSMEM read exchange: each CTA reads its 1/C share of every peer's 64 KB tile, 512 contiguous bytes per warp load
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/clbench tools/clbench.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int TILE = 4096, STAGES = 3, THREADS = 512;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void tma_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(s32(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(s32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t mapa_u32(const double2* base, int peer, int elem) {
    uint32_t local = s32(base + elem), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(peer));
    return remote;
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 1) gather_kernel(const __grid_constant__ CUtensorMap tm, const double2* in, double* sink, int C,
                                                            long long ntr) {
    extern __shared__ __align__(128) unsigned char raw[];
    double2* bufs = reinterpret_cast<double2*>(raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(bufs + STAGES * TILE);
    cg::cluster_group cl = cg::this_cluster();
    const int r = cl.block_rank();
    const long long cid = blockIdx.x / C, ncl = gridDim.x / C;
    const int mine = cid < ntr ? (int)((ntr - cid + ncl - 1) / ncl) : 0;
    double acc = 0;
    if (MODE == 0) {
        if (threadIdx.x == 0) {
            for (int b = 0; b < STAGES; b++) mbar_init(&full[b], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        auto issue = [&](int k, int b) {
            const long long tr = cid + (long long)k * ncl;
            mbar_expect(&full[b], TILE * 16);
            for (int i = 0; i < 16; i++) tma_2d(bufs + b * TILE + 256 * i, &tm, 2 * r, (int)(tr * TILE + 256 * i), &full[b]);
        };
        if (threadIdx.x == 0) for (int k = 0; k < STAGES && k < mine; k++) issue(k, k);
        for (int k = 0; k < mine; k++) {
            const int b = k % STAGES;
            mbar_wait(&full[b], (k / STAGES) & 1);
            acc += bufs[b * TILE + threadIdx.x].x;
            __syncthreads();
            if (threadIdx.x == 0 && k + STAGES < mine) issue(k + STAGES, b);
        }
    } else if (MODE == 1) {
        auto issue = [&](int k, int b) {
            const long long tr = cid + (long long)k * ncl;
            const double2* src = in + (tr * TILE) * C + r;
#pragma unroll
            for (int e = 0; e < TILE / THREADS; e++) {
                const int t = threadIdx.x + e * THREADS;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s32(bufs + b * TILE + t)), "l"(src + (long long)t * C) : "memory");
            }
        };
        for (int k = 0; k < STAGES - 1; k++) { if (k < mine) issue(k, k); asm volatile("cp.async.commit_group;" ::: "memory"); }
        for (int k = 0; k < mine; k++) {
            if (k + STAGES - 1 < mine) issue(k + STAGES - 1, (k + STAGES - 1) % STAGES);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
            __syncthreads();
            acc += bufs[(k % STAGES) * TILE + threadIdx.x].x;
            __syncthreads();
        }
    } else {
        for (int k = 0; k < mine; k++) {
            const long long tr = cid + (long long)k * ncl;
            const double2* src = in + (tr * TILE) * C + r;
            double2 v[TILE / THREADS];
#pragma unroll
            for (int e = 0; e < TILE / THREADS; e++) v[e] = __ldg(src + (long long)(threadIdx.x + e * THREADS) * C);
#pragma unroll
            for (int e = 0; e < TILE / THREADS; e++) acc += v[e].x;
        }
    }
    if (acc == 1.2345e-300) sink[blockIdx.x] = acc;
}

__global__ void __launch_bounds__(THREADS, 1) dsmem_kernel(double* sink, long long* cycles, int C, int iters) {
    extern __shared__ __align__(128) unsigned char raw[];
    double2* buf = reinterpret_cast<double2*>(raw);
    cg::cluster_group cl = cg::this_cluster();
    const int r = cl.block_rank();
    for (int i = threadIdx.x; i < TILE; i += THREADS) buf[i] = make_double2(i, r);
    cl.sync();
    const int share = TILE / C;   // this CTA's k range
    double acc = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        // element e of the share for every peer: TILE elements in total, 8 per thread
#pragma unroll
        for (int e = 0; e < TILE / THREADS; e++) {
            const int idx = threadIdx.x + e * THREADS;         // 0 .. 4095
            const int peer = idx / share, k = idx % share;
            double2 v;
            asm volatile("ld.shared::cluster.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(mapa_u32(buf, peer, r * share + ((k + it) & (share - 1)))));
            acc += v.x + v.y;
        }
    }
    const long long t1 = clock64();
    cl.sync();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 1.2345e-300) sink[blockIdx.x] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename K, typename... A>
static cudaError_t launch_cluster(K kern, int grid, int C, size_t smem, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}

int main(int argc, char** argv) {
    const long long total = 1LL << 27;   // complex elements: 2 GiB
    double2* in; double* sink; long long* cyc;
    CK(cudaMalloc(&in, total * 16)); CK(cudaMemset(in, 0, total * 16));
    CK(cudaMalloc(&sink, 4096 * 8)); CK(cudaMalloc(&cyc, 4096 * 8));
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)f;
    const size_t smem = STAGES * TILE * 16 + 64;
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int C : {1, 2, 4, 8, 16}) {
        int maxcl = 0;
        for (int mode = 0; mode < 3; mode++) {
            if (getenv("CLBENCH_DSMEM_ONLY") && mode != 0) continue;
            auto kern = mode == 0 ? gather_kernel<0> : mode == 1 ? gather_kernel<1> : gather_kernel<2>;
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (C > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(C * 148); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; cfg.attrs = at; cfg.numAttrs = 1;
            CK(cudaOccupancyMaxActiveClusters(&maxcl, kern, &cfg));
            CUtensorMap tm;
            const cuuint64_t gdim[2] = {(cuuint64_t)2 * C, (cuuint64_t)(total / C)};
            const cuuint64_t gstr[1] = {(cuuint64_t)16 * C};
            const cuuint32_t box[2] = {2, 256};
            const cuuint32_t estr[2] = {1, 1};
            CUresult rr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)rr); continue; }
            const long long ntr = total / ((long long)TILE * C);
            const int grid = maxcl * C;
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                CK(cudaEventRecord(e0));
                CK(launch_cluster(kern, grid, C, smem, tm, (const double2*)in, sink, C, ntr));
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("C=%2d clusters=%3d (SMs %3d) mode %d (%s): %.3f ms  %.0f GB/s read\n", C, maxcl, grid,
                   mode, mode == 0 ? "TMA box 16B x 256" : mode == 1 ? "cp.async 16B" : "ldg 16B", best, total * 16.0 / best * 1e-6);
        }
        if (C >= 2) {
            CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            if (C > 8) CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
            const int grid = maxcl * C, iters = 200;
            CK(cudaEventRecord(e0));
            CK(launch_cluster(dsmem_kernel, grid, C, smem, sink, cyc, C, iters));
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            long long h[4096]; CK(cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost));
            double mean = 0, mx = 0; for (int i = 0; i < grid; i++) { mean += h[i]; if (h[i] > mx) mx = h[i]; }
            mean /= grid;
            printf("C=%2d DSMEM exchange: %.0f cycles per 64 KB tile (mean), %.0f (max)  = %.1f B/clk/SM; %.3f ms for %d tiles per SM\n", C, mean / iters,
                   mx / iters, 65536.0 * iters / mean, ms, iters);
        }
    }
    return 0;
}
