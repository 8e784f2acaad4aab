// stridebench.cu - how fast can 64 KB tiles of W adjacent columns x H rows (W * H = 4096 complex doubles) of a row-major
// [H][ROWLEN] matrix move HBM -> SM -> HBM? (development tool: decides whether a 12 + 12 split of N = 2^24 with a one-visit strided
// first pass is viable.)  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/stridebench tools/stridebench.cu -lcuda
//   thread mode: every thread moves 8 elements with 128-bit loads / stores (rows of W * 16 bytes)
//   tma mode:    3 CTAs per SM, each a 64 KB buffer: tensor-box loads (box W x 256 rows) -> wait -> tensor-box stores -> wait
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
typedef double2 cd;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int LW>
__global__ void __launch_bounds__(512) thread_copy(const cd* __restrict__ in, cd* __restrict__ out, int log_rowlen, long long ntiles) {
    constexpr int W = 1 << LW, H = 4096 >> LW;
    const long long tiles_per_mat = (1LL << log_rowlen) >> LW;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long mat = tile / tiles_per_mat, cb = tile % tiles_per_mat;
        const size_t base = ((size_t)mat * H << log_rowlen) + (cb << LW);
        cd v[8];
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = threadIdx.x + 512 * i, r = e >> LW, c = e & (W - 1);
            v[i] = in[base + ((size_t)r << log_rowlen) + c];
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int e = threadIdx.x + 512 * i, r = e >> LW, c = e & (W - 1);
            out[base + ((size_t)r << log_rowlen) + c] = v[i];
        }
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) tma_copy(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int lw, int log_rowlen, long long ntiles) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    if (threadIdx.x != 0) return;
    const int W = 1 << lw, H = 4096 >> lw;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long tiles_per_mat = (1LL << log_rowlen) >> lw;
    uint32_t phase = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long mat = tile / tiles_per_mat, cb = tile % tiles_per_mat;
        const int c0 = (int)(cb << lw) * 2;   // doubles
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(65536u) : "memory");
        for (int r0 = 0; r0 < H; r0 += 256)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                             smem_u32(smem + (size_t)r0 * W * 16)),
                         "l"(&tin), "r"(c0), "r"((int)(mat * H + r0)), "r"(smem_u32(&bar))
                         : "memory");
        asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra W_%=;\n}\n" ::"r"(smem_u32(&bar)), "r"(phase) : "memory");
        phase ^= 1;
        for (int r0 = 0; r0 < H; r0 += 256)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&tout), "r"(c0), "r"((int)(mat * H + r0)),
                         "r"(smem_u32(smem + (size_t)r0 * W * 16))
                         : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int log_total = 28;   // 2^28 points = 4 GiB each way
    cd *in, *out;
    CK(cudaMalloc(&in, sizeof(cd) << log_total));
    CK(cudaMalloc(&out, sizeof(cd) << log_total));
    CK(cudaMemset(in, 1, sizeof(cd) << log_total));
    CK(cudaMemset(out, 0, sizeof(cd) << log_total));
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const long long ntiles = 1LL << (log_total - 12);
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    EncFn enc = (EncFn)f;
    CK(cudaFuncSetAttribute(tma_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    // contiguous reference: W = 4096 columns (a tile is one 64 KB row)
    for (int lw = 0; lw <= 4; lw++) {
        const int W = 1 << lw, H = 4096 >> lw;
        for (int log_rowlen = 12; log_rowlen <= 14; log_rowlen++) {   // 2^24 = 4096 x 4096 (lw 0), 2048 x 8192 (lw 1), ...
            if (log_rowlen != 12 + lw && log_rowlen != 12) continue;
            for (int cps = 2; cps <= 4; cps += 2) {
                float best = 1e9;
                for (int rep = 0; rep < 5; rep++) {
                    CK(cudaEventRecord(e0));
                    switch (lw) {
                        case 0: thread_copy<0><<<sms * cps, 512>>>(in, out, log_rowlen, ntiles); break;
                        case 1: thread_copy<1><<<sms * cps, 512>>>(in, out, log_rowlen, ntiles); break;
                        case 2: thread_copy<2><<<sms * cps, 512>>>(in, out, log_rowlen, ntiles); break;
                        case 3: thread_copy<3><<<sms * cps, 512>>>(in, out, log_rowlen, ntiles); break;
                        case 4: thread_copy<4><<<sms * cps, 512>>>(in, out, log_rowlen, ntiles); break;
                    }
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                printf("threads W=%2d H=%4d rowlen=2^%d ctas/sm=%d: %.3f ms  %.0f GB/s (r+w)\n", W, H, log_rowlen, cps, best, 2.0 * (16.0 * (1LL << log_total)) / best * 1e-6);
            }
            // TMA
            CUtensorMap tm[2];
            for (int i = 0; i < 2; i++) {
                const cuuint64_t gdim[2] = {(cuuint64_t)2 << log_rowlen, (cuuint64_t)1 << (log_total - log_rowlen)};
                const cuuint64_t gstr[1] = {(cuuint64_t)16 << log_rowlen};
                const cuuint32_t box[2] = {(cuuint32_t)2 * W, (cuuint32_t)(H < 256 ? H : 256)};
                const cuuint32_t estr[2] = {1, 1};
                CUresult r = enc(&tm[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, i ? (void*)out : (void*)in, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
            }
            float best = 1e9;
            for (int rep = 0; rep < 5; rep++) {
                CK(cudaEventRecord(e0));
                tma_copy<<<sms * 3, 32, 65536>>>(tm[0], tm[1], lw, log_rowlen, ntiles);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            CK(cudaGetLastError());
            printf("tma     W=%2d H=%4d rowlen=2^%d 3 ctas/sm:  %.3f ms  %.0f GB/s (r+w)\n", W, H, log_rowlen, best, 2.0 * (16.0 * (1LL << log_total)) / best * 1e-6);
        }
    }
    return 0;
}
