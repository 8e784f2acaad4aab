// fusedemu.cu - the data-movement skeleton of fft_fused_kernel without the arithmetic: NB ring buffers of TILE bytes per CTA,
// one manager lane per buffer (load -> wait for the "compute group" -> store -> next load chasing the store), G compute groups
// that hold a tile for SPIN cycles. Tiles alternate between the pass-A shape (streaming input -> L2-resident scratch ring) and
// the pass-B shape (scratch ring -> streaming output). Answers: what do smaller tiles / more buffers buy at a fixed 192 KB of
// shared memory and a fixed compute capacity (G groups x TILE bytes per SPIN cycles)? Development tool (profiles/r02_microbench.md).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fusedemu tools/fusedemu.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(s32(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ uint64_t pol_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t pol_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int x, int y, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(s32(dst)), "l"(tm), "r"(x), "r"(y), "r"(s32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, int x, int y, const void* src, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;"
                 ::"l"(tm), "r"(x), "r"(y), "r"(s32(src)), "l"(pol) : "memory");
}

struct Args {
    int lm, lr, boxes;     // boxes = 1: the real kernel's tensor boxes (A load / A store / B store), 64 KB tiles in 4 quarters
    long long nbatch;
    char* out; const char* in; char* ring;
    long long ntiles;      // tiles of the input (= of the output)
    long long ring_tiles;  // tiles of the scratch ring
    int tile_bytes, nb, groups, spin, parts, lag;
};

__global__ void __launch_bounds__(1024, 1) emu_kernel(const Args a, const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_sc, const __grid_constant__ CUtensorMap tm_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = (uint64_t*)(smem + (size_t)a.nb * a.tile_bytes);
    uint64_t* staged = full + a.nb;
    if (threadIdx.x == 0) {
        for (int b = 0; b < a.nb; b++) { mbar_init(&full[b], 1); mbar_init(&staged[b], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x & 31) != 0) return;
    const int w = threadIdx.x >> 5;
    // work list of this CTA: k even = pass-A tile, k odd = pass-B tile (of an earlier group: the ring is simply re-read `lag` A tiles later)
    const long long mine = blockIdx.x < a.ntiles ? (a.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long total = 2 * mine;
    if (w >= a.groups) {
        const int b = w - a.groups;
        if (b >= a.nb) return;
        unsigned char* buf = smem + (size_t)b * a.tile_bytes;
        const uint64_t pf = pol_first(), pl = pol_last();
        const uint32_t part = a.tile_bytes / a.parts;
        auto load = [&](long long k, bool chase) {
            const long long j = k >> 1, t = blockIdx.x + j * gridDim.x;
            const char* src = (k & 1) ? a.ring + ((t + a.ring_tiles - a.lag * (long long)gridDim.x % a.ring_tiles) % a.ring_tiles) * a.tile_bytes : a.in + t * a.tile_bytes;
            mbar_expect(&full[b], a.tile_bytes);
            if (a.boxes && !(k & 1)) {
                // pass-A tile t: transform tr = t >> (L - 12), column block blk: box C x M/4 per quarter out of the [nbatch * M][R] view
                const int L = a.lm + a.lr, lc = 12 - a.lm;
                const long long tr = t >> (L - 12); const int blk = (int)(t & ((1 << (L - 12)) - 1));
                for (int q = 0; q < 4; q++) {
                    if (chase) {
                        if (q == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                        else if (q == 1) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                        else if (q == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    }
                    tma_load_2d(buf + q * part, &tm_in, 2 * (blk << lc), (int)((tr << a.lm) + q * (1024 >> lc)), &full[b], pf);
                }
                return;
            }
            for (int q = 0; q < a.parts; q++) {
                if (chase) {   // part q only after the store of part q has been read out of the buffer
                    const int left = a.parts - 1 - q;
                    if (left >= 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
                    else if (left == 2) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
                    else if (left == 1) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                }
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                             ::"r"(s32(buf + q * part)), "l"(src + q * part), "r"(part), "r"(s32(&full[b])), "l"((k & 1) ? pl : pf) : "memory");
            }
        };
        if (b >= total) return;
        load(b, false);
        int n = 0;
        for (long long k = b; k < total; k += a.nb, n++) {
            mbar_wait(&staged[b], n & 1);
            const long long j = k >> 1, t = blockIdx.x + j * gridDim.x;
            char* dst = (k & 1) ? a.out + t * a.tile_bytes : a.ring + (t % a.ring_tiles) * a.tile_bytes;
            if (a.boxes) {
                const int L = a.lm + a.lr;
                const int blk = (int)(t & ((1 << (L - 12)) - 1));
                if (!(k & 1)) {   // pass-A store into the scratch ring: same box shape as the load, ring slot of the transform
                    const int lc = 12 - a.lm;
                    const long long trl = (t >> (L - 12)) % (a.ring_tiles >> (L - 12));
                    for (int q = 0; q < 4; q++) {
                        tma_store_2d(&tm_sc, 2 * (blk << lc), (int)((trl << a.lm) + q * (1024 >> lc)), buf + q * part, pl);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                } else {          // pass-B store: box C2 x R/4 per quarter into the [nbatch * R][M] view
                    const int lc2 = 12 - a.lr;
                    const long long tr = t >> (L - 12);
                    for (int q = 0; q < 4; q++) {
                        tma_store_2d(&tm_out, 2 * (blk << lc2), (int)((tr << a.lr) + q * (1024 >> lc2)), buf + q * part, pf);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else
            for (int q = 0; q < a.parts; q++) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + q * part), "r"(s32(buf + q * part)), "r"(part), "l"((k & 1) ? pf : pl) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (k + a.nb < total) load(k + a.nb, true);
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }
    // compute group w: tiles k = w, w + G, ...
    for (long long k = w; k < total; k += a.groups) {
        const int b = (int)(k % a.nb), n = (int)(k / a.nb);
        if (n >= 1) mbar_wait(&staged[b], (n - 1) & 1);
        mbar_wait(&full[b], n & 1);
        const long long t0 = clock64();
        while (clock64() - t0 < a.spin) {}
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive(&staged[b]);
    }
}

int main(int argc, char** argv) {
    const long long big = 4LL << 30;
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    char *in, *out, *ring;
    CK(cudaMalloc(&in, big)); CK(cudaMalloc(&out, big)); CK(cudaMalloc(&ring, 256ll << 20));
    CK(cudaMemset(in, 0, big)); CK(cudaMemset(out, 0, big)); CK(cudaMemset(ring, 0, 256ll << 20));
    CK(cudaFuncSetAttribute(emu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    struct Cfg { int kb, nb, groups; };
    const Cfg cfgs[] = {{64, 3, 2}};
    const int ring_mb = argc > 1 ? atoi(argv[1]) : 32;
    typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fptr = nullptr; cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fptr, cudaEnableDefault, &qres));
    EncodeTiledFn enc = (EncodeTiledFn)fptr;
    CUtensorMap tmz[3];
    memset(tmz, 0, sizeof(tmz));
    {
        const int pairs[][2] = {{7, 7}, {8, 8}, {9, 9}, {10, 10}};
        for (auto& pr : pairs)
            for (int spin64 : {0, 5830}) {
                const int lm = pr[0], lr = pr[1], L = lm + lr;
                const long long nbatch = (big / 16) >> L;
                Args a; memset(&a, 0, sizeof(a));
                a.in = in; a.out = out; a.ring = ring; a.tile_bytes = 65536; a.nb = 3; a.groups = 2; a.parts = 4; a.spin = spin64; a.lag = 3;
                a.ntiles = big / 65536; a.ring_tiles = ((long long)ring_mb << 20) / 65536; a.lm = lm; a.lr = lr; a.boxes = 1; a.nbatch = nbatch;
                const long long ring_tr = a.ring_tiles >> (L - 12);
                for (int i = 0; i < 3; i++) {
                    const int lcols = i == 2 ? lm : lr, lrows = i == 2 ? lr : lm;
                    const long long ntr = i == 1 ? (ring_tr > 0 ? ring_tr : 1) : nbatch;
                    void* base = i == 0 ? (void*)in : i == 1 ? (void*)ring : (void*)out;
                    const cuuint64_t gdim[2] = {(cuuint64_t)2 << lcols, (cuuint64_t)ntr << lrows};
                    const cuuint64_t gstr[1] = {(cuuint64_t)16 << lcols};
                    const cuuint32_t box[2] = {(cuuint32_t)2 << (12 - lrows), (cuuint32_t)1 << (lrows - 2)};
                    const cuuint32_t estr[2] = {1, 1};
                    const int promo = 12 - lm >= 4 ? 2 : 12 - lm >= 3 ? 1 : 0;
                    CUresult r = enc(&tmz[i], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                     i == 0 ? (promo >= 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE) : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                    if (r != CUDA_SUCCESS) { printf("encode %d failed %d\n", i, (int)r); return 1; }
                }
                if (ring_tr < 1) { printf("ring too small for 2^%d\n", L); continue; }
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    emu_kernel<<<sms, 32 * 5, 3 * 65536 + 16 * 3 + 64>>>(a, tmz[0], tmz[1], tmz[2]);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                CK(cudaGetLastError());
                printf("BOXES 2^%d = %d + %d (A rows %d B, B rows %d B), 64 KB x 3 bufs, 2 groups, hold %d: %.3f ms per 2^28 points, strict %.0f GB/s\n",
                       L, lm, lr, 16 << (12 - lm), 16 << (12 - lr), spin64, best, 2.0 * big / best * 1e-6);
            }
    }
    for (const Cfg& c : cfgs)
        for (int parts : {1, 4})
            for (int spin64 : {0, 3000, 5830, 8000}) {
                Args a; memset(&a, 0, sizeof(a));
                a.in = in; a.out = out; a.ring = ring;
                a.tile_bytes = c.kb << 10; a.nb = c.nb; a.groups = c.groups; a.parts = parts;
                a.ntiles = big / a.tile_bytes; a.ring_tiles = ((long long)ring_mb << 20) / a.tile_bytes;
                a.spin = (int)((long long)spin64 * c.kb / 64 * c.groups / 2);   // 512 compute threads: a group of 512 / groups threads holds its tile this long
                a.lag = 3;
                const int threads = 32 * (c.groups + c.nb);
                const size_t smem = (size_t)c.nb * a.tile_bytes + 16 * c.nb + 64;
                float best = 1e9;
                for (int rep = 0; rep < 3; rep++) {
                    CK(cudaEventRecord(e0));
                    emu_kernel<<<sms, threads, smem>>>(a, tmz[0], tmz[1], tmz[2]);
                    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                    if (ms < best) best = ms;
                }
                CK(cudaGetLastError());
                printf("tile %2d KB x %2d bufs, %d groups, %d parts, compute %4d cyc per 64 KB and 256 threads (hold %d): %.3f ms per 2^28 points, strict %.0f GB/s\n",
                       c.kb, c.nb, c.groups, parts, spin64, a.spin, best, 2.0 * big / best * 1e-6);
            }
    return 0;
}
