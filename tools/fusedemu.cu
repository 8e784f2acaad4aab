// fusedemu.cu - the data-movement skeleton of fft_fused_kernel without the arithmetic: NB ring buffers of TILE bytes per CTA,
// one manager lane per buffer (load -> wait for the "compute group" -> store -> next load chasing the store), G compute groups
// that hold a tile for SPIN cycles. Tiles alternate between the pass-A shape (streaming input -> L2-resident scratch ring) and
// the pass-B shape (scratch ring -> streaming output). Answers: what do smaller tiles / more buffers buy at a fixed 192 KB of
// shared memory and a fixed compute capacity (G groups x TILE bytes per SPIN cycles)? Development tool (profiles/r02_microbench.md).
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fusedemu tools/fusedemu.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>

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

struct Args {
    char* out; const char* in; char* ring;
    long long ntiles;      // tiles of the input (= of the output)
    long long ring_tiles;  // tiles of the scratch ring
    int tile_bytes, nb, groups, spin, parts, lag;
};

__global__ void __launch_bounds__(1024, 1) emu_kernel(const Args a) {
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
    const Cfg cfgs[] = {{64, 3, 2}, {32, 6, 4}, {32, 6, 2}, {32, 5, 4}, {32, 4, 4}, {16, 12, 8}, {16, 12, 4}, {16, 12, 2}};
    const int ring_mb = argc > 1 ? atoi(argv[1]) : 32;
    for (const Cfg& c : cfgs)
        for (int parts : {1, 4})
            for (int spin64 : {0, 3000, 5830, 8000}) {
                Args a;
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
                    emu_kernel<<<sms, threads, smem>>>(a);
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
