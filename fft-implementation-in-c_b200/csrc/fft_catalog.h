// fft_catalog.h - table of compiled tile-kernel variants shared between the kernel TUs and the planner.
#pragma once
#include "fft_tile.cuh"

namespace fftb200 {

struct KernelInfo {
    int mode, logp, logc, triv, nt, threads;
    size_t smem;
    const void* func;       // forward instantiation
    const void* func_inv;   // inverse instantiation (same resources; attributes are set on both)
    void (*launch)(const TileArgs&, int grid, cudaStream_t);
};

template <class C>
static void launch_tile(const TileArgs& a, int grid, cudaStream_t s) {
    if (a.inverse) fft_tile_kernel<C, true><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(a);
    else fft_tile_kernel<C, false><<<grid, C::THREADS, C::SMEM_BYTES, s>>>(a);
}

template <class C>
static KernelInfo make_info() {
    KernelInfo k;
    k.mode = C::MODE; k.logp = C::LOGP; k.logc = C::LOGC; k.triv = C::TRIV ? 1 : 0; k.nt = C::NT;
    k.threads = C::THREADS; k.smem = C::SMEM_BYTES;
    k.func = (const void*)fft_tile_kernel<C, false>;
    k.func_inv = (const void*)fft_tile_kernel<C, true>;
    k.launch = launch_tile<C>;
    return k;
}

// each returns a static array; defined in fft_kernels_{contig,strided,last}.cu
const KernelInfo* kernels_contig(int* count);
const KernelInfo* kernels_strided(int* count);
const KernelInfo* kernels_last(int* count);

// persistent TMA-fed kernels (fft_pipe.cuh), defined in fft_kernels_pipe.cu
struct PipeArgs;
const void* pipe_func(int logn, int inverse);                                  // nullptr if no variant
cudaError_t launch_pipe(int logn, const PipeArgs& a, int grid, cudaStream_t s);
const void* pipe_real_func(int logn, int kind);                                // PIPE_R2C / PIPE_C2R variants (fft_pipe.cuh)
cudaError_t launch_pipe_real(int logn, int kind, const PipeArgs& a, int grid, cudaStream_t s);

}  // namespace fftb200
