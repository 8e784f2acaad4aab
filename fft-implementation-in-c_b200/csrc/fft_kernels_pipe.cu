// Persistent TMA-fed whole-transform kernels, N = 512 .. 4096 (fft_pipe.cuh).
#include "fft_catalog.h"
#include "fft_pipe.cuh"
namespace fftb200 {

template <int LOGN>
static void launch_pipe_t(const PipeArgs& a, int grid, cudaStream_t s) {
    fft_pipe_kernel<LOGN><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
}

const void* pipe_func(int logn) {
    switch (logn) {
        case 9: return (const void*)fft_pipe_kernel<9>;
        case 10: return (const void*)fft_pipe_kernel<10>;
        case 11: return (const void*)fft_pipe_kernel<11>;
        case 12: return (const void*)fft_pipe_kernel<12>;
    }
    return nullptr;
}

void launch_pipe(int logn, const PipeArgs& a, int grid, cudaStream_t s) {
    switch (logn) {
        case 9: launch_pipe_t<9>(a, grid, s); break;
        case 10: launch_pipe_t<10>(a, grid, s); break;
        case 11: launch_pipe_t<11>(a, grid, s); break;
        case 12: launch_pipe_t<12>(a, grid, s); break;
    }
}
}  // namespace fftb200
