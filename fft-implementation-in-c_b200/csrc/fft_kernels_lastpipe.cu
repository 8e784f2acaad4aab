// Persistent TMA-ring last pass of multi-pass plans, P = 32 .. 512 points (fft_lastpipe.cuh).
#include "fft_lastpipe.cuh"
namespace fftb200 {
#define LASTPIPE_CASES(X) X(5) X(6) X(7) X(8) X(9)
const void* lastpipe_func(int lr, int inverse, int derive) {
    switch (lr) {
#define X(L) case L: return derive ? (inverse ? (const void*)fft_lastpipe_kernel<L, true, true> : (const void*)fft_lastpipe_kernel<L, false, true>) \
                                   : (inverse ? (const void*)fft_lastpipe_kernel<L, true, false> : (const void*)fft_lastpipe_kernel<L, false, false>);
        LASTPIPE_CASES(X)
#undef X
    }
    return nullptr;
}
cudaError_t launch_lastpipe(int lr, const LastPipeArgs& a, int grid, cudaStream_t s) {
    const void* f = lastpipe_func(lr, a.inverse, a.derive);
    if (!f) return cudaErrorInvalidValue;
    void* args[1] = {(void*)&a};
    return cudaLaunchKernel(f, dim3((unsigned)grid), dim3(2 * PIPE_GROUP), args, LASTPIPE_SMEM, s);
}
}  // namespace fftb200
