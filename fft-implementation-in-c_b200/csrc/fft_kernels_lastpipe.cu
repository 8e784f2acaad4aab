// Persistent TMA-ring last pass of multi-pass plans, P = 32 .. 512 points (fft_lastpipe.cuh).
#include "fft_lastpipe.cuh"
namespace fftb200 {
#define LASTPIPE_CASES(X) X(5) X(6) X(7) X(8) X(9)
const void* lastpipe_func(int lr, int inverse) {
    switch (lr) {
#define X(L) case L: return inverse ? (const void*)fft_lastpipe_kernel<L, true> : (const void*)fft_lastpipe_kernel<L, false>;
        LASTPIPE_CASES(X)
#undef X
    }
    return nullptr;
}
cudaError_t launch_lastpipe(int lr, const LastPipeArgs& a, int grid, cudaStream_t s) {
    switch (lr) {
#define X(L) case L: if (a.inverse) fft_lastpipe_kernel<L, true><<<grid, 2 * PIPE_GROUP, LASTPIPE_SMEM, s>>>(a); \
                     else fft_lastpipe_kernel<L, false><<<grid, 2 * PIPE_GROUP, LASTPIPE_SMEM, s>>>(a); break;
        LASTPIPE_CASES(X)
#undef X
    }
    return cudaGetLastError();
}
}  // namespace fftb200
