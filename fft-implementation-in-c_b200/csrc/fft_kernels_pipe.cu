// Persistent TMA-fed whole-transform kernels, N = 512 .. 4096 (fft_pipe.cuh).
#include "fft_catalog.h"
#include "fft_pipe.cuh"
#include "fft_pipe13.cuh"
#include "fft_pipe13t.cuh"
namespace fftb200 {

template <int LOGN, bool INV>
static void launch_pipe_t(const PipeArgs& a, int grid, cudaStream_t s) {
    fft_pipe_kernel<LOGN, INV><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
}

#define PIPE_CASES(X) X(9) X(10) X(11) X(12)

// real-input / real-output variants (fft_plan_r2c_1d / fft_plan_c2r_1d for N = 512 .. 4096)
const void* pipe_real_func(int logn, int kind) {
    switch (logn) {
#define X(L) case L: return kind == PIPE_R2C ? (const void*)fft_pipe_kernel<L, false, PIPE_R2C> : kind == PIPE_C2R ? (const void*)fft_pipe_kernel<L, true, PIPE_C2R> \
                          : kind == PIPE_BLUE_FWD ? (const void*)fft_pipe_kernel<L, false, PIPE_BLUE_FWD> : (const void*)fft_pipe_kernel<L, true, PIPE_BLUE_INV>;
        PIPE_CASES(X)
#undef X
    }
    return nullptr;
}
cudaError_t launch_pipe_real(int logn, int kind, const PipeArgs& a, int grid, cudaStream_t s) {
    switch (logn) {
#define X(L) case L: if (kind == PIPE_R2C) fft_pipe_kernel<L, false, PIPE_R2C><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a); \
                     else if (kind == PIPE_C2R) fft_pipe_kernel<L, true, PIPE_C2R><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a); \
                     else if (kind == PIPE_BLUE_FWD) fft_pipe_kernel<L, false, PIPE_BLUE_FWD><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a); \
                     else fft_pipe_kernel<L, true, PIPE_BLUE_INV><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a); break;
        PIPE_CASES(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

const void* pipe_func(int logn, int inverse) {
    switch (logn) {
#define X(L) case L: return inverse ? (const void*)fft_pipe_kernel<L, true> : (const void*)fft_pipe_kernel<L, false>;
        PIPE_CASES(X)
#undef X
    }
    return nullptr;
}

cudaError_t launch_pipe(int logn, const PipeArgs& a, int grid, cudaStream_t s) {
    switch (logn) {
#define X(L) case L: if (a.inverse) launch_pipe_t<L, true>(a, grid, s); else launch_pipe_t<L, false>(a, grid, s); break;
        PIPE_CASES(X)
#undef X
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// N = 8192: two 4096-point halves per transform (fft_pipe13.cuh)
const void* pipe13_func(int inverse) {
    return inverse ? (const void*)fft_pipe13_kernel<true> : (const void*)fft_pipe13_kernel<false>;
}
cudaError_t launch_pipe13(const PipeArgs& a, int grid, cudaStream_t s) {
    if (a.inverse) fft_pipe13_kernel<true><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
    else fft_pipe13_kernel<false><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
    return cudaGetLastError();
}
// N = 8192 with tensor memory as the parking space (fft_pipe13t.cuh)
const void* pipe13t_func(int inverse) {
    return inverse ? (const void*)fft_pipe13t_kernel<true> : (const void*)fft_pipe13t_kernel<false>;
}
cudaError_t launch_pipe13t(const PipeArgs& a, int grid, cudaStream_t s) {
    if (a.inverse) fft_pipe13t_kernel<true><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
    else fft_pipe13t_kernel<false><<<grid, 2 * PIPE_GROUP, PIPE_SMEM, s>>>(a);
    return cudaGetLastError();
}
}  // namespace fftb200
