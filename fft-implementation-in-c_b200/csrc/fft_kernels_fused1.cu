// Fused two-pass kernels (fft_fused.cuh), part 1 of 4: (log2 M, log2 R) pairs compiled in this unit.
#include "fft_fused.cuh"
namespace fftb200 {
#define FUSED_PAIRS(X) X(8,7) X(8,8)
const void* fused_func_1(int lm, int lr, int inverse) {
#define X(A, B) if (lm == A && lr == B) return inverse ? (const void*)fft_fused_kernel<A, B, true> : (const void*)fft_fused_kernel<A, B, false>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
bool launch_fused_1(int lm, int lr, const FusedArgs& a, const CUtensorMap* tm, int grid, cudaStream_t s) {
#define X(A, B)                                                                                                  \
    if (lm == A && lr == B) {                                                                                    \
        if (a.inverse) fft_fused_kernel<A, B, true><<<grid, FUSED_THREADS, FUSED_SMEM, s>>>(a, tm[0], tm[1], tm[2]);  \
        else fft_fused_kernel<A, B, false><<<grid, FUSED_THREADS, FUSED_SMEM, s>>>(a, tm[0], tm[1], tm[2]);           \
        return true;                                                                                             \
    }
    FUSED_PAIRS(X)
#undef X
    return false;
}
const void* fused_cols_func(int inverse) {
    return inverse ? (const void*)fft_fused_kernel<8, 8, true, true> : (const void*)fft_fused_kernel<8, 8, false, true>;
}
void launch_fused_cols(const FusedArgs& a, const CUtensorMap* tm, int grid, cudaStream_t s) {
    if (a.inverse) fft_fused_kernel<8, 8, true, true><<<grid, FUSED_THREADS, FUSED_SMEM, s>>>(a, tm[0], tm[1], tm[2]);
    else fft_fused_kernel<8, 8, false, true><<<grid, FUSED_THREADS, FUSED_SMEM, s>>>(a, tm[0], tm[1], tm[2]);
}
}  // namespace fftb200
