// Fused two-pass kernels (fft_fused.cuh), part 0 of 4: (log2 M, log2 R) pairs compiled in this unit.
#include "fft_fused.cuh"
namespace fftb200 {
#define FUSED_PAIRS(X) X(7,6) X(7,7)
const void* fused_func_0(int lm, int lr, int inverse) {
#define X(A, B) if (lm == A && lr == B) return inverse ? (const void*)fft_fused_kernel<A, B, true> : (const void*)fft_fused_kernel<A, B, false>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
bool launch_fused_0(int lm, int lr, const FusedArgs& a, const CUtensorMap* tm, int grid, cudaStream_t s) {
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
}  // namespace fftb200
