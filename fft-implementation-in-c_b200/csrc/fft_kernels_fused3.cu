// Fused two-pass kernels (fft_fused.cuh), part 3 of 4: (log2 M, log2 R) pairs compiled in this unit.
#include "fft_fused.cuh"
namespace fftb200 {
#define FUSED_PAIRS(X) X(10,9) X(10,10)
const void* fused_func_3(int lm, int lr, int inverse) {
#define X(A, B) if (lm == A && lr == B) return inverse ? (const void*)fft_fused_kernel<A, B, true> : (const void*)fft_fused_kernel<A, B, false>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
const void* fused_c2r_func_3(int lm, int lr, int herm) {
#define X(A, B) if (lm == A && lr == B) return herm ? (const void*)fft_fused_kernel<A, B, true, false, false, true, true> : (const void*)fft_fused_kernel<A, B, true, false, false, true>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
const void* fused_r2c_func_3(int lm, int lr, int herm) {
#define X(A, B) if (lm == A && lr == B) return herm ? (const void*)fft_fused_kernel<A, B, false, false, true, false, true> : (const void*)fft_fused_kernel<A, B, false, false, true>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
const void* fused_blue_func_3(int lm, int lr, int kind) {
#define X(A, B) if (lm == A && lr == B) return kind == FUSED_BLUE_FWD ? (const void*)fft_fused_kernel<A, B, false, false, false, false, false, FUSED_BLUE_FWD> \
                                                                        : (const void*)fft_fused_kernel<A, B, true, false, false, false, false, FUSED_BLUE_INV>;
    FUSED_PAIRS(X)
#undef X
    return nullptr;
}
}  // namespace fftb200
