// Last-pass variants of multi-pass plans: contiguous P-point columns in, natural-order rows out.
#include "fft_catalog.h"
namespace fftb200 {
// L5: the tail of N = 2^21 behind the fused column-mode head (2^16 x 32); 64 columns x 32 points per tile, 128 threads
typedef TileCfg<5, 6, 4, 1, MODE_LAST, false, 2, 3, 0, 0, 4, 4> L5;
typedef TileCfg<6, 4, 4, 1, MODE_LAST, false, 3, 3, 0, 0, 8, 4> L6;
typedef TileCfg<7, 4, 4, 1, MODE_LAST, false, 3, 4, 0, 0, 4, 4> L7;
typedef TileCfg<8, 4, 4, 1, MODE_LAST, false, 4, 4, 0, 0, 2, 4> L8;
typedef TileCfg<9, 3, 4, 1, MODE_LAST, false, 3, 3, 3, 0, 2, 3> L9;

const KernelInfo* kernels_last(int* count) {
    static KernelInfo tab[] = {make_info<L5>(), make_info<L6>(), make_info<L7>(), make_info<L8>(), make_info<L9>()};
    *count = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
}  // namespace fftb200
