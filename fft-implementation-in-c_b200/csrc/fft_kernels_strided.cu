// First / middle pass variants of multi-pass plans: P points of C adjacent columns, strided rows.
// TRIV = true: first pass of a transform (stages 1..LR0 use exact constants); false: middle pass.
#include "fft_catalog.h"
namespace fftb200 {
typedef TileCfg<6, 4, 4, 1, MODE_STRIDED, true, 3, 3, 0, 0, 8, 4> F6;
typedef TileCfg<7, 4, 4, 1, MODE_STRIDED, true, 3, 4, 0, 0, 4, 4> F7;
typedef TileCfg<8, 4, 4, 1, MODE_STRIDED, true, 4, 4, 0, 0, 2, 4> F8;
typedef TileCfg<9, 3, 4, 1, MODE_STRIDED, true, 3, 3, 3, 0, 2, 3> F9;
typedef TileCfg<6, 4, 4, 1, MODE_STRIDED, false, 3, 3, 0, 0, 8, 4> M6;
typedef TileCfg<7, 4, 4, 1, MODE_STRIDED, false, 3, 4, 0, 0, 4, 4> M7;
typedef TileCfg<8, 4, 4, 1, MODE_STRIDED, false, 4, 4, 0, 0, 2, 4> M8;
typedef TileCfg<9, 3, 4, 1, MODE_STRIDED, false, 3, 3, 3, 0, 2, 3> M9;

const KernelInfo* kernels_strided(int* count) {
    static KernelInfo tab[] = {make_info<F6>(), make_info<F7>(), make_info<F8>(), make_info<F9>(),
                               make_info<M6>(), make_info<M7>(), make_info<M8>(), make_info<M9>()};
    *count = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
}  // namespace fftb200
