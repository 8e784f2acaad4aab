// Whole-transform-per-CTA variants, N = 2 .. 8192 (register + shared-memory Stockham, unit stride).
#include "fft_catalog.h"
namespace fftb200 {
//                 LOGP C LOGE  NT  MODE         TRIV  radices      MINB PSH
typedef TileCfg<1, 0, 1, 128, MODE_CONTIG, true, 1, 0, 0, 0, 4, 4> C1;
typedef TileCfg<2, 0, 2, 128, MODE_CONTIG, true, 2, 0, 0, 0, 4, 4> C2;
typedef TileCfg<3, 0, 3, 128, MODE_CONTIG, true, 3, 0, 0, 0, 4, 4> C3;
typedef TileCfg<4, 0, 4, 64, MODE_CONTIG, true, 4, 0, 0, 0, 4, 4> C4;
typedef TileCfg<5, 0, 3, 32, MODE_CONTIG, true, 2, 3, 0, 0, 4, 4> C5;
typedef TileCfg<6, 0, 3, 16, MODE_CONTIG, true, 3, 3, 0, 0, 4, 3> C6;
typedef TileCfg<7, 0, 4, 16, MODE_CONTIG, true, 3, 4, 0, 0, 4, 4> C7;
typedef TileCfg<8, 0, 4, 8, MODE_CONTIG, true, 4, 4, 0, 0, 4, 4> C8;
typedef TileCfg<9, 0, 3, 4, MODE_CONTIG, true, 3, 3, 3, 0, 2, 3> C9;
typedef TileCfg<10, 0, 4, 4, MODE_CONTIG, true, 2, 4, 4, 0, 2, 4> C10;
typedef TileCfg<11, 0, 4, 2, MODE_CONTIG, true, 3, 4, 4, 0, 2, 4> C11;
typedef TileCfg<12, 0, 4, 1, MODE_CONTIG, true, 4, 4, 4, 0, 2, 4> C12;
typedef TileCfg<13, 0, 4, 1, MODE_CONTIG, true, 3, 3, 3, 4, 1, 4> C13;

const KernelInfo* kernels_contig(int* count) {
    static KernelInfo tab[] = {make_info<C1>(), make_info<C2>(), make_info<C3>(), make_info<C4>(),
                               make_info<C5>(), make_info<C6>(), make_info<C7>(), make_info<C8>(),
                               make_info<C9>(), make_info<C10>(), make_info<C11>(), make_info<C12>(),
                               make_info<C13>()};
    *count = (int)(sizeof(tab) / sizeof(tab[0]));
    return tab;
}
}  // namespace fftb200
