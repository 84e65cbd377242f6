// rows_big2.cu -- the power-of-two and 3*2^n widths of the large-frame row kernels (templates in rows_big.cuh).
#include "rows_big.cuh"

namespace admm {

int launch_rows_big_set_b(RowMode mode, const Geometry& g, const RowArgs& a, cudaStream_t st) {
    switch (g.W) {
        case 1024: return launch_rows_big_width<1024>(mode, g, a, st);
        case 2048: return launch_rows_big_width<2048>(mode, g, a, st);
        case 4096: return launch_rows_big_width<4096>(mode, g, a, st);
        case 768: return launch_rows_big_width<768>(mode, g, a, st);
        case 1536: return launch_rows_big_width<1536>(mode, g, a, st);
        case 3072: return launch_rows_big_width<3072>(mode, g, a, st);
        default: return fail(4, "no large-row kernel for this width");
    }
}

}  // namespace admm
