// rows_big.cu -- dispatch of the large-frame row kernels (templates in rows_big.cuh) + the widths 3840, 1920, 2560, 1280.
#include "rows_big.cuh"

namespace admm {

int launch_rows_big_set_b(RowMode mode, const Geometry& g, const RowArgs& a, cudaStream_t st);   // rows_big2.cu

bool rows_big_supported(const Geometry& g) {
    if (options().force_generic || !(options().use_big & 1)) return false;
    return (g.W == 3840 || g.W == 1920 || g.W == 1024 || g.W == 2048 || g.W == 4096 || g.W == 2560 || g.W == 1280 || g.W == 768 || g.W == 1536 || g.W == 3072) && (g.H % 2 == 0) && g.H >= 4;
}

static int launch_rows_big_any(RowMode mode, const Geometry& g, const RowArgs& a, cudaStream_t st) {
    switch (g.W) {
        case 3840: return launch_rows_big_width<3840>(mode, g, a, st);
        case 1920: return launch_rows_big_width<1920>(mode, g, a, st);
        case 2560: return launch_rows_big_width<2560>(mode, g, a, st);
        case 1280: return launch_rows_big_width<1280>(mode, g, a, st);
        default: return launch_rows_big_set_b(mode, g, a, st);
    }
}

int launch_rows_big(RowMode mode, const Geometry& g, const RowArgs& a, cudaStream_t st) {
    if (mode != ROWS_FULL && mode != ROWS_FULL_U) return fail(4, "large-row kernel: unsupported mode");
    return launch_rows_big_any(mode, g, a, st);
}

int launch_rows_big_plain(RowMode mode, const Geometry& g, const RowArgs& a, cudaStream_t st) {
    if (mode != ROWS_R2C && mode != ROWS_C2R) return fail(4, "large-row kernel: unsupported plain mode");
    return launch_rows_big_any(mode, g, a, st);
}

}  // namespace admm
