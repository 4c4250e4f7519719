// stencil_rules_test.cu -- the kernels' stencil arithmetic on the CPU.  Every stencil kernel of
// the library evaluates ONE expression per derivative -- d1_expr / d2_expr of csrc/o3d_common.cuh
// with the coefficients of make_coef -- at every point, reading ghost cells that hold the
// boundary closure as data (map_index + the sign of the odd closure).  This program does exactly
// that on the host for the 18 routines of src/derivation.f90 and writes the results, so that
// tests/test_host_rules_cpu.py can compare them, bit for bit, with the vectors obtained by
// executing the reference's Fortran source (tests/golden/hotpath.npz): the explicitly written
// boundary planes of the reference (e.g. a*(f(5)+f(3)) at derivation.f90:140, the literal 0.d0 at
// :87) must come out of the single interior expression.
//
//   stencil_rules_test in.bin nx ny nz dx dy dz out.bin
//   out = 18 fields: axis (x, y, z) x order (1, 2) x closure (00, p_11, i_11), in that nesting
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../osinco3d_b200/csrc/o3d_common.cuh"

namespace o3d {
void set_error(const char*, ...) {}
}  // namespace o3d
using namespace o3d;

int main(int argc, char** argv) {
    if (argc != 9) return 2;
    const int nx = atoi(argv[2]), ny = atoi(argv[3]), nz = atoi(argv[4]);
    const double dd[3] = {atof(argv[5]), atof(argv[6]), atof(argv[7])};
    const size_t N = (size_t)nx * ny * nz;
    std::vector<double> f(N), df(N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(f.data(), 8, N, fi) != N) return 3;
    fclose(fi);
    FILE* fo = fopen(argv[8], "wb");
    if (!fo) return 4;
    const int ext[3] = {nx, ny, nz};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    for (int axis = 0; axis < 3; ++axis)
        for (int order = 1; order <= 2; ++order)
            for (int closure = 0; closure < 3; ++closure) {  // 0: _00, 1: p_11 (even), 2: i_11 (odd)
                const int n = ext[axis];
                const int mode = (closure == 0) ? BM_WRAP : BM_MIRROR;
                const bool odd = (closure == 2);
                const Coef c = make_coef(dd[axis]);
                for (int k = 0; k < nz; ++k)
                    for (int j = 0; j < ny; ++j)
                        for (int i = 0; i < nx; ++i) {
                            const size_t m = i + (size_t)nx * (j + (size_t)ny * k);
                            const int p = (axis == 0) ? i : (axis == 1) ? j : k;
                            const size_t base = m - (size_t)p * stride[axis];
                            // what a kernel reads at line offset o: an interior cell, or the
                            // ghost cell the closure filled
                            auto at = [&](int o) {
                                bool refl;
                                const int src = map_index(p + o, n, mode, mode, refl);
                                const double v = f[base + (size_t)src * stride[axis]];
                                return (refl && odd) ? -v : v;
                            };
                            df[m] = (order == 1)
                                        ? d1_expr(c.a1, c.b1, c.c1, at(-3), at(-2), at(-1), at(1), at(2), at(3))
                                        : d2_expr(c.a2, c.b2, c.c2, at(-2), at(-1), at(0), at(1), at(2));
                        }
                if (fwrite(df.data(), 8, N, fo) != N) return 5;
            }
    fclose(fo);
    printf("stencil rules written\n");
    return 0;
}
