// der_kernel.cu -- the 18 single-axis derivative routines of src/derivation.f90 behind one
// kernel template (operator form of der_type, src/initialization.f90:86-91).  The closure
// (periodic / even / odd) is in the ghost cells of the padded input (o3d_common.cuh), so the
// kernel is the reference's interior formula at every point.
// Algorithmic traffic 16 B/pt (1 read + 1 write); neighbours come from L1/L2.  Used by the
// drop-in operator ABI; the time step itself runs the fused march kernels.
#include "kernels.h"

namespace o3d {
namespace {

template <int ORDER>
__global__ void __launch_bounds__(256) der_kernel(const Geom g, long long s, int zero, double a,
                                                  double b, double c,
                                                  const double* __restrict__ f,
                                                  double* __restrict__ df) {
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        for (int i = threadIdx.x; i < g.nx; i += blockDim.x) {
            const long long m = base + i;
            double r;
            if (zero)  // derz_2dsim / derzz_2dsim, src/derivation.f90:481,934
                r = 0.0;
            else if (ORDER == 1)
                r = d1_expr(a, b, c, __ldg(f + m - 3 * s), __ldg(f + m - 2 * s), __ldg(f + m - s),
                            __ldg(f + m + s), __ldg(f + m + 2 * s), __ldg(f + m + 3 * s));
            else
                r = d2_expr(a, b, c, __ldg(f + m - 2 * s), __ldg(f + m - s), __ldg(f + m),
                            __ldg(f + m + s), __ldg(f + m + 2 * s));
            df[m] = r;
        }
    }
}

}  // namespace

int launch_der(cudaStream_t st, const Geom& g, int axis, int order, int zero, double d,
               const double* f, double* df) {
    const Coef co = make_coef(d);
    const double a = (order == 1) ? co.a1 : co.a2;
    const double b = (order == 1) ? co.b1 : co.b2;
    const double c = (order == 1) ? co.c1 : co.c2;
    const long long s = (axis == 0) ? 1 : (axis == 1) ? g.sy : g.sz;
    long long nb = (long long)g.ny * g.nz;
    if (nb > 148 * 16) nb = 148 * 16;
    if (order == 1)
        der_kernel<1><<<(unsigned)nb, 256, 0, st>>>(g, s, zero, a, b, c, f, df);
    else
        der_kernel<2><<<(unsigned)nb, 256, 0, st>>>(g, s, zero, a, b, c, f, df);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
