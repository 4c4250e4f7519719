// der_kernel.cu -- the 18 single-axis derivative routines of src/derivation.f90 behind one
// kernel template (operator form of der_type, src/initialization.f90:86-91).
// Algorithmic traffic 16 B/pt (1 read + 1 write).  Used by the drop-in operator ABI, the
// initial-condition helpers and the non-fused diagnostics; the time step itself runs the fused
// kernels (rhs_kernel.cu, proj_kernels.cu).
#include "kernels.h"

namespace o3d {
namespace {

constexpr int DBX = 64, DBY = 4;

// value of line element q under the closure; `base` points at line element 0
__device__ __forceinline__ double line_val(const double* base, long long s, int q, int n,
                                           int mlo, int mhi, int parity) {
    bool refl;
    const int m = map_index(q, n, mlo, mhi, refl);
    const double v = __ldg(base + (long long)m * s);
    return (refl && parity) ? -v : v;
}

template <int AXIS, int ORDER>
__global__ void __launch_bounds__(DBX* DBY) der_kernel(const Dims g, int parity, int zero,
                                                        double a, double b, double c,
                                                        const double* __restrict__ f,
                                                        double* __restrict__ df, int zchunk) {
    const int i = blockIdx.x * DBX + threadIdx.x;
    const int j = blockIdx.y * DBY + threadIdx.y;
    if (i >= g.nx || j >= g.ny) return;
    const int kb = blockIdx.z * zchunk, ke = min(g.nz, kb + zchunk);
    const long long sy = g.nx, sz = (long long)g.nx * g.ny;
    const int n = (AXIS == 0) ? g.nx : (AXIS == 1) ? g.ny : g.nz;
    const long long s = (AXIS == 0) ? 1 : (AXIS == 1) ? sy : sz;
    const int mlo = (AXIS == 0) ? g.bx : (AXIS == 1) ? g.by : g.bz_lo;
    const int mhi = (AXIS == 0) ? g.bx : (AXIS == 1) ? g.by : g.bz_hi;
    constexpr int RR = (ORDER == 1) ? 3 : 2;
    for (int k = kb; k < ke; ++k) {
        const long long m = (long long)k * sz + (long long)j * sy + i;
        if (zero) {
            df[m] = 0.0;
            continue;
        }
        const int p = (AXIS == 0) ? i : (AXIS == 1) ? j : k;
        double r;
        if (p >= RR && p < n - RR) {
            if (ORDER == 1)
                r = d1_expr(a, b, c, __ldg(f + m - 3 * s), __ldg(f + m - 2 * s), __ldg(f + m - s),
                            __ldg(f + m + s), __ldg(f + m + 2 * s), __ldg(f + m + 3 * s));
            else
                r = d2_expr(a, b, c, __ldg(f + m - 2 * s), __ldg(f + m - s), __ldg(f + m),
                            __ldg(f + m + s), __ldg(f + m + 2 * s));
        } else {
            const double* base = f + (m - (long long)p * s);
            if (ORDER == 1) {
                const bool wall = !parity && ((p == 0 && mlo == BM_MIRROR) ||
                                              (p == n - 1 && mhi == BM_MIRROR));
                if (wall)
                    r = 0.0;  // src/derivation.f90:87,:105
                else
                    r = d1_expr(a, b, c, line_val(base, s, p - 3, n, mlo, mhi, parity),
                                line_val(base, s, p - 2, n, mlo, mhi, parity),
                                line_val(base, s, p - 1, n, mlo, mhi, parity),
                                line_val(base, s, p + 1, n, mlo, mhi, parity),
                                line_val(base, s, p + 2, n, mlo, mhi, parity),
                                line_val(base, s, p + 3, n, mlo, mhi, parity));
            } else {
                r = d2_expr(a, b, c, line_val(base, s, p - 2, n, mlo, mhi, parity),
                            line_val(base, s, p - 1, n, mlo, mhi, parity), __ldg(f + m),
                            line_val(base, s, p + 1, n, mlo, mhi, parity),
                            line_val(base, s, p + 2, n, mlo, mhi, parity));
            }
        }
        df[m] = r;
    }
}

}  // namespace

int launch_der(cudaStream_t st, const Dims& g, int axis, int order, int parity, int zero,
               double d, const double* f, double* df) {
    const Coef co = make_coef(d);
    const double a = (order == 1) ? co.a1 : co.a2;
    const double b = (order == 1) ? co.b1 : co.b2;
    const double c = (order == 1) ? co.c1 : co.c2;
    const int gx = (g.nx + DBX - 1) / DBX, gy = (g.ny + DBY - 1) / DBY;
    const int zchunk = pick_zchunk(gx * gy, g.nz);
    const dim3 grid(gx, gy, (g.nz + zchunk - 1) / zchunk), block(DBX, DBY, 1);
#define O3D_DER_CASE(AX, OR)                                                            \
    if (axis == AX && order == OR)                                                      \
        der_kernel<AX, OR><<<grid, block, 0, st>>>(g, parity, zero, a, b, c, f, df, zchunk);
    O3D_DER_CASE(0, 1)
    O3D_DER_CASE(1, 1)
    O3D_DER_CASE(2, 1)
    O3D_DER_CASE(0, 2)
    O3D_DER_CASE(1, 2)
    O3D_DER_CASE(2, 2)
#undef O3D_DER_CASE
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
