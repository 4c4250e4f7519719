// proj_kernels.cu -- the two stencil kernels of the projection step:
//   div_kernel : divergence (src/differential_operators.f90:7-38), optionally /dt -> Poisson
//             right-hand side (src/integration.f90:234-239).  3 sweeps + 3 temporaries in the
//             reference; here 3 reads + 1 write = 32 B/pt.  odd/even closure = ghost parity.
//   CorrEpi (march engine): u = u* - dt grad p with the NaN / >1000 guard fused in
//             (src/integration.f90:298-325).  pp(1)+u*(3) reads, u(3) writes = 56 B/pt.
#include "kernels.h"
#include "march.cuh"

namespace o3d {
namespace {

struct NoPre {};

// Divergence: each field is differentiated along ONE axis only, so a staged 3-field window
// would move 2x the needed bytes through shared memory.  Direct form instead: fz keeps its
// z-window in registers (one new load per plane), fx / fy neighbours are read straight from
// global memory -- they are the same 128-byte lines / the rows of the neighbouring warps of the
// CTA, i.e. L1 hits -- so there is no shared memory, no barrier, and with 40 registers per
// thread enough resident warps to cover the HBM latency.
constexpr int DTX = 32, DTY = 8;
__global__ void __launch_bounds__(DTX* DTY) div_kernel(const Geom g, const double* __restrict__ fx,
                                                        const double* __restrict__ fy,
                                                        const double* __restrict__ fz,
                                                        double* __restrict__ out, Coef cx, Coef cy,
                                                        Coef cz, int divide, double dt,
                                                        int zchunk) {
    const int i = blockIdx.x * DTX + threadIdx.x, j = blockIdx.y * DTY + threadIdx.y;
    if (i >= g.nx || j >= g.ny) return;
    const int kb = blockIdx.z * zchunk, ke = min(g.nz, kb + zchunk);
    const long long sy = g.sy, sz = g.sz;
    long long m = (long long)kb * sz + (long long)j * sy + i;
    double w[7];
#pragma unroll
    for (int q = 0; q < 6; ++q) w[q] = __ldg(fz + m + (long long)(q - 3) * sz);
    for (int k = kb; k < ke; ++k, m += sz) {
        w[6] = __ldg(fz + m + 3 * sz);
        const double dfx = d1_expr(cx.a1, cx.b1, cx.c1, __ldg(fx + m - 3), __ldg(fx + m - 2),
                                   __ldg(fx + m - 1), __ldg(fx + m + 1), __ldg(fx + m + 2),
                                   __ldg(fx + m + 3));
        const double dfy = d1_expr(cy.a1, cy.b1, cy.c1, __ldg(fy + m - 3 * sy),
                                   __ldg(fy + m - 2 * sy), __ldg(fy + m - sy), __ldg(fy + m + sy),
                                   __ldg(fy + m + 2 * sy), __ldg(fy + m + 3 * sy));
        const double dfz = g.sim2d ? 0.0
                                   : d1_expr(cz.a1, cz.b1, cz.c1, w[0], w[1], w[2], w[4], w[5], w[6]);
        double v = dfx + dfy + dfz;  // src/differential_operators.f90:35
        if (divide) v = v / dt;      // src/integration.f90:239
        out[m] = v;
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = w[q + 1];
    }
}

struct CorrEpi {
    const double* up[3];
    double* u[3];
    Coef cx, cy, cz;
    double dt;
    int* flag;
    int sim2d;
    int bad;
    struct Pre {
        double v[3];
    };
    __device__ __forceinline__ Pre prefetch(long long m, bool ok) const {
        Pre p;
#pragma unroll
        for (int c = 0; c < 3; ++c) p.v[c] = ok ? __ldg(up[c] + m) : 0.0;
        return p;
    }
    __device__ __forceinline__ void apply(const Ring<1>& r, long long m, int, int, int,
                                          const Pre& pre) {
        // src/integration.f90:298-300 (derxp, deryp, derzp)
        const double dpdx = r.d1x(0, cx);
        const double dpdy = r.d1y(0, cy);
        const double dpdz = sim2d ? 0.0 : r.d1z(0, cz);
        // src/integration.f90:304-306
        const double u0 = pre.v[0] - dt * dpdx;
        const double u1 = pre.v[1] - dt * dpdy;
        const double u2 = pre.v[2] - dt * dpdz;
        u[0][m] = u0;
        u[1][m] = u1;
        u[2][m] = u2;
        // src/integration.f90:309-325: contains_nan .or. maxval > 1000
        if (u0 != u0 || u1 != u1 || u2 != u2 || u0 > 1000. || u1 > 1000. || u2 > 1000.) bad = 1;
    }
    __device__ __forceinline__ void finish(int tid, double*) {
        const int b = warp_or(bad);
        if (b && (tid & 31) == 0) atomicOr(flag, 1);
    }
};

}  // namespace

int launch_div(cudaStream_t st, const Geom& g, const FieldRef* f, const Coef& cx, const Coef& cy,
               const Coef& cz, int divide_by_dt, double dt, double* out) {
    const int gx = (g.nx + DTX - 1) / DTX, gy = (g.ny + DTY - 1) / DTY;
    const int zchunk = pick_zchunk(gx * gy, g.nz);
    div_kernel<<<dim3(gx, gy, (g.nz + zchunk - 1) / zchunk), dim3(DTX, DTY, 1), 0, st>>>(
        g, f[0].p, f[1].p, f[2].p, out, cx, cy, cz, divide_by_dt, dt, zchunk);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_corr(cudaStream_t st, const Geom& g, const FieldRef& pp, const double* const* up,
                double* const* u, const Coef& cx, const Coef& cy, const Coef& cz, double dt,
                int* flag) {
    CorrEpi e;
    for (int c = 0; c < 3; ++c) e.up[c] = up[c], e.u[c] = u[c];
    e.cx = cx, e.cy = cy, e.cz = cz;
    e.dt = dt, e.flag = flag, e.sim2d = g.sim2d, e.bad = 0;
    MarchMaps<1> m;
    m.m[0] = *pp.tm;
    return launch_march<1, CorrEpi, 3>(st, g, m, e);
}

}  // namespace o3d
