// proj_kernels.cu -- the two stencil kernels of the projection step:
//   div_kernel : divergence (src/differential_operators.f90:7-38), optionally /dt -> Poisson
//                right-hand side (src/integration.f90:234-239).  3 sweeps + 3 temporaries in
//                the reference; here 3 reads + 1 write = 32 B/pt.
//   corr_kernel: u = u* - dt grad p with the NaN / >1000 guard fused in
//                (src/integration.f90:298-325).  pp(1)+u*(3) reads, u(3) writes = 56 B/pt.
#include "kernels.h"
#include "stencil_tile.cuh"

namespace o3d {
namespace {

// ------------------------------------------------------------------------------------------
struct DivArgs {
    const double* f[3];
    double* out;
    Coef cx, cy, cz;
    unsigned par[3];
    int divide;
    double dt;
    int zchunk;
};

constexpr unsigned DIV_XMASK = 0x1, DIV_YMASK = 0x2;
static_assert(halo_slots(DIV_XMASK, DIV_YMASK) == 1, "one halo slot per thread");

__global__ void __launch_bounds__(NT, 3) div_kernel(const Dims g, const DivArgs a) {
    __shared__ double sm[2][2][SH * SW];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int kb = blockIdx.z * a.zchunk, ke = min(g.nz, kb + a.zchunk);
    const long long sz = (long long)g.nx * g.ny;

    const OwnCell oc = own_cell(g, i, j);
    // fx is only differentiated along x, fy along y: ghost providers matter per field
    const double sx = psign(oc.rx, a.par[0], 0), sy = psign(oc.ry, a.par[1], 1);
    const bool load_x = oc.in_dom || (oc.loadable && i >= g.nx);
    const bool load_y = oc.in_dom || (oc.loadable && j >= g.ny);

    HaloSlot hs = halo_slot<2>(g, i0, j0, tid, DIV_XMASK, DIV_YMASK);
    if (tid >= halo_cells(DIV_XMASK, DIV_YMASK)) hs.off = -1;
    const double hsgn = psign(hs.rx, a.par[0], 0) * psign(hs.ry, a.par[1], 1);
    const double* hptr = hs.field == 0 ? a.f[0] : a.f[1];

    double wz[7];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        bool refl;
        const int pl = zplane(g, kb - R + m, refl);
        double v = 0.0;
        if (oc.in_dom) v = __ldg(a.f[2] + (long long)pl * sz + oc.off);
        wz[m] = v * psign(refl, a.par[2], 2);
    }
    double cxv = load_x ? __ldg(a.f[0] + (long long)kb * sz + oc.off) : 0.0;
    double cyv = load_y ? __ldg(a.f[1] + (long long)kb * sz + oc.off) : 0.0;
    double hreg = (hs.off >= 0) ? __ldg(hptr + (long long)kb * sz + hs.off) : 0.0;

    const bool wallx = even_wall(i, g.nx, g.bx, g.bx, a.par[0], 0);
    const bool wally = even_wall(j, g.ny, g.by, g.by, a.par[1], 1);

    for (int k = kb; k < ke; ++k) {
        const int buf = (k - kb) & 1;
        {
            bool refl;
            const int pl = zplane(g, k + R, refl);
            double v = 0.0;
            if (oc.in_dom) v = __ldg(a.f[2] + (long long)pl * sz + oc.off);
            wz[6] = v * psign(refl, a.par[2], 2);
        }
        double cxn = 0.0, cyn = 0.0, hn = 0.0;
        if (k + 1 < ke) {
            if (load_x) cxn = __ldg(a.f[0] + (long long)(k + 1) * sz + oc.off);
            if (load_y) cyn = __ldg(a.f[1] + (long long)(k + 1) * sz + oc.off);
            if (hs.off >= 0) hn = __ldg(hptr + (long long)(k + 1) * sz + hs.off);
        }
        if (load_x) sm[buf][0][(ty + R) * SW + tx + R] = cxv * sx;
        if (load_y) sm[buf][1][(ty + R) * SW + tx + R] = cyv * sy;
        if (hs.off >= 0) sm[buf][hs.field][hs.sm] = hreg * hsgn;
        __syncthreads();
        if (oc.in_dom) {
            const double* t0 = &sm[buf][0][(ty + R) * SW + tx + R];
            const double* t1 = &sm[buf][1][(ty + R) * SW + tx + R];
            const double dfx = wallx ? 0.0
                                     : d1_expr(a.cx.a1, a.cx.b1, a.cx.c1, t0[-3], t0[-2], t0[-1],
                                               t0[1], t0[2], t0[3]);
            const double dfy = wally ? 0.0
                                     : d1_expr(a.cy.a1, a.cy.b1, a.cy.c1, t1[-3 * SW],
                                               t1[-2 * SW], t1[-SW], t1[SW], t1[2 * SW],
                                               t1[3 * SW]);
            double dfz;
            if (g.sim2d) {
                dfz = 0.0;
            } else {
                const bool wallz = !((a.par[2] >> 2) & 1u) &&
                                   ((k == 0 && g.bz_lo == BM_MIRROR) ||
                                    (k == g.nz - 1 && g.bz_hi == BM_MIRROR));
                dfz = wallz ? 0.0
                            : d1_expr(a.cz.a1, a.cz.b1, a.cz.c1, wz[0], wz[1], wz[2], wz[4],
                                      wz[5], wz[6]);
            }
            double r = dfx + dfy + dfz;         // src/differential_operators.f90:35
            if (a.divide) r = r / a.dt;         // src/integration.f90:239
            a.out[(long long)k * sz + (long long)j * g.nx + i] = r;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) wz[q] = wz[q + 1];
        cxv = cxn;
        cyv = cyn;
        hreg = hn;
    }
}

// ------------------------------------------------------------------------------------------
struct CorrArgs {
    const double* pp;
    const double* up[3];
    double* u[3];
    Coef cx, cy, cz;
    double dt;
    int* flag;
    int zchunk;
};

constexpr unsigned CORR_XMASK = 0x1, CORR_YMASK = 0x1;
static_assert(halo_slots(CORR_XMASK, CORR_YMASK) == 1, "one halo slot per thread");

__global__ void __launch_bounds__(NT, 3) corr_kernel(const Dims g, const CorrArgs a) {
    __shared__ double sm[2][SH * SW];
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * TX + tx;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int kb = blockIdx.z * a.zchunk, ke = min(g.nz, kb + a.zchunk);
    const long long sz = (long long)g.nx * g.ny;

    const OwnCell oc = own_cell(g, i, j);  // pp is even everywhere: no signs
    HaloSlot hs = halo_slot<1>(g, i0, j0, tid, CORR_XMASK, CORR_YMASK);
    if (tid >= halo_cells(CORR_XMASK, CORR_YMASK)) hs.off = -1;

    double w[7];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        bool refl;
        const int pl = zplane(g, kb - R + m, refl);
        w[m] = oc.loadable ? __ldg(a.pp + (long long)pl * sz + oc.off) : 0.0;
    }
    double hreg = (hs.off >= 0) ? __ldg(a.pp + (long long)kb * sz + hs.off) : 0.0;

    const bool wallx = even_wall(i, g.nx, g.bx, g.bx, 0u, 0);
    const bool wally = even_wall(j, g.ny, g.by, g.by, 0u, 1);
    int bad = 0;

    for (int k = kb; k < ke; ++k) {
        const int buf = (k - kb) & 1;
        {
            bool refl;
            const int pl = zplane(g, k + R, refl);
            w[6] = oc.loadable ? __ldg(a.pp + (long long)pl * sz + oc.off) : 0.0;
        }
        double hn = 0.0;
        if (k + 1 < ke && hs.off >= 0) hn = __ldg(a.pp + (long long)(k + 1) * sz + hs.off);
        const long long m = (long long)k * sz + (long long)j * g.nx + i;
        double upv[3] = {0.0, 0.0, 0.0};
        if (oc.in_dom) {
#pragma unroll
            for (int c = 0; c < 3; ++c) upv[c] = __ldg(a.up[c] + m);
        }
        if (oc.loadable) sm[buf][(ty + R) * SW + tx + R] = w[3];
        if (hs.off >= 0) sm[buf][hs.sm] = hreg;
        __syncthreads();
        if (oc.in_dom) {
            const double* t = &sm[buf][(ty + R) * SW + tx + R];
            // src/integration.f90:298-300 (derxp, deryp, derzp)
            const double dpdx = wallx ? 0.0
                                      : d1_expr(a.cx.a1, a.cx.b1, a.cx.c1, t[-3], t[-2], t[-1],
                                                t[1], t[2], t[3]);
            const double dpdy = wally ? 0.0
                                      : d1_expr(a.cy.a1, a.cy.b1, a.cy.c1, t[-3 * SW], t[-2 * SW],
                                                t[-SW], t[SW], t[2 * SW], t[3 * SW]);
            double dpdz;
            if (g.sim2d) {
                dpdz = 0.0;
            } else {
                const bool wallz = (k == 0 && g.bz_lo == BM_MIRROR) ||
                                   (k == g.nz - 1 && g.bz_hi == BM_MIRROR);
                dpdz = wallz ? 0.0
                             : d1_expr(a.cz.a1, a.cz.b1, a.cz.c1, w[0], w[1], w[2], w[4], w[5],
                                       w[6]);
            }
            // src/integration.f90:304-306
            const double u0 = upv[0] - a.dt * dpdx;
            const double u1 = upv[1] - a.dt * dpdy;
            const double u2 = upv[2] - a.dt * dpdz;
            a.u[0][m] = u0;
            a.u[1][m] = u1;
            a.u[2][m] = u2;
            // src/integration.f90:309-325: contains_nan .or. maxval > 1000
            if (u0 != u0 || u1 != u1 || u2 != u2 || u0 > 1000. || u1 > 1000. || u2 > 1000.)
                bad = 1;
        }
#pragma unroll
        for (int q = 0; q < 6; ++q) w[q] = w[q + 1];
        hreg = hn;
    }
    bad = warp_or(bad);
    if (bad && (tid & 31) == 0) atomicOr(a.flag, 1);
}

}  // namespace

int launch_div(cudaStream_t st, const Dims& g, const double* fx, const double* fy,
               const double* fz, const Coef& cx, const Coef& cy, const Coef& cz, int odd,
               int divide_by_dt, double dt, double* out) {
    DivArgs a;
    a.f[0] = fx, a.f[1] = fy, a.f[2] = fz;
    a.out = out;
    a.cx = cx, a.cy = cy, a.cz = cz;
    // src/differential_operators.f90:25-33: odd -> derxi/deryi/derzi, else derxp/deryp/derzp
    a.par[0] = odd ? 0x1u : 0u;
    a.par[1] = odd ? 0x2u : 0u;
    a.par[2] = odd ? 0x4u : 0u;
    a.divide = divide_by_dt;
    a.dt = dt;
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    a.zchunk = pick_zchunk(gx * gy, g.nz);
    div_kernel<<<dim3(gx, gy, (g.nz + a.zchunk - 1) / a.zchunk), dim3(TX, TY, 1), 0, st>>>(g, a);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_corr(cudaStream_t st, const Dims& g, const double* pp, const double* const* up,
                double* const* u, const Coef& cx, const Coef& cy, const Coef& cz, double dt,
                int* flag) {
    CorrArgs a;
    a.pp = pp;
    for (int c = 0; c < 3; ++c) a.up[c] = up[c], a.u[c] = u[c];
    a.cx = cx, a.cy = cy, a.cz = cz;
    a.dt = dt;
    a.flag = flag;
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    a.zchunk = pick_zchunk(gx * gy, g.nz);
    corr_kernel<<<dim3(gx, gy, (g.nz + a.zchunk - 1) / a.zchunk), dim3(TX, TY, 1), 0, st>>>(g, a);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
