// vel_kernels.cu -- one z-marching tile kernel that evaluates, at every grid point, all nine
// first derivatives and nine second derivatives of (ux,uy,uz) with the reference's parity table
// (src/integration.f90:118-165, src/les_turbulence.f90:55-67, src/utils.f90:283-291,339-347),
// specialised by an epilogue:
//
//   RhsEpi   : fused convective+diffusive RHS, Smagorinsky nu_t and Euler/AB2/AB3 predictor.
//              ONE launch replaces reference predict_velocity (src/integration.f90:14-197) and
//              calculate_nu_t (src/les_turbulence.f90:10-97): 18 (DNS) / 27 (LES) derivative
//              sweeps, 19 temporaries and 6 history copies in the reference.
//              Algorithmic traffic: reads u(3)+f2,f3(6), writes f1(3)+u*(3) [+nu_t] = 120 (+8) B/pt
//   NutEpi   : calculate_nu_t alone (operator ABI)                       32 B/pt
//   RotEpi   : rotational (src/differential_operators.f90:40-77)        48 B/pt
//   QEpi     : calculate_Q_criterion (:79-108)                           32 B/pt
//   StatsEpi : statistics_calc (src/utils.f90:243-375), 17 sums          24 B/pt
//
// HBM-bound FP64 stencils: no tensor cores (nothing here is a contraction).  Unused derivative
// families are dead-code-eliminated per epilogue.
#include "kernels.h"
#include "stencil_tile.cuh"

namespace o3d {
namespace {

constexpr unsigned V_XMASK = 0x7, V_YMASK = 0x7;
constexpr int V_SLOTS = halo_slots(V_XMASK, V_YMASK);  // 3

struct VelIn {
    const double* u[3];
    Coef cx, cy, cz;
    unsigned par[3];  // parity bits per component (curl/Q use all-even for some terms)
    int zchunk;
};

struct Point {
    long long m;
    int i, j, k;
    double u[3];
    double d1[3][3];  // d1[c][axis] with the component's own parity table
    double d2[3][3];
};

template <class Epi, int MINB>
__global__ void __launch_bounds__(NT, MINB) vel_kernel(const Dims g, const VelIn a, Epi epi) {
    __shared__ double sm[2][3][SH * SW];

    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * TX + tx;
    const int i0 = blockIdx.x * TX, j0 = blockIdx.y * TY;
    const int i = i0 + tx, j = j0 + ty;
    const int kb = blockIdx.z * a.zchunk;
    const int ke = min(g.nz, kb + a.zchunk);
    const long long sz = (long long)g.nx * g.ny;

    const OwnCell oc = own_cell(g, i, j);
    double sown[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) sown[c] = psign(oc.rx, a.par[c], 0) * psign(oc.ry, a.par[c], 1);

    HaloSlot hs[V_SLOTS];
    double hsgn[V_SLOTS];
    const double* hptr[V_SLOTS];
#pragma unroll
    for (int s = 0; s < V_SLOTS; ++s) {
        const int idx = tid + s * NT;
        hs[s] = halo_slot<3>(g, i0, j0, idx, V_XMASK, V_YMASK);
        if (idx >= halo_cells(V_XMASK, V_YMASK)) hs[s].off = -1;
        const int c = hs[s].field;
        const unsigned pc = (c == 0) ? a.par[0] : (c == 1) ? a.par[1] : a.par[2];
        hsgn[s] = psign(hs[s].rx, pc, 0) * psign(hs[s].ry, pc, 1);
        hptr[s] = (c == 0) ? a.u[0] : (c == 1) ? a.u[1] : a.u[2];
    }

    // z window: w[c][m] = u_c(i,j,k-3+m), sign-carrying ghosts
    double w[3][7];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
        bool refl;
        const int pl = zplane(g, kb - R + m, refl);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = 0.0;
            if (oc.loadable) v = __ldg(a.u[c] + (long long)pl * sz + oc.off);
            w[c][m] = v * psign(refl, a.par[c], 2);
        }
    }
    double hreg[V_SLOTS];
#pragma unroll
    for (int s = 0; s < V_SLOTS; ++s)
        hreg[s] = (hs[s].off >= 0) ? __ldg(hptr[s] + (long long)kb * sz + hs[s].off) : 0.0;

    bool wallx[3], wally[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        wallx[c] = even_wall(i, g.nx, g.bx, g.bx, a.par[c], 0);
        wally[c] = even_wall(j, g.ny, g.by, g.by, a.par[c], 1);
    }

    for (int k = kb; k < ke; ++k) {
        const int buf = (k - kb) & 1;
        // ---- issue the long-latency loads of this iteration ----
        {
            bool refl;
            const int pl = zplane(g, k + R, refl);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double v = 0.0;
                if (oc.loadable) v = __ldg(a.u[c] + (long long)pl * sz + oc.off);
                w[c][6] = v * psign(refl, a.par[c], 2);
            }
        }
        double hn[V_SLOTS];
#pragma unroll
        for (int s = 0; s < V_SLOTS; ++s)
            hn[s] = (k + 1 < ke && hs[s].off >= 0)
                        ? __ldg(hptr[s] + (long long)(k + 1) * sz + hs[s].off)
                        : 0.0;
        const long long m = (long long)k * sz + (long long)j * g.nx + i;
        typename Epi::Pre pre = epi.prefetch(m, oc.in_dom);

        // ---- stage the centre plane ----
        if (oc.loadable) {
#pragma unroll
            for (int c = 0; c < 3; ++c) sm[buf][c][(ty + R) * SW + tx + R] = w[c][3] * sown[c];
        }
#pragma unroll
        for (int s = 0; s < V_SLOTS; ++s)
            if (hs[s].off >= 0) sm[buf][hs[s].field][hs[s].sm] = hreg[s] * hsgn[s];
        __syncthreads();

        if (oc.in_dom) {
            const bool wz = (k == 0 && g.bz_lo == BM_MIRROR) ||
                            (k == g.nz - 1 && g.bz_hi == BM_MIRROR);
            Point P;
            P.m = m, P.i = i, P.j = j, P.k = k;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double* t = &sm[buf][c][(ty + R) * SW + tx + R];
                const double f0 = w[c][3];
                P.u[c] = f0;
                {
                    const double m3 = t[-3], m2 = t[-2], m1 = t[-1], p1 = t[1], p2 = t[2],
                                 p3 = t[3];
                    P.d1[c][0] = wallx[c] ? 0.0
                                          : d1_expr(a.cx.a1, a.cx.b1, a.cx.c1, m3, m2, m1, p1,
                                                    p2, p3);
                    P.d2[c][0] = d2_expr(a.cx.a2, a.cx.b2, a.cx.c2, m2, m1, f0, p1, p2);
                }
                {
                    const double m3 = t[-3 * SW], m2 = t[-2 * SW], m1 = t[-SW], p1 = t[SW],
                                 p2 = t[2 * SW], p3 = t[3 * SW];
                    P.d1[c][1] = wally[c] ? 0.0
                                          : d1_expr(a.cy.a1, a.cy.b1, a.cy.c1, m3, m2, m1, p1,
                                                    p2, p3);
                    P.d2[c][1] = d2_expr(a.cy.a2, a.cy.b2, a.cy.c2, m2, m1, f0, p1, p2);
                }
                if (g.sim2d) {  // derz_2dsim / derzz_2dsim, src/derivation.f90:481,934
                    P.d1[c][2] = 0.0;
                    P.d2[c][2] = 0.0;
                } else {
                    const bool wallz = wz && !((a.par[c] >> 2) & 1u);
                    P.d1[c][2] = wallz ? 0.0
                                       : d1_expr(a.cz.a1, a.cz.b1, a.cz.c1, w[c][0], w[c][1],
                                                 w[c][2], w[c][4], w[c][5], w[c][6]);
                    P.d2[c][2] = d2_expr(a.cz.a2, a.cz.b2, a.cz.c2, w[c][1], w[c][2], f0,
                                         w[c][4], w[c][5]);
                }
            }
            epi.apply(P, pre);
        }
        // ---- advance the z window ----
#pragma unroll
        for (int c = 0; c < 3; ++c) {
#pragma unroll
            for (int q = 0; q < 6; ++q) w[c][q] = w[c][q + 1];
        }
#pragma unroll
        for (int s = 0; s < V_SLOTS; ++s) hreg[s] = hn[s];
    }
    epi.finish(tid, &sm[0][0][0]);
}

// Smagorinsky viscosity, src/les_turbulence.f90:70-88
__device__ __forceinline__ double smagorinsky(const Point& P, double csd2) {
    const double s11 = P.d1[0][0], s22 = P.d1[1][1], s33 = P.d1[2][2];
    const double s12 = 0.5 * (P.d1[0][1] + P.d1[1][0]);
    const double s13 = 0.5 * (P.d1[0][2] + P.d1[2][0]);
    const double s23 = 0.5 * (P.d1[1][2] + P.d1[2][1]);
    const double smag = sqrt(2.0 * (s11 * s11 + s22 * s22 + s33 * s33 +
                                    2.0 * (s12 * s12 + s13 * s13 + s23 * s23)));
    return csd2 * smag;
}

// ---------------------------------------------------------------------------------------
struct RhsEpi {
    const double* f2[3];
    const double* f3[3];
    double* f1[3];
    double* up[3];
    double* nu_t;
    double onere, adu, bdu, cdu, csd2;
    int iles;
    struct Pre {
        double f2v[3], f3v[3];
    };
    __device__ __forceinline__ Pre prefetch(long long m, bool in_dom) const {
        Pre p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p.f2v[c] = in_dom ? __ldg(f2[c] + m) : 0.0;
            p.f3v[c] = in_dom ? f3[c][m] : 0.0;  // plain load: f3 may alias f1
        }
        return p;
    }
    __device__ __forceinline__ void apply(const Point& P, const Pre& pre) {
        double nut = 0.0;
        if (iles) {
            nut = smagorinsky(P, csd2);
            nu_t[P.m] = nut;
        }
        const double nu_eff = onere + nut;  // src/integration.f90:114
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // src/integration.f90:129-134 (and :149-154, :169-174)
            const double f =
                nu_eff * (P.d2[c][0] + P.d2[c][1] + P.d2[c][2]) -
                (P.u[0] * P.d1[c][0] + P.u[1] * P.d1[c][1] + P.u[2] * P.d1[c][2]);
            const double upv = P.u[c] + adu * f + bdu * pre.f2v[c] + cdu * pre.f3v[c];
            f1[c][P.m] = f;
            up[c][P.m] = upv;
        }
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

struct NoPre {};

struct NutEpi {
    double* nu_t;
    double csd2;
    typedef NoPre Pre;
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Point& P, const Pre&) {
        nu_t[P.m] = smagorinsky(P, csd2);
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// rotational: every term uses the even closure (src/differential_operators.f90:64-74); the
// launcher passes par = {0,0,0}
struct RotEpi {
    double *rx, *ry, *rz;
    typedef NoPre Pre;
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Point& P, const Pre&) {
        rx[P.m] = P.d1[2][1] - P.d1[1][2];  // duzdy - duydz
        ry[P.m] = P.d1[0][2] - P.d1[2][0];  // duxdz - duzdx
        rz[P.m] = P.d1[1][0] - P.d1[0][1];  // duydx - duxdy
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

struct QEpi {
    double* q;
    typedef NoPre Pre;
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Point& P, const Pre&) {
        // src/differential_operators.f90:103-104
        q[P.m] = -(0.5 * (P.d1[0][0] * P.d1[0][0] + P.d1[1][1] * P.d1[1][1] +
                          P.d1[2][2] * P.d1[2][2])) -
                 P.d1[0][1] * P.d1[1][0] - P.d1[0][2] * P.d1[2][0] - P.d1[1][2] * P.d1[2][1];
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// statistics_calc: 16 sums (columns 2..17 of stats.dat), src/utils.f90:277-361
constexpr int NSTAT = 16;
struct StatsEpi {
    double* partial;  // [NSTAT][nblocks]
    double xnu;
    double acc[NSTAT];
    typedef NoPre Pre;
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Point& P, const Pre&) {
        // order of acc: e_k, eps, eps2, dzeta, ux2, uy2, uz2, 9 x d1^2 (c major, axis minor)
        acc[0] += 0.5 * (P.u[0] * P.u[0] + P.u[1] * P.u[1] + P.u[2] * P.u[2]);
        const double a = 2.0 * P.d1[0][0], b = 2.0 * P.d1[1][1], c = 2.0 * P.d1[2][2];
        const double sxy = P.d1[0][1] + P.d1[1][0], sxz = P.d1[0][2] + P.d1[2][0],
                     syz = P.d1[1][2] + P.d1[2][1];
        acc[1] += 0.5 * xnu *
                  (a * a + b * b + c * c + 2.0 * (sxy * sxy) + 2.0 * (sxz * sxz) +
                   2.0 * (syz * syz));
        acc[2] += (-xnu) * (P.u[0] * (P.d2[0][0] + P.d2[0][1] + P.d2[0][2]) +
                            P.u[1] * (P.d2[1][0] + P.d2[1][1] + P.d2[1][2]) +
                            P.u[2] * (P.d2[2][0] + P.d2[2][1] + P.d2[2][2]));
        const double wx = P.d1[2][1] - P.d1[1][2], wy = P.d1[0][2] - P.d1[2][0],
                     wz = P.d1[1][0] - P.d1[0][1];
        acc[3] += 0.5 * (wx * wx + wy * wy + wz * wz);
        acc[4] += P.u[0] * P.u[0];
        acc[5] += P.u[1] * P.u[1];
        acc[6] += P.u[2] * P.u[2];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) acc[7 + 3 * cc + ax] += P.d1[cc][ax] * P.d1[cc][ax];
    }
    __device__ __forceinline__ void finish(int tid, double* smem) {
        // deterministic block reduction: warp shuffle, then warp 0 over the 8 warp sums
        __syncthreads();
        const int nblocks = gridDim.x * gridDim.y * gridDim.z;
        const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
#pragma unroll
        for (int s = 0; s < NSTAT; ++s) {
            const double v = warp_sum(acc[s]);
            if ((tid & 31) == 0) smem[s * 8 + (tid >> 5)] = v;
        }
        __syncthreads();
        if (tid < NSTAT) {
            double t = 0.0;
            for (int q = 0; q < NT / 32; ++q) t += smem[tid * 8 + q];
            partial[(long long)tid * nblocks + b] = t;
        }
    }
};

template <class Epi, int MINB>
int run_vel(cudaStream_t st, const Dims& g, VelIn a, const Epi& epi) {
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    a.zchunk = pick_zchunk(gx * gy, g.nz);
    const int gz = (g.nz + a.zchunk - 1) / a.zchunk;
    vel_kernel<Epi, MINB><<<dim3(gx, gy, gz), dim3(TX, TY, 1), 0, st>>>(g, a, epi);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

VelIn make_in(const double* ux, const double* uy, const double* uz, const Coef& cx,
              const Coef& cy, const Coef& cz, bool natural_parity) {
    VelIn a;
    a.u[0] = ux, a.u[1] = uy, a.u[2] = uz;
    a.cx = cx, a.cy = cy, a.cz = cz;
    // natural parity: the component normal to an axis is odd along it
    a.par[0] = natural_parity ? 0x1u : 0u;
    a.par[1] = natural_parity ? 0x2u : 0u;
    a.par[2] = natural_parity ? 0x4u : 0u;
    a.zchunk = 0;
    return a;
}

}  // namespace

int launch_rhs(cudaStream_t st, const Dims& g, const RhsArgs& r) {
    RhsEpi e;
    for (int c = 0; c < 3; ++c) e.f2[c] = r.f2[c], e.f3[c] = r.f3[c], e.f1[c] = r.f1[c], e.up[c] = r.up[c];
    e.nu_t = r.nu_t;
    e.onere = r.onere, e.adu = r.adu, e.bdu = r.bdu, e.cdu = r.cdu, e.csd2 = r.csd2;
    e.iles = r.iles;
    return run_vel<RhsEpi, 2>(st, g, make_in(r.u[0], r.u[1], r.u[2], r.cx, r.cy, r.cz, true), e);
}

int launch_nu_t(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double csd2,
                double* nu_t) {
    NutEpi e;
    e.nu_t = nu_t, e.csd2 = csd2;
    return run_vel<NutEpi, 2>(st, g, make_in(ux, uy, uz, cx, cy, cz, true), e);
}

int launch_rot(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
               const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double* rotx,
               double* roty, double* rotz) {
    RotEpi e;
    e.rx = rotx, e.ry = roty, e.rz = rotz;
    return run_vel<RotEpi, 2>(st, g, make_in(ux, uy, uz, cx, cy, cz, false), e);
}

int launch_qcrit(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                 const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double* q) {
    QEpi e;
    e.q = q;
    // derxi/deryi/derzi on the diagonal, p closures elsewhere = natural parity
    return run_vel<QEpi, 2>(st, g, make_in(ux, uy, uz, cx, cy, cz, true), e);
}

int stats_blocks(const Dims& g) {
    const int gx = (g.nx + TX - 1) / TX, gy = (g.ny + TY - 1) / TY;
    const int zc = pick_zchunk(gx * gy, g.nz);
    return gx * gy * ((g.nz + zc - 1) / zc);
}

int launch_stats(cudaStream_t st, const Dims& g, const double* ux, const double* uy,
                 const double* uz, const Coef& cx, const Coef& cy, const Coef& cz, double xnu,
                 double* partial) {
    StatsEpi e;
    e.partial = partial, e.xnu = xnu;
    for (int s = 0; s < NSTAT; ++s) e.acc[s] = 0.0;
    return run_vel<StatsEpi, 1>(st, g, make_in(ux, uy, uz, cx, cy, cz, true), e);
}

}  // namespace o3d
