// vel_kernels.cu -- march-engine epilogues that differentiate the three velocity components
// with the reference's parity table (src/integration.f90:118-165, src/les_turbulence.f90:55-67,
// src/utils.f90:283-291,339-347; the parity lives in the ghost cells, see o3d_common.cuh):
//
//   RhsEpi   : fused convective+diffusive RHS, Smagorinsky nu_t and Euler/AB2/AB3 predictor.
//              ONE launch replaces reference predict_velocity (src/integration.f90:14-197) and
//              calculate_nu_t (src/les_turbulence.f90:10-97): 18 (DNS) / 27 (LES) derivative
//              sweeps, 19 temporaries and 6 history copies in the reference.
//              Algorithmic traffic: reads u(3)+f2,f3(6), writes f1(3)+u*(3) [+nu_t] = 120 (+8) B/pt
//   NutEpi   : calculate_nu_t alone (operator ABI)                       32 B/pt
//   RotEpi   : rotational (src/differential_operators.f90:40-77)        48 B/pt
//   QEpi     : calculate_Q_criterion (:79-108)                           32 B/pt
//   StatsEpi : statistics_calc (src/utils.f90:243-375), 16 sums          24 B/pt
//
// HBM-bound FP64 stencils: no tensor cores (nothing here is a contraction).
#include "kernels.h"
#include "march.cuh"

namespace o3d {
namespace {

struct Coefs3 {
    Coef x, y, z;
};

// the nine first derivatives d1[c][axis] of (ux,uy,uz)
struct Grad {
    double d[3][3];
};
__device__ __forceinline__ Grad gradient(const Ring<3>& r, const Coefs3& q, int sim2d) {
    Grad G;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        G.d[c][0] = r.d1x(c, q.x);
        G.d[c][1] = r.d1y(c, q.y);
        G.d[c][2] = sim2d ? 0.0 : r.d1z(c, q.z);  // derz_2dsim, src/derivation.f90:481
    }
    return G;
}

// Smagorinsky viscosity, src/les_turbulence.f90:70-88
__device__ __forceinline__ double smagorinsky(const Grad& G, double csd2) {
    const double s11 = G.d[0][0], s22 = G.d[1][1], s33 = G.d[2][2];
    const double s12 = 0.5 * (G.d[0][1] + G.d[1][0]);
    const double s13 = 0.5 * (G.d[0][2] + G.d[2][0]);
    const double s23 = 0.5 * (G.d[1][2] + G.d[2][1]);
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
    Coefs3 q;
    double onere, adu, bdu, cdu, csd2;
    int iles, sim2d;
    // ghost images of u* along each component's own axis (all that divergence(odd=1) needs,
    // src/differential_operators.f90:30-32): odd closure -> mirrored copies change sign
    Img2 ix, iy;
    int nz, bz_lo, bz_hi;
    long long sy_, sz_;
    double sgx, sgy, sgz_lo, sgz_hi;
    __device__ __forceinline__ void setup(const MarchGeom& g, int i, int j) {
        ix = image_offsets(i, g.nx, g.bx, g.bx);
        iy = image_offsets(j, g.ny, g.by, g.by);
        nz = g.nz, bz_lo = g.bz_lo, bz_hi = g.bz_hi;
        sy_ = g.sy, sz_ = g.sz;
        sgx = (g.bx == BM_MIRROR) ? -1.0 : 1.0;
        sgy = (g.by == BM_MIRROR) ? -1.0 : 1.0;
        sgz_lo = (g.bz_lo == BM_MIRROR) ? -1.0 : 1.0;
        sgz_hi = (g.bz_hi == BM_MIRROR) ? -1.0 : 1.0;
    }
    struct Pre {
        double f2v[3], f3v[3];
    };
    __device__ __forceinline__ Pre prefetch(long long m, bool ok) const {
        Pre p;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            p.f2v[c] = ok ? __ldg(f2[c] + m) : 0.0;
            p.f3v[c] = ok ? f3[c][m] : 0.0;  // plain load: f3 may alias f1 (read before written)
        }
        return p;
    }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int k,
                                          const Pre& pre) {
        const Grad G = gradient(r, q, sim2d);
        double nut = 0.0;
        if (iles) {
            nut = smagorinsky(G, csd2);
            nu_t[m] = nut;
        }
        const double nu_eff = onere + nut;  // src/integration.f90:114
        const double u0 = r.c(0), u1 = r.c(1), u2 = r.c(2);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double lx = r.d2x(c, q.x), ly = r.d2y(c, q.y);
            const double lz = sim2d ? 0.0 : r.d2z(c, q.z);
            // src/integration.f90:129-134 (and :149-154, :169-174)
            const double f = nu_eff * (lx + ly + lz) -
                             (u0 * G.d[c][0] + u1 * G.d[c][1] + u2 * G.d[c][2]);
            const double uc = (c == 0) ? u0 : (c == 1) ? u1 : u2;
            const double upv = uc + adu * f + bdu * pre.f2v[c] + cdu * pre.f3v[c];
            f1[c][m] = f;
            up[c][m] = upv;
            if (c == 0) {
                if (ix.lo) up[0][m + ix.lo] = sgx * upv;
                if (ix.hi) up[0][m + ix.hi] = sgx * upv;
            } else if (c == 1) {
                if (iy.lo) up[1][m + iy.lo * sy_] = sgy * upv;
                if (iy.hi) up[1][m + iy.hi * sy_] = sgy * upv;
            } else {
                const Img2 iz = image_offsets(k, nz, bz_lo, bz_hi);
                if (iz.lo) up[2][m + iz.lo * sz_] = sgz_lo * upv;
                if (iz.hi) up[2][m + iz.hi * sz_] = sgz_hi * upv;
            }
        }
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

struct NoPre {};

struct NutEpi {
    double* nu_t;
    Coefs3 q;
    double csd2;
    int sim2d;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int, const Pre&) {
        nu_t[m] = smagorinsky(gradient(r, q, sim2d), csd2);
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// rotational: every term uses the even closure (src/differential_operators.f90:64-74); the
// caller fills the ghost cells with even parity before the launch
struct RotEpi {
    double *rx, *ry, *rz;
    Coefs3 q;
    int sim2d;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int, const Pre&) {
        const Grad G = gradient(r, q, sim2d);
        rx[m] = G.d[2][1] - G.d[1][2];  // duzdy - duydz
        ry[m] = G.d[0][2] - G.d[2][0];  // duxdz - duzdx
        rz[m] = G.d[1][0] - G.d[0][1];  // duydx - duxdy
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

struct QEpi {
    double* qc;
    Coefs3 q;
    int sim2d;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int, const Pre&) {
        const Grad G = gradient(r, q, sim2d);
        // src/differential_operators.f90:103-104
        qc[m] = -(0.5 * (G.d[0][0] * G.d[0][0] + G.d[1][1] * G.d[1][1] + G.d[2][2] * G.d[2][2])) -
                G.d[0][1] * G.d[1][0] - G.d[0][2] * G.d[2][0] - G.d[1][2] * G.d[2][1];
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// statistics_calc: 16 sums (columns 2..17 of stats.dat), src/utils.f90:277-361
constexpr int NSTAT = 16;
struct StatsEpi {
    double* partial;  // [NSTAT][nblocks]
    Coefs3 q;
    double xnu;
    int sim2d;
    double acc[NSTAT];
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long, int, int, int, const Pre&) {
        const Grad G = gradient(r, q, sim2d);
        const double u0 = r.c(0), u1 = r.c(1), u2 = r.c(2);
        // order of acc: e_k, eps, eps2, dzeta, ux2, uy2, uz2, 9 x d1^2 (c major, axis minor)
        acc[0] += 0.5 * (u0 * u0 + u1 * u1 + u2 * u2);
        const double a = 2.0 * G.d[0][0], b = 2.0 * G.d[1][1], c = 2.0 * G.d[2][2];
        const double sxy = G.d[0][1] + G.d[1][0], sxz = G.d[0][2] + G.d[2][0],
                     syz = G.d[1][2] + G.d[2][1];
        acc[1] += 0.5 * xnu *
                  (a * a + b * b + c * c + 2.0 * (sxy * sxy) + 2.0 * (sxz * sxz) +
                   2.0 * (syz * syz));
        double lap[3];
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
            lap[cc] = r.d2x(cc, q.x) + r.d2y(cc, q.y) + (sim2d ? 0.0 : r.d2z(cc, q.z));
        acc[2] += (-xnu) * (u0 * lap[0] + u1 * lap[1] + u2 * lap[2]);
        const double wx = G.d[2][1] - G.d[1][2], wy = G.d[0][2] - G.d[2][0],
                     wz = G.d[1][0] - G.d[0][1];
        acc[3] += 0.5 * (wx * wx + wy * wy + wz * wz);
        acc[4] += u0 * u0;
        acc[5] += u1 * u1;
        acc[6] += u2 * u2;
#pragma unroll
        for (int cc = 0; cc < 3; ++cc)
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) acc[7 + 3 * cc + ax] += G.d[cc][ax] * G.d[cc][ax];
    }
    __device__ __forceinline__ void finish(int tid, double* smem) {
        // deterministic block reduction: warp shuffle, then fixed-order sum of the 8 warp sums
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
            for (int w = 0; w < MNT / 32; ++w) t += smem[tid * 8 + w];
            partial[(long long)tid * nblocks + b] = t;
        }
    }
};

MarchMaps<3> maps3(const FieldRef& a, const FieldRef& b, const FieldRef& c) {
    MarchMaps<3> m;
    m.m[0] = *a.tm, m.m[1] = *b.tm, m.m[2] = *c.tm;
    return m;
}

Coefs3 coefs(const Coef& cx, const Coef& cy, const Coef& cz) {
    Coefs3 q;
    q.x = cx, q.y = cy, q.z = cz;
    return q;
}

}  // namespace

int launch_rhs(cudaStream_t st, const Geom& g, const RhsArgs& r, int zmode, int zedge) {
    RhsEpi e;
    for (int c = 0; c < 3; ++c)
        e.f2[c] = r.f2[c], e.f3[c] = r.f3[c], e.f1[c] = r.f1[c], e.up[c] = r.up[c];
    e.nu_t = r.nu_t;
    e.q = coefs(r.cx, r.cy, r.cz);
    e.onere = r.onere, e.adu = r.adu, e.bdu = r.bdu, e.cdu = r.cdu, e.csd2 = r.csd2;
    e.iles = r.iles, e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, RhsEpi, 2>(st, g, maps3(r.u[0], r.u[1], r.u[2]), e, zmode, zedge);
}

int launch_nu_t(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                const Coef& cy, const Coef& cz, double csd2, double* nu_t) {
    NutEpi e;
    e.nu_t = nu_t, e.csd2 = csd2, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, NutEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

int launch_rot(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
               const Coef& cz, double* rotx, double* roty, double* rotz) {
    RotEpi e;
    e.rx = rotx, e.ry = roty, e.rz = rotz, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, RotEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

int launch_qcrit(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                 const Coef& cy, const Coef& cz, double* q) {
    QEpi e;
    e.qc = q, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, QEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

int stats_blocks(const Geom& g) {
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    const int zc = pick_zchunk(gx * gy, g.nz);
    return gx * gy * ((g.nz + zc - 1) / zc);
}

int launch_stats(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                 const Coef& cy, const Coef& cz, double xnu, double* partial) {
    StatsEpi e;
    e.partial = partial, e.xnu = xnu, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    for (int s = 0; s < NSTAT; ++s) e.acc[s] = 0.0;
    return launch_march<3, 0, 1, StatsEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

}  // namespace o3d
