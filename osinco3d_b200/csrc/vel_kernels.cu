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
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "march.cuh"

namespace o3d {
namespace {

struct Coefs3 {
    Coef x, y, z;
};

// (struct Grad, smagorinsky, rhs_expr, predictor_expr: o3d_common.cuh)
template <class RG>
__device__ __forceinline__ Grad gradient(const RG& r, const Coefs3& q, int sim2d) {
    Grad G;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        G.d[c][0] = r.d1x(c, q.x);
        G.d[c][1] = r.d1y(c, q.y);
        G.d[c][2] = sim2d ? 0.0 : r.d1z(c, q.z);  // derz_2dsim, src/derivation.f90:481
    }
    return G;
}

// ---------------------------------------------------------------------------------------
// FAST: DNS in 3-D with sim2d / iles resolved at compile time (straight-line code, no basic-block
// breaks between the derivative evaluations).  !FAST: both are runtime flags; the branches keep
// the live ranges of the LES instantiation inside 128 registers (no spills).
// MODE 0: general (sim2d / iles runtime flags); 1: DNS in 3-D; 2: Smagorinsky LES in 3-D -- both
// compile-time, straight-line code on the split ring.
template <int MODE>
struct RhsEpi {
    static constexpr bool FAST = MODE != 0;
    static constexpr int STREAMS = 15;
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
    // image stores are the exception (points within 3 cells of a boundary): one flag per thread
    // for x/y and one plane test for z keep them out of the instruction stream of interior warps
    bool edge_xy;
    int zimg_lo, zimg_hi;
    __device__ __forceinline__ void setup(const MarchGeom& g, int i, int j) {
        ix = image_offsets(i, g.nx, g.bx, g.bx);
        iy = image_offsets(j, g.ny, g.by, g.by);
        nz = g.nz, bz_lo = g.bz_lo, bz_hi = g.bz_hi;
        sy_ = g.sy, sz_ = g.sz;
        sgx = (g.bx == BM_MIRROR) ? -1.0 : 1.0;
        sgy = (g.by == BM_MIRROR) ? -1.0 : 1.0;
        sgz_lo = (g.bz_lo == BM_MIRROR) ? -1.0 : 1.0;
        sgz_hi = (g.bz_hi == BM_MIRROR) ? -1.0 : 1.0;
        edge_xy = (ix.lo | ix.hi | iy.lo | iy.hi) != 0;
        // planes whose points have z images on this rank (walls / local periodic wrap)
        zimg_lo = (g.bz_lo == BM_MIRROR || g.bz_hi == BM_WRAP) ? R : -1;
        zimg_hi = (g.bz_hi == BM_MIRROR || g.bz_lo == BM_WRAP) ? g.nz - 1 - R : g.nz;
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
    template <class RG>
    __device__ __forceinline__ void apply(const RG& r, long long m, int, int, int k,
                                          const Pre& pre) {
        const Grad G = gradient(r, q, FAST ? 0 : sim2d);
        double nut = 0.0;
        if (MODE == 2 || (MODE == 0 && iles)) {
            nut = smagorinsky(G, csd2);
            nu_t[m] = nut;
        }
        const double nu_eff = onere + nut;  // src/integration.f90:114
        const double u0 = r.c(0), u1 = r.c(1), u2 = r.c(2);
        double ups[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double lx = r.d2x(c, q.x), ly = r.d2y(c, q.y);
            const double lz = (!FAST && sim2d) ? 0.0 : r.d2z(c, q.z);
            // src/integration.f90:129-134 (and :149-154, :169-174)
            const double f = rhs_expr(nu_eff, lx, ly, lz, u0, u1, u2, G.d[c][0], G.d[c][1], G.d[c][2]);
            const double uc = (c == 0) ? u0 : (c == 1) ? u1 : u2;
            const double upv = predictor_expr(uc, adu, f, bdu, pre.f2v[c], cdu, pre.f3v[c]);
            f1[c][m] = f;
            up[c][m] = upv;
            ups[c] = upv;
        }
        if (edge_xy || k <= zimg_lo || k >= zimg_hi) {  // boundary-adjacent points only
            if (ix.lo) up[0][m + ix.lo] = sgx * ups[0];
            if (ix.hi) up[0][m + ix.hi] = sgx * ups[0];
            if (iy.lo) up[1][m + iy.lo * sy_] = sgy * ups[1];
            if (iy.hi) up[1][m + iy.hi * sy_] = sgy * ups[1];
            const Img2 iz = image_offsets(k, nz, bz_lo, bz_hi);
            if (iz.lo) up[2][m + iz.lo * sz_] = sgz_lo * ups[2];
            if (iz.hi) up[2][m + iz.hi * sz_] = sgz_hi * ups[2];
        }
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// ---------------------------------------------------------------------------------------
// Role-split RHS (march_roles_kernel): thread role c in {0,1,2} owns velocity component c of its
// grid point -- grad(u_c), lap(u_c), f_c and u_c* -- exactly the per-component blocks of the
// reference (src/integration.f90:118-134, :138-154, :158-174).  The roles only meet for the
// Smagorinsky viscosity, which needs all nine first derivatives (src/les_turbulence.f90:55-88):
// each role publishes its three through shared memory.
struct RhsRoleEpi {
    const double* f2[3];
    const double* f3[3];
    double* f1[3];
    double* up[3];
    double* nu_t;
    Coefs3 q;
    double onere, adu, bdu, cdu, csd2;
    int iles, sim2d;
    // per thread
    int c, fo;                 // component, field offset in a ring stage
    const double *f2c, *f3c;
    double *f1c, *upc;
    long long img_lo, img_hi;  // own-axis ghost images of u* (x for c = 0, y for c = 1)
    double sg_xy;
    int nz, bz_lo, bz_hi;
    long long sz_;
    __device__ __forceinline__ void setup(const MarchGeom& g, int i, int j, int role) {
        c = role, fo = role * MFIELD;
        f2c = f2[role], f3c = f3[role], f1c = f1[role], upc = up[role];
        img_lo = img_hi = 0;
        sg_xy = 1.0;
        if (role == 0) {
            const Img2 ix = image_offsets(i, g.nx, g.bx, g.bx);
            img_lo = ix.lo, img_hi = ix.hi;
            sg_xy = (g.bx == BM_MIRROR) ? -1.0 : 1.0;
        } else if (role == 1) {
            const Img2 iy = image_offsets(j, g.ny, g.by, g.by);
            img_lo = iy.lo * g.sy, img_hi = iy.hi * g.sy;
            sg_xy = (g.by == BM_MIRROR) ? -1.0 : 1.0;
        }
        nz = g.nz, bz_lo = g.bz_lo, bz_hi = g.bz_hi, sz_ = g.sz;
    }
    struct Pre {
        double f2v, f3v;
    };
    __device__ __forceinline__ Pre prefetch(long long m, bool ok) const {
        Pre p;
        p.f2v = ok ? __ldg(f2c + m) : 0.0;
        p.f3v = ok ? f3c[m] : 0.0;  // plain load: f3 may alias f1 (read before written, per point)
        return p;
    }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int k, const Pre& pre,
                                          bool ok, int pt, double* xch) {
        const double g0 = r.d1x(c, q.x);
        const double g1 = r.d1y(c, q.y);
        const double g2 = sim2d ? 0.0 : r.d1z(c, q.z);  // derz_2dsim, src/derivation.f90:481
        double nut = 0.0;
        if (iles) {  // uniform across the CTA
            xch[(3 * c + 0) * MNT + pt] = g0;
            xch[(3 * c + 1) * MNT + pt] = g1;
            xch[(3 * c + 2) * MNT + pt] = g2;
            __syncthreads();
            Grad G;
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) G.d[a][b] = xch[(3 * a + b) * MNT + pt];
            nut = smagorinsky(G, csd2);
            if (ok && c == 0) nu_t[m] = nut;
        }
        const double nu_eff = onere + nut;  // src/integration.f90:114
        const double u0 = r.c(0), u1 = r.c(1), u2 = r.c(2);
        const double lx = r.d2x(c, q.x), ly = r.d2y(c, q.y);
        const double lz = sim2d ? 0.0 : r.d2z(c, q.z);
        // src/integration.f90:129-134 (and :149-154, :169-174)
        const double f = rhs_expr(nu_eff, lx, ly, lz, u0, u1, u2, g0, g1, g2);
        const double uc = (c == 0) ? u0 : (c == 1) ? u1 : u2;
        const double upv = predictor_expr(uc, adu, f, bdu, pre.f2v, cdu, pre.f3v);
        if (!ok) return;
        f1c[m] = f;
        upc[m] = upv;
        // ghost images of u* along the component's own axis (odd closure: mirrored copies change
        // sign) -- all that divergence(odd = 1) needs, src/differential_operators.f90:30-32
        if (c < 2) {
            if (img_lo) upc[m + img_lo] = sg_xy * upv;
            if (img_hi) upc[m + img_hi] = sg_xy * upv;
        } else {
            const Img2 iz = image_offsets(k, nz, bz_lo, bz_hi);
            if (iz.lo) upc[m + iz.lo * sz_] = (bz_lo == BM_MIRROR) ? -upv : upv;
            if (iz.hi) upc[m + iz.hi * sz_] = (bz_hi == BM_MIRROR) ? -upv : upv;
        }
    }
};

struct NoPre {};

struct NutEpi {
    static constexpr int STREAMS = 4;
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
    static constexpr int STREAMS = 6;
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

// vorticity magnitude as written by write_all_data (src/visualization.f90:258-259):
// sqrt(rotx**2 + roty**2 + rotz**2) of the curl above, fused -- 32 B/pt instead of curl (48) + a
// 32 B/pt magnitude pass; same even closure as RotEpi
struct VortEpi {
    static constexpr int STREAMS = 4;
    double* vm;
    Coefs3 q;
    int sim2d;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int, const Pre&) {
        const Grad G = gradient(r, q, sim2d);
        const double rx = G.d[2][1] - G.d[1][2];
        const double ry = G.d[0][2] - G.d[2][0];
        const double rz = G.d[1][0] - G.d[0][1];
        vm[m] = sqrt(rx * rx + ry * ry + rz * rz);
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

struct QEpi {
    static constexpr int STREAMS = 4;
    double* qc;
    Coefs3 q;
    int sim2d;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom&, int, int) {}
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    __device__ __forceinline__ void apply(const Ring<3>& r, long long m, int, int, int, const Pre&) {
        const Grad G = gradient(r, q, sim2d);
        qc[m] = q_criterion_expr(G);  // src/differential_operators.f90:103-104
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

// statistics_calc: 16 sums (columns 2..17 of stats.dat), src/utils.f90:277-361
constexpr int NSTAT = 16;
struct StatsEpi {
    static constexpr int STREAMS = 3;
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

// Per-step diagnostics of the driver in ONE pass over a velocity triple:
// divergence(odd = 1) (src/differential_operators.f90:25-35) fed straight into function_stats
// (src/functions.f90:27-63: min, max with first-occurrence position, sum), plus minval / maxval
// of each component (print_velocity_values, src/IOfunctions.f90:322) and maxval(abs())
// (compute_cfl, src/utils.f90:199-201).  24 B/pt instead of divergence (32) + function_stats (8)
// + nine reductions (72).
constexpr int NDIAG = 10;  // dmin dmax dsum dlin | umin[3] | umax[3]   (max|u| = max(-min, max))
struct DiagEpi {
    static constexpr int STREAMS = 3;
    double* partial;  // [NDIAG][nblocks]
    Coefs3 q;
    int sim2d, nx, ny, gz0;
    double dmin, dmax, dsum, umin[3], umax[3];
    long long dlin;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom& g, int, int) { nx = g.nx, ny = g.ny; }
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    // z-field 0 = uz (7-plane window); centre-only fields 0, 1 = ux, uy (x / y halo of the plane
    // being computed): the ring layout of the divergence kernel, 3 CTAs per SM
    template <class RG>
    __device__ __forceinline__ void apply(const RG& r, long long, int i, int j, int k,
                                          const Pre&) {
        const double dfx = r.c_d1x(0, q.x), dfy = r.c_d1y(1, q.y);
        const double dfz = sim2d ? 0.0 : r.d1z(0, q.z);
        const double dv = dfx + dfy + dfz;  // src/differential_operators.f90:35
        const long long lin = ((long long)(gz0 + k) * ny + j) * nx + i;  // array-order position
        dmin = fmin(dmin, dv);
        dsum += dv;
        if (dv > dmax) dmax = dv, dlin = lin;  // k ascending per thread: first occurrence
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double u = (c == 2) ? r.z(0, 0) : r.cx(c, 0);
            umin[c] = fmin(umin[c], u);
            umax[c] = fmax(umax[c], u);
        }
    }
    __device__ __forceinline__ void finish(int tid, double* smem) {
        const int nblocks = gridDim.x * gridDim.y * gridDim.z;
        const int b = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
        // warp level: min / max / sum by shuffles; the arg-max keeps the smaller position on ties
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            const double om = __shfl_xor_sync(0xffffffffu, dmax, o);
            const long long ol = __shfl_xor_sync(0xffffffffu, dlin, o);
            if (om > dmax || (om == dmax && ol < dlin)) dmax = om, dlin = ol;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                umin[c] = fmin(umin[c], __shfl_xor_sync(0xffffffffu, umin[c], o));
                umax[c] = fmax(umax[c], __shfl_xor_sync(0xffffffffu, umax[c], o));
            }
        }
        dsum = warp_sum(dsum);
        if ((tid & 31) == 0) {
            double* w = smem + (tid >> 5) * NDIAG;
            w[0] = dmin, w[1] = dmax, w[2] = dsum, w[3] = (double)dlin;
#pragma unroll
            for (int c = 0; c < 3; ++c) w[4 + c] = umin[c], w[7 + c] = umax[c];
        }
        __syncthreads();
        if (tid == 0) {
            double v[NDIAG];
            for (int s = 0; s < NDIAG; ++s) v[s] = smem[s];
            for (int w = 1; w < MNT / 32; ++w) {  // fixed order: deterministic
                const double* x = smem + w * NDIAG;
                v[0] = fmin(v[0], x[0]);
                if (x[1] > v[1] || (x[1] == v[1] && x[3] < v[3])) v[1] = x[1], v[3] = x[3];
                v[2] += x[2];
                for (int c = 0; c < 3; ++c) {
                    v[4 + c] = fmin(v[4 + c], x[4 + c]);
                    v[7 + c] = fmax(v[7 + c], x[7 + c]);
                }
            }
            for (int s = 0; s < NDIAG; ++s) partial[(long long)s * nblocks + b] = v[s];
        }
    }
};

// merge the per-CTA partials of DiagEpi in a fixed order -> out13 = dmin dmax dsum dlin |
// umin[3] | umax[3] | max|u|[3]
__global__ void __launch_bounds__(256) diag_stage2(const double* partial, int nblocks,
                                                   double* out) {
    __shared__ double red[8][NDIAG];
    const int tid = threadIdx.x;
    double v[NDIAG];
    v[0] = 1.7976931348623157e308, v[1] = -1.7976931348623157e308, v[2] = 0.0;
    v[3] = 9.0e18;
    for (int c = 0; c < 3; ++c) v[4 + c] = 1.7976931348623157e308, v[7 + c] = -1.7976931348623157e308;
    auto merge = [&](const double* x) {
        v[0] = fmin(v[0], x[0]);
        if (x[1] > v[1] || (x[1] == v[1] && x[3] < v[3])) v[1] = x[1], v[3] = x[3];
        v[2] += x[2];
        for (int c = 0; c < 3; ++c) {
            v[4 + c] = fmin(v[4 + c], x[4 + c]);
            v[7 + c] = fmax(v[7 + c], x[7 + c]);
        }
    };
    for (int b = tid; b < nblocks; b += 256) {
        double x[NDIAG];
        for (int s = 0; s < NDIAG; ++s) x[s] = partial[(long long)s * nblocks + b];
        merge(x);
    }
    // butterfly inside the warp (the lower lane of a pair merges first: fixed association)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double x[NDIAG];
#pragma unroll
        for (int s = 0; s < NDIAG; ++s) x[s] = __shfl_xor_sync(0xffffffffu, v[s], o);
        if (tid & o) {  // keep "lower lane first" so that both partners compute the same sum
            double t[NDIAG];
            for (int s = 0; s < NDIAG; ++s) t[s] = v[s], v[s] = x[s];
            merge(t);
        } else {
            merge(x);
        }
    }
    if ((tid & 31) == 0)
        for (int s = 0; s < NDIAG; ++s) red[tid >> 5][s] = v[s];
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w) merge(red[w]);
        for (int s = 0; s < NDIAG; ++s) out[s] = v[s];
        for (int c = 0; c < 3; ++c) out[10 + c] = fmax(-v[4 + c], v[7 + c]);  // maxval(abs(u))
    }
}

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

template <int MODE>
static int launch_rhs_t(cudaStream_t st, const Geom& g, const RhsArgs& r, int zmode, int zedge) {
    constexpr bool FAST = MODE != 0;
    RhsEpi<MODE> e;
    for (int c = 0; c < 3; ++c)
        e.f2[c] = r.f2[c], e.f3[c] = r.f3[c], e.f1[c] = r.f1[c], e.up[c] = r.up[c];
    e.nu_t = r.nu_t;
    e.q = coefs(r.cx, r.cy, r.cz);
    e.onere = r.onere, e.adu = r.adu, e.bdu = r.bdu, e.cdu = r.cdu, e.csd2 = r.csd2;
    e.iles = r.iles, e.sim2d = g.sim2d;
    // O3D_RHS_RING=split<P>: split-ring staging (march.cuh RingCW: halo'd box for plane k only,
    // tile-only boxes for the z window), P = 2 or 3 planes of prefetch instead of 1
    // Measured (profiles/r1r_variants.txt): split2 0.405 vs 0.435 ms at 256^3 and 3.13 vs 3.38 ms
    // at 512^3 for the DNS instantiation (default there); the LES instantiation spills in the
    // split layout and stays on the classic ring.  O3D_RHS_RING=classic | split2 | split3 forces.
    static const char* ring = getenv("O3D_RHS_RING");
    const bool split = ring ? !strncmp(ring, "split", 5) : FAST;
    if (split) {
        MarchMaps<6> m6;
        for (int c = 0; c < 3; ++c) m6.m[c] = *r.u[c].tm, m6.m[3 + c] = *r.u[c].tms;
        if (ring && ring[5] == '3')
            return launch_march<0, 3, 3, RhsEpi<MODE>, 2, 0, 0, 1, 3>(st, g, m6, e, zmode, zedge);
        return launch_march<0, 3, 2, RhsEpi<MODE>, 2, 0, 0, 1, 3>(st, g, m6, e, zmode, zedge);
    }
    // O3D_RHS_UNROLL=1: plane loop unrolled over the 8-stage ring (march.cuh, UNR = 8: ring
    // positions become immediates).  Measured equal at 256^3 and 1-2 % slower at 512^3 than the
    // rolled loop (the kernel is not instruction-bound), so the rolled loop is the default.
    static const bool unrolled = getenv("O3D_RHS_UNROLL") && atoi(getenv("O3D_RHS_UNROLL")) == 1;
    if (unrolled)
        return launch_march<3, 0, 1, RhsEpi<MODE>, 2, 0, 0, 8>(st, g, maps3(r.u[0], r.u[1], r.u[2]),
                                                               e, zmode, zedge);
    return launch_march<3, 0, 1, RhsEpi<MODE>, 2>(st, g, maps3(r.u[0], r.u[1], r.u[2]), e, zmode,
                                                  zedge);
}

int launch_rhs(cudaStream_t st, const Geom& g, const RhsArgs& r, int zmode, int zedge) {
    static const int variant = getenv("O3D_RHS_VARIANT") ? atoi(getenv("O3D_RHS_VARIANT")) : 0;
    if (variant >= 2) {
        RhsRoleEpi w;
        for (int c = 0; c < 3; ++c)
            w.f2[c] = r.f2[c], w.f3[c] = r.f3[c], w.f1[c] = r.f1[c], w.up[c] = r.up[c];
        w.nu_t = r.nu_t, w.q = coefs(r.cx, r.cy, r.cz);
        w.onere = r.onere, w.adu = r.adu, w.bdu = r.bdu, w.cdu = r.cdu, w.csd2 = r.csd2;
        w.iles = r.iles, w.sim2d = g.sim2d;
        const MarchMaps<3> mm = maps3(r.u[0], r.u[1], r.u[2]);
        return launch_march_roles<3, 2, 3, RhsRoleEpi>(st, g, mm, w, zmode, zedge);
    }
    if (!g.sim2d && !r.iles) return launch_rhs_t<1>(st, g, r, zmode, zedge);
    // MODE 2 (LES resolved at compile time, split ring) is selectable with O3D_RHS_LES=static for
    // A/B runs only: straight-line LES code holds the nine first derivatives across the Smagorinsky
    // evaluation AND the three convective terms, ptxas spills 120 B at the 128-register cap and the
    // kernel is slower than the branchy general instantiation on the classic ring (512^3, round 2:
    // 4.48 ms split2 / 4.81 split3 / 4.94 classic against 3.81 ms general, profiles/r2g_les_variants.txt)
    static const bool les_static = getenv("O3D_RHS_LES") && !strcmp(getenv("O3D_RHS_LES"), "static");
    if (!g.sim2d && r.iles && les_static) return launch_rhs_t<2>(st, g, r, zmode, zedge);
    return launch_rhs_t<0>(st, g, r, zmode, zedge);
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

int launch_vort(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
                const Coef& cz, double* vm) {
    VortEpi e;
    e.vm = vm, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, VortEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

int launch_qcrit(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx,
                 const Coef& cy, const Coef& cz, double* q) {
    QEpi e;
    e.qc = q, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d;
    return launch_march<3, 0, 1, QEpi, 2>(st, g, maps3(u[0], u[1], u[2]), e);
}

int diag_blocks(const Geom& g) {
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    const int zc = pick_zchunk(gx * gy, g.nz, 3, 1, DiagEpi::STREAMS);  // as launch_march below
    return gx * gy * ((g.nz + zc - 1) / zc);
}

int launch_diag(cudaStream_t st, const Geom& g, const FieldRef* u, const Coef& cx, const Coef& cy,
                const Coef& cz, double* partial, double* out13) {
    DiagEpi e;
    e.partial = partial, e.q = coefs(cx, cy, cz), e.sim2d = g.sim2d, e.gz0 = g.gz0;
    e.dmin = 1.7976931348623157e308, e.dmax = -1.7976931348623157e308, e.dsum = 0.0;
    e.dlin = 0x7fffffffffffffffLL;
    for (int c = 0; c < 3; ++c)
        e.umin[c] = 1.7976931348623157e308, e.umax[c] = -1.7976931348623157e308;
    MarchMaps<3> ms;  // ux, uy with halo (c-ring); uz tile-only (w-ring), as the divergence kernel
    ms.m[0] = *u[0].tm, ms.m[1] = *u[1].tm, ms.m[2] = *u[2].tms;
    if (launch_march<0, 2, 2, DiagEpi, 3, 0, 0, 1, 1>(st, g, ms, e)) return 1;
    diag_stage2<<<1, 256, 0, st>>>(partial, diag_blocks(g), out13);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int stats_blocks(const Geom& g) {
    const int gx = (g.nx + MTX - 1) / MTX, gy = (g.ny + MTY - 1) / MTY;
    const int zc = pick_zchunk(gx * gy, g.nz, 2, 3, StatsEpi::STREAMS);  // as launch_march<3,0,1,StatsEpi,2>
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
