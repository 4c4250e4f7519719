// proj_kernels.cu -- the two stencil kernels of the projection step:
//   div_kernel : divergence (src/differential_operators.f90:7-38), optionally /dt -> Poisson
//             right-hand side (src/integration.f90:234-239).  3 sweeps + 3 temporaries in the
//             reference; here 3 reads + 1 write = 32 B/pt.  odd/even closure = ghost parity.
//   CorrEpi (march engine): u = u* - dt grad p with the NaN / >1000 guard fused in
//             (src/integration.f90:298-325).  pp(1)+u*(3) reads, u(3) writes = 56 B/pt.
//             u* arrives through TMA stream fields (32 x 8 boxes, 3 planes ahead).
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "march.cuh"

namespace o3d {
namespace {

struct NoPre {};

// Divergence: fz needs the z window, fx and fy only the plane being computed (centre-only ring
// of the march engine), prefetched 2 planes ahead.
// S2 = sim2d resolved at compile time (no basic-block break in the common 3-D case)
template <bool S2>
struct DivEpi {
    static constexpr int STREAMS = 4;
    double* out;
    Coef cx, cy, cz;
    double dt;
    int divide, sim2d;
    // even ghost images of the result (faces): the Poisson right-hand side is read by the SOR
    // pass through its ghost cells (the index rule of src/poisson.f90:57-92 as data)
    Img2 ix, iy;
    int nz, bz_lo, bz_hi;
    long long sy_, sz_;
    // image stores are the exception: one flag per thread for x/y, one plane test for z
    bool edge_xy;
    int zimg_lo, zimg_hi;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom& g, int i, int j) {
        ix = image_offsets(i, g.nx, g.bx, g.bx);
        iy = image_offsets(j, g.ny, g.by, g.by);
        nz = g.nz, bz_lo = g.bz_lo, bz_hi = g.bz_hi;
        sy_ = g.sy, sz_ = g.sz;
        edge_xy = (ix.lo | ix.hi | iy.lo | iy.hi) != 0;
        zimg_lo = (g.bz_lo == BM_MIRROR || g.bz_hi == BM_WRAP) ? R : -1;
        zimg_hi = (g.bz_hi == BM_MIRROR || g.bz_lo == BM_WRAP) ? g.nz - 1 - R : g.nz;
    }
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    template <class RG>
    __device__ __forceinline__ void apply(const RG& r, long long m, int, int, int k,
                                          const Pre&) {
        const double dfx = r.c_d1x(0, cx);
        const double dfy = r.c_d1y(1, cy);
        const double dfz = S2 ? 0.0 : r.d1z(0, cz);
        // src/differential_operators.f90:35, src/integration.f90:239
        const double v = div_expr(dfx, dfy, dfz, divide, dt);
        out[m] = v;
        if (edge_xy || k <= zimg_lo || k >= zimg_hi) {  // boundary-adjacent points only
            const Img2 iz = image_offsets(k, nz, bz_lo, bz_hi);
            if (ix.lo) out[m + ix.lo] = v;
            if (ix.hi) out[m + ix.hi] = v;
            if (iy.lo) out[m + iy.lo * sy_] = v;
            if (iy.hi) out[m + iy.hi * sy_] = v;
            if (iz.lo) out[m + iz.lo * sz_] = v;
            if (iz.hi) out[m + iz.hi * sz_] = v;
        }
    }
    __device__ __forceinline__ void finish(int, double*) {}
};

template <bool S2>
struct CorrEpi {
    static constexpr int STREAMS = 7;
    double* u[3];
    Coef cx, cy, cz;
    double dt;
    int* flag;
    int sim2d;
    int bad;
    // ghost images of the corrected velocity, natural parity (component c is odd along axis c):
    // the next RHS / statistics launch finds its closure already in the ghost cells
    Img2 ix, iy;
    int nz, bz_lo, bz_hi;
    long long sy_, sz_;
    bool mx, my, mz_lo, mz_hi;  // mirrored sides
    bool edge_xy;
    int zimg_lo, zimg_hi;
    typedef NoPre Pre;
    __device__ __forceinline__ void setup(const MarchGeom& g, int i, int j) {
        ix = image_offsets(i, g.nx, g.bx, g.bx);
        iy = image_offsets(j, g.ny, g.by, g.by);
        nz = g.nz, bz_lo = g.bz_lo, bz_hi = g.bz_hi;
        sy_ = g.sy, sz_ = g.sz;
        mx = g.bx == BM_MIRROR, my = g.by == BM_MIRROR;
        mz_lo = g.bz_lo == BM_MIRROR, mz_hi = g.bz_hi == BM_MIRROR;
        edge_xy = (ix.lo | ix.hi | iy.lo | iy.hi) != 0;
        zimg_lo = (g.bz_lo == BM_MIRROR || g.bz_hi == BM_WRAP) ? R : -1;
        zimg_hi = (g.bz_hi == BM_MIRROR || g.bz_lo == BM_WRAP) ? g.nz - 1 - R : g.nz;
    }
    __device__ __forceinline__ Pre prefetch(long long, bool) const { return Pre(); }
    // z-field 0 = pp (7-plane window); stream fields 0..2 = u* (TMA, 3 planes ahead)
    __device__ __forceinline__ void apply(const Ring<1>& r, long long m, int, int, int k,
                                          const Pre&) {
        // src/integration.f90:298-300 (derxp, deryp, derzp)
        const double dpdx = r.d1x(0, cx);
        const double dpdy = r.d1y(0, cy);
        const double dpdz = S2 ? 0.0 : r.d1z(0, cz);
        // src/integration.f90:304-306
        const double u0 = corr_expr(r.st(0), dt, dpdx);
        const double u1 = corr_expr(r.st(1), dt, dpdy);
        const double u2 = corr_expr(r.st(2), dt, dpdz);
        u[0][m] = u0;
        u[1][m] = u1;
        u[2][m] = u2;
        if (edge_xy || k <= zimg_lo || k >= zimg_hi) {  // boundary-adjacent points only
            const Img2 iz = image_offsets(k, nz, bz_lo, bz_hi);
            const double v[3] = {u0, u1, u2};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const double vx = (mx && c == 0) ? -v[c] : v[c];
                if (ix.lo) u[c][m + ix.lo] = vx;
                if (ix.hi) u[c][m + ix.hi] = vx;
                const double vy = (my && c == 1) ? -v[c] : v[c];
                if (iy.lo) u[c][m + iy.lo * sy_] = vy;
                if (iy.hi) u[c][m + iy.hi * sy_] = vy;
                if (iz.lo) u[c][m + iz.lo * sz_] = (mz_lo && c == 2) ? -v[c] : v[c];
                if (iz.hi) u[c][m + iz.hi * sz_] = (mz_hi && c == 2) ? -v[c] : v[c];
            }
        }
        // src/integration.f90:309-325: contains_nan .or. maxval > 1000
        if (u0 != u0 || u1 != u1 || u2 != u2 || u0 > 1000. || u1 > 1000. || u2 > 1000.) bad = 1;
    }
    __device__ __forceinline__ void finish(int tid, double*) {
        const int b = warp_or(bad);
        if (b && (tid & 31) == 0) atomicOr(flag, 1);
    }
};

}  // namespace

template <bool S2>
static int launch_div_t(cudaStream_t st, const Geom& g, const FieldRef* f, const Coef& cx,
                        const Coef& cy, const Coef& cz, int divide_by_dt, double dt, double* out,
                        int zmode, int zedge) {
    DivEpi<S2> e;
    e.out = out;
    e.cx = cx, e.cy = cy, e.cz = cz;
    e.divide = divide_by_dt, e.dt = dt, e.sim2d = g.sim2d;
    MarchMaps<3> m;
    m.m[0] = *f[2].tm;  // z-window field first
    m.m[1] = *f[0].tm;
    m.m[2] = *f[1].tm;
    // fz is only differentiated in z: staged without halo (32 x 8 boxes, march.cuh RingCW) --
    // 18 KB instead of 40 KB of ring and no halo re-reads for that field: 0.126 -> 0.113 ms at
    // 256^3, 0.885 -> 0.803 ms at 512^3.  O3D_DIV_RING=classic: all three fields with halo.
    static const bool split = !(getenv("O3D_DIV_RING") && !strcmp(getenv("O3D_DIV_RING"), "classic"));
    if (split) {
        MarchMaps<3> ms;
        ms.m[0] = *f[0].tm, ms.m[1] = *f[1].tm, ms.m[2] = *f[2].tms;
        return launch_march<0, 2, 2, DivEpi<S2>, 3, 0, 0, 1, 1>(st, g, ms, e, zmode, zedge);
    }
    // O3D_DIV_UNROLL=1: plane loop unrolled over the 9-stage ring (march.cuh UNR)
    static const bool unr = getenv("O3D_DIV_UNROLL") && atoi(getenv("O3D_DIV_UNROLL")) == 1;
    if (unr) return launch_march<1, 2, 2, DivEpi<S2>, 3, 0, 0, 9>(st, g, m, e, zmode, zedge);
    return launch_march<1, 2, 2, DivEpi<S2>, 3>(st, g, m, e, zmode, zedge);
}

int launch_div(cudaStream_t st, const Geom& g, const FieldRef* f, const Coef& cx, const Coef& cy,
               const Coef& cz, int divide_by_dt, double dt, double* out, int zmode, int zedge) {
    return g.sim2d ? launch_div_t<true>(st, g, f, cx, cy, cz, divide_by_dt, dt, out, zmode, zedge)
                   : launch_div_t<false>(st, g, f, cx, cy, cz, divide_by_dt, dt, out, zmode, zedge);
}

template <bool S2>
static int launch_corr_t(cudaStream_t st, const Geom& g, const FieldRef& pp, const FieldRef* up,
                         double* const* u, const Coef& cx, const Coef& cy, const Coef& cz,
                         double dt, int* flag, int zmode, int zedge, const FieldRef* pp_alt,
                         const SorCtrl* gate) {
    CorrEpi<S2> e;
    for (int c = 0; c < 3; ++c) e.u[c] = u[c];
    e.cx = cx, e.cy = cy, e.cz = cz;
    e.dt = dt, e.flag = flag, e.sim2d = g.sim2d, e.bad = 0;
    if (gate) {  // gated on the SOR control block; pp or pp_alt by the parity of the pass count
        MarchMaps<5> m;
        m.m[0] = *pp.tm;
        for (int c = 0; c < 3; ++c) m.m[1 + c] = *up[c].tms;
        m.m[4] = *pp_alt->tm;
        return launch_march<1, 0, 3, CorrEpi<S2>, 3, 3, 1>(st, g, m, e, zmode, zedge, gate);
    }
    MarchMaps<4> m;
    m.m[0] = *pp.tm;
    for (int c = 0; c < 3; ++c) m.m[1 + c] = *up[c].tms;
    // O3D_CORR_VARIANT: 0 = 3 planes ahead, rolled loop (default); 1 = same, unrolled over the
    // rings (20 planes); 2 = 2 planes ahead, unrolled (9 planes); 3 = 2 planes ahead, rolled
    static const int variant = getenv("O3D_CORR_VARIANT") ? atoi(getenv("O3D_CORR_VARIANT")) : 0;
    if (variant == 1)
        return launch_march<1, 0, 3, CorrEpi<S2>, 3, 3, 0, 20>(st, g, m, e, zmode, zedge);
    if (variant == 2)
        return launch_march<1, 0, 2, CorrEpi<S2>, 3, 3, 0, 9>(st, g, m, e, zmode, zedge);
    if (variant == 3) return launch_march<1, 0, 2, CorrEpi<S2>, 3, 3>(st, g, m, e, zmode, zedge);
    return launch_march<1, 0, 3, CorrEpi<S2>, 3, 3>(st, g, m, e, zmode, zedge);
}

int launch_corr(cudaStream_t st, const Geom& g, const FieldRef& pp, const FieldRef* up,
                double* const* u, const Coef& cx, const Coef& cy, const Coef& cz, double dt,
                int* flag, int zmode, int zedge, const FieldRef* pp_alt, const SorCtrl* gate) {
    return g.sim2d ? launch_corr_t<true>(st, g, pp, up, u, cx, cy, cz, dt, flag, zmode, zedge,
                                         pp_alt, gate)
                   : launch_corr_t<false>(st, g, pp, up, u, cx, cy, cz, dt, flag, zmode, zedge,
                                          pp_alt, gate);
}

}  // namespace o3d
