// sor_common.cuh -- point-level rules of the red-black / wavefront SOR sweeps shared by
// sor_kernels.cu (one launch per colour class / pass) and sor_persist_kernel.cu (all iterations
// of a solve in one persistent launch): the reference's neighbour-index rule, point update,
// relaxation, exit tests and dynamic omega (src/poisson.f90:41-123).
#pragma once
#include "kernels.h"

namespace o3d {
namespace {

// (__host__ too: tests/cpu/sor_classes_test.cu checks on the CPU that every colour / seam class
// is an independent set under this neighbour rule)
__host__ __device__ __forceinline__ void nbr_idx(int p, int n, int mlo, int mhi, int& m1, int& p1) {
    // src/poisson.f90:57-66 (periodic) / :197-206 (mirrored); BM_HALO: stored ghost plane
    m1 = p - 1;
    p1 = p + 1;
    if (p == 0) {
        if (mlo == BM_WRAP) m1 = n - 1;
        else if (mlo == BM_MIRROR) m1 = 1;
    }
    if (p == n - 1) {
        if (mhi == BM_WRAP) p1 = 0;
        else if (mhi == BM_MIRROR) p1 = n - 2;
    }
}

__host__ __device__ __forceinline__ int seam_pop(const SorArgs& a, int i, int j, int gk) {
    return (a.seam_x && i == a.nx - 1) + (a.seam_y && j == a.ny - 1) +
           (a.seam_z && gk == a.gnz - 1);
}

// The reference's point update and relaxation, src/poisson.f90:95-98 and :102 (division by A: the
// bit-parity form used by the verification ordering), and the bit pattern -> double view of the
// residual accumulator.  __host__ too: tests/cpu/sor_sweep_test.cu runs the reference's
// lexicographic sweep with them on the CPU.
__host__ __device__ __forceinline__ double sor_pnew_ref(double ox, double oy, double oz, double pw,
                                                        double pe, double ps, double pn, double pb,
                                                        double pt, double rhs, double A) {
    return (-(ox * (pw + pe)) - oy * (ps + pn) - oz * (pb + pt) + rhs) / A;
}
__host__ __device__ __forceinline__ double sor_relax(double omega, double pc, double p_new) {
    return (1.0 - omega) * pc + omega * p_new;
}
__host__ __device__ __forceinline__ double bits_as_double(unsigned long long b) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    union {
        unsigned long long u;
        double d;
    } v;
    v.u = b;
    return v.d;
#endif
}

template <bool IMAGES = false>
__device__ __forceinline__ double sor_point(const SorArgs& a, int i, int j, int k,
                                            double omega) {
    const long long sy = a.sy, sz = a.sz;
    int im1, ip1, jm1, jp1, km1, kp1;
    nbr_idx(i, a.nx, a.mx, a.mx, im1, ip1);
    nbr_idx(j, a.ny, a.my, a.my, jm1, jp1);
    nbr_idx(k, a.nz, a.mz_lo, a.mz_hi, km1, kp1);
    const long long row = (long long)k * sz + (long long)j * sy;
    const long long m = row + i;
    const double pw = a.pp[row + im1], pe = a.pp[row + ip1];
    const double ps = a.pp[(long long)k * sz + (long long)jm1 * sy + i];
    const double pn = a.pp[(long long)k * sz + (long long)jp1 * sy + i];
    const double pb = a.pp[(long long)km1 * sz + (long long)j * sy + i];
    const double pt = a.pp[(long long)kp1 * sz + (long long)j * sy + i];
    const double pc = a.pp[m];
    // src/poisson.f90:95-98 with "/ A" replaced by "* (1/A)" (this ordering is not bit-parity)
    const double p_new = (-(a.oneondx2 * (pw + pe)) - a.oneondy2 * (ps + pn) -
                          a.oneondz2 * (pb + pt) + __ldg(a.rhs + m)) *
                         a.invA;
    const double v = (1.0 - omega) * pc + omega * p_new;  // :102
    a.pp[m] = v;
    if (IMAGES) {
        // keep the ghost cells of pp coherent (the fused TMA pass reads its closure from them)
        const Img2 ix = image_offsets(i, a.nx, a.mx, a.mx);
        const Img2 iy = image_offsets(j, a.ny, a.my, a.my);
        const Img2 iz = image_offsets(k, a.nz, a.mz_lo, a.mz_hi);
        store_images(a.pp, m, v, ix, iy.lo * sy, iy.hi * sy, iz.lo * sz, iz.hi * sz);
    }
    return fabs(p_new - pc);  // :100
}

__device__ __forceinline__ void seam_point_of(const SorArgs& a, long long t, long long nxf,
                                              long long nyf, long long nzf, int& i, int& j,
                                              int& k, bool& ok) {
    ok = false;
    i = j = k = -1;
    if (t < nxf) {  // x seam plane: (nx-1, j, k)
        i = a.nx - 1;
        j = (int)(t % a.ny);
        k = (int)(t / a.ny);
        ok = true;
    } else if ((t -= nxf) < nyf) {  // y seam plane, excluding points on the x seam
        j = a.ny - 1;
        i = (int)(t % a.nx);
        k = (int)(t / a.nx);
        ok = !(a.seam_x && i == a.nx - 1);
    } else if ((t -= nyf) < nzf) {  // z seam plane
        k = a.nz - 1;
        i = (int)(t % a.nx);
        j = (int)(t / a.nx);
        ok = !(a.seam_x && i == a.nx - 1) && !(a.seam_y && j == a.ny - 1);
    }
}

// src/poisson.f90:110-122, evaluated once per completed sweep
__host__ __device__ __forceinline__ void sor_control_step(SorCtrl* c, double eps, int kmax,
                                                          int idyn, double factor) {
    if (c->done) return;
    const double dmax = bits_as_double(c->dmax_bits);
    c->dmax_bits = 0ull;
    const int iter = c->iter + 1;
    c->iter = iter;
    c->dmax_last = dmax;
    if (dmax < eps) {  // :110
        c->done = 1;
        return;
    }
    if (fabs(c->dmax_old - dmax) < eps / 1000.0) {  // :111-114
        c->done = 2;
        return;
    }
    if (iter > 1 && idyn == 1) {  // :115-121
        if (dmax > c->dmax_old)
            c->omega = c->omega * (2.0 - factor);
        else if (dmax < 0.1 * c->dmax_old)
            c->omega = fmin(c->omega * factor, 2.0);
    }
    c->dmax_old = dmax;
    if (iter >= kmax) c->done = 3;  // loop exhausted
}
}  // namespace
}  // namespace o3d
