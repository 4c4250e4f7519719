// mg_kernels.cu -- device kernels of the geometric multigrid V-cycle that replaces the
// reference's solve_poisson_multigrid (src/poisson_multigrid.f90:10-189; that routine is
// undefined behaviour as called).  Same 7-point operator and neighbour rule
// as the SOR solvers (src/poisson.f90:42-51,57-92), on every level.
//
//   mg_residual  : r = rhs - L p  (+ max|r| for the stopping test)      24 B/pt
//   mg_restrict  : coarse rhs = R r (tensor product of 1-D tables), coarse p = 0
//   mg_prolong   : p += P e       (tensor-product linear interpolation)
//   mg_coarse    : coarsest level, one CTA: compatibility projection of the singular system +
//                  red-black Gauss-Seidel sweeps entirely inside the CTA
// The smoother on every other level is the red-black SOR half-sweep of sor_kernels.cu with
// omega = 1 (seam classes included: coarse periodic extents are usually odd).
// HBM-bound integer/FP64 streaming work: no tensor cores.
#include "kernels.h"

namespace o3d {
namespace {

__device__ __forceinline__ void mg_nbr(int p, int n, int mlo, int mhi, int& m1, int& p1) {
    // src/poisson.f90:57-66 (periodic) / :197-206 (mirrored); BM_HALO: stored ghost plane
    m1 = p - 1, p1 = p + 1;
    if (p == 0 && mlo != BM_HALO) m1 = (mlo == BM_WRAP) ? n - 1 : 1;
    if (p == n - 1 && mhi != BM_HALO) p1 = (mhi == BM_WRAP) ? 0 : n - 2;
}

__global__ void __launch_bounds__(256) mg_residual_kernel(const MgGrid g,
                                                          const double* __restrict__ p,
                                                          const double* __restrict__ rhs,
                                                          double* __restrict__ res,
                                                          unsigned long long* maxbits) {
    __shared__ double red[32];
    double dmax = 0.0;
    const long long rows = (long long)g.ny * g.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % g.ny), k = (int)(row / g.ny);
        int jm, jp, km, kp;
        mg_nbr(j, g.ny, g.my, g.my, jm, jp);
        mg_nbr(k, g.nz, g.mz_lo, g.mz_hi, km, kp);
        const long long base = (long long)k * g.sz + (long long)j * g.sy;
        const double* ps = p + (long long)k * g.sz + (long long)jm * g.sy;
        const double* pn = p + (long long)k * g.sz + (long long)jp * g.sy;
        const double* pb = p + (long long)km * g.sz + (long long)j * g.sy;
        const double* pt = p + (long long)kp * g.sz + (long long)j * g.sy;
        for (int i = threadIdx.x; i < g.nx; i += blockDim.x) {
            int im, ip;
            mg_nbr(i, g.nx, g.mx, g.mx, im, ip);
            const double lp = g.ox * (p[base + im] + p[base + ip]) + g.oy * (ps[i] + pn[i]) +
                              g.oz * (pb[i] + pt[i]) + g.A * p[base + i];
            const double r = __ldg(rhs + base + i) - lp;
            if (res) res[base + i] = r;
            dmax = fmax(dmax, fabs(r));
        }
    }
    if (maxbits) {
        const double bm = block_max(dmax, red);
        if (threadIdx.x == 0 && bm > 0.0) atomic_max_nonneg(maxbits, bm);
    }
}

// one thread per coarse point; 4 taps per axis (weights may be zero)
__global__ void __launch_bounds__(256) mg_restrict_kernel(const MgGrid f, const MgGrid c,
                                                          const MgTables t,
                                                          const double* __restrict__ res,
                                                          double* __restrict__ rhs_c,
                                                          double* __restrict__ p_c, int ck0,
                                                          int nck) {
    const long long rows = (long long)c.ny * nck;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int cj = (int)(row % c.ny), ck = ck0 + (int)(row / c.ny);
        const long long cbase = (long long)ck * c.sz + (long long)cj * c.sy;
        for (int ci = threadIdx.x; ci < c.nx; ci += blockDim.x) {
            double acc = 0.0;
#pragma unroll
            for (int tz = 0; tz < 4; ++tz) {
                const double wz = __ldg(t.rw[2] + 4 * ck + tz);
                if (wz == 0.0) continue;
                const long long oz = (long long)__ldg(t.ridx[2] + 4 * ck + tz) * f.sz;
                double accy = 0.0;
#pragma unroll
                for (int ty = 0; ty < 4; ++ty) {
                    const double wy = __ldg(t.rw[1] + 4 * cj + ty);
                    if (wy == 0.0) continue;
                    const long long oy = oz + (long long)__ldg(t.ridx[1] + 4 * cj + ty) * f.sy;
                    double accx = 0.0;
#pragma unroll
                    for (int tx = 0; tx < 4; ++tx) {
                        const double wx = __ldg(t.rw[0] + 4 * ci + tx);
                        if (wx == 0.0) continue;
                        accx += wx * __ldg(res + oy + __ldg(t.ridx[0] + 4 * ci + tx));
                    }
                    accy += wy * accx;
                }
                acc += wz * accy;
            }
            rhs_c[cbase + ci] = acc;
            p_c[cbase + ci] = 0.0;
        }
    }
}

// one thread per fine point: p += trilinear interpolation of the coarse correction
__global__ void __launch_bounds__(256) mg_prolong_kernel(const MgGrid f, const MgGrid c,
                                                         const MgTables t,
                                                         const double* __restrict__ e,
                                                         double* __restrict__ p, int kz0) {
    const long long rows = (long long)f.ny * f.nz;
    for (long long row = blockIdx.x; row < rows; row += gridDim.x) {
        const int j = (int)(row % f.ny), k = (int)(row / f.ny);
        const int j0 = __ldg(t.c0[1] + j), k0 = __ldg(t.c0[2] + kz0 + k);
        int j1 = j0 + 1, k1 = k0 + 1;
        if (j1 >= c.ny) j1 = (c.my == BM_WRAP) ? 0 : c.ny - 1;
        if (k1 >= c.nz) k1 = (c.mz_lo == BM_WRAP) ? 0 : c.nz - 1;
        const double wy = __ldg(t.w[1] + j), wz = __ldg(t.w[2] + kz0 + k);
        const double* e00 = e + (long long)k0 * c.sz + (long long)j0 * c.sy;
        const double* e10 = e + (long long)k0 * c.sz + (long long)j1 * c.sy;
        const double* e01 = e + (long long)k1 * c.sz + (long long)j0 * c.sy;
        const double* e11 = e + (long long)k1 * c.sz + (long long)j1 * c.sy;
        const long long base = (long long)k * f.sz + (long long)j * f.sy;
        for (int i = threadIdx.x; i < f.nx; i += blockDim.x) {
            const int i0 = __ldg(t.c0[0] + i);
            int i1 = i0 + 1;
            if (i1 >= c.nx) i1 = (c.mx == BM_WRAP) ? 0 : c.nx - 1;
            const double wx = __ldg(t.w[0] + i);
            const double a00 = (1.0 - wx) * __ldg(e00 + i0) + wx * __ldg(e00 + i1);
            const double a10 = (1.0 - wx) * __ldg(e10 + i0) + wx * __ldg(e10 + i1);
            const double a01 = (1.0 - wx) * __ldg(e01 + i0) + wx * __ldg(e01 + i1);
            const double a11 = (1.0 - wx) * __ldg(e11 + i0) + wx * __ldg(e11 + i1);
            const double b0 = (1.0 - wy) * a00 + wy * a10;
            const double b1 = (1.0 - wy) * a01 + wy * a11;
            p[base + i] += (1.0 - wz) * b0 + wz * b1;
        }
    }
}

// Coarsest level in ONE CTA: (1) make the singular system compatible -- subtract the mean of rhs
// weighted by the left null vector (1/2 per mirrored wall the point lies on); (2) `sweeps`
// red-black Gauss-Seidel sweeps over the four (colour, seam-parity) classes.
constexpr int CNT = 1024;
__global__ void __launch_bounds__(CNT) mg_coarse_kernel(const MgGrid g, double* p, double* rhs,
                                                        int sweeps) {
    __shared__ double red[2][32];
    __shared__ double mean_s;
    const int tid = threadIdx.x;
    const long long n = (long long)g.nx * g.ny * g.nz;
    const int seam_x = (g.mx == BM_WRAP) && (g.nx & 1), seam_y = (g.my == BM_WRAP) && (g.ny & 1),
              seam_z = (g.mz_lo == BM_WRAP) && (g.nz & 1);
    double sw = 0.0, sr = 0.0;
    for (long long m = tid; m < n; m += CNT) {
        const int i = (int)(m % g.nx), j = (int)((m / g.nx) % g.ny), k = (int)(m / ((long long)g.nx * g.ny));
        double w = 1.0;
        if (g.mx == BM_MIRROR && (i == 0 || i == g.nx - 1)) w *= 0.5;
        if (g.my == BM_MIRROR && (j == 0 || j == g.ny - 1)) w *= 0.5;
        if (g.mz_lo == BM_MIRROR && (k == 0 || k == g.nz - 1)) w *= 0.5;
        sw += w;
        sr += w * rhs[(long long)k * g.sz + (long long)j * g.sy + i];
    }
    sw = warp_sum(sw), sr = warp_sum(sr);
    if ((tid & 31) == 0) red[0][tid >> 5] = sw, red[1][tid >> 5] = sr;
    __syncthreads();
    if (tid == 0) {
        double a = 0.0, b = 0.0;
        for (int q = 0; q < CNT / 32; ++q) a += red[0][q], b += red[1][q];
        mean_s = b / a;
    }
    __syncthreads();
    const double mean = mean_s;
    for (long long m = tid; m < n; m += CNT) {
        const int i = (int)(m % g.nx), j = (int)((m / g.nx) % g.ny), k = (int)(m / ((long long)g.nx * g.ny));
        rhs[(long long)k * g.sz + (long long)j * g.sy + i] -= mean;
    }
    __syncthreads();
    for (int s = 0; s < sweeps; ++s) {
        for (int cls = 0; cls < 4; ++cls) {
            const int colour = cls & 1, sp = cls >> 1;
            for (long long m = tid; m < n; m += CNT) {
                const int i = (int)(m % g.nx), j = (int)((m / g.nx) % g.ny),
                          k = (int)(m / ((long long)g.nx * g.ny));
                const int pop = (seam_x && i == g.nx - 1) + (seam_y && j == g.ny - 1) +
                                (seam_z && k == g.nz - 1);
                if (((i + j + k) & 1) != colour || (pop & 1) != sp) continue;
                int im, ip, jm, jp, km, kp;
                mg_nbr(i, g.nx, g.mx, g.mx, im, ip);
                mg_nbr(j, g.ny, g.my, g.my, jm, jp);
                mg_nbr(k, g.nz, g.mz_lo, g.mz_hi, km, kp);
                const long long kz = (long long)k * g.sz, jy = (long long)j * g.sy;
                volatile double* vp = p;
                const double s6 = g.ox * (vp[kz + jy + im] + vp[kz + jy + ip]) +
                                  g.oy * (vp[kz + (long long)jm * g.sy + i] +
                                          vp[kz + (long long)jp * g.sy + i]) +
                                  g.oz * (vp[(long long)km * g.sz + jy + i] +
                                          vp[(long long)kp * g.sz + jy + i]);
                vp[kz + jy + i] = (rhs[kz + jy + i] - s6) * g.invA;
            }
            __syncthreads();
        }
    }
}

int blocks_for(long long rows) {
    long long b = rows;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace

int launch_mg_residual(cudaStream_t st, const MgGrid& g, const double* p, const double* rhs,
                       double* res, unsigned long long* maxbits) {
    mg_residual_kernel<<<blocks_for((long long)g.ny * g.nz), 256, 0, st>>>(g, p, rhs, res, maxbits);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_mg_restrict(cudaStream_t st, const MgGrid& f, const MgGrid& c, const MgTables& t,
                       const double* res, double* rhs_c, double* p_c, int ck0, int nck) {
    if (nck <= 0) return 0;
    mg_restrict_kernel<<<blocks_for((long long)c.ny * nck), 256, 0, st>>>(f, c, t, res, rhs_c,
                                                                         p_c, ck0, nck);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_mg_prolong(cudaStream_t st, const MgGrid& f, const MgGrid& c, const MgTables& t,
                      const double* e, double* p, int kz0) {
    mg_prolong_kernel<<<blocks_for((long long)f.ny * f.nz), 256, 0, st>>>(f, c, t, e, p, kz0);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

int launch_mg_coarse(cudaStream_t st, const MgGrid& g, double* p, double* rhs, int sweeps) {
    mg_coarse_kernel<<<1, CNT, 0, st>>>(g, p, rhs, sweeps);
    count_launch();
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

}  // namespace o3d
