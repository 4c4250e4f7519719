// sor_sweep_test.cu -- the reference's lexicographic SOR solve on the CPU with the PRODUCT's
// functions (csrc/sor_kernels.cu): nbr_idx (neighbour rule), sor_pnew_ref / sor_relax (point
// update of the verification ordering), sor_control_step (exit tests and dynamic omega,
// src/poisson.f90:110-122, the code the device runs after every sweep), fed the operator
// constants exactly as make_sor_args (csrc/poisson.cu) computes them.  tests/test_host_rules_cpu.py
// compares iterates, iteration count, omega and dmax bit for bit with poisson_solver_0000 / _0011 /
// _111111 as executed from the reference source (tests/golden/hotpath.npz).  On the GPU the same
// functions run hyperplane by hyperplane (sor_wavefront_kernel), which visits every point with
// exactly the old / new neighbours of this loop nest.
//
//   sor_sweep_test in.bin out.bin nx ny nz dx dy dz variant omega eps kmax idyn
//   in = pp | rhs      out = pp | iter (Fortran loop variable after the loop), omega, dmax
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace o3d {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace o3d

#include "../../osinco3d_b200/csrc/sor_kernels.cu"

using namespace o3d;

int main(int argc, char** argv) {
    if (argc != 14) return 2;
    const int nx = atoi(argv[3]), ny = atoi(argv[4]), nz = atoi(argv[5]);
    const double dx = atof(argv[6]), dy = atof(argv[7]), dz = atof(argv[8]);
    const int variant = atoi(argv[9]);
    const double omega0 = atof(argv[10]), eps = atof(argv[11]);
    const int kmax = atoi(argv[12]), idyn = atoi(argv[13]);
    const size_t N = (size_t)nx * ny * nz;
    std::vector<double> in(2 * N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(in.data(), 8, 2 * N, fi) != 2 * N) return 3;
    fclose(fi);
    double* pp = in.data();
    const double* rhs = in.data() + N;

    SorArgs a;  // as make_sor_args, csrc/poisson.cu (src/poisson.f90:42-51), unpadded strides
    const double dx2 = dx * dx, dy2 = dy * dy, dz2 = dz * dz;
    a.oneondx2 = 1.0 / dx2, a.oneondy2 = 1.0 / dy2, a.oneondz2 = 1.0 / dz2;
    const double twoondx2 = 2.0 * a.oneondx2, twoondy2 = 2.0 * a.oneondy2, twoondz2 = 2.0 * a.oneondz2;
    a.A = -(twoondx2 + twoondy2 + twoondz2);
    a.mx = (variant == 2) ? BM_MIRROR : BM_WRAP;
    a.my = (variant >= 1) ? BM_MIRROR : BM_WRAP;
    a.mz_lo = a.mz_hi = (variant == 2) ? BM_MIRROR : BM_WRAP;
    a.nx = nx, a.ny = ny, a.nz = nz;
    a.sy = nx, a.sz = (long long)nx * ny;
    const double factor = (variant == 1) ? 1.01 : 1.05;  // sor_factor, csrc/poisson.cu

    SorCtrl c;  // reset as sor_solve does
    c.dmax_bits = 0ull, c.omega = omega0, c.dmax_old = 1609.0, c.dmax_last = 0.0;
    c.iter = 0, c.done = (kmax < 1) ? 3 : 0;
    while (!c.done) {
        double dmax = 0.0;
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    int im1, ip1, jm1, jp1, km1, kp1;
                    nbr_idx(i, nx, a.mx, a.mx, im1, ip1);
                    nbr_idx(j, ny, a.my, a.my, jm1, jp1);
                    nbr_idx(k, nz, a.mz_lo, a.mz_hi, km1, kp1);
                    const long long m = (long long)k * a.sz + (long long)j * a.sy + i;
                    const double pw = pp[(long long)k * a.sz + (long long)j * a.sy + im1];
                    const double pe = pp[(long long)k * a.sz + (long long)j * a.sy + ip1];
                    const double ps = pp[(long long)k * a.sz + (long long)jm1 * a.sy + i];
                    const double pn = pp[(long long)k * a.sz + (long long)jp1 * a.sy + i];
                    const double pb = pp[(long long)km1 * a.sz + (long long)j * a.sy + i];
                    const double pt = pp[(long long)kp1 * a.sz + (long long)j * a.sy + i];
                    const double pc = pp[m];
                    const double p_new = sor_pnew_ref(a.oneondx2, a.oneondy2, a.oneondz2, pw, pe, ps,
                                                      pn, pb, pt, rhs[m], a.A);
                    dmax = fmax(dmax, fabs(p_new - pc));
                    pp[m] = sor_relax(c.omega, pc, p_new);
                }
        // the device accumulates max|p_new - p| through its bit pattern (monotone for x >= 0)
        union {
            double d;
            unsigned long long u;
        } v;
        v.d = dmax;
        c.dmax_bits = v.u;
        sor_control_step(&c, eps, kmax, idyn, factor);
    }
    const double scal[3] = {(double)((c.done == 3 || c.done == 0) ? kmax + 1 : c.iter), c.omega,
                            c.dmax_last};  // iters as reported by sor_solve
    FILE* fo = fopen(argv[2], "wb");
    if (!fo || fwrite(pp, 8, N, fo) != N || fwrite(scal, 8, 3, fo) != 3) return 4;
    fclose(fo);
    printf("sor sweep written\n");
    return 0;
}
