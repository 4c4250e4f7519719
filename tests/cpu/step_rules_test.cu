// step_rules_test.cu -- the arithmetic of the fused RHS + nu_t + predictor kernel on the CPU.
// RhsEpi::apply (csrc/vel_kernels.cu) computes, per point, the nine first and nine second
// derivatives of (ux, uy, uz) with d1_expr / d2_expr on ghost cells of natural parity (component c
// odd along axis c: the parity table of src/integration.f90:118-165), then smagorinsky(),
// rhs_expr() and predictor_expr() of csrc/o3d_common.cuh.  This program performs the same
// sequence with the SAME functions on the host; tests/test_host_rules_cpu.py compares the result
// bit for bit with predict_velocity as executed from the reference source
// (tests/golden/hotpath.npz).  What it cannot cover is the kernels' data movement (TMA staging,
// rings, ghost images) -- that is what the GPU parity tests are for.
//
//   step_rules_test in.bin out.bin nx ny nz dx dy dz bx by bz sim2d iles re cs delta adu bdu cdu
//   in  = ux uy uz | f2x f2y f2z | f3x f3y f3z
//   out = upx upy upz | nu_t | f1x f1y f1z | Q | rotx roty rotz
// (Q = q_criterion_expr on the same natural-parity gradient, src/differential_operators.f90:79-108;
//  the curl as RotEpi forms it from the off-diagonal -- even -- derivatives, :64-73)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../osinco3d_b200/csrc/o3d_common.cuh"

namespace o3d {
void set_error(const char*, ...) {}
}  // namespace o3d
using namespace o3d;

int main(int argc, char** argv) {
    if (argc != 20) return 2;
    const int nx = atoi(argv[3]), ny = atoi(argv[4]), nz = atoi(argv[5]);
    const double dd[3] = {atof(argv[6]), atof(argv[7]), atof(argv[8])};
    const int bc[3] = {atoi(argv[9]), atoi(argv[10]), atoi(argv[11])};  // 0 periodic, 1 free-slip
    const int sim2d = atoi(argv[12]), iles = atoi(argv[13]);
    const double re = atof(argv[14]), cs = atof(argv[15]), delta = atof(argv[16]);
    const double adu = atof(argv[17]), bdu = atof(argv[18]), cdu = atof(argv[19]);
    const size_t N = (size_t)nx * ny * nz;
    std::vector<double> in(9 * N), out(11 * N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(in.data(), 8, 9 * N, fi) != 9 * N) return 3;
    fclose(fi);
    const double* u[3] = {&in[0], &in[N], &in[2 * N]};
    const double* f2[3] = {&in[3 * N], &in[4 * N], &in[5 * N]};
    const double* f3[3] = {&in[6 * N], &in[7 * N], &in[8 * N]};
    const int ext[3] = {nx, ny, nz};
    const size_t stride[3] = {1, (size_t)nx, (size_t)nx * ny};
    const Coef q[3] = {make_coef(dd[0]), make_coef(dd[1]), make_coef(dd[2])};
    const double onere = 1.0 / re;         // as o3d_s_predict_velocity (csrc/api.cu)
    const double csd = cs * delta;
    const double csd2 = csd * csd;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const size_t m = i + (size_t)nx * (j + (size_t)ny * k);
                const int p3[3] = {i, j, k};
                // ghost-cell read of component c at offset o along `axis`: natural parity
                auto at = [&](int c, int axis, int o) {
                    const int mode = bc[axis] ? BM_MIRROR : BM_WRAP;
                    bool refl;
                    const int src = map_index(p3[axis] + o, ext[axis], mode, mode, refl);
                    const double v = u[c][m + ((long long)src - p3[axis]) * (long long)stride[axis]];
                    return (refl && c == axis) ? -v : v;
                };
                Grad G;
                double lap[3][3];
                for (int c = 0; c < 3; ++c)
                    for (int a = 0; a < 3; ++a) {
                        const bool zero = (a == 2 && sim2d);  // derz_2dsim / derzz_2dsim
                        G.d[c][a] = zero ? 0.0
                                         : d1_expr(q[a].a1, q[a].b1, q[a].c1, at(c, a, -3), at(c, a, -2),
                                                   at(c, a, -1), at(c, a, 1), at(c, a, 2), at(c, a, 3));
                        lap[c][a] = zero ? 0.0
                                         : d2_expr(q[a].a2, q[a].b2, q[a].c2, at(c, a, -2), at(c, a, -1),
                                                   at(c, a, 0), at(c, a, 1), at(c, a, 2));
                    }
                out[7 * N + m] = q_criterion_expr(G);
                out[8 * N + m] = G.d[2][1] - G.d[1][2];   // duzdy - duydz
                out[9 * N + m] = G.d[0][2] - G.d[2][0];   // duxdz - duzdx
                out[10 * N + m] = G.d[1][0] - G.d[0][1];  // duydx - duxdy
                double nut = 0.0;
                if (iles) nut = smagorinsky(G, csd2);
                out[3 * N + m] = nut;
                const double nu_eff = onere + nut;  // src/integration.f90:114
                const double u0 = u[0][m], u1 = u[1][m], u2 = u[2][m];
                for (int c = 0; c < 3; ++c) {
                    const double f = rhs_expr(nu_eff, lap[c][0], lap[c][1], lap[c][2], u0, u1, u2,
                                              G.d[c][0], G.d[c][1], G.d[c][2]);
                    out[(4 + c) * N + m] = f;
                    out[c * N + m] = predictor_expr(u[c][m], adu, f, bdu, f2[c][m], cdu, f3[c][m]);
                }
            }
    FILE* fo = fopen(argv[2], "wb");
    if (!fo || fwrite(out.data(), 8, 11 * N, fo) != 11 * N) return 4;
    fclose(fo);
    printf("step rules written\n");
    return 0;
}
