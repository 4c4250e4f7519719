// scalar_rules_test.cu -- the arithmetic of the divergence, projection-correction and
// scalar-transport kernels on the CPU, with the SAME point-wise functions the kernels call
// (csrc/o3d_common.cuh: d1_expr / d2_expr on ghost cells via map_index, div_expr, corr_expr,
// transeq_alpha, transeq_rhs_expr, predictor_expr, clip01, transeq_weight,
// transeq_redistribute).  tests/test_host_rules_cpu.py compares the results bit for bit with
// divergence / correct_velocity / transeq as executed from the reference source
// (tests/golden/hotpath.npz).  The three global sums of transeq are accumulated here in
// array-element order like the reference's sum(); on the GPU they are tree reductions, which is
// why the device result is compared to a tolerance instead.
//
//   scalar_rules_test div    in out nx ny nz dx dy dz bx by bz sim2d odd
//       in = fx fy fz                      out = divf
//   scalar_rules_test corr   in out nx ny nz dx dy dz bx by bz sim2d dt
//       in = upx upy upz pp                out = ux uy uz
//   scalar_rules_test transeq in out nx ny nz dx dy dz bx by bz sim2d iles re sc adu bdu cdu
//       in = phi ux uy uz nu_t f2 f3       out = phi f1
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../osinco3d_b200/csrc/o3d_common.cuh"

namespace o3d {
void set_error(const char*, ...) {}
}  // namespace o3d
using namespace o3d;

static int nx, ny, nz, bc[3], sim2d;
static Coef q[3];

static size_t idx(int i, int j, int k) { return i + (size_t)nx * (j + (size_t)ny * k); }

// ghost-cell read along `axis` at offset o from point (i,j,k); odd: mirrored copies change sign
static double at(const double* f, int i, int j, int k, int axis, int o, bool odd) {
    const int ext[3] = {nx, ny, nz};
    int p[3] = {i, j, k};
    const int mode = bc[axis] ? BM_MIRROR : BM_WRAP;
    bool refl;
    p[axis] = map_index(p[axis] + o, ext[axis], mode, mode, refl);
    const double v = f[idx(p[0], p[1], p[2])];
    return (refl && odd) ? -v : v;
}
static double d1(const double* f, int i, int j, int k, int a, bool odd) {
    if (a == 2 && sim2d) return 0.0;
    return d1_expr(q[a].a1, q[a].b1, q[a].c1, at(f, i, j, k, a, -3, odd), at(f, i, j, k, a, -2, odd),
                   at(f, i, j, k, a, -1, odd), at(f, i, j, k, a, 1, odd), at(f, i, j, k, a, 2, odd),
                   at(f, i, j, k, a, 3, odd));
}
static double d2(const double* f, int i, int j, int k, int a, bool odd) {
    if (a == 2 && sim2d) return 0.0;
    return d2_expr(q[a].a2, q[a].b2, q[a].c2, at(f, i, j, k, a, -2, odd), at(f, i, j, k, a, -1, odd),
                   at(f, i, j, k, a, 0, odd), at(f, i, j, k, a, 1, odd), at(f, i, j, k, a, 2, odd));
}

int main(int argc, char** argv) {
    if (argc < 14) return 2;
    const char* mode = argv[1];
    nx = atoi(argv[4]), ny = atoi(argv[5]), nz = atoi(argv[6]);
    for (int a = 0; a < 3; ++a) q[a] = make_coef(atof(argv[7 + a])), bc[a] = atoi(argv[10 + a]);
    sim2d = atoi(argv[13]);
    const size_t N = (size_t)nx * ny * nz;
    const int nin = !strcmp(mode, "div") ? 3 : !strcmp(mode, "corr") ? 4 : 7;
    const int nout = !strcmp(mode, "div") ? 1 : !strcmp(mode, "corr") ? 3 : 2;
    std::vector<double> in(nin * N), out(nout * N);
    FILE* fi = fopen(argv[2], "rb");
    if (!fi || fread(in.data(), 8, nin * N, fi) != nin * N) return 3;
    fclose(fi);
    if (!strcmp(mode, "div")) {
        const int odd = atoi(argv[14]);
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i)  // derxi / deryi / derzi (odd) or derxp / deryp / derzp
                    out[idx(i, j, k)] = div_expr(d1(&in[0], i, j, k, 0, odd), d1(&in[N], i, j, k, 1, odd),
                                                 d1(&in[2 * N], i, j, k, 2, odd), 0, 1.0);
    } else if (!strcmp(mode, "corr")) {
        const double dt = atof(argv[14]);
        const double* pp = &in[3 * N];
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i)
                    for (int c = 0; c < 3; ++c)  // derxp / deryp / derzp of pp
                        out[c * N + idx(i, j, k)] = corr_expr(in[c * N + idx(i, j, k)], dt, d1(pp, i, j, k, c, false));
    } else {
        const int iles = atoi(argv[14]);
        const double re = atof(argv[15]), sc = atof(argv[16]);
        const double adu = atof(argv[17]), bdu = atof(argv[18]), cdu = atof(argv[19]);
        const double resc = re * sc;  // as o3d_s_transeq (csrc/api.cu)
        const double *phi = &in[0], *u0 = &in[N], *u1 = &in[2 * N], *u2 = &in[3 * N];
        const double *nut = &in[4 * N], *f2 = &in[5 * N], *f3 = &in[6 * N];
        double* pn = &out[0];
        double* f1 = &out[N];
        double s_old = 0.0, s_clip = 0.0, s_w = 0.0;
        for (int k = 0; k < nz; ++k)
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const size_t m = idx(i, j, k);
                    const double alpha = transeq_alpha(resc, nut[m], sc, iles);
                    const double f = transeq_rhs_expr(alpha, d2(phi, i, j, k, 0, false), d2(phi, i, j, k, 1, false),
                                                      d2(phi, i, j, k, 2, false), u0[m], u1[m], u2[m],
                                                      d1(phi, i, j, k, 0, false), d1(phi, i, j, k, 1, false),
                                                      d1(phi, i, j, k, 2, false), 0.0);
                    f1[m] = f;
                    pn[m] = predictor_expr(phi[m], adu, f, bdu, f2[m], cdu, f3[m]);
                    s_old += pn[m];
                    const double pc = clip01(pn[m]);
                    s_clip += pc;
                    s_w += transeq_weight(pc);
                }
        const double count = (double)((long long)nx * ny * nz);
        const double excess = s_old / count - s_clip / count;  // as transeq_clip_kernel
        for (size_t m = 0; m < N; ++m) pn[m] = transeq_redistribute(clip01(pn[m]), excess, s_w);
    }
    FILE* fo = fopen(argv[3], "wb");
    if (!fo || fwrite(out.data(), 8, nout * N, fo) != nout * N) return 4;
    fclose(fo);
    printf("scalar rules written\n");
    return 0;
}
