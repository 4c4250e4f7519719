// sor_classes_test.cu -- the race-freedom claim of the red-black SOR path, checked on the CPU with
// the kernels' own index helpers (csrc/sor_kernels.cu: nbr_idx = the neighbour rule of
// src/poisson.f90:57-92 / :185-218 / :310-344, seam_pop = the seam planes of odd periodic extents).
//
// A sweep updates the points of one class (colour = (i+j+k) mod 2, seam parity) in parallel and
// in place.  That is only deterministic -- and only a Gauss-Seidel sweep -- if no point of a class
// reads another point of the same class: every neighbour of a point must lie in a different class,
// for every boundary variant (_0000, _0011, _111111) and every parity of the extents.
#include <cstdarg>
#include <cstdio>

namespace o3d {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace o3d

#include "../../osinco3d_b200/csrc/sor_kernels.cu"

using namespace o3d;

static int failures = 0;

static int class_of(const SorArgs& a, int i, int j, int k) {
    return ((i + j + k) & 1) | ((seam_pop(a, i, j, k) & 1) << 1);
}

static void check(int nx, int ny, int nz, int variant) {
    SorArgs a;
    a.nx = nx, a.ny = ny, a.nz = nz, a.gnz = nz, a.gz0 = 0;
    // neighbour rule per variant, as make_sor_args (csrc/poisson.cu)
    a.mx = (variant == 2) ? BM_MIRROR : BM_WRAP;
    a.my = (variant >= 1) ? BM_MIRROR : BM_WRAP;
    a.mz_lo = a.mz_hi = (variant == 2) ? BM_MIRROR : BM_WRAP;
    a.seam_x = (a.mx == BM_WRAP) && (nx & 1);
    a.seam_y = (a.my == BM_WRAP) && (ny & 1);
    a.seam_z = (a.mz_lo == BM_WRAP) && (nz & 1);
    long long per_class[4] = {0, 0, 0, 0};
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int c = class_of(a, i, j, k);
                ++per_class[c];
                int m1, p1;
                int nb[6][3];
                nbr_idx(i, nx, a.mx, a.mx, m1, p1);
                nb[0][0] = m1, nb[0][1] = j, nb[0][2] = k;
                nb[1][0] = p1, nb[1][1] = j, nb[1][2] = k;
                nbr_idx(j, ny, a.my, a.my, m1, p1);
                nb[2][0] = i, nb[2][1] = m1, nb[2][2] = k;
                nb[3][0] = i, nb[3][1] = p1, nb[3][2] = k;
                nbr_idx(k, nz, a.mz_lo, a.mz_hi, m1, p1);
                nb[4][0] = i, nb[4][1] = j, nb[4][2] = m1;
                nb[5][0] = i, nb[5][1] = j, nb[5][2] = p1;
                for (int q = 0; q < 6; ++q) {
                    const int ii = nb[q][0], jj = nb[q][1], kk = nb[q][2];
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) {
                        if (++failures < 10)
                            printf("FAIL %dx%dx%d v%d: neighbour (%d,%d,%d) of (%d,%d,%d) outside\n", nx,
                                   ny, nz, variant, ii, jj, kk, i, j, k);
                        continue;
                    }
                    if (class_of(a, ii, jj, kk) == c && ++failures < 10)
                        printf("FAIL %dx%dx%d v%d: (%d,%d,%d) and its neighbour (%d,%d,%d) share class %d\n",
                               nx, ny, nz, variant, i, j, k, ii, jj, kk, c);
                }
            }
    // the seam classes only exist on odd periodic extents
    if (!(a.seam_x || a.seam_y || a.seam_z) && (per_class[2] || per_class[3]) && ++failures < 10)
        printf("FAIL %dx%dx%d v%d: seam classes on a 2-colourable grid\n", nx, ny, nz, variant);
}

int main() {
    const int ext[] = {7, 8, 9, 12, 13};
    for (int variant = 0; variant < 3; ++variant)
        for (int nx : ext)
            for (int ny : ext)
                for (int nz : ext) check(nx, ny, nz, variant);
    if (failures) {
        printf("%d check(s) failed\n", failures);
        return 1;
    }
    printf("sor classes OK\n");
    return 0;
}
