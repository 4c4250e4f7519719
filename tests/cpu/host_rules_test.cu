// host_rules_test.cu -- CPU checks of the __host__ __device__ index rules every kernel shares
// (csrc/o3d_common.cuh) and of the host-side launch planning (csrc/kernels.h).  Compiled by
// nvcc and run on the host by tests/test_host_rules_cpu.py; makes no CUDA call.
//
// The design rests on one invariant: PRODUCER kernels store the ghost images of every interior
// point they write (image_offsets), so that CONSUMER kernels can apply the interior stencil at
// every point without a ghost-fill pass.  That is only right if the set of images equals, cell
// for cell, what the ghost-fill rule (map_index: periodic wrap / mirror reflection, the closures
// of src/derivation.f90) would have put there.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <set>
#include <utility>

#include "../../osinco3d_b200/csrc/kernels.h"

using namespace o3d;

static int failures = 0;
#define CHECK(cond, ...)                          \
    do {                                          \
        if (!(cond)) {                            \
            ++failures;                           \
            if (failures < 20) {                  \
                printf("FAIL %s:%d: ", __FILE__, __LINE__); \
                printf(__VA_ARGS__);              \
                printf("\n");                     \
            }                                     \
        }                                         \
    } while (0)

// (ghost index q) -> (source interior index, reflected?) by the fill rule
static void images_match_fill_rule(int n, int mode_lo, int mode_hi) {
    std::map<int, int> by_fill;  // ghost cell -> source
    for (int g = 1; g <= R; ++g) {
        for (int side = 0; side < 2; ++side) {
            const int q = side ? n - 1 + g : -g;
            const int mode = side ? mode_hi : mode_lo;
            if (mode == BM_HALO) continue;  // filled by the neighbouring rank
            bool refl;
            const int src = map_index(q, n, mode_lo, mode_hi, refl);
            CHECK(src >= 0 && src < n, "n=%d modes=%d,%d ghost %d maps outside: %d", n, mode_lo, mode_hi, q, src);
            CHECK(refl == (mode == BM_MIRROR), "n=%d ghost %d: reflection flag", n, q);
            by_fill[q] = src;
        }
    }
    std::map<int, int> by_image;  // ghost cell -> producer point
    for (int p = 0; p < n; ++p) {
        const Img2 im = image_offsets(p, n, mode_lo, mode_hi);
        const int offs[2] = {im.lo, im.hi};
        for (int s = 0; s < 2; ++s) {
            if (!offs[s]) continue;
            const int q = p + offs[s];
            CHECK(q < 0 || q >= n, "n=%d p=%d: image %d is an interior cell", n, p, q);
            CHECK(q >= -R && q <= n - 1 + R, "n=%d p=%d: image %d beyond the ghost layers", n, p, q);
            CHECK(!by_image.count(q), "n=%d: ghost %d written by two points", n, q);
            by_image[q] = p;
        }
    }
    CHECK(by_fill == by_image, "n=%d modes=%d,%d: producer images != ghost-fill rule (%zu vs %zu cells)", n,
          mode_lo, mode_hi, by_image.size(), by_fill.size());
}

static void zchunks_cover_the_slab() {
    for (int tiles : {1, 12, 128, 256, 1024, 4096})
        for (int nz : {7, 8, 9, 16, 31, 81, 129, 255, 256, 512, 1024, 2041})
            for (int ctas : {1, 2, 3})
                for (int nfz : {1, 3})
                    for (int streams : {3, 4, 7, 15}) {
                        const int zc = pick_zchunk(tiles, nz, ctas, nfz, streams);
                        CHECK(zc >= 1 && zc <= nz, "zchunk %d for nz %d", zc, nz);
                        const int nch = (nz + zc - 1) / zc;
                        CHECK(nch >= 1 && nch <= 128, "nz %d: %d chunks", nz, nch);
                        // never thinner than 8 planes unless the slab itself is (the stencil
                        // window costs 6 extra planes per chunk)
                        CHECK(nch == 1 || zc >= 8, "nz %d tiles %d: chunks of %d planes", nz, tiles, zc);
                    }
}

static void layout_rules() {
    for (int nx : {7, 9, 16, 33, 241, 256, 257, 512, 1024}) {
        const int px = pitch_for(nx);
        CHECK(px % 16 == 0, "pitch %d of nx %d is not a multiple of 128 B", px, nx);
        CHECK(px >= GX + nx + R, "pitch %d too small for nx %d", px, nx);
        CHECK(px < GX + nx + R + 16, "pitch %d wastes more than a line for nx %d", px, nx);
    }
    Geom g;
    g.nx = 33, g.ny = 21, g.nz = 17;
    g.px = pitch_for(g.nx), g.py = g.ny + 2 * GH;
    g.sy = g.px, g.sz = (long long)g.px * g.py;
    CHECK(field_elems(g) == g.sz * (g.nz + 2 * GH), "field_elems");
    CHECK(interior_offset(g) == GX + g.sy * GH + g.sz * GH, "interior_offset");
    // the lowest ghost cell of the lowest ghost plane is still inside the allocation
    CHECK(interior_offset(g) - R - R * g.sy - R * g.sz >= 0, "ghost corner below the allocation");
    const long long last = interior_offset(g) + (g.nx - 1 + R) + (g.ny - 1 + R) * g.sy + (g.nz - 1 + R) * g.sz;
    CHECK(last < field_elems(g), "ghost corner beyond the allocation");
}

int main() {
    for (int n : {7, 8, 9, 10, 16, 33, 257})
        for (int lo : {BM_WRAP, BM_MIRROR, BM_HALO})
            for (int hi : {BM_WRAP, BM_MIRROR, BM_HALO}) {
                // an axis is periodic on both sides or on neither (src/initialization.f90:228-242);
                // BM_HALO replaces either side at a rank boundary, and both for a periodic axis
                // split over ranks
                if ((lo == BM_WRAP) != (hi == BM_WRAP) && lo != BM_HALO && hi != BM_HALO) continue;
                images_match_fill_rule(n, lo, hi);
            }
    zchunks_cover_the_slab();
    layout_rules();
    if (failures) {
        printf("%d check(s) failed\n", failures);
        return 1;
    }
    printf("host rules OK\n");
    return 0;
}
